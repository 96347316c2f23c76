// Forward selective scan, tiled path for sm_100a: persistent CTAs, TMA-staged shared-memory tiles, an mbarrier ring and
// packed f32x2 arithmetic.  Replaces selective_scan_fwd_kernel
// (/root/reference/mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303, incl. the SiLU(z) gate :280-298) for the
// aligned shapes Mamba-UNet produces (channels per group a multiple of 64, 16-byte aligned rows); everything else takes
// selscan_fwd.cu.
//
// Work item = 64 channels of one (batch, group) x the whole sequence (or one segment of it), walked in tiles of 32 positions.
// The grid is persistent (at most kCtasPerSm CTAs per SM); CTA i takes items i, i + grid, ...  The producer runs ahead of the
// consumers across item boundaries, so that short sequences (stage 4: L = 49 = two tiles) do not pay a pipeline fill per item.
//   warp 4 (producer): per tile, one elected lane issues the TMA loads (u, delta [, z]: box 64 rows x 32 positions, 128-byte
//       swizzle) into a ring; all 32 lanes gather the tile's B and C values (any strides: the (N, L) layout and the l-major
//       x_dbl layout both coalesce) into a [position][B0..15 C0..15] tile.
//   warps 0-3 (consumers): 16 channels each, TWO lanes per channel (lane = channel + 16 * state half, 8 states per lane as four
//       packed pairs).  Round 1 used four lanes x four states: every per-element scalar (delta, delta*u, their packing), the
//       B/C loads and the partial-sum shuffles were paid four times per element; two lanes halve all of that (135 -> ~90
//       thread-instructions per element, 1.47 -> ~0.8 shared-memory wavefronts) and the 16 lanes of a half warp read the SAME
//       B/C address (2 distinct 16-byte chunks per warp instruction: 2 wavefronts instead of 4).  softplus(delta + bias) is
//       evaluated once per element (lane half h takes positions 2h, 2h+1 of a quad) and exchanged with the partner lane; the
//       two partial y of a position are reduce-scattered over the pair (one shuffle per position); outputs are staged in a
//       swizzled 16 x 32 tile per warp and written back by a per-warp TMA store, so HBM only ever sees full 128-byte rows.
//       The recurrence is thread-serial: per position and state pair one FMUL2, two MUFU.EX2, one FMUL2 and two FFMA2.
//       The state after every 8th position (what the backward restarts from) is staged in two swizzled half tiles per warp and
//       leaves as 32-float x 16-row TMA boxes, half a tile apart: as 16-byte global stores, one row per lane, it was 32 distinct
//       lines per warp instruction and 10-17 % of the kernel (profiles/r02_bwd_whatif.json -> forward).
// Full/empty mbarriers per stage are the only synchronisation; all waits are bounded (trap instead of hang).
//
// Low-parallelism shapes (small batch: the reference validates slice by slice, val_2D.py:35-47) split the SEQUENCE into
// segments so that the whole chip works on a call: pass A (kMode 1) runs every segment from a zero state and keeps only
// its end state and sum of delta; a tiny combine kernel turns those into the true state at every segment start
// (h_s = exp2(A2 * sum_delta_{s-1}) * h_{s-1} + xend_{s-1}); pass C (kMode 2) re-runs the segments from those states and
// produces the outputs.  1.7x the arithmetic for n_segs x the parallelism; results identical to rounding.
#include <atomic>
#include <type_traits>

#include "selscan_common.cuh"
#include "selscan_kernels.h"
#include "selscan_ptx.cuh"
#include "selscan_tma_host.h"

#ifndef SELSCAN_FWD_STAGES
#define SELSCAN_FWD_STAGES 3
#endif
#ifndef SELSCAN_FWD_CTAS
#define SELSCAN_FWD_CTAS 2
#endif

namespace selscan {

namespace {

constexpr int kTL = 32;          // positions per tile (128-byte rows)
constexpr int kRows = 64;        // channels per work item
constexpr int kConsWarps = 4;    // 16 channels each, 2 lanes per channel
constexpr int kWarpRows = kRows / kConsWarps;
constexpr int kBCPitch = 36;     // floats per position in the B/C tile (32 + pad, keeps 16-byte alignment)
constexpr int kThreads = (kConsWarps + 1) * 32;
constexpr int kLaneStates = kStatePad / 2;  // states per lane
constexpr int kTileBytes = kRows * kTL * 4;          // 8 KB
constexpr int kOutBytes = kWarpRows * kTL * 4;       // 2 KB
// Saved states ("checkpoints", one per 8 positions) are staged in shared memory and leave by TMA: 32 scattered 16-byte global stores per
// warp instruction cost the forward 10-17 % (what-if build without them, round 2).  Two half buffers per warp (saved states 0-1 / 2-3 of
// a tile), each [16 rows][2 x 16 states] = 128-byte rows with the 128-byte swizzle (a quarter-warp = 8 rows x one 16-byte chunk
// touches all banks), one 32-float x 16-row box each.
constexpr int kCkHalfBytes = kWarpRows * 2 * kStatePad * 4;   // 2 KB

// shared-memory carve-up (bytes from a 1024-byte aligned base; every TMA tile is a multiple of 1024 bytes)
// kDt > 0: fused dt_proj -- the stage carries a [kDt ranks][32 positions] tile of x_dbl's dt rows instead of a delta tile
template <bool kHasZ, int kDt = 0>
struct Lay {
  static constexpr int kStages = kHasZ ? 2 : SELSCAN_FWD_STAGES;
  static constexpr int kCtas = kHasZ ? 2 : SELSCAN_FWD_CTAS;
  static constexpr int kIn = kHasZ ? 3 : 2;      // input tiles per stage: u, delta [, z]
  static constexpr int kOut = kHasZ ? 2 : 1;     // output tiles per buffer: out [, out_z]
  static constexpr uint32_t oIn = 0;                                            // [stage][which]
  static constexpr uint32_t oOut = oIn + kStages * kIn * kTileBytes;           // [warp][buffer][which]
  static constexpr uint32_t oCk = oOut + kConsWarps * 2 * kOut * kOutBytes;    // [warp][half][16 rows][128 bytes]  (TMA, 1024-byte aligned)
  static constexpr uint32_t oBC = oCk + kConsWarps * 2 * kCkHalfBytes;         // [stage][position][36]
  static constexpr uint32_t oDt = oBC + kStages * kTL * kBCPitch * 4;          // [stage][rank][32 positions] (TMA, dense 128-byte rows)
  static constexpr uint32_t kDtBytes = kDt * kTL * 4;
  static constexpr uint32_t oBar = oDt + kStages * kDtBytes;                   // full[stage], empty[stage]
  static constexpr uint32_t kBytes = oBar + 2 * kStages * 8;
};

struct Item {
  int b, g, d0, seg, t_begin, n_tiles;
  int ds0;    // first row of this item in u / out (mirrored pairs: the even group's rows)
  bool rev;   // mirrored pairs: odd group, walks the rows back to front
};

template <int kMode>
__device__ __forceinline__ Item decode_item(const FwdLaunch& p, int item) {
  Item it;
  const int tiles_per_group = p.dim_per_group / kRows;
  int bid = item;
  it.seg = 0;
  if (kMode != 0) {
    it.seg = bid % p.n_segs;
    bid /= p.n_segs;
  }
  const int tile_g = bid % tiles_per_group;
  bid /= tiles_per_group;
  it.g = bid % p.a.ngroups;
  it.b = bid / p.a.ngroups;
  it.d0 = it.g * p.dim_per_group + tile_g * kRows;
  it.rev = p.a.mirror_pairs != 0 && (it.g & 1);
  it.ds0 = p.a.mirror_pairs != 0 ? (it.g >> 1) * p.dim_per_group + tile_g * kRows : it.d0;
  const int n_tiles_all = (p.a.seqlen + kTL - 1) / kTL;
  it.t_begin = (kMode == 0) ? 0 : it.seg * p.seg_tiles;
  const int t_end = (kMode == 0) ? n_tiles_all : min(n_tiles_all, it.t_begin + p.seg_tiles);
  it.n_tiles = t_end - it.t_begin;
  return it;
}

// kMode 0: whole sequence per item.  1: segment aggregates only (no outputs).  2: segment with an initial state.
// kMir: mirrored direction pairs (selscan_b200.h: mirror_pairs): odd groups read the even group's u rows back to front -- mirrored
// TMA coordinates, mirrored tile columns in the consumers -- and both groups of a pair ADD their outputs into the same rows
template <int kMode, bool kHasZ, int kDt, bool kMir>
__global__ void __launch_bounds__(kThreads, Lay<kHasZ, kDt>::kCtas)
selscan_fwd_tma_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_dt,
                       const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_out,
                       const __grid_constant__ CUtensorMap map_outz, const __grid_constant__ CUtensorMap map_ck, const int ck_tma,
                       const FwdLaunch p, const int n_items) {
  // ck_tma (kernel-uniform): the saved states leave by TMA (map_ck: (n_ckpt * 16 floats, batch * dim rows), box 32 x 16, 128-byte swizzle)
  // kDt > 0: map_dt is the 4-D map over dt_x (seqlen, rank, group, batch) and the raw step is formed here (mamba_sys.py:409)
  using LY = Lay<kHasZ, kDt>;
  constexpr int kStages = LY::kStages;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sm0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const selscan_fwd_args& a = p.a;
  const int L = a.seqlen, N = a.dstate;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto full_bar = [&](int s) { return sm0 + LY::oBar + (uint32_t)s * 8u; };
  auto empty_bar = [&](int s) { return sm0 + LY::oBar + (uint32_t)(kStages + s) * 8u; };
  auto in_tile = [&](int s, int which) { return sm0 + LY::oIn + (uint32_t)(s * LY::kIn + which) * kTileBytes; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 32);            // the 32 producer lanes (+ the TMA transaction bytes)
      mbar_init(empty_bar(s), kConsWarps);   // one arrival per consumer warp
    }
    mbar_fence_init();
    tma_prefetch_desc(&map_u);
    tma_prefetch_desc(&map_dt);
    if (kMode != 1) tma_prefetch_desc(&map_out);
    if (kMode != 1 && ck_tma) tma_prefetch_desc(&map_ck);
    if (kHasZ) {
      tma_prefetch_desc(&map_z);
      tma_prefetch_desc(&map_outz);
    }
  }
  __syncthreads();

  if (warp == kConsWarps) {
    // ================================ producer ================================
    const bool lanes_along_l = (a.B_l_stride == 1 && a.C_l_stride == 1);
    const bool small_n_strides = a.B_n_stride >= 0 && a.B_n_stride < (1 << 26) && a.C_n_stride >= 0 && a.C_n_stride < (1 << 26);   // 16 * stride fits an int
    uint32_t it = 0;   // tiles issued by this CTA: ring position
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const Item w = decode_item<kMode>(p, item);
      const float* __restrict__ Bg = a.B + (int64_t)w.b * a.B_batch_stride + (int64_t)w.g * a.B_group_stride;
      const float* __restrict__ Cg = a.C + (int64_t)w.b * a.C_batch_stride + (int64_t)w.g * a.C_group_stride;
      for (int t = 0; t < w.n_tiles; ++t, ++it) {
        const int s = it % kStages;
        const uint32_t k = it / kStages;
        if (k > 0) mbar_wait(empty_bar(s), (k - 1) & 1);
        const int l0 = (w.t_begin + t) * kTL;
        const bool rev = kMir && w.rev;
        // mirrored: the tile's source positions.  TMA coordinates are kept non-negative: the partial tile of a mirrored item loads the
        // row's first 32 positions and the consumers shift their columns by the missing count (a multiple of 4: seqlen % 4 == 0)
        const int lc = rev ? max(L - l0 - kTL, 0) : l0;
        const uint32_t full = full_bar(s);
        if (lane == 0) {
          if (kDt > 0) {
            mbar_expect_tx(full, (uint32_t)kTileBytes + LY::kDtBytes);
            tma_load_3d(in_tile(s, 0), &map_u, lc, kMir ? w.ds0 : w.d0, w.b, full);
            tma_load_4d(sm0 + LY::oDt + (uint32_t)s * LY::kDtBytes, &map_dt, lc, 0, w.g, w.b, full);   // ranks >= dt_rank: zero fill
          } else {
            mbar_expect_tx(full, (uint32_t)((kHasZ && kMode != 1) ? 3 : 2) * kTileBytes);
            tma_load_3d(in_tile(s, 0), &map_u, lc, kMir ? w.ds0 : w.d0, w.b, full);
            tma_load_3d(in_tile(s, 1), &map_dt, lc, w.d0, w.b, full);
          }
          if (kHasZ && kMode != 1) tma_load_3d(in_tile(s, 2), &map_z, l0, w.d0, w.b, full);
        }
        const uint32_t bc = sm0 + LY::oBC + (uint32_t)s * (kTL * kBCPitch * 4);
        float v[32];
        if (lanes_along_l && N == kStatePad && l0 + kTL <= L && !rev && small_n_strides) {
          // the common case (all 16 states, a full tile, forward direction): no bounds or direction arithmetic, one 32-bit
          // multiply-add per address -- this warp's gather cost the forward ~6 % (what-if build), most of it instructions on the
          // consumers' scheduler
          const float* pb = Bg + l0 + lane;
          const float* pc = Cg + l0 + lane;
          const int sb = (int)a.B_n_stride, sc = (int)a.C_n_stride;
#pragma unroll
          for (int n = 0; n < 16; ++n) {
            v[n] = __ldg(pb + n * sb);
            v[16 + n] = __ldg(pc + n * sc);
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            sts_f4(bc + (uint32_t)(lane * kBCPitch + i) * 4, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        } else if (lanes_along_l) {  // (.., N, L) layout: a warp reads 128 contiguous bytes of one state row
          const int l = l0 + lane;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int n = i & 15;
            const float* src = (i < 16) ? (Bg + (int64_t)n * a.B_n_stride) : (Cg + (int64_t)n * a.C_n_stride);
            v[i] = (n < N && l < L) ? __ldg(src + (rev ? L - 1 - l : l)) : 0.f;
          }
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            sts_f4(bc + (uint32_t)(lane * kBCPitch + i) * 4, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
        } else {              // l-major layout (x_dbl): a warp reads the 16 B and 16 C values of one position
          const int n = lane & 15;
          const float* src = (lane < 16) ? (Bg + (int64_t)n * a.B_n_stride) : (Cg + (int64_t)n * a.C_n_stride);
          const int64_t ls = (lane < 16) ? a.B_l_stride : a.C_l_stride;
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (n < N && l0 + j < L) ? __ldg(src + (int64_t)(rev ? L - 1 - (l0 + j) : l0 + j) * ls) : 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(bc + (uint32_t)(j * kBCPitch + lane) * 4), "f"(v[j]) : "memory");
        }
        mbar_arrive(full);
      }
    }
    return;
  }

  // ================================ consumers ================================
  const int ch = lane & 15;            // channel inside the warp
  const int h = lane >> 4;             // which 8 states; also: which 2 positions of a quad I discretise and finally own
  const bool hi = h != 0;
  const int rr = warp * kWarpRows + ch;   // channel inside the item
  const uint32_t swz = (uint32_t)(rr & 7) << 4;   // 128B swizzle: 16-byte chunk index ^= row & 7 (same key in my 16-row out tile)
  const bool softplus = a.delta_softplus != 0;
  const bool A_vec = (a.A_n_stride == 1) && ((a.A_d_stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15u) == 0) && N == kStatePad;

  struct Params {
    u64 A2p[4];
    float Dv, bias;
    float Wd[kDt > 0 ? kDt : 1];   // my channel's row of the dt_proj weight (zero beyond dt_rank)
  };
  auto load_params = [&](const Item& w, Params& q) {
    const int d = w.d0 + rr;
    float av[8];
    if (A_vec) {
      const float4 v0 = ldg4(a.A + (int64_t)d * a.A_d_stride + h * kLaneStates);
      const float4 v1 = ldg4(a.A + (int64_t)d * a.A_d_stride + h * kLaneStates + 4);
      av[0] = v0.x; av[1] = v0.y; av[2] = v0.z; av[3] = v0.w;
      av[4] = v1.x; av[5] = v1.y; av[6] = v1.z; av[7] = v1.w;
    } else {
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const int nn = h * kLaneStates + n;
        av[n] = (nn < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)nn * a.A_n_stride) : 0.f;
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) q.A2p[n] = pk2(av[2 * n] * kLog2e, av[2 * n + 1] * kLog2e);
    q.Dv = a.D ? __ldg(a.D + d) : 0.f;
    q.bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.f;
    if (kDt > 0) {
#pragma unroll
      for (int r = 0; r < kDt; ++r) q.Wd[r] = (r < a.dt_rank) ? __ldg(a.dt_w + (int64_t)d * a.dt_w_d_stride + r) : 0.f;
    }
  };

  uint32_t it = 0;        // tiles consumed by this CTA: ring position
  uint32_t n_out = 0;     // output tiles written by this warp: double-buffer index
  auto out_tile_of = [&](int wp, uint32_t n) { return sm0 + LY::oOut + (uint32_t)((wp * 2 + (n & 1)) * LY::kOut) * kOutBytes; };
  Params cur, nxt;
  Item w = decode_item<kMode>(p, blockIdx.x < n_items ? blockIdx.x : 0);
  if (blockIdx.x < n_items) load_params(w, cur);
  nxt = cur;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int d = w.d0 + rr;
    const int64_t row = (int64_t)w.b * a.dim + d;
    u64 x2[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) x2[q] = pk2(0.f, 0.f);
    float* __restrict__ ws_state = nullptr;   // [row][segment][16]: end states (pass A) -> start states (after the combine)
    if (kMode != 0) {
      ws_state = p.seg_ws + (row * p.n_segs + w.seg) * kStatePad + h * kLaneStates;
      if (kMode == 2) {
        const float4 h0 = *reinterpret_cast<const float4*>(ws_state);
        const float4 h1 = *reinterpret_cast<const float4*>(ws_state + 4);
        x2[0] = pk2(h0.x, h0.y); x2[1] = pk2(h0.z, h0.w);
        x2[2] = pk2(h1.x, h1.y); x2[3] = pk2(h1.z, h1.w);
      }
    }
    float sum_delta = 0.f;
    float* __restrict__ ck = (kMode != 1 && a.ckpt) ? a.ckpt + row * p.n_ckpt * kStatePad + h * kLaneStates : nullptr;
    const bool ckt = kMode != 1 && ck_tma != 0;     // saved states through the staging tiles + TMA (else: the global stores below)
    const uint32_t ck_stage = sm0 + LY::oCk + (uint32_t)warp * (2 * kCkHalfBytes) + (uint32_t)ch * (2 * kStatePad * 4);   // my row; + half * kCkHalfBytes
    const uint32_t ck_key = (uint32_t)(ch & 7) << 4;   // 128-byte swizzle of the staging tiles
    const int ck_row0 = (int)((int64_t)w.b * a.dim + w.d0 + warp * kWarpRows);
    const float Dv = cur.Dv, bias = cur.bias;
    const int next_item = item + (int)gridDim.x;
    Item wn = w;
    if (next_item < n_items) wn = decode_item<kMode>(p, next_item);

    for (int t = 0; t < w.n_tiles; ++t, ++it) {
      if (t == w.n_tiles - 1 && next_item < n_items) load_params(wn, nxt);   // in flight while the item's last tile is processed
      const int s = it % kStages;
      const uint32_t k = it / kStages;
      const int l0 = (w.t_begin + t) * kTL;
      mbar_wait(full_bar(s), k & 1);
      const uint32_t u_row = in_tile(s, 0) + rr * (kTL * 4);
      const uint32_t dt_row = in_tile(s, 1) + rr * (kTL * 4);
      const uint32_t z_row = in_tile(s, kHasZ ? 2 : 0) + rr * (kTL * 4) + h * 8;
      const uint32_t bc_base = sm0 + LY::oBC + (uint32_t)s * (kTL * kBCPitch * 4) + h * (kLaneStates * 4);
      const uint32_t xdt_base = sm0 + LY::oDt + (uint32_t)s * LY::kDtBytes;
      const bool rev = kMir && w.rev;          // mirrored item: scan position p of the tile sits in column 31 - p - miss
      const int miss = rev ? max(kTL - (L - l0), 0) : 0;   // positions the (last, partial) tile lacks; its box starts at source 0
      const int qmiss = miss >> 2;
      if (kMir && miss > 0) {                  // columns of the out tile past the sequence must add zeros: clear the tile first
#pragma unroll
        for (int i = 0; i < kOutBytes / (32 * 16); ++i) sts_f4(out_tile_of(warp, n_out) + (uint32_t)(lane + 32 * i) * 16, make_float4(0.f, 0.f, 0.f, 0.f));
        __syncwarp();
      }
      const uint32_t out_tile = out_tile_of(warp, n_out);
      const uint32_t out_row = out_tile + ch * (kTL * 4) + h * 8;
      const uint32_t out_row_rev = out_tile + ch * (kTL * 4) + (1 - h) * 8;
      // Software pipeline over the 8 quads of the tile: while quad q's recurrence runs, quad q+1's row data is loaded,
      // discretised and exchanged, and quad q-1's partial sums are reduce-scattered and stored.
      float uv_n[4], dl_n[4];              // quad q+1 (prefetched)
      float yp[4], up0 = 0.f, up1 = 0.f;   // quad q-1: partial sums and my two u values
      auto prefetch = [&](int q, float (&uv)[4], float (&dl)[4]) {
        const uint32_t qsrc = (uint32_t)(rev ? max(7 - q - qmiss, 0) : q) << 4;   // (clamped: a partial tile prefetches one quad past its end)
        const float4 u4 = lds_f4(u_row + (qsrc ^ swz));
        uv[0] = rev ? u4.w : u4.x; uv[1] = rev ? u4.z : u4.y; uv[2] = rev ? u4.y : u4.z; uv[3] = rev ? u4.x : u4.w;
        // my two positions of the quad: discretise delta once per element, then exchange with the partner lane
        float m0, m1;
        if (kDt > 0) {   // delta = dt_w[d, :] . dt_x[:, l] + bias: rank-R expansion in registers, no (batch, dim, seqlen) step tensor
          m0 = m1 = bias;
#pragma unroll
          for (int r = 0; r < kDt; ++r) {
            float x0, x1;   // my two positions of the quad (mirrored: columns 30 - p, 31 - p, swapped)
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(x0), "=f"(x1)
                         : "r"(xdt_base + (uint32_t)(r * (kTL * 4)) + (rev ? (uint32_t)max(120 - 16 * q - 8 * h - 4 * miss, 0) : (uint32_t)(16 * q + 8 * h))));
            m0 = fmaf(cur.Wd[r], rev ? x1 : x0, m0);
            m1 = fmaf(cur.Wd[r], rev ? x0 : x1, m1);
          }
        } else {
          const float4 d4 = lds_f4(dt_row + (qsrc ^ swz));
          m0 = (rev ? (hi ? d4.y : d4.w) : (hi ? d4.z : d4.x)) + bias;
          m1 = (rev ? (hi ? d4.x : d4.z) : (hi ? d4.w : d4.y)) + bias;
        }
        {   // branch-free on the (uniform) softplus flag: a branch here would split the tile into basic blocks and
            // serialise this latency chain (LDS -> EX2 -> RCP -> polynomial -> SHFL) against the recurrence
          float w_unused;
          const float s0 = softplus_fast(m0, w_unused), s1 = softplus_fast(m1, w_unused);
          m0 = softplus ? s0 : m0;
          m1 = softplus ? s1 : m1;
        }
        const int pos = l0 + 4 * q + 2 * h;
        m0 = (pos < L) ? m0 : 0.f;           // past the end: a = 1, b = 0
        m1 = (pos + 1 < L) ? m1 : 0.f;
        const float o0 = __shfl_xor_sync(0xffffffffu, m0, 16), o1 = __shfl_xor_sync(0xffffffffu, m1, 16);
        dl[0] = hi ? o0 : m0; dl[1] = hi ? o1 : m1;
        dl[2] = hi ? m0 : o0; dl[3] = hi ? m1 : o1;
      };
      auto finish = [&](int q, const float (&y)[4], float um0, float um1) {
        // reduce-scatter the partial sums over the two lanes of the channel: lane half h ends with positions 4q + 2h, 4q + 2h + 1
        float k0 = hi ? y[2] : y[0], k1 = hi ? y[3] : y[1];
        k0 += __shfl_xor_sync(0xffffffffu, hi ? y[0] : y[2], 16);
        k1 += __shfl_xor_sync(0xffffffffu, hi ? y[1] : y[3], 16);
        const float o0 = fmaf(Dv, um0, k0), o1 = fmaf(Dv, um1, k1);
        const uint32_t off = ((uint32_t)q << 4) ^ swz;
        if (rev) {   // back into source order: positions p, p + 1 -> columns 31 - p, 30 - p
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(out_row_rev + (((uint32_t)(7 - q - qmiss) << 4) ^ swz)), "f"(o1), "f"(o0) : "memory");
        } else {
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(out_row + off), "f"(o0), "f"(o1) : "memory");
        }
        if (kHasZ) {   // out_z = out * silu(z)   (fwd_kernel.cuh:293)
          float z0, z1;
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(z0), "=f"(z1) : "r"(z_row + off));
          const float g0 = z0 * sigmoidf_fast(z0), g1 = z1 * sigmoidf_fast(z1);
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(out_row + kOutBytes + off), "f"(o0 * g0), "f"(o1 * g1) : "memory");
        }
      };
      // saved states 2 hb, 2 hb + 1 of the tile -> rows of the checkpoint tensor (the tensor map clips the interval a sequence ends in)
      auto send_ck = [&](int hb) {
        tma_store_2d_if(&map_ck, sm0 + LY::oCk + (uint32_t)(warp * 2 + hb) * kCkHalfBytes, (l0 / kCkptInterval + 2 * hb) * kStatePad, ck_row0, ckt && lane == 0);
      };
      prefetch(0, uv_n, dl_n);
      // quads of the tile that hold at least one position of the sequence: only the last tile of a sequence can be partial, and it
      // alone takes the instantiation with the early exit.  The full tile is ONE basic block (every condition below is a
      // compile-time constant or a predicated store), so that ptxas interleaves quad q+1's discretisation chain and quad q-1's
      // reduce-scatter with quad q's recurrence; with only two consumer warps per scheduler that static overlap is what hides
      // the MUFU / shared-memory / shuffle latencies.
      const int n_q = min(kTL / 4, (L - l0 + 3) >> 2);
      auto quads = [&](auto PARTIAL) {
        constexpr bool kPartial = decltype(PARTIAL)::value;
#pragma unroll
        for (int q = 0; q < kTL / 4; ++q) {
          if (kPartial && q >= n_q) break;
          float uv[4], dl[4], y[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) { uv[j] = uv_n[j]; dl[j] = dl_n[j]; }
          if (q + 1 < kTL / 4) prefetch(q + 1, uv_n, dl_n);   // (a partial tile prefetches one quad past its end: unused, harmless)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t bc = bc_base + (uint32_t)(4 * q + j) * (kBCPitch * 4);
            u64 Bp[4], Cp[4];
            lds_2x64(bc, Bp[0], Bp[1]);
            lds_2x64(bc + 16, Bp[2], Bp[3]);
            if (kMode != 1) {
              lds_2x64(bc + 64, Cp[0], Cp[1]);
              lds_2x64(bc + 80, Cp[2], Cp[3]);
            }
            const u64 dd = pk2(dl[j], dl[j]);
            const float du = dl[j] * uv[j];
            const u64 duu = pk2(du, du);
#pragma unroll
            for (int n = 0; n < 4; ++n) {
              float t0, t1;
              upk2(mul2(dd, cur.A2p[n]), t0, t1);
              const u64 e = pk2(ex2(t0), ex2(t1));
              x2[n] = fma2(e, x2[n], mul2(duu, Bp[n]));
            }
            if (kMode != 1)
              y[j] = hsum2(add2(fma2(Cp[1], x2[1], mul2(Cp[0], x2[0])), fma2(Cp[3], x2[3], mul2(Cp[2], x2[2]))));
          }
          if (kMode == 1) {                      // aggregates only: the segment's sum of delta (identical in the 2 lanes)
            sum_delta += (dl[0] + dl[1]) + (dl[2] + dl[3]);
            continue;
          }
          if (q > 0) finish(q - 1, yp, up0, up1);
          if (!kPartial && q == 4) {   // saved states 0-1 of the tile were staged and fenced in quad 3 and the shuffles of finish() put every
            send_ck(0);                // lane's fence before this point: send them half a tile ahead of the rest
            tma_store_commit();
            tma_store_wait_read<1>();  // the group before this one (previous tile: out + saved states 2-3) has been read: half buffer 1 is
          }                            // free for quads 5 and 7 (the shuffles in between order lane 0's wait before the other lanes' stores)
#pragma unroll
          for (int j = 0; j < 4; ++j) yp[j] = y[j];
          up0 = hi ? uv[2] : uv[0];
          up1 = hi ? uv[3] : uv[1];
          if (q & 1) {   // position l0 + 4q + 3 closes an interval of 8: save the state (predicated stores, no branch: a
                         // branch would end the basic block and put a divergence check in front of the next shuffle)
            const int done = l0 + 4 * q + 4;
            float xs[8];
#pragma unroll
            for (int n = 0; n < 4; ++n) upk2(x2[n], xs[2 * n], xs[2 * n + 1]);
            float* dst = ck + (int64_t)(done / kCkptInterval - 1) * kStatePad;
            stg_f4_if(dst, xs[0], xs[1], xs[2], xs[3], !ckt && ck != nullptr && done < L);
            stg_f4_if(dst + 4, xs[4], xs[5], xs[6], xs[7], !ckt && ck != nullptr && done < L);
            const uint32_t cb = ck_stage + (uint32_t)(q >> 2) * kCkHalfBytes;
            const uint32_t c16 = (uint32_t)((q >> 1) & 1) * (kStatePad * 4) + (uint32_t)h * (kLaneStates * 4);
            sts_f4(cb + (c16 ^ ck_key), make_float4(xs[0], xs[1], xs[2], xs[3]));
            sts_f4(cb + ((c16 + 16) ^ ck_key), make_float4(xs[4], xs[5], xs[6], xs[7]));
            if ((q & 3) == 3) fence_proxy_async_smem();   // both saved states of a half buffer written: visible to the TMA store
          }
        }
      };
      if (n_q == kTL / 4) quads(std::false_type{});
      else quads(std::true_type{});
      if (kMode == 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(s));
        continue;
      }
      finish(n_q - 1, yp, up0, up1);
      fence_proxy_async_smem();   // my OUT writes -> visible to the TMA store
      __syncwarp();
      if (n_q != kTL / 4) send_ck(0);   // (a partial tile sends both halves here)
      send_ck(1);
      if (lane == 0) {
        mbar_arrive(empty_bar(s));   // the warp is done reading stage s
        if (kMir) tma_reduce_add_3d(&map_out, out_tile, rev ? max(L - l0 - kTL, 0) : l0, w.ds0 + warp * kWarpRows, w.b);   // both groups of a pair add
        else tma_store_3d(&map_out, out_tile, l0, w.d0 + warp * kWarpRows, w.b);
        if (kHasZ) tma_store_3d(&map_outz, out_tile + kOutBytes, l0, w.d0 + warp * kWarpRows, w.b);
        tma_store_commit();
        // the group before this one has been read: the other OUT buffer (previous tile) and half buffer 0 of the saved states (this
        // tile's quads 0-3) are free for the next tile; a partial tile (the last of a sequence) sent both halves just now: drain
        if (n_q == kTL / 4) tma_store_wait_read<1>();
        else tma_store_wait_read<0>();
      }
      ++n_out;
      __syncwarp();
    }
    // ---- end of the item ----
    if (kMode == 1) {   // end state of the segment (from a zero start) and its sum of delta
      float xs[8];
#pragma unroll
      for (int n = 0; n < 4; ++n) upk2(x2[n], xs[2 * n], xs[2 * n + 1]);
      *reinterpret_cast<float4*>(ws_state) = make_float4(xs[0], xs[1], xs[2], xs[3]);
      *reinterpret_cast<float4*>(ws_state + 4) = make_float4(xs[4], xs[5], xs[6], xs[7]);
      if (h == 0) p.seg_ws[(int64_t)a.batch * a.dim * p.n_segs * kStatePad + row * p.n_segs + w.seg] = sum_delta;
    } else if (a.last_state != nullptr && (kMode == 0 || w.seg == p.n_segs - 1)) {
      float xs[8];
#pragma unroll
      for (int n = 0; n < 4; ++n) upk2(x2[n], xs[2 * n], xs[2 * n + 1]);
#pragma unroll
      for (int n = 0; n < kLaneStates; ++n)
        if (h * kLaneStates + n < N) a.last_state[row * N + h * kLaneStates + n] = xs[n];
    }
    w = wn;
    cur = nxt;
  }
  if (kMode != 1 && lane == 0) tma_store_wait_read<0>();   // sources read: the CTA may exit (the writes are complete when the grid is)
}

// Carry combine of the segmented forward: thread = (row, state); sequential over the (few) segments.
__global__ void selscan_fwd_combine_kernel(const FwdLaunch p) {
  const selscan_fwd_args& a = p.a;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t rows = (int64_t)a.batch * a.dim;
  if (idx >= rows * kStatePad) return;
  const int64_t row = idx / kStatePad;
  const int n = (int)(idx % kStatePad);
  const int d = (int)(row % a.dim);
  const float A2 = (n < a.dstate) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)n * a.A_n_stride) * kLog2e : 0.f;
  float* st = p.seg_ws + row * p.n_segs * kStatePad + n;
  const float* sd = p.seg_ws + rows * p.n_segs * kStatePad + row * p.n_segs;
  float h = 0.f;
  for (int s = 0; s < p.n_segs; ++s) {
    const float xe = st[(int64_t)s * kStatePad];
    st[(int64_t)s * kStatePad] = h;                 // state at the start of segment s
    h = fmaf(ex2(A2 * sd[s]), h, xe);
  }
}

template <int kMode, bool kHasZ, int kDt = 0, bool kMir = false>
cudaError_t launch_one(const CUtensorMap& mu, const CUtensorMap& mdt, const CUtensorMap& mz, const CUtensorMap& mout,
                       const CUtensorMap& moutz, const FwdLaunch& p, int n_items, cudaStream_t stream) {
  // saved states by TMA: rows of n_ckpt * 16 floats, box = 2 saved states (128 bytes) x 16 rows, 128-byte swizzle (the staging tiles' layout;
  // with a 64-byte inner box the hardware pads every inner row to the swizzle span)
  CUtensorMap mck = mu;
  int ck_tma = 0;
  if (kMode != 1 && p.a.ckpt != nullptr && p.n_ckpt >= 1 && (reinterpret_cast<uintptr_t>(p.a.ckpt) & 15u) == 0) {
    if (auto enc = tensor_map_encoder()) {
      const cuuint64_t gdim[2] = {(cuuint64_t)p.n_ckpt * kStatePad, (cuuint64_t)p.a.batch * (cuuint64_t)p.a.dim};
      const cuuint64_t gstr[1] = {(cuuint64_t)p.n_ckpt * kStatePad * 4};
      const cuuint32_t box[2] = {2 * (cuuint32_t)kStatePad, (cuuint32_t)kWarpRows};
      const cuuint32_t estr[2] = {1, 1};
      ck_tma = enc(&mck, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.a.ckpt, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
  }
  constexpr int smem = (int)Lay<kHasZ, kDt>::kBytes + 1024;
  static std::atomic<unsigned long long> configured{0};   // one cudaFuncSetAttribute per device, not per launch
  if (const cudaError_t e = set_smem_once(configured, selscan_fwd_tma_kernel<kMode, kHasZ, kDt, kMir>, smem)) return e;
  const int slots = sm_count() * Lay<kHasZ, kDt>::kCtas;
  const unsigned grid = (unsigned)(n_items < slots ? n_items : slots);
  selscan_fwd_tma_kernel<kMode, kHasZ, kDt, kMir><<<grid, kThreads, smem, stream>>>(mu, mdt, mz, mout, moutz, mck, ck_tma, p, n_items);
  return cudaGetLastError();
}

// 4-D map over dt_x (fastest first: seqlen, rank, group, batch), box = 32 positions x kDt ranks, dense rows; ranks >= dt_rank are
// out of bounds and read as zeros
inline bool make_dtx_map(CUtensorMap* map, const selscan_fwd_args& a, int box_l, int box_r) {
  auto enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[4] = {(cuuint64_t)a.seqlen, (cuuint64_t)a.dt_rank, (cuuint64_t)a.ngroups, (cuuint64_t)a.batch};
  const cuuint64_t gstr[3] = {(cuuint64_t)a.dt_x_r_stride * 4, (cuuint64_t)(a.ngroups > 1 ? a.dt_x_group_stride : a.dt_x_r_stride * a.dt_rank) * 4,
                              (cuuint64_t)(a.batch > 1 ? a.dt_x_batch_stride : a.dt_x_r_stride * a.dt_rank * a.ngroups) * 4};
  const cuuint32_t box[4] = {(cuuint32_t)box_l, (cuuint32_t)box_r, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.dt_x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Segment plan shared by the workspace query and the launcher: split only when the call cannot fill the chip.
void fwd_plan_segments(int batch, int dim, int seqlen, int ngroups, int* n_segs, int* seg_tiles) {
  *n_segs = 1;
  *seg_tiles = (seqlen + kTL - 1) / kTL;
  if (batch <= 0 || dim <= 0 || ngroups <= 0 || dim % ngroups || (dim / ngroups) % kRows) return;
  const int64_t n_ctas = (int64_t)batch * (dim / kRows);
  const int n_tiles = (seqlen + kTL - 1) / kTL;
  const int slots = 2 * sm_count();
  if (n_ctas * 3 > slots || n_tiles < 4) return;      // at least a third of the 2-per-SM CTA slots is busy anyway
  int want = (int)(slots / n_ctas);
  if (want > n_tiles / 2) want = n_tiles / 2;
  if (want < 2) return;
  const int st = (n_tiles + want - 1) / want;
  // Two passes over st tiles + two extra launches (~10 tiles' worth of time) must beat one pass over n_tiles (measured on
  // B200: profiles/r01_roofline_sweep.json; stage 3/4 lengths never qualify).
  if (n_tiles - 2 * st < 12) return;
  *seg_tiles = st;
  *n_segs = (n_tiles + st - 1) / st;
}

bool fwd_tma_eligible(const FwdLaunch& p) {
  const selscan_fwd_args& a = p.a;
  if (a.dstate > kStatePad) return false;
  if (p.dim_per_group % kRows != 0) return false;
  if (a.seqlen < 1 || a.batch < 1) return false;
  if (!tma_row_ok(a.u, a.u_d_stride, a.batch > 1 ? a.u_batch_stride : 4)) return false;
  if (a.dt_w == nullptr && !tma_row_ok(a.delta, a.delta_d_stride, a.batch > 1 ? a.delta_batch_stride : 4)) return false;
  if (!tma_row_ok(a.out, a.out_d_stride, a.batch > 1 ? a.out_batch_stride : 4)) return false;
  if (a.z != nullptr) {
    if (!tma_row_ok(a.z, a.z_d_stride, a.batch > 1 ? a.z_batch_stride : 4)) return false;
    if (!tma_row_ok(a.out_z, a.out_z_d_stride, a.batch > 1 ? a.out_z_batch_stride : 4)) return false;
  }
  return tensor_map_encoder() != nullptr;
}

// Returns cudaErrorNotSupported when a tensor map cannot be encoded for this layout: the caller then takes the generic kernel.
cudaError_t launch_fwd_tma(const FwdLaunch& p, cudaStream_t stream) {
  const selscan_fwd_args& a = p.a;
  CUtensorMap mu, mdt, mout, mz, moutz;
  const int dt_box = a.dt_w == nullptr ? 0 : (a.dt_rank <= 6 ? 6 : kMaxFusedDtRank);
  const int src_rows = a.mirror_pairs ? a.dim / 2 : a.dim;   // rows per batch of u / out
  if (!make_row_map(&mu, a.u, a.seqlen, src_rows, a.batch, a.u_d_stride, a.u_batch_stride, kTL, kRows) ||
      !(dt_box ? make_dtx_map(&mdt, a, kTL, dt_box)
               : make_row_map(&mdt, a.delta, a.seqlen, a.dim, a.batch, a.delta_d_stride, a.delta_batch_stride, kTL, kRows)) ||
      !make_row_map(&mout, a.out, a.seqlen, src_rows, a.batch, a.out_d_stride, a.out_batch_stride, kTL, kWarpRows))
    return cudaErrorNotSupported;
  const bool has_z = a.z != nullptr;
  if (has_z) {
    if (!make_row_map(&mz, a.z, a.seqlen, a.dim, a.batch, a.z_d_stride, a.z_batch_stride, kTL, kRows) ||
        !make_row_map(&moutz, a.out_z, a.seqlen, a.dim, a.batch, a.out_z_d_stride, a.out_z_batch_stride, kTL, kWarpRows))
      return cudaErrorNotSupported;
  } else {
    mz = mu;
    moutz = mout;
  }
  const int n_base = (int)((int64_t)a.batch * a.ngroups * (p.dim_per_group / kRows));
  FwdLaunch q = p;
  q.n_segs = 1;
  fwd_plan_segments(a.batch, a.dim, a.seqlen, a.ngroups, &q.n_segs, &q.seg_tiles);
  if (a.mirror_pairs) {   // mirrored pairs: whole sequences only (selscan_b200_mirror_ok), no z
    q.n_segs = 1;
    if (has_z) return cudaErrorNotSupported;
    if (dt_box == 6) return launch_one<0, false, 6, true>(mu, mdt, mz, mout, moutz, q, n_base, stream);
    if (dt_box) return launch_one<0, false, kMaxFusedDtRank, true>(mu, mdt, mz, mout, moutz, q, n_base, stream);
    return launch_one<0, false, 0, true>(mu, mdt, mz, mout, moutz, q, n_base, stream);
  }
  if (dt_box) {   // fused dt_proj: whole sequences only (selscan_b200_dt_fusable excludes the segmented sizes), no z
    q.n_segs = 1;
    return dt_box == 6 ? launch_one<0, false, 6>(mu, mdt, mz, mout, moutz, q, n_base, stream)
                       : launch_one<0, false, kMaxFusedDtRank>(mu, mdt, mz, mout, moutz, q, n_base, stream);
  }
  if (p.seg_ws == nullptr || q.n_segs < 2) {
    q.n_segs = 1;
    return has_z ? launch_one<0, true>(mu, mdt, mz, mout, moutz, q, n_base, stream)
                 : launch_one<0, false>(mu, mdt, mz, mout, moutz, q, n_base, stream);
  }
  cudaError_t e = launch_one<1, false>(mu, mdt, mz, mout, moutz, q, n_base * q.n_segs, stream);
  if (e != cudaSuccess) return e;
  const int64_t n_comb = (int64_t)a.batch * a.dim * kStatePad;
  selscan_fwd_combine_kernel<<<(unsigned)((n_comb + 255) / 256), 256, 0, stream>>>(q);
  return has_z ? launch_one<2, true>(mu, mdt, mz, mout, moutz, q, n_base * q.n_segs, stream)
               : launch_one<2, false>(mu, mdt, mz, mout, moutz, q, n_base * q.n_segs, stream);
}

}  // namespace selscan
