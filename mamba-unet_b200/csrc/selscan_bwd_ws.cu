// Backward selective scan, warp-specialised tiled path for sm_100a.  Replaces selective_scan_bwd_kernel
// (/root/reference/mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489) for the aligned shapes Mamba-UNet produces
// (channels per group a multiple of 64, 16-byte aligned rows, no z, B and C with the same position stride); everything else takes
// selscan_bwd.cu.  (Its single-role predecessor and a TMEM-pipelined successor that did not beat it are not part of the library:
// experiments/, profiles/r02_ncu_bwd_tmem_*.)
//
// Why two roles.  In the single-role kernel every thread ran ~1240 instructions per chunk of 8 positions, of which only ~600 were the
// recurrences (FFMA2 / FMUL2 / MUFU.EX2); the rest -- softplus / sigmoid of delta, the B/C gather, TMA issue, the reduce-scatter and
// finalisation of du / ddelta, the contraction of the dB / dC products over channels, atomics -- are latency-bound chains
// (MUFU -> RCP -> polynomial, shared-memory round trips, shuffles) and with 254 registers per thread only 2 warps per SM
// sub-partition were there to hide them: issue slots 45 % busy.  Here the CTA is two warpgroups with different register budgets
// (setmaxnreg):
//   warps 0-3  COMPUTE (184 registers): a thread owns 2 channels x 4 states; per chunk it restarts the forward recurrence from the
//              saved state, runs the reverse recurrence in registers (B and delta*u of the chunk stay in registers; the decays are
//              evaluated a second time rather than kept: MUFU has slack, the shared-memory return path has none), and leaves
//              (a) the channel-pair products for dB / dC in the P tile and (b) its 4-state partial sums of dx*B and dx*a*x*A per
//              (channel, position) in the S12 tile.  Nothing else: no transcendental besides the decays, no shuffles, no global
//              memory.  784 instructions per chunk.
//   warps 4-7  HELPER (72 registers): thread (row, half chunk) discretises delta once per element (softplus, sigmoid) one chunk
//              ahead and publishes delta and delta*u; gathers B/C (any strides, loaded two chunks ahead) into the
//              [position][B0..15 C0..15] tile; one lane issues the TMA loads two chunks ahead (u, delta, dout: 64 x 8 boxes with the
//              32-byte swizzle; saved state: 64 x 16) into a 3-stage ring; after the compute warps finish a half chunk it sums P
//              over the 32 channel pairs (one sum per (state, position) per CTA, sent as 16 x 4 TMA reduce-add boxes where the rows
//              of dB / dC are 16-byte aligned, as scalar atomics otherwise -- the reference issues one atomic per (channel, state,
//              position), bwd_kernel.cuh:298-316), sums the S12 partials over the 4 lanes of a channel and finalises du, ddelta
//              (through softplus'), dD, ddelta_bias; du / ddelta leave by per-half TMA stores.
// Hand-over is by mbarriers only (full/empty per stage and per half chunk); the two roles drift by up to a chunk.  All waits are
// bounded (trap instead of hang).  16 warps per SM (2 CTAs): 8 dense compute warps + 8 helper warps that fill their stalls.
// Measured (stage 1, batch 24): 0.888 ms (single-role kernel: 0.98); issue slots 54 % busy, shared-memory pipe 71 %, MUFU 51 %.
// No pipe is the limit by itself: what-if builds (profiles/r02_bwd_whatif.json) show that the time follows the HELPER warps' chain
// of dependent instructions (busy ~90 % of the time; the compute warps wait 10 % for prep_done) -- hence dB / dC as TMA reduce-adds
// of staged tiles and packed f32x2 helper math.  One CTA per SM (batch <= 8): 0.69 against 0.85 ms.
#include <atomic>
#include <type_traits>

#include "selscan_common.cuh"
#include "selscan_kernels.h"
#include "selscan_ptx.cuh"
#include "selscan_tma_host.h"

namespace selscan {

namespace {

constexpr int kR = 64;            // channels per CTA
constexpr int kNP = kR / 2;       // channel pairs per CTA
constexpr int kC = kCkptInterval; // positions per chunk (8)
constexpr int kHP = kC / 2;       // positions per half chunk
constexpr int kStg = 3;
constexpr int kPitch = 36;        // B/C tile pitch (floats)
constexpr int kGroupThr = 128;    // threads per role
constexpr int kThr = 2 * kGroupThr;
constexpr int kLS = kStatePad / 4;  // states per lane
constexpr int kCompRegs = 184;    // 128 * (184 + 72) * 2 CTAs = the whole register file
constexpr int kHelpRegs = 72;
constexpr int kLaunchRegs = 128;  // what ptxas must report for the kernel (checked by the launcher: setmaxnreg.inc would block otherwise)

constexpr int kPP = kHP * 32 + 16;   // floats per channel pair in a P half: 4 positions x 32 values + 64 bytes, so that the two pairs of
                                     // a quarter-warp store to disjoint banks and readers need no swizzle
constexpr int kS12P = 36;            // floats per row in an S12 half (32 + 16 bytes: conflict-free STS.128 by (row, lane), LDS.128 by row)

// One stage of the ring: every per-stage tile in ONE block, so that a stage costs one base address per thread (the tiles used to be
// separate [kStg] arrays: one multiply-add each, ~15 instructions per chunk and warp).  Multiples of 256 bytes throughout.
struct WsStage {
  float CK[kR * kStatePad];         // [row][16 states]                                    (TMA)
  float U[kR * kC];                 // [row][8 positions] (one 64 x 8 box)                  (TMA)
  float DT[kR * kC];                //   raw delta
  float DY[kR * kC];
  float SD[kR * kC];                // delta   [row][8]                                     (helper)
  float SDU[kR * kC];               // delta*u
  float BC[kC * kPitch + 32];       // [position][B0..15 C0..15]                            (helper)   (+ 128 bytes: block size % 256 == 0)
};
static_assert(sizeof(WsStage) % 256 == 0, "stage tiles keep their 256-byte alignment");

template <bool kHasZ, int kDt>
struct WsSmemT {
  WsStage st[kStg];
  // gated calls only (bwd_kernel.cuh:171-207): z and the ungated forward output, [row][8] like U (TMA: the hardware swizzle is a
  // function of the shared-memory ADDRESS, so these tiles sit with the other TMA tiles at multiples of 256 bytes); the helpers
  // overwrite dout with dout * silu(z) and `out` with dz in place, and dz leaves by a TMA store of the whole stage tile
  float Z[kHasZ ? kStg : 1][kHasZ ? kR * kC : 64];
  float O[kHasZ ? kStg : 1][kHasZ ? kR * kC : 64];
  // fused dt_proj only (mamba_sys.py:409): the dt rows of x_dbl, [rank][8 positions] dense (TMA), instead of the DT tile
  float XDT[kStg][kDt > 0 ? (kDt * kC + 31) / 32 * 32 : 32];   // (multiples of 128 bytes: TMA tiles before and after stay aligned)
  float P[2][kNP * kPP];            // per half chunk: [pair][position][dB 0..15 | dC 0..15]   (compute)
  float S12[2][kR * kS12P];         // per half chunk: [row][lane 0..3][s1 x4 positions | s2 x4] (compute)
  float DU[2][kR * kC];             // output tiles [half][row][4] (one 64 x 4 TMA store per half), double-buffered by chunk parity (helper)
  float DDT[2][kR * kC];
  // dB / dC sums of a half chunk, [chunk parity][half][dB | dC][16 states][4 positions]: two 16 x 4 TMA reduce-add boxes per half
  // instead of 128 scattered global REDs (profiles/r02_bwd_whatif.json: the REDs alone cost 5 % of the kernel)
  float DBC[2][2][2 * kStatePad * kHP];
  u64 tma_full[kStg];               // TMA transaction bytes of a stage
  u64 prep_done[kStg];              // helper warps: delta / delta*u / B/C tiles of a stage written
  u64 stage_free[kStg];             // compute + helper warps: stage no longer read
  u64 half_full[2];                 // compute warps: P / S12 of a half chunk written
  u64 half_free[2];                 // helper warps: P / S12 of a half chunk consumed
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void sts_2x64(uint32_t addr, u64 a, u64 b) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ float ex2v(float x) {   // MUFU.EX2 that the compiler may not merge with an identical earlier one
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ u64 mul2v(u64 a, u64 b) {
  u64 d;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// kMir: mirrored direction pairs (selscan_b200.h: mirror_pairs) -- odd groups read u / dout of the even group of their pair back to
// front and add their du into the same rows.  Only the helper warps know: they put the tiles into scan order before anyone else
// reads them and reverse the outputs on the way out.
template <bool kHasZ, int kDt, bool kMir>
__global__ void __launch_bounds__(kThr, 2)
selscan_bwd_ws_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_dt,
                      const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_ck,
                      const __grid_constant__ CUtensorMap map_du, const __grid_constant__ CUtensorMap map_ddt,
                      const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_o,
                      const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_db,
                      const __grid_constant__ CUtensorMap map_dc, const int tma_bc, const BwdLaunch p) {
  // tma_bc (kernel-uniform): dB / dC leave as TMA reduce-adds of staged tiles (map_db / map_dc: (seqlen, 16, batch * groups), box 4 x 16)
  // kDt > 0: map_dt is the 4-D map over dt_x (seqlen, rank, group, batch); the helpers form the raw step themselves
  using WsSmem = WsSmemT<kHasZ, kDt>;
  extern __shared__ unsigned char smem_raw[];
  WsSmem& sm = *reinterpret_cast<WsSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 255) & ~(uintptr_t)255);   // 32B-swizzle atom = 256 B
  const selscan_bwd_args& a = p.a;
  const int L = a.seqlen, N = a.dstate;
  // One volatile read of %tid.x: with 72 registers the compiler otherwise re-reads the special register inside the helper loop
  // (S2R + a short-scoreboard wait, 6 % of the helper warps' time in profiles/r02_ncu_scan_summary.txt's source view)
  int tid;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid));
  const int warp = tid >> 5, lane = tid & 31;
  const int tiles_per_group = p.dim_per_group / kR;
  int bid = blockIdx.x;
  const int tile_g = bid % tiles_per_group; bid /= tiles_per_group;
  const int g = bid % a.ngroups;
  const int b = bid / a.ngroups;
  const int d0 = g * p.dim_per_group + tile_g * kR;
  const int n_tiles = (L + kC - 1) / kC;
  const bool rev = kMir && (g & 1);                                           // CTA-uniform
  const int ds0 = kMir ? (g >> 1) * p.dim_per_group + tile_g * kR : d0;       // first row of the u / dout / du tensors

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kStg; ++s) {
      mbar_init(smem_u32(&sm.tma_full[s]), 1);
      mbar_init(smem_u32(&sm.prep_done[s]), 4);
      mbar_init(smem_u32(&sm.stage_free[s]), 8);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mbar_init(smem_u32(&sm.half_full[h]), 4);
      mbar_init(smem_u32(&sm.half_free[h]), 4);
    }
    mbar_fence_init();
    tma_prefetch_desc(&map_u);
    tma_prefetch_desc(&map_dt);
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_ck);
    tma_prefetch_desc(&map_du);
    tma_prefetch_desc(&map_ddt);
    if (kHasZ) {
      tma_prefetch_desc(&map_z);
      tma_prefetch_desc(&map_o);
      tma_prefetch_desc(&map_dz);
    }
    if (tma_bc) {
      tma_prefetch_desc(&map_db);
      tma_prefetch_desc(&map_dc);
    }
  }
  __syncthreads();

  // [row][8] tiles carry the TMA 32-byte swizzle (16-byte half index ^= bit 2 of the row): 8 consecutive rows x one half are then
  // 8 distinct bank groups for the helper's per-row 128-bit accesses; the compute warps' broadcast reads do not care.
  constexpr uint32_t kCT = kNP * kC * 4;    // byte offset of a pair's second channel (row + 32) in such a tile

  if (warp >= 4) {
    // =========================================== helper warpgroup ===========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kHelpRegs));
    const int htid = tid - kGroupThr;
    constexpr int kLoadThr = 32;            // the thread that issues the TMA loads (and the dz stores): lane 0 of the second helper warp --
                                            // threads 0 and 64 issue the output stores of their halves, so three warps share the issue work
    const int hw = warp - 4;
    const int row = htid & (kR - 1);        // my channel inside the CTA ...
    const int hf = htid >> 6;               // ... and my half of every chunk (warp-uniform)
    const int d = d0 + row;
    const float Dv = a.D ? __ldg(a.D + d) : 0.f;
    const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.f;
    const bool softplus = a.delta_softplus != 0;
    float Wd[kDt > 0 ? kDt : 1];            // my channel's row of the dt_proj weight (zero beyond dt_rank)
    if (kDt > 0) {
#pragma unroll
      for (int r = 0; r < kDt; ++r) Wd[r] = (r < a.dt_rank) ? __ldg(a.dt_w + (int64_t)d * a.dt_w_d_stride + r) : 0.f;
    }
    const uint32_t my16 = (uint32_t)row * (kC * 4) + (uint32_t)((hf ^ ((row >> 2) & 1)) << 4);   // my 4 elements in every [row][8] tile
    const uint32_t out16 = (uint32_t)hf * (kR * 16) + (uint32_t)row * 16;    // ... and in the [half][row][4] output tiles
    // Mirrored groups: the TMA box of a chunk holds the chunk's SOURCE positions, i.e. my 4 scan positions sit in the other half,
    // reversed.  TMA coordinates are kept non-negative: the partial chunk (seqlen % 8 == 4, the first one processed) loads source
    // positions 0..7, of which 3..0 are scan half 0 -- the same half, reversed.
    const bool part_first = rev && (L & 4);
    auto src16_of = [&](int j) { return (rev && !(part_first && j == 0)) ? (my16 ^ 16u) : my16; };
    auto srchalf_of = [&](int j) { return (rev && !(part_first && j == 0)) ? 1 - hf : hf; };
    auto unrev = [&](float4 v) { return rev ? make_float4(v.w, v.z, v.y, v.x) : v; };

    // ---- B/C gather: elements htid and htid + 128 of a chunk's [8 positions][32 values] tile; the pointers walk backwards ----
    // (B and C have the same position stride on this path: bwd_ws_eligible)
    const int64_t bc_step = (rev ? -(int64_t)kC : (int64_t)kC) * a.B_l_stride;
    const bool along_l = a.B_l_stride == 1;   // (.., N, L) layout: 8 consecutive threads read 8 consecutive positions of a state row;
                                              // l-major x_dbl layout: a warp reads the 16 B and 16 C values of one position
    const int bc_e1 = htid + kGroupThr;
    const int bc_pos0 = along_l ? (htid & 7) : (htid >> 5), bc_val0 = along_l ? (htid >> 3) : (htid & 31);
    const int bc_pos1 = along_l ? (bc_e1 & 7) : (bc_e1 >> 5), bc_val1 = along_l ? (bc_e1 >> 3) : (bc_e1 & 31);
    auto bc_src = [&](int pos, int val) -> const float* {
      const int n = val & 15;
      const float* base = (val >= 16) ? (a.C + (int64_t)b * a.C_batch_stride + (int64_t)g * a.C_group_stride + (int64_t)n * a.C_n_stride)
                                      : (a.B + (int64_t)b * a.B_batch_stride + (int64_t)g * a.B_group_stride + (int64_t)n * a.B_n_stride);
      const int sp = (n_tiles - 1) * kC + pos;                           // scan position; first chunk processed = last of the sequence
      return base + (int64_t)(rev ? L - 1 - sp : sp) * a.B_l_stride;
    };
    const float* bcp0 = bc_src(bc_pos0, bc_val0);
    const float* bcp1 = bc_src(bc_pos1, bc_val1);
    const uint32_t bc_i0 = (uint32_t)(bc_pos0 * kPitch + bc_val0) * 4, bc_i1 = (uint32_t)(bc_pos1 * kPitch + bc_val1) * 4;   // byte offsets in a tile
    const bool bc_ok0 = (bc_val0 & 15) < N, bc_ok1 = (bc_val1 & 15) < N;
    // ---- contraction role: warp = position of the half chunk, lane = (pair quarter q, 16-byte chunk c of the 32 values) ----
    const int c_c = lane & 7, c_q = lane >> 3;
    const int c_n = (c_c & 3) * 4 + 2 * (c_q >> 1) + (c_q & 1);   // the state whose sum this lane ends up with
    const bool c_ok = c_n < N;
    const bool tbc = !kMir && tma_bc != 0;
    const uint32_t tbc_off = (uint32_t)(((c_c >= 4) ? kStatePad * kHP : 0) + c_n * kHP + hw) * 4;   // my (tensor, state, position) in a DBC tile
    // address of my (state, position hw of half 0) in the chunk being processed; walks backwards by one chunk per iteration
    const int dbc_sgn = rev ? -1 : 1;        // mirrored groups: dB / dC in source order, i.e. at seqlen-1-position
    float* dbc = ((c_c >= 4) ? a.dC : a.dB) + (((int64_t)b * a.ngroups + g) * N + (c_ok ? c_n : 0)) * (int64_t)L +
                 (rev ? L - 1 - ((n_tiles - 1) * kC + hw) : (n_tiles - 1) * kC + hw);

    auto issue_tma = [&](int j) {            // chunk j (processing order) -> stage j % kStg
      const int t = n_tiles - 1 - j, s = j % kStg, l0 = t * kC;
      if (j >= kStg) mbar_wait(smem_u32(&sm.stage_free[s]), (uint32_t)((j / kStg - 1) & 1));
      if (kHasZ) tma_store_wait_read<0>();   // my dz store out of this stage's O tile (three chunks ago) has read it
      const uint32_t full = smem_u32(&sm.tma_full[s]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full),
                   "r"((uint32_t)(((kHasZ ? 5 : 3) - (kDt > 0 ? 1 : 0)) * kR * kC * 4 + kDt * kC * 4 + kR * kStatePad * 4)) : "memory");
      if (kHasZ) {
        tma_load_3d(smem_u32(sm.Z[s]), &map_z, l0, d0, b, full);
        tma_load_3d(smem_u32(sm.O[s]), &map_o, l0, d0, b, full);
      }
      const int lc = rev ? max(L - l0 - kC, 0) : l0;    // mirrored groups: the chunk's source positions
      tma_load_3d(smem_u32(sm.st[s].U), &map_u, lc, ds0, b, full);
      if (kDt > 0) tma_load_4d(smem_u32(sm.XDT[s]), &map_dt, lc, 0, g, b, full);   // ranks >= dt_rank / positions outside: zero fill
      else tma_load_3d(smem_u32(sm.st[s].DT), &map_dt, lc, d0, b, full);
      tma_load_3d(smem_u32(sm.st[s].DY), &map_dy, lc, ds0, b, full);
      // saved state t-1 = state before the chunk's first position; state "-1" is out of bounds -> zeros
      tma_load_2d(smem_u32(sm.st[s].CK), &map_ck, (t - 1) * kStatePad, b * a.dim + d0, full);
    };
    // discretise my 4 elements of chunk j, publish delta, delta*u, softplus' and the chunk's B/C values (loaded one iteration earlier)
    auto prep = [&](int j, float cb0, float cb1, float4& sg_out) {
      const int t = n_tiles - 1 - j, s = j % kStg, l0 = t * kC + hf * kHP;
      mbar_wait(smem_u32(&sm.tma_full[s]), (uint32_t)((j / kStg) & 1));
      const uint32_t src16 = src16_of(j);
      const float4 u4 = unrev(lds_f4(smem_u32(sm.st[s].U) + src16));
      float4 t4;
      if (kDt > 0) {   // raw step of my 4 positions: dt_w[d, :] . dt_x[:, l]  (all rows of a warp read the same 16 bytes: broadcast)
        u64 ta = pk2(0.f, 0.f), tb = ta;   // two positions per FFMA2 (the weight is a broadcast operand)
#pragma unroll
        for (int r = 0; r < kDt; ++r) {
          u64 xa, xb;
          lds_2x64(smem_u32(sm.XDT[s]) + (uint32_t)(r * (kC * 4) + srchalf_of(j) * 16), xa, xb);
          const u64 w = pk2(Wd[r], Wd[r]);
          ta = fma2(w, xa, ta);
          tb = fma2(w, xb, tb);
        }
        upk2(ta, t4.x, t4.y);
        upk2(tb, t4.z, t4.w);
        t4 = unrev(t4);                    // (mirrored groups: the box holds source order)
      } else {
        t4 = unrev(lds_f4(smem_u32(sm.st[s].DT) + src16));
      }
      if (kMir && rev) {   // put u and dout of the stage into scan order, in place, before anyone else reads them: every thread reads
                           // the half its partner writes, hence the barrier between the reads and the writes
        float4 y4 = unrev(lds_f4(smem_u32(sm.st[s].DY) + src16));
        float4 uz = u4;
        const int lp = (n_tiles - 1 - j) * kC + hf * kHP;   // past the end the box holds other positions' data, not zero fill: clear
        if (lp + 0 >= L) { uz.x = 0.f; y4.x = 0.f; }
        if (lp + 1 >= L) { uz.y = 0.f; y4.y = 0.f; }
        if (lp + 2 >= L) { uz.z = 0.f; y4.z = 0.f; }
        if (lp + 3 >= L) { uz.w = 0.f; y4.w = 0.f; }
        named_bar_sync(1, kGroupThr);
        sts_f4(smem_u32(sm.st[s].U) + my16, uz);
        sts_f4(smem_u32(sm.st[s].DY) + my16, y4);
        fence_proxy_async_smem();    // ordered before the TMA load that refills this stage
      }
      // Two elements per instruction wherever a packed form exists (FMUL2 / FADD2 / FFMA2): this role is issue-bound like the other,
      // and softplus + its derivative are 2/5 of its instructions (profiles/r02_bwd_whatif.json).  Same arithmetic as softplus_fast /
      // sigmoid_from_w (selscan_common.cuh), element by element.
      const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
      float v[4], vu[4], sg[4];
#pragma unroll
      for (int e = 0; e < 4; e += 2) {
        const float x0 = tt[e] + bias, x1 = tt[e + 1] + bias;
        float d0v = x0, d1v = x1, g0 = 1.f, g1 = 1.f;
        if (softplus) {
          float w0, w1, r0, r1, q0, q1;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(-fabsf(x0) * 1.4426950408889634f));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(-fabsf(x1) * 1.4426950408889634f));
          const u64 w2 = pk2(w0, w1);
          float a0, a1, b0, b1;
          upk2(add2(w2, pk2(2.f, 2.f)), a0, a1);
          upk2(add2(w2, pk2(1.f, 1.f)), b0, b1);
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(a0));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(a1));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q0) : "f"(b0));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q1) : "f"(b1));
          const u64 t = mul2(w2, pk2(r0, r1));                 // w / (2 + w)
          const u64 t2 = mul2(t, t);
          u64 pl = fma2(t2, pk2(kAtanhC4, kAtanhC4), pk2(kAtanhC3, kAtanhC3));
          pl = fma2(pl, t2, pk2(kAtanhC2, kAtanhC2));
          pl = fma2(pl, t2, pk2(kAtanhC1, kAtanhC1));
          pl = fma2(pl, t2, pk2(1.f, 1.f));
          upk2(fma2(add2(t, t), pl, pk2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), d0v, d1v);
          float wq0, wq1;
          upk2(mul2(w2, pk2(q0, q1)), wq0, wq1);               // softplus' (bwd_kernel.cuh:446-450; == 1 to rounding for x > 20)
          g0 = x0 >= 0.f ? q0 : wq0;
          g1 = x1 >= 0.f ? q1 : wq1;
        }
        if (j == 0) {                         // only the first chunk processed can reach past the end: a = 1, b = 0 there
          d0v = (l0 + e < L) ? d0v : 0.f;     // (u and dout are TMA zero fill)
          d1v = (l0 + e + 1 < L) ? d1v : 0.f;
        }
        v[e] = d0v; v[e + 1] = d1v;
        upk2(mul2(pk2(d0v, d1v), pk2(uu[e], uu[e + 1])), vu[e], vu[e + 1]);
        sg[e] = g0; sg[e + 1] = g1;
      }
      sts_f4(smem_u32(sm.st[s].SD) + my16, make_float4(v[0], v[1], v[2], v[3]));
      sts_f4(smem_u32(sm.st[s].SDU) + my16, make_float4(vu[0], vu[1], vu[2], vu[3]));
      sg_out = make_float4(sg[0], sg[1], sg[2], sg[3]);
      if (kHasZ) {   // bwd_kernel.cuh:186-191: dz = dout * out * sigmoid(z) * (1 + z * (1 - sigmoid(z))), then dout <- dout * silu(z)
        const float4 z4 = lds_f4(smem_u32(sm.Z[s]) + my16);
        const float4 o4 = lds_f4(smem_u32(sm.O[s]) + my16);
        const float4 y4 = lds_f4(smem_u32(sm.st[s].DY) + my16);
        const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, oo[4] = {o4.x, o4.y, o4.z, o4.w}, yy[4] = {y4.x, y4.y, y4.z, y4.w};
        float dzv[4], dyg[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float sz = sigmoidf_fast(zz[e]);
          dzv[e] = yy[e] * oo[e] * sz * (1.f + zz[e] * (1.f - sz));
          dyg[e] = yy[e] * zz[e] * sz;
        }
        sts_f4(smem_u32(sm.O[s]) + my16, make_float4(dzv[0], dzv[1], dzv[2], dzv[3]));     // in place: dz
        sts_f4(smem_u32(sm.st[s].DY) + my16, make_float4(dyg[0], dyg[1], dyg[2], dyg[3]));    // in place: the gated dout everyone else reads
        fence_proxy_async_smem();            // generic writes -> visible to the TMA store below, ordered before the stage's next TMA load
        named_bar_sync(1, kGroupThr);        // the whole dz tile is written
        if (htid == kLoadThr) {
          tma_store_3d(&map_dz, smem_u32(sm.O[s]), t * kC, d0, b);      // positions past the end are clipped by the tensor map
          tma_store_commit();
        }
      }
      sts_f1(smem_u32(sm.st[s].BC) + bc_i0, cb0);
      sts_f1(smem_u32(sm.st[s].BC) + bc_i1, cb1);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.prep_done[s]));
    };

    u64 dD2 = pk2(0.f, 0.f), dbias2 = pk2(0.f, 0.f);   // packed partial sums of dD / ddelta_bias
    float cb0, cb1;
    float4 sg_cur, sg_nxt = make_float4(0.f, 0.f, 0.f, 0.f);   // softplus' of my 4 elements: chunk being finalised / prepared ahead
    {
      if (htid == kLoadThr) {
        issue_tma(0);
        if (n_tiles > 1) issue_tma(1);
      }
      const int lbase = (n_tiles - 1) * kC;   // the only chunk that can be partial
      cb0 = (bc_ok0 && lbase + bc_pos0 < L) ? __ldg(bcp0) : 0.f;
      cb1 = (bc_ok1 && lbase + bc_pos1 < L) ? __ldg(bcp1) : 0.f;
      bcp0 -= bc_step;
      bcp1 -= bc_step;
      prep(0, cb0, cb1, sg_cur);
      cb0 = (bc_ok0 && n_tiles > 1) ? __ldg(bcp0) : 0.f;
      cb1 = (bc_ok1 && n_tiles > 1) ? __ldg(bcp1) : 0.f;
      bcp0 -= bc_step;
      bcp1 -= bc_step;
    }
    for (int i = 0; i < n_tiles; ++i) {
      const int t = n_tiles - 1 - i, s = i % kStg, c0 = t * kC, ob = i & 1;
      float nb0 = 0.f, nb1 = 0.f;
      if (i + 2 < n_tiles) {                 // B/C of chunk i+2: in flight while chunk i is finalised
        if (bc_ok0) nb0 = __ldg(bcp0);
        if (bc_ok1) nb1 = __ldg(bcp1);
        bcp0 -= bc_step;
        bcp1 -= bc_step;
      }
      // first the discretisation the compute warps wait for next (its TMA data landed an iteration ago), then the TMA loads of
      // chunk i+2, whose stage must first be released by every warp's work on chunk i-1
      if (i + 1 < n_tiles) prep(i + 1, cb0, cb1, sg_nxt);
      if (htid == kLoadThr && i + 2 < n_tiles) issue_tma(i + 2);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 1 - hh;
        mbar_wait_relaxed(smem_u32(&sm.half_full[h]), (uint32_t)(i & 1));
        // ---------------- contraction of the pair products over the CTA's 32 channel pairs ----------------
        {
          const uint32_t src = smem_u32(sm.P[h]) + (uint32_t)(c_q * 8) * (kPP * 4) + (uint32_t)hw * 128 + (uint32_t)c_c * 16;
          u64 v[8][2];
#pragma unroll
          for (int e = 0; e < 8; ++e) lds_2x64(src + (uint32_t)e * (kPP * 4), v[e][0], v[e][1]);   // all in flight together
          const u64 acc0 = add2(add2(add2(v[0][0], v[1][0]), add2(v[2][0], v[3][0])), add2(add2(v[4][0], v[5][0]), add2(v[6][0], v[7][0])));
          const u64 acc1 = add2(add2(add2(v[0][1], v[1][1]), add2(v[2][1], v[3][1])), add2(add2(v[4][1], v[5][1]), add2(v[6][1], v[7][1])));
          float o0, o1, o2, o3;
          upk2(acc0, o0, o1);
          upk2(acc1, o2, o3);
          // reduce-scatter over the 4 pair quarters: lane q ends with value 2*(q>>1) + (q&1) of its chunk
          const bool q1 = (c_q & 2) != 0, q0 = (c_q & 1) != 0;
          float k0 = q1 ? o2 : o0, k1 = q1 ? o3 : o1;
          k0 += __shfl_xor_sync(0xffffffffu, q1 ? o0 : o2, 16);
          k1 += __shfl_xor_sync(0xffffffffu, q1 ? o1 : o3, 16);
          float kk = q0 ? k1 : k0;
          kk += __shfl_xor_sync(0xffffffffu, q0 ? k0 : k1, 8);
          if (tbc) sts_f1(smem_u32(sm.DBC[ob][h]) + tbc_off, kk);    // positions past the end are clipped by the tensor map
          else if (c_ok && (i > 0 || c0 + h * kHP + hw < L)) atomicAdd(dbc + dbc_sgn * (h * kHP), kk);
        }
        // ---------------- du / ddelta of my (row, half): sum the 4 lanes' partials, finalise ----------------
        if (hf == h) {
          const uint32_t srow = smem_u32(sm.S12[h]) + (uint32_t)row * (kS12P * 4);
          u64 s1a, s1b, s2a, s2b;           // s1 / s2 of my 4 positions as two packed pairs each
          lds_2x64(srow, s1a, s1b);
          lds_2x64(srow + 16, s2a, s2b);
#pragma unroll
          for (int q = 1; q < 4; ++q) {
            u64 xa, xb, ya, yb;
            lds_2x64(srow + q * 32, xa, xb);
            lds_2x64(srow + q * 32 + 16, ya, yb);
            s1a = add2(s1a, xa); s1b = add2(s1b, xb);
            s2a = add2(s2a, ya); s2b = add2(s2b, yb);
          }
          u64 ua, ub, ya, yb, la, lb;
          lds_2x64(smem_u32(sm.st[s].U) + my16, ua, ub);
          lds_2x64(smem_u32(sm.st[s].DY) + my16, ya, yb);
          lds_2x64(smem_u32(sm.st[s].SD) + my16, la, lb);
          const u64 Dv2 = pk2(Dv, Dv), ln2 = pk2(kLn2, kLn2);
          const u64 dua = fma2(la, s1a, mul2(Dv2, ya)), dub = fma2(lb, s1b, mul2(Dv2, yb));      // bwd_kernel.cuh:211, :280
          const u64 dda = mul2(fma2(ua, s1a, mul2(s2a, ln2)), pk2(sg_cur.x, sg_cur.y));          // :281-284, :446-450 (s2 in units of log2 e)
          const u64 ddb = mul2(fma2(ub, s1b, mul2(s2b, ln2)), pk2(sg_cur.z, sg_cur.w));
          // positions past the end contribute exact zeros (delta forced to 0, u / dout / B / C zero fill), so no masking is needed
          dbias2 = add2(dbias2, add2(dda, ddb));
          dD2 = fma2(ya, ua, fma2(yb, ub, dD2));                                                 // :213
          float4 o_du, o_dd;
          upk2(dua, o_du.x, o_du.y); upk2(dub, o_du.z, o_du.w);
          upk2(dda, o_dd.x, o_dd.y); upk2(ddb, o_dd.z, o_dd.w);
          sts_f4(smem_u32(sm.DU[ob]) + out16, unrev(o_du));     // mirrored groups: back into source order
          sts_f4(smem_u32(sm.DDT[ob]) + out16, unrev(o_dd));
        }
        __syncwarp();
        // With staged dB / dC tiles all four warps feed the store of a half: the other half's warps only arrive at its barrier,
        // and P / S12 of this half are handed back to the compute warps AFTER the barrier -- the next chunk's arrivals cannot mix
        // with this one's, and the owner's wait_read below (store of chunk i-1) is ordered before anyone's DBC writes of chunk i+1.
        if (!tbc && lane == 0) mbar_arrive(smem_u32(&sm.half_free[h]));
        if (hf == h) {                       // the two warps of this half hand their tiles to the TMA store
          if (hh == 1 && lane == 0) mbar_arrive(smem_u32(&sm.stage_free[s]));   // (warps of half 0 are done with the stage here ...
          fence_proxy_async_smem();          // my du / ddelta (/ dB / dC) writes -> visible to the async proxy
          if ((htid & 63) == 0) tma_store_wait_read<0>();   // my store of chunk i-1 has read its tiles: buffer ob^1 is free for chunk i+1
          if (tbc) {
            named_bar_sync(2 + h, kGroupThr);
            if (lane == 0) mbar_arrive(smem_u32(&sm.half_free[h]));
          } else {
            named_bar_sync(2 + h, 64);
          }
          if ((htid & 63) == 0) {
            if (tbc) {
              tma_reduce_add_3d(&map_db, smem_u32(sm.DBC[ob][h]), c0 + kHP * h, 0, b * a.ngroups + g);
              tma_reduce_add_3d(&map_dc, smem_u32(sm.DBC[ob][h]) + kStatePad * kHP * 4, c0 + kHP * h, 0, b * a.ngroups + g);
            }
            const int lo = rev ? L - (c0 + kHP * h) - kHP : c0 + kHP * h;   // mirrored: a half past the end (lo < 0) has nothing to store
            if (lo >= 0) {
              if (kMir) tma_reduce_add_3d(&map_du, smem_u32(sm.DU[ob]) + h * (kR * 16), lo, ds0, b);   // both groups of a pair add
              else tma_store_3d(&map_du, smem_u32(sm.DU[ob]) + h * (kR * 16), lo, d0, b);
              tma_store_3d(&map_ddt, smem_u32(sm.DDT[ob]) + h * (kR * 16), lo, d0, b);
            }
            tma_store_commit();
          }
        } else {
          if (tbc) {
            fence_proxy_async_smem();
            named_bar_arrive(2 + h, kGroupThr);
            if (lane == 0) mbar_arrive(smem_u32(&sm.half_free[h]));
          }
          if (hh == 1 && lane == 0) mbar_arrive(smem_u32(&sm.stage_free[s]));   //  ... and the warps of half 1 after contracting half 0)
        }
      }
      dbc -= dbc_sgn * kC;
      sg_cur = sg_nxt;
      cb0 = nb0;
      cb1 = nb1;
    }
    // every thread that committed TMA stores: their shared-memory sources have been read (the CTA may exit; the writes themselves
    // are complete when the grid is -- waiting for them here cost short sequences ~1 us per CTA)
    if ((htid & 63) == 0 || htid == kLoadThr) tma_store_wait_read<0>();
    if (a.dD != nullptr) atomicAdd(a.dD + d, hsum2(dD2));                  // two threads (halves) per channel, summed over batch
    if (a.ddelta_bias != nullptr) atomicAdd(a.ddelta_bias + d, hsum2(dbias2));
    return;
  }

  // =========================================== compute warpgroup ===========================================
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kCompRegs));
  const int sq = lane & 3;                // which 4 states
  const int pr = lane >> 2;               // channel pair inside the warp
  const int pp = warp * 8 + pr;           // channel pair inside the CTA: rows pp and pp + 32
  u64 A2p[2][2], dA2[2][2], w2[2][2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int d = d0 + pp + c * kNP;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n0 = sq * kLS + 2 * q;
      const float a0 = (n0 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)n0 * a.A_n_stride) * kLog2e : 0.f;
      const float a1 = (n0 + 1 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)(n0 + 1) * a.A_n_stride) * kLog2e : 0.f;
      A2p[c][q] = pk2(a0, a1);
      dA2[c][q] = pk2(0.f, 0.f);
      w2[c][q] = pk2(0.f, 0.f);   // a_{l+1} * dx_{l+1}: zero beyond the last position
    }
  }
  const uint32_t hoff0 = (uint32_t)((pp >> 2) & 1) << 4;   // byte offset of half 0 inside my rows of the swizzled [row][8] tiles (half 1: ^ 16)
  const uint32_t p_off = (uint32_t)pp * (kPP * 4) + (uint32_t)sq * 16;          // dB chunk; dC chunk at + 64
  const uint32_t s_off = (uint32_t)pp * (kS12P * 4) + (uint32_t)sq * 32;       // s1 x4; s2 x4 at + 16; second channel at + kNP rows

  for (int i = 0; i < n_tiles; ++i) {
    const int s = i % kStg;
    const uint32_t par = (uint32_t)((i / kStg) & 1);
    mbar_wait(smem_u32(&sm.prep_done[s]), par);
    mbar_wait(smem_u32(&sm.tma_full[s]), par);     // completed long ago: makes the TMA-written dout / state tiles visible to me
    const uint32_t dy_row = smem_u32(sm.st[s].DY) + pp * (kC * 4);   // + hoff[half] + c * kCT
    const uint32_t ck_row = smem_u32(sm.st[s].CK) + pp * (kStatePad * 4) + sq * 16;
    const uint32_t bc_base = smem_u32(sm.st[s].BC) + sq * (kLS * 4);
    const uint32_t sd_row = smem_u32(sm.st[s].SD) + pp * (kC * 4);
    const uint32_t sdu_row = smem_u32(sm.st[s].SDU) + pp * (kC * 4);

    float dl[2][kC];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float4 t0 = lds_f4(sd_row + hoff0 + c * kCT), t1 = lds_f4(sd_row + (hoff0 ^ 16u) + c * kCT);
      dl[c][0] = t0.x; dl[c][1] = t0.y; dl[c][2] = t0.z; dl[c][3] = t0.w;
      dl[c][4] = t1.x; dl[c][5] = t1.y; dl[c][6] = t1.z; dl[c][7] = t1.w;
    }
    // ---------------- forward recompute from the saved state ----------------
    u64 x0[2][2], xs[2][kC][2], Bk[kC][2];   // B of the chunk stays in registers for the reverse pass (the decays do not: see below)
    float duk[2][kC];                        // delta*u likewise
    lds_2x64(ck_row, x0[0][0], x0[0][1]);
    lds_2x64(ck_row + kNP * kStatePad * 4, x0[1][0], x0[1][1]);
    auto fwd_half = [&](auto HF) {
      constexpr int h = decltype(HF)::value;   // compile-time half: keeps xs[][][] in registers
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 v4 = lds_f4(sdu_row + (hoff0 ^ (h * 16u)) + c * kCT);
        duk[c][h * 4 + 0] = v4.x; duk[c][h * 4 + 1] = v4.y; duk[c][h * 4 + 2] = v4.z; duk[c][h * 4 + 3] = v4.w;
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = h * 4 + jj;
        lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bk[j][0], Bk[j][1]);
        const u64* Bp = Bk[j];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const u64 dd = pk2(dl[c][j], dl[c][j]);
          const u64 duu = pk2(duk[c][j], duk[c][j]);
          float t0, t1, t2, t3;
          upk2(mul2(dd, A2p[c][0]), t0, t1);
          upk2(mul2(dd, A2p[c][1]), t2, t3);
          const u64 e0 = pk2(ex2(t0), ex2(t1)), e1 = pk2(ex2(t2), ex2(t3));
          xs[c][j][0] = fma2(e0, j == 0 ? x0[c][0] : xs[c][j - 1][0], mul2(duu, Bp[0]));
          xs[c][j][1] = fma2(e1, j == 0 ? x0[c][1] : xs[c][j - 1][1], mul2(duu, Bp[1]));
        }
      }
    };
    fwd_half(std::integral_constant<int, 0>{});
    fwd_half(std::integral_constant<int, 1>{});
    // ---------------- reverse recurrence, one half chunk at a time ----------------
    auto rev_half = [&](auto HF) {
      constexpr int h = decltype(HF)::value;
      if (i > 0) mbar_wait(smem_u32(&sm.half_free[h]), (uint32_t)((i - 1) & 1));   // the helpers have consumed this half of chunk i-1
      const uint32_t p_row = smem_u32(sm.P[h]) + p_off;
      float dyh[2][4], s1p[2][4], s2p[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 y4 = lds_f4(dy_row + (hoff0 ^ (h * 16u)) + c * kCT);
        dyh[c][0] = y4.x; dyh[c][1] = y4.y; dyh[c][2] = y4.z; dyh[c][3] = y4.w;
      }
#pragma unroll
      for (int jj = 3; jj >= 0; --jj) {
        const int j = h * 4 + jj;
        const u64* Bp = Bk[j];
        u64 Cp[2];
        lds_2x64(bc_base + (uint32_t)j * (kPitch * 4) + 64, Cp[0], Cp[1]);
        u64 pB0 = 0, pB1 = 0, pC0 = 0, pC1 = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const u64 dyy = pk2(dyh[c][jj], dyh[c][jj]);
          const u64 dd = pk2(dl[c][j], dl[c][j]);
          const u64 duu = pk2(duk[c][j], duk[c][j]);
          // The decays are evaluated again (volatile: no reuse of the forward half's values).  What binds this kernel is the
          // shared-memory return path, MUFU has slack: 64 registers of decays would otherwise push B and delta*u out of the
          // register file and cost 12 more 128-bit shared loads per chunk.
          float t0, t1, t2, t3;
          upk2(mul2v(dd, A2p[c][0]), t0, t1);
          upk2(mul2v(dd, A2p[c][1]), t2, t3);
          const u64 e0 = pk2(ex2v(t0), ex2v(t1)), e1 = pk2(ex2v(t2), ex2v(t3));
          const u64 dx0 = fma2(Cp[0], dyy, w2[c][0]);                    // dx_{l,n}
          const u64 dx1 = fma2(Cp[1], dyy, w2[c][1]);
          s1p[c][jj] = hsum2(fma2(dx1, Bp[1], mul2(dx0, Bp[0])));        // sum_n dx * B          (bwd_kernel.cuh:280-281)
          w2[c][0] = mul2(e0, dx0);                                      // a_l * dx_l: carried to position l-1 ...
          w2[c][1] = mul2(e1, dx1);
          const u64 wg0 = mul2(w2[c][0], j == 0 ? x0[c][0] : xs[c][j - 1][0]);   // ... and dx * a_l * x_{l-1}  (:283)
          const u64 wg1 = mul2(w2[c][1], j == 0 ? x0[c][1] : xs[c][j - 1][1]);
          s2p[c][jj] = hsum2(fma2(wg1, A2p[c][1], mul2(wg0, A2p[c][0])));  // in units of log2(e)
          dA2[c][0] = fma2(wg0, dd, dA2[c][0]);                          // :286
          dA2[c][1] = fma2(wg1, dd, dA2[c][1]);
          if (c == 0) {                                                  // channel-pair products for dB / dC
            pB0 = mul2(duu, dx0); pB1 = mul2(duu, dx1);
            pC0 = mul2(dyy, xs[c][j][0]); pC1 = mul2(dyy, xs[c][j][1]);
          } else {
            pB0 = fma2(duu, dx0, pB0); pB1 = fma2(duu, dx1, pB1);
            pC0 = fma2(dyy, xs[c][j][0], pC0); pC1 = fma2(dyy, xs[c][j][1], pC1);
          }
        }
        sts_2x64(p_row + (uint32_t)jj * 128, pB0, pB1);
        sts_2x64(p_row + (uint32_t)jj * 128 + 64, pC0, pC1);
      }
      const uint32_t s_row = smem_u32(sm.S12[h]) + s_off;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        sts_f4(s_row + c * (kNP * kS12P * 4), make_float4(s1p[c][0], s1p[c][1], s1p[c][2], s1p[c][3]));
        sts_f4(s_row + c * (kNP * kS12P * 4) + 16, make_float4(s2p[c][0], s2p[c][1], s2p[c][2], s2p[c][3]));
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.half_full[h]));
    };
    rev_half(std::integral_constant<int, 1>{});
    rev_half(std::integral_constant<int, 0>{});
    if (lane == 0) mbar_arrive(smem_u32(&sm.stage_free[s]));   // ordered after my reads by the __syncwarp in rev_half
  }

#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int d = d0 + pp + c * kNP;
    float da[4];
    upk2(dA2[c][0], da[0], da[1]);
    upk2(dA2[c][1], da[2], da[3]);
#pragma unroll
    for (int n = 0; n < kLS; ++n)
      if (sq * kLS + n < N) atomicAdd(a.dA + (int64_t)d * N + sq * kLS + n, da[n]);   // sum over batch
  }
}

inline bool make_ckpt_map(CUtensorMap* map, const float* base, int64_t rows, int n_ckpt) {
  auto enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)n_ckpt * kStatePad, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)n_ckpt * kStatePad * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kStatePad, (cuuint32_t)kR};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// The warp-specialised kernel is usable when ptxas gave the kernel exactly the launch register count the setmaxnreg
// arithmetic assumes (otherwise setmaxnreg.inc could wait for registers that never come).
bool bwd_ws_usable() {
  static const bool ok = [] {
    cudaFuncAttributes fa[7];
    if (cudaFuncGetAttributes(&fa[0], selscan_bwd_ws_kernel<false, 0, false>) != cudaSuccess ||
        cudaFuncGetAttributes(&fa[1], selscan_bwd_ws_kernel<true, 0, false>) != cudaSuccess ||
        cudaFuncGetAttributes(&fa[2], selscan_bwd_ws_kernel<false, 6, false>) != cudaSuccess ||
        cudaFuncGetAttributes(&fa[3], selscan_bwd_ws_kernel<false, kMaxFusedDtRank, false>) != cudaSuccess ||
        cudaFuncGetAttributes(&fa[4], selscan_bwd_ws_kernel<false, 0, true>) != cudaSuccess ||
        cudaFuncGetAttributes(&fa[5], selscan_bwd_ws_kernel<false, 6, true>) != cudaSuccess ||
        cudaFuncGetAttributes(&fa[6], selscan_bwd_ws_kernel<false, kMaxFusedDtRank, true>) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    for (int i = 0; i < 7; ++i)
      if (fa[i].numRegs != kLaunchRegs) return false;
    return true;
  }();
  return ok;
}

bool bwd_ws_eligible(const BwdLaunch& p) {
  const selscan_bwd_args& a = p.a;
  if (a.dstate > kStatePad) return false;
  if (p.dim_per_group % kR != 0) return false;
  if (p.n_ckpt < 1) return false;                      // seqlen > 8
  if (a.B_l_stride != a.C_l_stride) return false;      // the B/C gather walks both with one step
  const int64_t one_batch = 4;   // batch 1: the batch stride is unused (any positive multiple of 4 passes the check)
  if (!tma_row_ok(a.u, a.u_d_stride, a.batch > 1 ? a.u_batch_stride : one_batch)) return false;
  if (a.dt_w == nullptr && !tma_row_ok(a.delta, a.delta_d_stride, a.batch > 1 ? a.delta_batch_stride : one_batch)) return false;
  if (!tma_row_ok(a.dout, a.dout_d_stride, a.batch > 1 ? a.dout_batch_stride : one_batch)) return false;
  if (!tma_row_ok(a.du, a.du_d_stride, a.batch > 1 ? a.du_batch_stride : one_batch)) return false;
  if (!tma_row_ok(a.ddelta, a.ddelta_d_stride, a.batch > 1 ? a.ddelta_batch_stride : one_batch)) return false;
  if (a.z != nullptr) {   // the gated path streams z, the ungated forward output and dz as well
    if (a.out == nullptr || a.dz == nullptr) return false;
    if (!tma_row_ok(a.z, a.z_d_stride, a.batch > 1 ? a.z_batch_stride : one_batch)) return false;
    if (!tma_row_ok(a.out, a.out_d_stride, a.batch > 1 ? a.out_batch_stride : one_batch)) return false;
    if (!tma_row_ok(a.dz, a.dz_d_stride, a.batch > 1 ? a.dz_batch_stride : one_batch)) return false;
  }
  if ((reinterpret_cast<uintptr_t>(a.ckpt) & 15u) != 0) return false;
  return tensor_map_encoder() != nullptr && bwd_ws_usable();
}

// 4-D map over dt_x (fastest first: seqlen, rank, group, batch), box = 8 positions x kDt ranks, dense 32-byte rows
inline bool make_dtx_map_bwd(CUtensorMap* map, const selscan_bwd_args& a, int box_r) {
  auto enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[4] = {(cuuint64_t)a.seqlen, (cuuint64_t)a.dt_rank, (cuuint64_t)a.ngroups, (cuuint64_t)a.batch};
  const cuuint64_t gstr[3] = {(cuuint64_t)a.dt_x_r_stride * 4, (cuuint64_t)(a.ngroups > 1 ? a.dt_x_group_stride : a.dt_x_r_stride * a.dt_rank) * 4,
                              (cuuint64_t)(a.batch > 1 ? a.dt_x_batch_stride : a.dt_x_r_stride * a.dt_rank * a.ngroups) * 4};
  const cuuint32_t box[4] = {(cuuint32_t)kC, (cuuint32_t)box_r, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a.dt_x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool kHasZ, int kDt, bool kMir = false>
cudaError_t launch_ws_variant(const CUtensorMap (&m)[11], int tma_bc, const BwdLaunch& p, unsigned grid, cudaStream_t stream) {
  constexpr int smem = (int)sizeof(WsSmemT<kHasZ, kDt>) + 256;
  static_assert(kHasZ || sizeof(WsSmemT<kHasZ, kDt>) + 256 + 1024 <= 116736, "two CTAs per SM");
  static_assert(sizeof(WsSmemT<kHasZ, kDt>) + 256 <= 232448, "fits one CTA per SM");
  static std::atomic<unsigned long long> configured{0};   // one cudaFuncSetAttribute per device, not per launch
  if (const cudaError_t e = set_smem_once(configured, selscan_bwd_ws_kernel<kHasZ, kDt, kMir>, smem)) return e;
  selscan_bwd_ws_kernel<kHasZ, kDt, kMir><<<grid, kThr, smem, stream>>>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10],
                                                                        tma_bc, p);
  return cudaGetLastError();
}

// Returns cudaErrorNotSupported when a tensor map cannot be encoded for this layout: the caller then takes the generic kernel.
cudaError_t launch_bwd_ws(const BwdLaunch& p, cudaStream_t stream) {
  const selscan_bwd_args& a = p.a;
  CUtensorMap m[11];  // u, delta (or dt_x), dout, saved states, du, ddelta, z, out, dz, dB, dC
  const int dt_box = a.dt_w == nullptr ? 0 : (a.dt_rank <= 6 ? 6 : kMaxFusedDtRank);
  const int src_rows = a.mirror_pairs ? a.dim / 2 : a.dim;   // rows per batch of u / dout / du
  if (!make_row_map_sw(&m[0], a.u, a.seqlen, src_rows, a.batch, a.u_d_stride, a.u_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B) ||
      !(dt_box ? make_dtx_map_bwd(&m[1], a, dt_box)
               : make_row_map_sw(&m[1], a.delta, a.seqlen, a.dim, a.batch, a.delta_d_stride, a.delta_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B)) ||
      !make_row_map_sw(&m[2], a.dout, a.seqlen, src_rows, a.batch, a.dout_d_stride, a.dout_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B) ||
      !make_ckpt_map(&m[3], a.ckpt, (int64_t)a.batch * a.dim, p.n_ckpt) ||
      !make_row_map(&m[4], a.du, a.seqlen, src_rows, a.batch, a.du_d_stride, a.du_batch_stride, kHP, kR) ||
      !make_row_map(&m[5], a.ddelta, a.seqlen, a.dim, a.batch, a.ddelta_d_stride, a.ddelta_batch_stride, kHP, kR))
    return cudaErrorNotSupported;
  m[6] = m[7] = m[8] = m[9] = m[10] = m[0];
  // dB / dC as TMA reduce-adds: 16 states, rows of seqlen floats 16-byte aligned, source-order positions (not the mirrored groups)
  const int tma_bc = !a.mirror_pairs && a.dstate == kStatePad && a.seqlen % 4 == 0 && ((reinterpret_cast<uintptr_t>(a.dB) | reinterpret_cast<uintptr_t>(a.dC)) & 15u) == 0 &&
                     make_row_map(&m[9], a.dB, a.seqlen, a.dstate, a.batch * a.ngroups, a.seqlen, (int64_t)a.dstate * a.seqlen, kHP, kStatePad, false) &&
                     make_row_map(&m[10], a.dC, a.seqlen, a.dstate, a.batch * a.ngroups, a.seqlen, (int64_t)a.dstate * a.seqlen, kHP, kStatePad, false);
  const unsigned grid = (unsigned)((int64_t)a.batch * a.ngroups * (p.dim_per_group / kR));
  if (a.z != nullptr) {
    if (dt_box || a.mirror_pairs) return cudaErrorNotSupported;   // (the C ABI rejects the combinations earlier)
    if (!make_row_map_sw(&m[6], a.z, a.seqlen, a.dim, a.batch, a.z_d_stride, a.z_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B) ||
        !make_row_map_sw(&m[7], a.out, a.seqlen, a.dim, a.batch, a.out_d_stride, a.out_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B) ||
        !make_row_map_sw(&m[8], a.dz, a.seqlen, a.dim, a.batch, a.dz_d_stride, a.dz_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B))
      return cudaErrorNotSupported;
    return launch_ws_variant<true, 0>(m, tma_bc, p, grid, stream);
  }
  if (a.mirror_pairs) {
    if (dt_box == 6) return launch_ws_variant<false, 6, true>(m, tma_bc, p, grid, stream);
    if (dt_box) return launch_ws_variant<false, kMaxFusedDtRank, true>(m, tma_bc, p, grid, stream);
    return launch_ws_variant<false, 0, true>(m, tma_bc, p, grid, stream);
  }
  if (dt_box == 6) return launch_ws_variant<false, 6>(m, tma_bc, p, grid, stream);
  if (dt_box) return launch_ws_variant<false, kMaxFusedDtRank>(m, tma_bc, p, grid, stream);
  return launch_ws_variant<false, 0>(m, tma_bc, p, grid, stream);
}

}  // namespace selscan
