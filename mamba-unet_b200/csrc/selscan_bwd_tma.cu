// Backward selective scan, tiled path for sm_100a: TMA-staged tiles walked from the end of the sequence,
// mbarrier pipeline, packed f32x2 arithmetic.  Replaces selective_scan_bwd_kernel
// (/root/reference/mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489) for the aligned shapes Mamba-UNet
// produces (channels per group a multiple of 32, 16-byte aligned rows, no z); everything else takes selscan_bwd.cu.
//
// CTA = 32 channels of one (batch, group); the sequence is walked backwards in tiles of 16 positions = 2 chunks of 8
// (the checkpoint interval).
//   warp 4 (producer): per tile one elected lane issues TMA loads of u, delta, dout (box 32 rows x 16 positions) and of
//       the two saved scan states the tile's chunks restart from (box 32 rows x 2 states; the state "before position
//       0" is the tensor map's out-of-bounds zero fill); all lanes gather the tile's B/C into a [position][32] tile.
//   warps 0-3 (compute): 8 channels each, four lanes per channel, 4 states per lane.  Per chunk:
//       prep     lane j of a channel discretises positions 2j, 2j+1 (softplus, sigmoid) and publishes delta and
//                delta*u for its channel; nothing is computed twice;
//       forward  restart from the saved state: per position and state pair FMUL2, 2 MUFU.EX2, FMUL2, FFMA2; decays and
//                states stay in registers (packed), states also go to the swizzled X tile;
//       reverse  dx = C*dy + a*dx', in registers; dx goes to the DX tile; partial sums of du / ddelta over the lane's 4
//                states, dA accumulated in registers;
//       reduce   the partial sums are reduce-scattered over the 4 lanes (lane j finalises positions 2j, 2j+1: du, ddelta
//                through softplus', dD, ddelta_bias) into du/ddelta tiles that leave by per-warp TMA stores.
//   contraction: after a named barrier over the 4 compute warps, the CTA contracts X and DX over its 32 channels
//       (dC = sum_d dy*x, dB = sum_d delta*u*dx; thread = (tensor, position, 4 states, row parity)) and adds ONE value
//       per (state, position) to global memory -- the reference issues one atomic per (channel, state, position)
//       (bwd_kernel.cuh:298-316).
#include "selscan_common.cuh"
#include "selscan_kernels.h"
#include "selscan_ptx.cuh"
#include "selscan_tma_host.h"

namespace selscan {

namespace {

constexpr int kR = 32;            // channels per CTA
constexpr int kW = 4;             // compute warps
constexpr int kT = 16;            // positions per staged tile
constexpr int kC = kCkptInterval; // positions per chunk (8)
constexpr int kStg = 2;
constexpr int kPitch = 36;        // B/C tile pitch (floats)
constexpr int kThr = (kW + 1) * 32;
constexpr int kLS = kStatePad / 4;  // states per lane

struct BwdTmaSmem {
  float CK[kStg][kR * 32];          // 4 KB: [row][2 states x 16], 128B swizzle
  float U[kStg][kR * kT];           // 2 KB: [row][16]
  float DT[kStg][kR * kT];
  float DY[kStg][kR * kT];
  float BC[kStg][kT * kPitch];
  float X[kR * kC * kStatePad];     // 16 KB, swizzled [row][pos][state]
  float DX[kR * kC * kStatePad];
  float SD[kR * kC];                // delta   per (row, position of the chunk)
  float SDU[kR * kC];               // delta*u
  float DU[kW][2][8 * kT];          // per-warp output tiles, double-buffered (512 B each)
  float DDT[kW][2][8 * kT];
  u64 full[kStg];
  u64 empty[kStg];
};

__device__ __forceinline__ int swz_row(int r) { return ((r & 1) << 2) | ((r >> 1) & 3); }
// float index of the 4-state group nq of position j of row r in the X / DX tiles
__device__ __forceinline__ int xt_idx(int r, int j, int nq) { return r * (kC * kStatePad) + (((j * 4 + nq) ^ swz_row(r)) << 2); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}

__global__ void __launch_bounds__(kThr, 3)
selscan_bwd_tma_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_dt,
                       const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_ck,
                       const __grid_constant__ CUtensorMap map_du, const __grid_constant__ CUtensorMap map_ddt, const BwdLaunch p) {
  extern __shared__ unsigned char smem_raw[];
  BwdTmaSmem& sm = *reinterpret_cast<BwdTmaSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const selscan_bwd_args& a = p.a;
  const int L = a.seqlen, N = a.dstate;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_per_group = p.dim_per_group / kR;
  int bid = blockIdx.x;
  const int tile_g = bid % tiles_per_group; bid /= tiles_per_group;
  const int g = bid % a.ngroups;
  const int b = bid / a.ngroups;
  const int d0 = g * p.dim_per_group + tile_g * kR;
  const int n_tiles = (L + kT - 1) / kT;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStg; ++s) {
      mbar_init(smem_u32(&sm.full[s]), 32);
      mbar_init(smem_u32(&sm.empty[s]), kW);
    }
    mbar_fence_init();
    tma_prefetch_desc(&map_u);
    tma_prefetch_desc(&map_dt);
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_ck);
    tma_prefetch_desc(&map_du);
    tma_prefetch_desc(&map_ddt);
  }
  __syncthreads();

  if (warp == kW) {
    // ================================ producer ================================
    const float* __restrict__ Bg = a.B + (int64_t)b * a.B_batch_stride + (int64_t)g * a.B_group_stride;
    const float* __restrict__ Cg = a.C + (int64_t)b * a.C_batch_stride + (int64_t)g * a.C_group_stride;
    const bool lanes_along_l = (a.B_l_stride == 1 && a.C_l_stride == 1);
    const int row0 = b * a.dim + d0;
    for (int i = 0; i < n_tiles; ++i) {
      const int t = n_tiles - 1 - i;                 // tiles are consumed last -> first
      const int s = i % kStg, k = i / kStg;
      if (k > 0) mbar_wait(smem_u32(&sm.empty[s]), (k - 1) & 1);
      const int l0 = t * kT;
      const uint32_t full = smem_u32(&sm.full[s]);
      if (lane == 0) {
        mbar_expect_tx(full, (uint32_t)(3 * kR * kT * 4 + kR * 32 * 4));
        tma_load_3d(smem_u32(sm.U[s]), &map_u, l0, d0, b, full);
        tma_load_3d(smem_u32(sm.DT[s]), &map_dt, l0, d0, b, full);
        tma_load_3d(smem_u32(sm.DY[s]), &map_dy, l0, d0, b, full);
        // saved states 2t-1 (start of chunk 2t) and 2t (start of chunk 2t+1); state "-1" is out of bounds -> zeros
        tma_load_2d(smem_u32(sm.CK[s]), &map_ck, (2 * t - 1) * kStatePad, row0, full);
      }
      float* bc = sm.BC[s];
      float v[16];
      if (lanes_along_l) {   // lane = (B|C, position): 64 contiguous bytes per half warp
        const int l = l0 + (lane & 15);
        const int which = lane >> 4;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          const float* src = which ? (Cg + (int64_t)n * a.C_n_stride) : (Bg + (int64_t)n * a.B_n_stride);
          v[n] = (n < N && l < L) ? __ldg(src + l) : 0.f;
        }
#pragma unroll
        for (int n = 0; n < 16; ++n) bc[(lane & 15) * kPitch + which * 16 + n] = v[n];
      } else {               // lane = (B|C, state): the 16 B and 16 C values of one position are contiguous
        const int n = lane & 15;
        const float* src = (lane < 16) ? (Bg + (int64_t)n * a.B_n_stride) : (Cg + (int64_t)n * a.C_n_stride);
        const int64_t ls = (lane < 16) ? a.B_l_stride : a.C_l_stride;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (n < N && l0 + j < L) ? __ldg(src + (int64_t)(l0 + j) * ls) : 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) bc[j * kPitch + lane] = v[j];
      }
      mbar_arrive(full);
    }
    return;
  }

  // ================================ compute warps ================================
  const int tid = threadIdx.x;            // 0..127
  const int sq = lane & 3;                // which 4 states
  const int rw = lane >> 2;               // channel inside the warp
  const int r = warp * 8 + rw;            // channel inside the CTA
  const int d = d0 + r;
  const bool hi1 = (sq & 2) != 0, hi0 = (sq & 1) != 0;
  u64 A2p[2], dA2[2], w2[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int n0 = sq * kLS + 2 * q;
    const float a0 = (n0 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)n0 * a.A_n_stride) * kLog2e : 0.f;
    const float a1 = (n0 + 1 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)(n0 + 1) * a.A_n_stride) * kLog2e : 0.f;
    A2p[q] = pk2(a0, a1);
    dA2[q] = pk2(0.f, 0.f);
    w2[q] = pk2(0.f, 0.f);   // a_{l+1} * dx_{l+1}: zero beyond the last position
  }
  const float Dv = a.D ? __ldg(a.D + d) : 0.f;
  const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.f;
  const bool softplus = a.delta_softplus != 0;
  float dD_acc = 0.f, dbias_acc = 0.f;
  // contraction role of this thread
  const int c_which = tid >> 6;                       // 0: dB from DX and delta*u, 1: dC from X and dy
  const int c_j = ((tid >> 5) & 1) * 4 + ((lane >> 2) & 3);
  const int c_nq = lane & 3;
  const int c_par = lane >> 4;                        // row parity handled by this lane
  float* __restrict__ dBC = (c_which ? a.dC : a.dB) + ((int64_t)b * a.ngroups + g) * N * (int64_t)L;

  for (int i = 0; i < n_tiles; ++i) {
    const int t = n_tiles - 1 - i;
    const int s = i % kStg, k = i / kStg;
    const int l0 = t * kT;
    mbar_wait(smem_u32(&sm.full[s]), k & 1);
    const uint32_t u_row = smem_u32(sm.U[s]) + r * (kT * 4);
    const uint32_t dt_row = smem_u32(sm.DT[s]) + r * (kT * 4);
    const uint32_t dy_row = smem_u32(sm.DY[s]) + r * (kT * 4);
    const uint32_t ck_row = smem_u32(sm.CK[s]) + r * 128;
    const uint32_t du_tile = smem_u32(sm.DU[warp][i & 1]);
    const uint32_t ddt_tile = smem_u32(sm.DDT[warp][i & 1]);

#pragma unroll 1
    for (int cc = 1; cc >= 0; --cc) {       // chunk 2t+1 first, then 2t
      const int c0 = l0 + cc * kC;          // first position of the chunk
      if (c0 >= L) continue;                // (only the second chunk of the last tile can be empty; uniform)
      // ---------------- prep: my two positions of the chunk ----------------
      const int pj = 2 * sq;                // chunk-local positions pj, pj+1
      float my_dl[2], my_sg[2], my_u[2], my_dy[2];
      {
        float2 uu, dd, yy;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(uu.x), "=f"(uu.y) : "r"(u_row + (cc * kC + pj) * 4));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(dd.x), "=f"(dd.y) : "r"(dt_row + (cc * kC + pj) * 4));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(yy.x), "=f"(yy.y) : "r"(dy_row + (cc * kC + pj) * 4));
        const float uin[2] = {uu.x, uu.y}, din[2] = {dd.x, dd.y}, yin[2] = {yy.x, yy.y};
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float xb = din[e] + bias;
          float wexp = 0.f;
          float v = xb;
          float sgm = 1.f;
          if (softplus) {
            v = softplus_fast(xb, wexp);
            sgm = sigmoid_from_w(xb, wexp);   // softplus' (bwd_kernel.cuh:446-450; == 1 to rounding for x > 20)
          }
          const bool valid = (c0 + pj + e) < L;
          my_dl[e] = valid ? v : 0.f;          // past the end: a = 1, b = 0 (u and dout are TMA zero fill there)
          my_sg[e] = sgm;
          my_u[e] = uin[e];
          my_dy[e] = yin[e];
        }
        float* sd = &sm.SD[r * kC + pj];
        float* sdu = &sm.SDU[r * kC + pj];
        *reinterpret_cast<float2*>(sd) = make_float2(my_dl[0], my_dl[1]);
        *reinterpret_cast<float2*>(sdu) = make_float2(my_dl[0] * my_u[0], my_dl[1] * my_u[1]);
      }
      __syncwarp();
      float dl[kC], du_[kC], dy[kC];
      {
        const float4 t0 = *reinterpret_cast<const float4*>(&sm.SD[r * kC]);
        const float4 t1 = *reinterpret_cast<const float4*>(&sm.SD[r * kC + 4]);
        const float4 v0 = *reinterpret_cast<const float4*>(&sm.SDU[r * kC]);
        const float4 v1 = *reinterpret_cast<const float4*>(&sm.SDU[r * kC + 4]);
        const float4 y0 = lds_f4(dy_row + (cc * kC) * 4);
        const float4 y1 = lds_f4(dy_row + (cc * kC + 4) * 4);
        dl[0] = t0.x; dl[1] = t0.y; dl[2] = t0.z; dl[3] = t0.w; dl[4] = t1.x; dl[5] = t1.y; dl[6] = t1.z; dl[7] = t1.w;
        du_[0] = v0.x; du_[1] = v0.y; du_[2] = v0.z; du_[3] = v0.w; du_[4] = v1.x; du_[5] = v1.y; du_[6] = v1.z; du_[7] = v1.w;
        dy[0] = y0.x; dy[1] = y0.y; dy[2] = y0.z; dy[3] = y0.w; dy[4] = y1.x; dy[5] = y1.y; dy[6] = y1.z; dy[7] = y1.w;
      }
      // ---------------- forward recompute from the saved state ----------------
      u64 x0[2], ea[kC][2], xs[kC][2];
      lds_2x64(ck_row + ((uint32_t)((cc * 4 + sq) ^ (r & 7)) << 4), x0[0], x0[1]);
      const uint32_t bc_base = smem_u32(sm.BC[s]) + (uint32_t)(cc * kC) * (kPitch * 4) + sq * (kLS * 4);
#pragma unroll
      for (int j = 0; j < kC; ++j) {
        u64 Bp[2];
        lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bp[0], Bp[1]);
        const u64 dd = pk2(dl[j], dl[j]);
        const u64 duu = pk2(du_[j], du_[j]);
        float t0, t1, t2, t3;
        upk2(mul2(dd, A2p[0]), t0, t1);
        upk2(mul2(dd, A2p[1]), t2, t3);
        ea[j][0] = pk2(ex2(t0), ex2(t1));
        ea[j][1] = pk2(ex2(t2), ex2(t3));
        xs[j][0] = fma2(ea[j][0], j == 0 ? x0[0] : xs[j - 1][0], mul2(duu, Bp[0]));
        xs[j][1] = fma2(ea[j][1], j == 0 ? x0[1] : xs[j - 1][1], mul2(duu, Bp[1]));
        float4 xv;
        upk2(xs[j][0], xv.x, xv.y);
        upk2(xs[j][1], xv.z, xv.w);
        *reinterpret_cast<float4*>(&sm.X[xt_idx(r, j, sq)]) = xv;
      }
      // ---------------- reverse recurrence ----------------
      float s1p[kC], s2p[kC];
#pragma unroll
      for (int j = kC - 1; j >= 0; --j) {
        u64 Bp[2], Cp[2];
        lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bp[0], Bp[1]);
        lds_2x64(bc_base + (uint32_t)j * (kPitch * 4) + 64, Cp[0], Cp[1]);
        const u64 dyy = pk2(dy[j], dy[j]);
        const u64 dd = pk2(dl[j], dl[j]);
        const u64 dx0 = fma2(Cp[0], dyy, w2[0]);                      // dx_{l,n}
        const u64 dx1 = fma2(Cp[1], dyy, w2[1]);
        float4 dv;
        upk2(dx0, dv.x, dv.y);
        upk2(dx1, dv.z, dv.w);
        *reinterpret_cast<float4*>(&sm.DX[xt_idx(r, j, sq)]) = dv;
        s1p[j] = hsum2(fma2(dx1, Bp[1], mul2(dx0, Bp[0])));           // sum_n dx * B          (bwd_kernel.cuh:280-281)
        const u64 g0 = mul2(ea[j][0], j == 0 ? x0[0] : xs[j - 1][0]);  // a_l * x_{l-1}         (:283, x - b form)
        const u64 g1 = mul2(ea[j][1], j == 0 ? x0[1] : xs[j - 1][1]);
        const u64 wg0 = mul2(dx0, g0), wg1 = mul2(dx1, g1);
        s2p[j] = hsum2(fma2(wg1, A2p[1], mul2(wg0, A2p[0])));          // in units of log2(e)
        dA2[0] = fma2(wg0, dd, dA2[0]);                                 // :286
        dA2[1] = fma2(wg1, dd, dA2[1]);
        w2[0] = mul2(ea[j][0], dx0);                                    // carried to position l-1
        w2[1] = mul2(ea[j][1], dx1);
      }
      // ---------------- reduce-scatter s1 / s2 over the 4 lanes: lane sq finalises positions 2sq, 2sq+1 ----------------
      float f1[2], f2[2];
      {
        float k1[4], k2[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float snd1 = hi1 ? s1p[e] : s1p[4 + e], snd2 = hi1 ? s2p[e] : s2p[4 + e];
          k1[e] = (hi1 ? s1p[4 + e] : s1p[e]) + __shfl_xor_sync(0xffffffffu, snd1, 2);
          k2[e] = (hi1 ? s2p[4 + e] : s2p[e]) + __shfl_xor_sync(0xffffffffu, snd2, 2);
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float snd1 = hi0 ? k1[e] : k1[2 + e], snd2 = hi0 ? k2[e] : k2[2 + e];
          f1[e] = (hi0 ? k1[2 + e] : k1[e]) + __shfl_xor_sync(0xffffffffu, snd1, 1);
          f2[e] = (hi0 ? k2[2 + e] : k2[e]) + __shfl_xor_sync(0xffffffffu, snd2, 1);
        }
      }
      {
        float o_du[2], o_dd[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          o_du[e] = fmaf(my_dl[e], f1[e], Dv * my_dy[e]);                          // :211, :280
          o_dd[e] = fmaf(my_u[e], f1[e], f2[e] * kLn2) * my_sg[e];                 // :281-284, :446-450
          dbias_acc += ((c0 + pj + e) < L) ? o_dd[e] : 0.f;
          dD_acc = fmaf(my_dy[e], my_u[e], dD_acc);                                // :213
        }
        const uint32_t off = (uint32_t)(rw * kT + cc * kC + pj) * 4;
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(du_tile + off), "f"(o_du[0]), "f"(o_du[1]) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ddt_tile + off), "f"(o_dd[0]), "f"(o_dd[1]) : "memory");
      }
      named_bar_sync(1, kW * 32);   // X, DX, SDU complete for the CTA's 32 channels
      // ---------------- contraction over the channels ----------------
      {
        const float* __restrict__ src = c_which ? sm.X : sm.DX;
        u64 acc0 = pk2(0.f, 0.f), acc1 = pk2(0.f, 0.f);
#pragma unroll 8
        for (int rr = 0; rr < kR / 2; ++rr) {
          const int r2 = rr * 2 + c_par;
          u64 v0, v1;
          lds_2x64(smem_u32(&src[xt_idx(r2, c_j, c_nq)]), v0, v1);
          const float sc = c_which ? sm.DY[s][r2 * kT + cc * kC + c_j] : sm.SDU[r2 * kC + c_j];
          const u64 ss = pk2(sc, sc);
          acc0 = fma2(ss, v0, acc0);
          acc1 = fma2(ss, v1, acc1);
        }
        float o[4];
        upk2(acc0, o[0], o[1]);
        upk2(acc1, o[2], o[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
        if (c_par == 0 && c0 + c_j < L) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c_nq * 4 + e < N) atomicAdd(dBC + (int64_t)(c_nq * 4 + e) * L + (c0 + c_j), o[e]);
        }
      }
      named_bar_sync(1, kW * 32);   // tiles free for the next chunk
    }
    // ---------------- tile done: release the stage, ship du / ddelta ----------------
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(smem_u32(&sm.empty[s]));
      tma_store_3d(&map_du, du_tile, l0, d0 + warp * 8, b);
      tma_store_3d(&map_ddt, ddt_tile, l0, d0 + warp * 8, b);
      tma_store_commit();
      tma_store_wait_read<1>();
    }
    __syncwarp();
  }
  if (lane == 0) tma_store_wait_all<0>();

  // dA: sum over batch through atomics (one per (channel, state) per CTA)
  {
    float da[4];
    upk2(dA2[0], da[0], da[1]);
    upk2(dA2[1], da[2], da[3]);
#pragma unroll
    for (int n = 0; n < kLS; ++n)
      if (sq * kLS + n < N) atomicAdd(a.dA + (int64_t)d * N + sq * kLS + n, da[n]);
  }
  dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, 1);
  dD_acc += __shfl_xor_sync(0xffffffffu, dD_acc, 2);
  dbias_acc += __shfl_xor_sync(0xffffffffu, dbias_acc, 1);
  dbias_acc += __shfl_xor_sync(0xffffffffu, dbias_acc, 2);
  if (sq == 0) {
    if (a.dD != nullptr) atomicAdd(a.dD + d, dD_acc);
    if (a.ddelta_bias != nullptr) atomicAdd(a.ddelta_bias + d, dbias_acc);
  }
}

inline bool make_ckpt_map(CUtensorMap* map, const float* base, int64_t rows, int n_ckpt) {
  auto enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)n_ckpt * kStatePad, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)n_ckpt * kStatePad * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)kR};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool bwd_tma_eligible(const BwdLaunch& p) {
  const selscan_bwd_args& a = p.a;
  if (a.z != nullptr) return false;
  if (p.dim_per_group % kR != 0) return false;
  if (p.n_ckpt < 2 || (a.seqlen & 3) != 0) return false;   // the checkpoint box spans two saved states
  const int64_t zero = 0;
  if (!tma_row_ok(a.u, a.u_d_stride, a.batch > 1 ? a.u_batch_stride : zero)) return false;
  if (!tma_row_ok(a.delta, a.delta_d_stride, a.batch > 1 ? a.delta_batch_stride : zero)) return false;
  if (!tma_row_ok(a.dout, a.dout_d_stride, a.batch > 1 ? a.dout_batch_stride : zero)) return false;
  if (!tma_row_ok(a.du, a.seqlen, zero) || !tma_row_ok(a.ddelta, a.seqlen, zero)) return false;
  if ((reinterpret_cast<uintptr_t>(a.ckpt) & 15u) != 0) return false;
  return tensor_map_encoder() != nullptr;
}

cudaError_t launch_bwd_tma(const BwdLaunch& p, cudaStream_t stream) {
  const selscan_bwd_args& a = p.a;
  CUtensorMap mu, mdt, mdy, mck, mdu, mddt;
  const int64_t cs = (int64_t)a.dim * a.seqlen;
  if (!make_row_map(&mu, a.u, a.seqlen, a.dim, a.batch, a.u_d_stride, a.u_batch_stride, kT, kR) ||
      !make_row_map(&mdt, a.delta, a.seqlen, a.dim, a.batch, a.delta_d_stride, a.delta_batch_stride, kT, kR) ||
      !make_row_map(&mdy, a.dout, a.seqlen, a.dim, a.batch, a.dout_d_stride, a.dout_batch_stride, kT, kR) ||
      !make_row_map(&mdu, a.du, a.seqlen, a.dim, a.batch, a.seqlen, cs, kT, 8) ||
      !make_row_map(&mddt, a.ddelta, a.seqlen, a.dim, a.batch, a.seqlen, cs, kT, 8) ||
      !make_ckpt_map(&mck, a.ckpt, (int64_t)a.batch * a.dim, p.n_ckpt))
    return cudaErrorInvalidValue;
  const int smem = (int)sizeof(BwdTmaSmem) + 1024;
  cudaError_t e = cudaFuncSetAttribute(selscan_bwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((int64_t)a.batch * a.ngroups * (p.dim_per_group / kR));
  selscan_bwd_tma_kernel<<<grid, kThr, smem, stream>>>(mu, mdt, mdy, mck, mdu, mddt, p);
  return cudaGetLastError();
}

}  // namespace selscan
