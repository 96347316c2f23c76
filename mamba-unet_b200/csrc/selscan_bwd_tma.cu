// Backward selective scan, tiled path for sm_100a: TMA-staged tiles walked from the end of the sequence,
// mbarrier pipeline, packed f32x2 arithmetic.  Replaces selective_scan_bwd_kernel
// (/root/reference/mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489) for the aligned shapes Mamba-UNet
// produces (channels per group a multiple of 32, 16-byte aligned rows, no z); everything else takes selscan_bwd.cu.
//
// CTA = 32 channels of one (batch, group); the sequence is walked backwards in chunks of 8 positions (the checkpoint
// interval).  52 KB of shared memory and 96 registers per thread: four CTAs = 16 compute warps per SM, which makes a
// batch-24 stage-1 call exactly one wave.
//   warp 4 (producer): per chunk one elected lane issues TMA loads of u, delta, dout (box 32 rows x 8 positions) and of
//       the saved scan state the chunk restarts from (box 32 rows x 16 states; the state "before position 0" is the
//       tensor map's out-of-bounds zero fill); all lanes gather the chunk's B/C into a [position][32] tile.
//   warps 0-3 (compute): 8 channels each, four lanes per channel, 4 states per lane.  Per chunk:
//       prep     lane j of a channel discretises positions 2j, 2j+1 (softplus, sigmoid) and publishes delta and
//                delta*u for its channel; nothing is computed twice;
//       forward  restart from the saved state: per position and state pair FMUL2, 2 MUFU.EX2, FMUL2, FFMA2; the
//                states go to the swizzled X tile (decays are recomputed in the reverse pass: MUFU has slack,
//                registers do not);
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
constexpr int kT = 8;             // positions per staged tile = one chunk
constexpr int kC = kCkptInterval; // positions per chunk (8)
constexpr int kStg = 2;
constexpr int kPitch = 36;        // B/C tile pitch (floats)
constexpr int kThr = (kW + 1) * 32;
constexpr int kLS = kStatePad / 4;  // states per lane
constexpr int kSP = 12;           // row pitch of the per-row scalar scratch (floats)
constexpr int kTP = 36;           // position pitch of the transposed scalar scratch (floats)

struct BwdTmaSmem {
  float CK[kStg][kR * kStatePad];   // 2 KB: [row][16 states]
  float U[kStg][kR * kT];           // 1 KB: [row][8]
  float DT[kStg][kR * kT];
  float DY[kStg][kR * kT];
  float BC[kStg][kT * kPitch];
  float X[kR * kC * kStatePad];     // 16 KB, swizzled [row][pos][state]
  float DX[kR * kC * kStatePad];
  float SD[kR * kSP];               // delta   per (row, position of the chunk): [row][pos], pitch 12 (conflict-free LDS.128)
  float SDU[kR * kSP];              // delta*u                                    [row][pos]
  float TDU[kC * kTP];              // delta*u, transposed for the contraction    [pos][row], pitch 36
  float TDY[kC * kTP];              // dout                                       [pos][row]
  float DU[kW][2][8 * kT];          // per-warp output tiles, double-buffered (256 B each)
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

__global__ void __launch_bounds__(kThr, 4)
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
        mbar_expect_tx(full, (uint32_t)(3 * kR * kT * 4 + kR * kStatePad * 4));
        tma_load_3d(smem_u32(sm.U[s]), &map_u, l0, d0, b, full);
        tma_load_3d(smem_u32(sm.DT[s]), &map_dt, l0, d0, b, full);
        tma_load_3d(smem_u32(sm.DY[s]), &map_dy, l0, d0, b, full);
        // saved state t-1 = state before the chunk's first position; state "-1" is out of bounds -> zeros
        tma_load_2d(smem_u32(sm.CK[s]), &map_ck, (t - 1) * kStatePad, row0, full);
      }
      float* bc = sm.BC[s];
      float v[8];
      if (lanes_along_l) {   // lane = (group of 8 values, position): 32 contiguous bytes per 8 lanes
        const int pos = lane & 7, grp = lane >> 3;       // values 8*grp .. 8*grp+7 of [B0..15 C0..15]
        const int l = l0 + pos;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int vi = grp * 8 + q, n = vi & 15;
          const float* src = (vi >= 16) ? (Cg + (int64_t)n * a.C_n_stride) : (Bg + (int64_t)n * a.B_n_stride);
          v[q] = (n < N && l < L) ? __ldg(src + l) : 0.f;
        }
        const uint32_t dst = smem_u32(bc) + (uint32_t)(pos * kPitch + grp * 8) * 4;
        sts_f4(dst, make_float4(v[0], v[1], v[2], v[3]));
        sts_f4(dst + 16, make_float4(v[4], v[5], v[6], v[7]));
      } else {               // lane = (B|C, state): the 16 B and 16 C values of one position are contiguous
        const int n = lane & 15;
        const float* src = (lane < 16) ? (Bg + (int64_t)n * a.B_n_stride) : (Cg + (int64_t)n * a.C_n_stride);
        const int64_t ls = (lane < 16) ? a.B_l_stride : a.C_l_stride;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (n < N && l0 + j < L) ? __ldg(src + (int64_t)(l0 + j) * ls) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) bc[j * kPitch + lane] = v[j];
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
  // shared-space addresses of this thread's slots
  const uint32_t x_row = smem_u32(sm.X) + (uint32_t)r * (kC * kStatePad * 4);
  const uint32_t dx_row = smem_u32(sm.DX) + (uint32_t)r * (kC * kStatePad * 4);
  const uint32_t xsw = (uint32_t)swz_row(r);
  const uint32_t sd_row = smem_u32(sm.SD) + (uint32_t)r * (kSP * 4);
  const uint32_t sdu_row = smem_u32(sm.SDU) + (uint32_t)r * (kSP * 4);
  // contraction role of this thread: (tensor, position, 4 states, half of the rows)
  const int c_which = tid >> 6;                       // 0: dB from DX and delta*u, 1: dC from X and dy
  const int c_j = ((tid >> 5) & 1) * 4 + ((lane >> 2) & 3);
  const int c_nq = lane & 3;
  const int c_half = lane >> 4;                       // rows 16*c_half .. 16*c_half+15
  const uint32_t c_src = smem_u32(c_which ? sm.X : sm.DX);
  const uint32_t c_scal = smem_u32(c_which ? sm.TDY : sm.TDU) + (uint32_t)(c_j * kTP + c_half * 16) * 4;
  float* __restrict__ dBC = (c_which ? a.dC : a.dB) + ((int64_t)b * a.ngroups + g) * N * (int64_t)L + (int64_t)(c_nq * 4) * L + c_j;

  for (int i = 0; i < n_tiles; ++i) {
    const int t = n_tiles - 1 - i;
    const int s = i % kStg, k = i / kStg;
    const int l0 = t * kT;
    mbar_wait(smem_u32(&sm.full[s]), k & 1);
    const uint32_t u_row = smem_u32(sm.U[s]) + r * (kT * 4);
    const uint32_t dt_row = smem_u32(sm.DT[s]) + r * (kT * 4);
    const uint32_t dy_row = smem_u32(sm.DY[s]) + r * (kT * 4);
    const uint32_t ck_row = smem_u32(sm.CK[s]) + r * (kStatePad * 4);
    const uint32_t du_tile = smem_u32(sm.DU[warp][i & 1]);
    const uint32_t ddt_tile = smem_u32(sm.DDT[warp][i & 1]);

    {
      const int c0 = l0;                    // first position of the chunk
      // ---------------- prep: my two positions (2sq, 2sq+1) of the chunk ----------------
      const int pj = 2 * sq;
      float my_dl[2], my_sg[2], my_u[2], my_dy[2];
      {
        const uint32_t off = (uint32_t)pj * 4;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(my_u[0]), "=f"(my_u[1]) : "r"(u_row + off));
        float din[2];
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(din[0]), "=f"(din[1]) : "r"(dt_row + off));
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(my_dy[0]), "=f"(my_dy[1]) : "r"(dy_row + off));
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const float xb = din[e] + bias;
          float v = xb, sgm = 1.f;
          if (softplus) {
            float wexp;
            v = softplus_fast(xb, wexp);
            sgm = sigmoid_from_w(xb, wexp);   // softplus' (bwd_kernel.cuh:446-450; == 1 to rounding for x > 20)
          }
          my_dl[e] = ((c0 + pj + e) < L) ? v : 0.f;   // past the end: a = 1, b = 0 (u and dout are TMA zero fill there)
          my_sg[e] = sgm;
        }
        const float du0 = my_dl[0] * my_u[0], du1 = my_dl[1] * my_u[1];
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sd_row + pj * 4), "f"(my_dl[0]), "f"(my_dl[1]) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(sdu_row + pj * 4), "f"(du0), "f"(du1) : "memory");
        // transposed copies for the contraction: [position][row]
        const uint32_t tdu = smem_u32(sm.TDU) + (uint32_t)(pj * kTP + r) * 4, tdy = smem_u32(sm.TDY) + (uint32_t)(pj * kTP + r) * 4;
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(tdu), "f"(du0) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(tdu + kTP * 4), "f"(du1) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(tdy), "f"(my_dy[0]) : "memory");
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(tdy + kTP * 4), "f"(my_dy[1]) : "memory");
      }
      __syncwarp();
      // ---------------- forward recompute from the saved state (states kept, decays recomputed later) ----------------
      u64 x0[2];
      lds_2x64(ck_row + (uint32_t)sq * 16, x0[0], x0[1]);
      const uint32_t bc_base = smem_u32(sm.BC[s]) + sq * (kLS * 4);
      {
        const float4 dA_ = lds_f4(sd_row), dB_ = lds_f4(sd_row + 16);
        const float4 uA_ = lds_f4(sdu_row), uB_ = lds_f4(sdu_row + 16);
        const float dl[kC] = {dA_.x, dA_.y, dA_.z, dA_.w, dB_.x, dB_.y, dB_.z, dB_.w};
        const float du_[kC] = {uA_.x, uA_.y, uA_.z, uA_.w, uB_.x, uB_.y, uB_.z, uB_.w};
        u64 xr0 = x0[0], xr1 = x0[1];   // running state; every x_j goes to the X tile (also read back by the reverse pass)
#pragma unroll
        for (int j = 0; j < kC; ++j) {
          u64 Bp[2];
          lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bp[0], Bp[1]);
          const u64 dd = pk2(dl[j], dl[j]);
          const u64 duu = pk2(du_[j], du_[j]);
          float t0, t1, t2, t3;
          upk2(mul2(dd, A2p[0]), t0, t1);
          upk2(mul2(dd, A2p[1]), t2, t3);
          const u64 e0 = pk2(ex2(t0), ex2(t1)), e1 = pk2(ex2(t2), ex2(t3));
          xr0 = fma2(e0, xr0, mul2(duu, Bp[0]));
          xr1 = fma2(e1, xr1, mul2(duu, Bp[1]));
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(x_row + ((((uint32_t)(j * 4 + sq)) ^ xsw) << 4)), "l"(xr0), "l"(xr1) : "memory");
        }
      }
      // ---------------- reverse recurrence ----------------
      float s1p[kC], s2p[kC];
      {
        const float4 dA_ = lds_f4(sd_row), dB_ = lds_f4(sd_row + 16);
        const float4 yA_ = lds_f4(dy_row), yB_ = lds_f4(dy_row + 16);
        const float dl[kC] = {dA_.x, dA_.y, dA_.z, dA_.w, dB_.x, dB_.y, dB_.z, dB_.w};
        const float dy[kC] = {yA_.x, yA_.y, yA_.z, yA_.w, yB_.x, yB_.y, yB_.z, yB_.w};
#pragma unroll
        for (int j = kC - 1; j >= 0; --j) {
          u64 Bp[2], Cp[2];
          lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bp[0], Bp[1]);
          lds_2x64(bc_base + (uint32_t)j * (kPitch * 4) + 64, Cp[0], Cp[1]);
          const u64 dyy = pk2(dy[j], dy[j]);
          const u64 dd = pk2(dl[j], dl[j]);
          float t0, t1, t2, t3;                                          // decays again: MUFU has slack, registers do not
          upk2(mul2(dd, A2p[0]), t0, t1);
          upk2(mul2(dd, A2p[1]), t2, t3);
          const u64 e0 = pk2(ex2(t0), ex2(t1)), e1 = pk2(ex2(t2), ex2(t3));
          const u64 dx0 = fma2(Cp[0], dyy, w2[0]);                      // dx_{l,n}
          const u64 dx1 = fma2(Cp[1], dyy, w2[1]);
          asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(dx_row + ((((uint32_t)(j * 4 + sq)) ^ xsw) << 4)), "l"(dx0), "l"(dx1) : "memory");
          s1p[j] = hsum2(fma2(dx1, Bp[1], mul2(dx0, Bp[0])));           // sum_n dx * B          (bwd_kernel.cuh:280-281)
          u64 xp0 = x0[0], xp1 = x0[1];                                   // x_{l-1}: my own slot of the X tile
          if (j > 0) lds_2x64(x_row + ((((uint32_t)((j - 1) * 4 + sq)) ^ xsw) << 4), xp0, xp1);
          const u64 wg0 = mul2(dx0, mul2(e0, xp0));                       // dx * a_l * x_{l-1}  (:283, x - b form)
          const u64 wg1 = mul2(dx1, mul2(e1, xp1));
          s2p[j] = hsum2(fma2(wg1, A2p[1], mul2(wg0, A2p[0])));          // in units of log2(e)
          dA2[0] = fma2(wg0, dd, dA2[0]);                                 // :286
          dA2[1] = fma2(wg1, dd, dA2[1]);
          w2[0] = mul2(e0, dx0);                                          // carried to position l-1
          w2[1] = mul2(e1, dx1);
        }
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
        const uint32_t off = (uint32_t)(rw * kT + pj) * 4;
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(du_tile + off), "f"(o_du[0]), "f"(o_du[1]) : "memory");
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(ddt_tile + off), "f"(o_dd[0]), "f"(o_dd[1]) : "memory");
      }
      named_bar_sync(1, kW * 32);   // X, DX, TDU, TDY complete for the CTA's 32 channels
      // ---------------- contraction over the channels ----------------
      {
        u64 acc0 = pk2(0.f, 0.f), acc1 = pk2(0.f, 0.f);
#pragma unroll
        for (int r4 = 0; r4 < 4; ++r4) {
          const float4 sc4 = lds_f4(c_scal + r4 * 16);
          const float sc[4] = {sc4.x, sc4.y, sc4.z, sc4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r2 = c_half * 16 + r4 * 4 + e;
            u64 v0, v1;
            lds_2x64(c_src + (uint32_t)r2 * (kC * kStatePad * 4) + ((((uint32_t)(c_j * 4 + c_nq)) ^ (uint32_t)swz_row(r2)) << 4), v0, v1);
            const u64 ss = pk2(sc[e], sc[e]);
            acc0 = fma2(ss, v0, acc0);
            acc1 = fma2(ss, v1, acc1);
          }
        }
        float o[4];
        upk2(acc0, o[0], o[1]);
        upk2(acc1, o[2], o[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
        if (c_half == 0 && c0 + c_j < L) {
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (c_nq * 4 + e < N) atomicAdd(dBC + (int64_t)e * L + c0, o[e]);
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
  const cuuint32_t box[2] = {(cuuint32_t)kStatePad, (cuuint32_t)kR};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool bwd_tma_eligible(const BwdLaunch& p) {
  const selscan_bwd_args& a = p.a;
  if (a.z != nullptr) return false;
  if (p.dim_per_group % kR != 0) return false;
  if (p.n_ckpt < 1) return false;
  const int64_t zero = 0;
  if (!tma_row_ok(a.u, a.u_d_stride, a.batch > 1 ? a.u_batch_stride : zero)) return false;
  if (!tma_row_ok(a.delta, a.delta_d_stride, a.batch > 1 ? a.delta_batch_stride : zero)) return false;
  if (!tma_row_ok(a.dout, a.dout_d_stride, a.batch > 1 ? a.dout_batch_stride : zero)) return false;
  if (!tma_row_ok(a.du, a.du_d_stride, a.batch > 1 ? a.du_batch_stride : zero)) return false;
  if (!tma_row_ok(a.ddelta, a.ddelta_d_stride, a.batch > 1 ? a.ddelta_batch_stride : zero)) return false;
  if ((reinterpret_cast<uintptr_t>(a.ckpt) & 15u) != 0) return false;
  return tensor_map_encoder() != nullptr;
}

cudaError_t launch_bwd_tma(const BwdLaunch& p, cudaStream_t stream) {
  const selscan_bwd_args& a = p.a;
  CUtensorMap mu, mdt, mdy, mck, mdu, mddt;
  if (!make_row_map(&mu, a.u, a.seqlen, a.dim, a.batch, a.u_d_stride, a.u_batch_stride, kT, kR) ||
      !make_row_map(&mdt, a.delta, a.seqlen, a.dim, a.batch, a.delta_d_stride, a.delta_batch_stride, kT, kR) ||
      !make_row_map(&mdy, a.dout, a.seqlen, a.dim, a.batch, a.dout_d_stride, a.dout_batch_stride, kT, kR) ||
      !make_row_map(&mdu, a.du, a.seqlen, a.dim, a.batch, a.du_d_stride, a.du_batch_stride, kT, 8) ||
      !make_row_map(&mddt, a.ddelta, a.seqlen, a.dim, a.batch, a.ddelta_d_stride, a.ddelta_batch_stride, kT, 8) ||
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
