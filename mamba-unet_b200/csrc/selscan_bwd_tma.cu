// Backward selective scan, tiled path for sm_100a: TMA-staged tiles walked from the end of the sequence,
// mbarrier pipeline, packed f32x2 arithmetic.  Replaces selective_scan_bwd_kernel
// (/root/reference/mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489) for the aligned shapes Mamba-UNet
// produces (channels per group a multiple of 64, 16-byte aligned rows, no z); everything else takes selscan_bwd.cu.
//
// What bounds this kernel on B200 is the shared-memory -> register return path (128 B/clk/SM): a 128-bit shared load
// costs 4 wavefronts even when the whole warp reads the same 64 bytes (measured, profiles/r01_ncu_bwd_*), so the design
// minimises BYTES PER THREAD through that pipe:
//   * a thread owns TWO channels x 4 states: every B/C value it loads serves both channels, and the channel-pair
//     products for dB/dC are summed in registers before they touch shared memory (half the contraction traffic);
//   * the chunk's states stay in registers (packed pairs); decays are recomputed in the reverse pass (MUFU has slack).
//
// CTA = 64 channels of one (batch, group); the sequence is walked backwards in chunks of 8 positions (the checkpoint
// interval).
//   staging: one elected thread issues, one chunk ahead, TMA loads of u, delta, dout (box 64 rows x 8 positions) and of the
//       saved scan state the chunk restarts from (box 64 rows x 16 states; the state "before position 0" is the tensor
//       map's out-of-bounds zero fill) into a 2-stage ring; all threads prefetch the next chunk's B/C (any strides) into a
//       [position][32] tile.  There is no dedicated producer warp: 4-warp CTAs put at most 2 warps on an SM
//       sub-partition, which is what lets a thread hold ~200 registers at 2 CTAs/SM.
//   4 warps, 8 channel pairs each, four lanes per pair.  Per chunk:
//       prep     lane j of a pair discretises positions j and 4+j of both channels (softplus, sigmoid) and publishes
//                delta and delta*u; nothing is computed twice;
//       forward  restart from the saved state: per position, channel and state pair FMUL2, 2 MUFU.EX2, FMUL2, FFMA2;
//       reverse  dx = C*dy + a*dx' in registers; partial sums of du / ddelta over the lane's 4 states, dA in registers;
//                pair products (delta*u)*dx and dy*x go to the swizzled P tile;
//       reduce   per half chunk the partial sums are reduce-scattered over the 4 lanes (lane j finalises position j of
//                the half for both channels: du, ddelta through softplus', dD, ddelta_bias) into du/ddelta tiles that
//                leave by per-warp TMA stores.
//   contraction: after a named barrier over the 4 compute warps the CTA sums P over its 32 channel pairs and adds ONE
//       value per (state, position) to global memory -- the reference issues one atomic per (channel, state, position)
//       (bwd_kernel.cuh:298-316).
#include <type_traits>

#include "selscan_common.cuh"
#include "selscan_kernels.h"
#include "selscan_ptx.cuh"
#include "selscan_tma_host.h"

namespace selscan {

namespace {

constexpr int kR = 64;            // channels per CTA
constexpr int kNP = kR / 2;       // channel pairs per CTA
constexpr int kW = 4;             // compute warps (8 pairs = 16 channels each)
constexpr int kC = kCkptInterval; // positions per chunk = per staged tile (8)
constexpr int kStg = 2;
constexpr int kPitch = 36;        // B/C tile pitch (floats)
constexpr int kThr = kW * 32;     // no dedicated producer warp: 4-warp CTAs keep <= 2 warps per SM sub-partition (255 regs)
constexpr int kLS = kStatePad / 4;  // states per lane
constexpr int kSP = 12;           // row pitch of the per-row scalar scratch (floats): conflict-free LDS.128

struct BwdTmaSmem {
  float CK[kStg][kR * kStatePad];   // 4 KB: [row][16 states]
  float U[kStg][kR * kC];           // 2 KB: [half chunk][row][4 positions] (two 16-byte-wide TMA boxes): lane (pair, quad)
  float DT[kStg][kR * kC];          //       reads bank 4*pair + quad -> no conflicts on the scalar accesses
  float DY[kStg][kR * kC];
  float BC[kStg][kC * kPitch];
  float P[kNP * kC * 32];           // 32 KB, swizzled [pair][pos][dB 0..15 | dC 0..15]
  float SD[kR * kSP];               // delta   per (row, position of the chunk)
  float SDU[kR * kSP];              // delta*u
  float DU[2][kR * kC];             // CTA-wide output tiles [half][row][4], double-buffered (2 KB each)
  float DDT[2][kR * kC];
  u64 full[kStg];
};

__device__ __forceinline__ uint32_t swz8(int p) { return (uint32_t)(((p & 1) << 2) | ((p >> 1) & 3)); }

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_2x64(uint32_t addr, u64 a, u64 b) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}

__global__ void __launch_bounds__(kThr, 2)
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
  const int n_tiles = (L + kC - 1) / kC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStg; ++s) {
      mbar_init(smem_u32(&sm.full[s]), 1);   // the elected thread's arrive.expect_tx; TMA completes the bytes
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

  // ---- staging (no dedicated producer warp): thread 0 issues the TMA loads of a chunk, all 128 threads gather its B/C ----
  const float* __restrict__ Bg = a.B + (int64_t)b * a.B_batch_stride + (int64_t)g * a.B_group_stride;
  const float* __restrict__ Cg = a.C + (int64_t)b * a.C_batch_stride + (int64_t)g * a.C_group_stride;
  const bool lanes_along_l = (a.B_l_stride == 1 && a.C_l_stride == 1);
  const int row0 = b * a.dim + d0;
  auto issue_tma = [&](int i) {           // chunk index in processing order -> stage i % kStg
    const int t = n_tiles - 1 - i, s = i % kStg, l0 = t * kC;
    const uint32_t full = smem_u32(&sm.full[s]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((uint32_t)(3 * kR * kC * 4 + kR * kStatePad * 4)) : "memory");
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      tma_load_3d(smem_u32(sm.U[s]) + hf * (kR * 16), &map_u, l0 + 4 * hf, d0, b, full);
      tma_load_3d(smem_u32(sm.DT[s]) + hf * (kR * 16), &map_dt, l0 + 4 * hf, d0, b, full);
      tma_load_3d(smem_u32(sm.DY[s]) + hf * (kR * 16), &map_dy, l0 + 4 * hf, d0, b, full);
    }
    // saved state t-1 = state before the chunk's first position; state "-1" is out of bounds -> zeros
    tma_load_2d(smem_u32(sm.CK[s]), &map_ck, (t - 1) * kStatePad, row0, full);
  };
  // element e (0..255) of a chunk's [8 positions][B0..15 C0..15] tile handled by thread (e & 127), two per thread
  auto bc_load = [&](int i, int e) -> float {
    const int t = n_tiles - 1 - i, l0 = t * kC;
    const int pos = lanes_along_l ? (e & 7) : (e >> 5), val = lanes_along_l ? (e >> 3) : (e & 31);
    const int n = val & 15, l = l0 + pos;
    const float* src = (val >= 16) ? (Cg + (int64_t)n * a.C_n_stride + (int64_t)l * a.C_l_stride)
                                   : (Bg + (int64_t)n * a.B_n_stride + (int64_t)l * a.B_l_stride);
    return (n < N && l < L) ? __ldg(src) : 0.f;
  };
  auto bc_store = [&](int i, int e, float v) {
    const int pos = lanes_along_l ? (e & 7) : (e >> 5), val = lanes_along_l ? (e >> 3) : (e & 31);
    sm.BC[i % kStg][pos * kPitch + val] = v;
  };
  {
    if (threadIdx.x == 0) issue_tma(0);
    const float v0 = bc_load(0, threadIdx.x), v1 = bc_load(0, threadIdx.x + kThr);
    bc_store(0, threadIdx.x, v0);
    bc_store(0, threadIdx.x + kThr, v1);
  }
  __syncthreads();

  // ================================ compute warps ================================
  const int sq = lane & 3;                // which 4 states
  const int pr = lane >> 2;               // channel pair inside the warp
  const int pp = warp * 8 + pr;           // channel pair inside the CTA
  const int rA = pp;                      // rows (channels inside the CTA) rA and rA + 32
  const bool hi1 = (sq & 2) != 0, hi0 = (sq & 1) != 0;
  u64 A2p[2][2], dA2[2][2], w2[2][2];
  float Dv[2], bias[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int d = d0 + rA + c * kNP;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n0 = sq * kLS + 2 * q;
      const float a0 = (n0 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)n0 * a.A_n_stride) * kLog2e : 0.f;
      const float a1 = (n0 + 1 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)(n0 + 1) * a.A_n_stride) * kLog2e : 0.f;
      A2p[c][q] = pk2(a0, a1);
      dA2[c][q] = pk2(0.f, 0.f);
      w2[c][q] = pk2(0.f, 0.f);   // a_{l+1} * dx_{l+1}: zero beyond the last position
    }
    Dv[c] = a.D ? __ldg(a.D + d) : 0.f;
    bias[c] = a.delta_bias ? __ldg(a.delta_bias + d) : 0.f;
  }
  const bool softplus = a.delta_softplus != 0;
  float dD_acc[2] = {0.f, 0.f}, dbias_acc[2] = {0.f, 0.f};
  // shared-space addresses of this thread's slots
  const uint32_t p_row = smem_u32(sm.P) + (uint32_t)pp * (kC * 32 * 4);
  const uint32_t psw = swz8(pr);
  const uint32_t sd_row = smem_u32(sm.SD) + (uint32_t)rA * (kSP * 4);       // channel c: + c * kNP * kSP * 4
  const uint32_t sdu_row = smem_u32(sm.SDU) + (uint32_t)rA * (kSP * 4);
  constexpr uint32_t kCS = kNP * kSP * 4;   // byte offset of the pair's second channel in SD / SDU
  constexpr uint32_t kCT = kNP * 16;        // ... in a [half][row][4] tile
  constexpr uint32_t kHT = kR * 16;         // byte offset of the second half chunk in a tile
  // contraction role: 16-byte chunk q of the [8 positions][32 values] output, half h of the channel pairs
  const int c_q = warp * 16 + (lane & 15);
  const int c_j = c_q >> 3, c_c = c_q & 7;
  const int c_h = lane >> 4;
  float* __restrict__ dBC = ((c_c >= 4) ? a.dC : a.dB) + ((int64_t)b * a.ngroups + g) * N * (int64_t)L + (int64_t)((c_c & 3) * 4) * L + c_j;

  for (int i = 0; i < n_tiles; ++i) {
    const int t = n_tiles - 1 - i;
    const int s = i % kStg, k = i / kStg;
    const int c0 = t * kC;                  // first position of the chunk
    // prefetch the next chunk: its stage was released by the barrier that ended chunk i-1
    const bool has_next = (i + 1 < n_tiles);
    float nb0 = 0.f, nb1 = 0.f;
    if (has_next) {
      if (threadIdx.x == 0) issue_tma(i + 1);
      nb0 = bc_load(i + 1, threadIdx.x);
      nb1 = bc_load(i + 1, threadIdx.x + kThr);
    }
    mbar_wait(smem_u32(&sm.full[s]), k & 1);
    const uint32_t u_row = smem_u32(sm.U[s]) + rA * 16;      // + hf * kHT + c * kCT + position-in-half * 4
    const uint32_t dt_row = smem_u32(sm.DT[s]) + rA * 16;
    const uint32_t dy_row = smem_u32(sm.DY[s]) + rA * 16;
    const uint32_t ck_row = smem_u32(sm.CK[s]) + rA * (kStatePad * 4) + sq * 16;
    const uint32_t bc_base = smem_u32(sm.BC[s]) + sq * (kLS * 4);
    const uint32_t du_tile = smem_u32(sm.DU[i & 1]);
    const uint32_t ddt_tile = smem_u32(sm.DDT[i & 1]);

    // ---------------- prep: positions sq and 4+sq of both channels ----------------
    float my_sg[2][2];   // [channel][half]; u, dy and delta of my elements are re-read from shared memory when finalising
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const uint32_t off = (uint32_t)hf * kHT + (uint32_t)c * kCT + (uint32_t)sq * 4;
        const float my_u = lds_f1(u_row + off);
        const float xb = lds_f1(dt_row + off) + bias[c];
        float v = xb, sgm = 1.f;
        if (softplus) {
          float wexp;
          v = softplus_fast(xb, wexp);
          sgm = sigmoid_from_w(xb, wexp);   // softplus' (bwd_kernel.cuh:446-450; == 1 to rounding for x > 20)
        }
        v = ((c0 + hf * 4 + sq) < L) ? v : 0.f;   // past the end: a = 1, b = 0 (u and dout are TMA zero fill there)
        my_sg[c][hf] = sgm;
        const uint32_t so = (uint32_t)c * kCS + (uint32_t)(hf * 4 + sq) * 4;
        sts_f1(sd_row + so, v);
        sts_f1(sdu_row + so, v * my_u);
      }
    }
    __syncwarp();
    float dl[2][kC];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float4 t0 = lds_f4(sd_row + c * kCS), t1 = lds_f4(sd_row + c * kCS + 16);
      dl[c][0] = t0.x; dl[c][1] = t0.y; dl[c][2] = t0.z; dl[c][3] = t0.w;
      dl[c][4] = t1.x; dl[c][5] = t1.y; dl[c][6] = t1.z; dl[c][7] = t1.w;
    }
    // ---------------- forward recompute from the saved state ----------------
    u64 x0[2][2], xs[2][kC][2];
    lds_2x64(ck_row, x0[0][0], x0[0][1]);
    lds_2x64(ck_row + kNP * kStatePad * 4, x0[1][0], x0[1][1]);
    auto fwd_half = [&](auto HF) {
      constexpr int hf = decltype(HF)::value;   // compile-time half: keeps xs[][][] in registers
      float duh[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 v4 = lds_f4(sdu_row + c * kCS + hf * 16);
        duh[c][0] = v4.x; duh[c][1] = v4.y; duh[c][2] = v4.z; duh[c][3] = v4.w;
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = hf * 4 + jj;
        u64 Bp[2];
        lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bp[0], Bp[1]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const u64 dd = pk2(dl[c][j], dl[c][j]);
          const u64 duu = pk2(duh[c][jj], duh[c][jj]);
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
      constexpr int hf = decltype(HF)::value;
      float duh[2][4], dyh[2][4], s1p[2][4], s2p[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 v4 = lds_f4(sdu_row + c * kCS + hf * 16);
        const float4 y4 = lds_f4(dy_row + hf * kHT + c * kCT);
        duh[c][0] = v4.x; duh[c][1] = v4.y; duh[c][2] = v4.z; duh[c][3] = v4.w;
        dyh[c][0] = y4.x; dyh[c][1] = y4.y; dyh[c][2] = y4.z; dyh[c][3] = y4.w;
      }
#pragma unroll
      for (int jj = 3; jj >= 0; --jj) {
        const int j = hf * 4 + jj;
        u64 Bp[2], Cp[2];
        lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bp[0], Bp[1]);
        lds_2x64(bc_base + (uint32_t)j * (kPitch * 4) + 64, Cp[0], Cp[1]);
        u64 pB0 = 0, pB1 = 0, pC0 = 0, pC1 = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const u64 dyy = pk2(dyh[c][jj], dyh[c][jj]);
          const u64 dd = pk2(dl[c][j], dl[c][j]);
          const u64 duu = pk2(duh[c][jj], duh[c][jj]);
          float t0, t1, t2, t3;                                          // decays again: MUFU has slack, registers do not
          upk2(mul2(dd, A2p[c][0]), t0, t1);
          upk2(mul2(dd, A2p[c][1]), t2, t3);
          const u64 e0 = pk2(ex2(t0), ex2(t1)), e1 = pk2(ex2(t2), ex2(t3));
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
        const uint32_t prow = p_row + (uint32_t)j * 128;
        sts_2x64(prow + ((((uint32_t)sq) ^ psw) << 4), pB0, pB1);
        sts_2x64(prow + ((((uint32_t)(4 + sq)) ^ psw) << 4), pC0, pC1);
      }
      // ---- reduce-scatter over the 4 lanes: lane sq finalises position hf*4 + sq of both channels ----
      float f1[2], f2[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float a1 = hi1 ? s1p[c][0] : s1p[c][2], b1 = hi1 ? s1p[c][1] : s1p[c][3];
        const float a2 = hi1 ? s2p[c][0] : s2p[c][2], b2 = hi1 ? s2p[c][1] : s2p[c][3];
        const float k1a = (hi1 ? s1p[c][2] : s1p[c][0]) + __shfl_xor_sync(0xffffffffu, a1, 2);
        const float k1b = (hi1 ? s1p[c][3] : s1p[c][1]) + __shfl_xor_sync(0xffffffffu, b1, 2);
        const float k2a = (hi1 ? s2p[c][2] : s2p[c][0]) + __shfl_xor_sync(0xffffffffu, a2, 2);
        const float k2b = (hi1 ? s2p[c][3] : s2p[c][1]) + __shfl_xor_sync(0xffffffffu, b2, 2);
        f1[c] = (hi0 ? k1b : k1a) + __shfl_xor_sync(0xffffffffu, hi0 ? k1a : k1b, 1);
        f2[c] = (hi0 ? k2b : k2a) + __shfl_xor_sync(0xffffffffu, hi0 ? k2a : k2b, 1);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const uint32_t eo = (uint32_t)hf * kHT + (uint32_t)c * kCT + (uint32_t)sq * 4;
        const float e_u = lds_f1(u_row + eo), e_dy = lds_f1(dy_row + eo);
        const float e_dl = lds_f1(sd_row + (uint32_t)c * kCS + (uint32_t)(hf * 4 + sq) * 4);
        const float o_du = fmaf(e_dl, f1[c], Dv[c] * e_dy);                                 // :211, :280
        const float o_dd = fmaf(e_u, f1[c], f2[c] * kLn2) * my_sg[c][hf];                   // :281-284, :446-450
        dbias_acc[c] += ((c0 + hf * 4 + sq) < L) ? o_dd : 0.f;
        dD_acc[c] = fmaf(e_dy, e_u, dD_acc[c]);                                             // :213
        sts_f1(du_tile + eo + rA * 16, o_du);     // same [half][row][4] layout as the input tiles
        sts_f1(ddt_tile + eo + rA * 16, o_dd);
      }
    };
    rev_half(std::integral_constant<int, 1>{});
    rev_half(std::integral_constant<int, 0>{});
    fence_proxy_async_smem();     // my du / ddelta writes -> visible to the TMA store
    named_bar_sync(1, kW * 32);   // P and the output tiles complete for the CTA's 64 channels
    if (threadIdx.x == 0) {
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        tma_store_3d(&map_du, du_tile + hf * kHT, c0 + 4 * hf, d0, b);
        tma_store_3d(&map_ddt, ddt_tile + hf * kHT, c0 + 4 * hf, d0, b);
      }
      tma_store_commit();
      tma_store_wait_read<1>();   // the other output buffer (chunk i-1) has been read: free for chunk i+1
    }
    if (has_next) {               // the other stage's B/C tile: last read in chunk i-1, next read after the barrier below
      bc_store(i + 1, threadIdx.x, nb0);
      bc_store(i + 1, threadIdx.x + kThr, nb1);
    }
    // ---------------- contraction over the channel pairs ----------------
    {
      u64 acc0 = pk2(0.f, 0.f), acc1 = pk2(0.f, 0.f);
      const uint32_t src = smem_u32(sm.P) + (uint32_t)c_j * 128;
#pragma unroll
      for (int e = 0; e < kNP / 2; ++e) {
        const int p2 = c_h * (kNP / 2) + e;
        u64 v0, v1;
        lds_2x64(src + (uint32_t)p2 * (kC * 32 * 4) + ((((uint32_t)c_c) ^ swz8(p2 & 7)) << 4), v0, v1);
        acc0 = add2(acc0, v0);
        acc1 = add2(acc1, v1);
      }
      float o[4];
      upk2(acc0, o[0], o[1]);
      upk2(acc1, o[2], o[3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] += __shfl_xor_sync(0xffffffffu, o[e], 16);
      if (c_h == 0 && c0 + c_j < L) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if ((c_c & 3) * 4 + e < N) atomicAdd(dBC + (int64_t)e * L + c0, o[e]);
      }
    }
    named_bar_sync(1, kW * 32);   // P free for the next chunk
  }
  if (threadIdx.x == 0) tma_store_wait_all<0>();

#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int d = d0 + rA + c * kNP;
    float da[4];
    upk2(dA2[c][0], da[0], da[1]);
    upk2(dA2[c][1], da[2], da[3]);
#pragma unroll
    for (int n = 0; n < kLS; ++n)
      if (sq * kLS + n < N) atomicAdd(a.dA + (int64_t)d * N + sq * kLS + n, da[n]);   // sum over batch
    float dDs = dD_acc[c], dbs = dbias_acc[c];
    dDs += __shfl_xor_sync(0xffffffffu, dDs, 1);
    dDs += __shfl_xor_sync(0xffffffffu, dDs, 2);
    dbs += __shfl_xor_sync(0xffffffffu, dbs, 1);
    dbs += __shfl_xor_sync(0xffffffffu, dbs, 2);
    if (sq == 0) {
      if (a.dD != nullptr) atomicAdd(a.dD + d, dDs);
      if (a.ddelta_bias != nullptr) atomicAdd(a.ddelta_bias + d, dbs);
    }
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
  if (a.z != nullptr || a.dstate > kStatePad) return false;
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
  if (bwd_ws_usable() && bwd_ws_eligible(p)) return launch_bwd_ws(p, stream);
  const selscan_bwd_args& a = p.a;
  CUtensorMap mu, mdt, mdy, mck, mdu, mddt;
  if (!make_row_map(&mu, a.u, a.seqlen, a.dim, a.batch, a.u_d_stride, a.u_batch_stride, 4, kR) ||
      !make_row_map(&mdt, a.delta, a.seqlen, a.dim, a.batch, a.delta_d_stride, a.delta_batch_stride, 4, kR) ||
      !make_row_map(&mdy, a.dout, a.seqlen, a.dim, a.batch, a.dout_d_stride, a.dout_batch_stride, 4, kR) ||
      !make_row_map(&mdu, a.du, a.seqlen, a.dim, a.batch, a.du_d_stride, a.du_batch_stride, 4, kR) ||
      !make_row_map(&mddt, a.ddelta, a.seqlen, a.dim, a.batch, a.ddelta_d_stride, a.ddelta_batch_stride, 4, kR) ||
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
