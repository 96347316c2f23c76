// Forward selective scan, warp-specialised variant for sm_100a -- OPT-IN (SELSCAN_B200_FWD=ws), measured and not the default.
// Same contract and arithmetic as selscan_fwd_tma.cu (replaces /root/reference/mamba/csrc/selective_scan/
// selective_scan_fwd_kernel.cuh:67-303 for channels per group % 64 == 0, 16-byte aligned rows, no z; calls that split the sequence
// into segments stay on selscan_fwd_tma.cu).
//
// The idea (it is what made the backward faster, selscan_bwd_ws.cu): the scan is bound by the shared-memory -> register return
// path, so give a thread TWO channels (every B / C value it loads serves both: 65 M instead of 85 M wavefronts at stage 1, batch 24)
// and make up for the halved number of recurrence warps with helper warps:
//   warps 0-3  RECURRENCE: thread = 2 adjacent channels x 4 states.  Per position and channel: FMUL2, 2 x MUFU.EX2, FMUL2,
//              2 x FFMA2 for the state pair update, FMUL2 + FFMA2 + FADD for its share of y; per quad of positions the 4 partial
//              y of a channel are reduce-scattered over its 4 lanes (3 shuffles) and D*u is added.  Inputs arrive discretised.
//   warps 4-7  HELPER: one lane issues the TMA loads (u, delta: 64 rows x 16 positions, 64-byte swizzle) into a 5-stage ring;
//              every thread discretises 8 elements per tile one tile ahead (softplus once per element), overwrites delta with
//              softplus(delta + bias) and u with D*u, publishes delta*u; gathers B / C (any strides); one lane hands finished
//              output tiles to TMA stores.
// Result on B200 (stage 1, batch 24): 0.44 ms against 0.37 ms for the one-channel kernel.  With the B/C bytes halved the shared-memory
// pipe drops from 84 % to 59 % busy, but the forward then sits on two limits at once -- MUFU (16 ex2 per element: 8 pipe cycles per
// warp instruction) and issue slots (7.5 instructions per MUFU) both need ~2300 cycles per 2048 elements and SM sub-partition -- and
// 2 recurrence + 2 helper warps per sub-partition reach 55 % of that, where 4-5 one-channel warps reach 62 %.  The backward differs:
// its recurrence code was register-capped at 2 warps per sub-partition either way.  Kept selectable and under test.
#include "selscan_common.cuh"
#include "selscan_kernels.h"
#include "selscan_ptx.cuh"
#include "selscan_tma_host.h"

namespace selscan {

namespace {

constexpr int kR = 64;            // channels per CTA
constexpr int kT = 16;            // positions per tile (64-byte rows)
constexpr int kQ = kT / 4;        // quads per tile
constexpr int kStg = 5;            // TMA loads run kStg - 1 tiles ahead of the discretisation, which runs one tile ahead of the recurrence
constexpr int kPitch = 36;        // B/C tile pitch (floats)
constexpr int kGroupThr = 128;
constexpr int kThr = 2 * kGroupThr;
constexpr int kLS = kStatePad / 4;

struct FwsSmem {
  float U[kStg][kR * kT];      // TMA: u      -> helper overwrites with D*u            [row][16], 64-byte swizzle
  float DT[kStg][kR * kT];     // TMA: delta  -> helper overwrites with softplus(delta + bias) (0 past the end)
  float SDU[kStg][kR * kT];    // helper: delta*u, same layout
  float OUT[2][kR * kT];       // recurrence warps: y + D*u, same swizzled layout, double-buffered by tile parity -> TMA store
                               // (every swizzled tile starts on a multiple of 512 bytes: the pattern uses absolute address bits)
  float BC[kStg][kT * kPitch]; // helper: [position][B0..15 C0..15]
  u64 tma_full[kStg];
  u64 prep_done[kStg];
  u64 stage_free[kStg];
  u64 out_full[2];
  u64 out_free[2];
};

// byte offset of 16-byte chunk `q` (a quad of positions) of row `row` in a [row][16] tile with the TMA 64-byte swizzle
__device__ __forceinline__ uint32_t swz64(int row, int q) { return (uint32_t)row * (kT * 4) + (uint32_t)((q ^ ((row >> 1) & 3)) << 4); }

__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float lds_f1(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__global__ void __launch_bounds__(kThr, 2)
selscan_fwd_ws_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_dt,
                      const __grid_constant__ CUtensorMap map_out, const FwdLaunch p) {
  extern __shared__ unsigned char smem_raw[];
  FwsSmem& sm = *reinterpret_cast<FwsSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const selscan_fwd_args& a = p.a;
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
#pragma unroll
    for (int s = 0; s < kStg; ++s) {
      mbar_init(smem_u32(&sm.tma_full[s]), 1);
      mbar_init(smem_u32(&sm.prep_done[s]), 4);
      mbar_init(smem_u32(&sm.stage_free[s]), 4);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mbar_init(smem_u32(&sm.out_full[h]), 4);
      mbar_init(smem_u32(&sm.out_free[h]), 1);
    }
    mbar_fence_init();
    tma_prefetch_desc(&map_u);
    tma_prefetch_desc(&map_dt);
    tma_prefetch_desc(&map_out);
  }
  __syncthreads();

  if (warp >= 4) {
    // =========================================== helper warps ===========================================
    const int htid = threadIdx.x - kGroupThr;
    const int row = htid & (kR - 1);          // my channel inside the CTA; my quads of every tile: 2*qh, 2*qh + 1
    const int qh = htid >> 6;
    const int d = d0 + row;
    const float Dv = a.D ? __ldg(a.D + d) : 0.f;
    const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.f;
    const bool softplus = a.delta_softplus != 0;
    const uint32_t my0 = swz64(row, 2 * qh), my1 = swz64(row, 2 * qh + 1);

    // ---- B/C gather: elements htid + 128 r (r = 0..3) of a tile's [16 positions][32 values]; the pointers walk forwards ----
    const bool along_l = (a.B_l_stride == 1 && a.C_l_stride == 1);   // (.., N, L) layout vs the l-major x_dbl layout
    const float* bcp[4];
    int64_t bcs[4];
    uint32_t bci[4];
    int bcpos[4];
    bool bcok[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = htid + r * kGroupThr;
      const int pos = along_l ? (e & 15) : (e >> 5), val = along_l ? (e >> 4) : (e & 31);
      const int n = val & 15;
      const bool isC = val >= 16;
      const float* base = isC ? (a.C + (int64_t)b * a.C_batch_stride + (int64_t)g * a.C_group_stride + (int64_t)n * a.C_n_stride)
                              : (a.B + (int64_t)b * a.B_batch_stride + (int64_t)g * a.B_group_stride + (int64_t)n * a.B_n_stride);
      const int64_t ls = isC ? a.C_l_stride : a.B_l_stride;
      bcp[r] = base + (int64_t)pos * ls;
      bcs[r] = (int64_t)kT * ls;
      bci[r] = (uint32_t)(pos * kPitch + val) * 4;
      bcpos[r] = pos;
      bcok[r] = n < N;
    }
    auto bc_load = [&](int j, float (&v)[4]) {           // B/C values of tile j (predicated past the end of the sequence)
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        v[r] = (bcok[r] && j * kT + bcpos[r] < L) ? __ldg(bcp[r]) : 0.f;
        bcp[r] += bcs[r];
      }
    };
    auto issue_tma = [&](int j) {
      const int s = j % kStg;
      if (j >= kStg) mbar_wait(smem_u32(&sm.stage_free[s]), (uint32_t)((j / kStg - 1) & 1));
      const uint32_t full = smem_u32(&sm.tma_full[s]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((uint32_t)(2 * kR * kT * 4)) : "memory");
      tma_load_3d(smem_u32(sm.U[s]), &map_u, j * kT, d0, b, full);
      tma_load_3d(smem_u32(sm.DT[s]), &map_dt, j * kT, d0, b, full);
    };
    auto prep = [&](int j, const float (&bcv)[4]) {
      const int s = j % kStg;
      mbar_wait(smem_u32(&sm.tma_full[s]), (uint32_t)((j / kStg) & 1));
      const uint32_t ub = smem_u32(sm.U[s]), tb = smem_u32(sm.DT[s]), sb = smem_u32(sm.SDU[s]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t off = h ? my1 : my0;
        const float4 u4 = lds_f4(ub + off), t4 = lds_f4(tb + off);
        const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
        float v[4], vu[4], du[4];
        const int l0 = j * kT + (2 * qh + h) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float dv = tt[e] + bias;
          if (softplus) {
            float w_unused;
            dv = softplus_fast(dv, w_unused);
          }
          dv = (l0 + e < L) ? dv : 0.f;     // past the end: a = 1, b = 0 (u is TMA zero fill there)
          v[e] = dv;
          vu[e] = dv * uu[e];
          du[e] = Dv * uu[e];
        }
        sts_f4(tb + off, make_float4(v[0], v[1], v[2], v[3]));
        sts_f4(sb + off, make_float4(vu[0], vu[1], vu[2], vu[3]));
        sts_f4(ub + off, make_float4(du[0], du[1], du[2], du[3]));
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) sts_f1(smem_u32(sm.BC[s]) + bci[r], bcv[r]);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.prep_done[s]));
    };

    float cb[4], nb[4];
    if (htid == 0) {
#pragma unroll
      for (int j = 0; j < kStg - 1; ++j)
        if (j < n_tiles) issue_tma(j);
    }
    bc_load(0, cb);
    prep(0, cb);
    bc_load(1, cb);                                   // (all zeros when there is no tile 1)
    for (int i = 0; i < n_tiles; ++i) {
      if (htid == 0 && i + kStg - 1 < n_tiles) issue_tma(i + kStg - 1);
      bc_load(i + 2, nb);                             // in flight while tile i + 1 is prepared
      if (i + 1 < n_tiles) prep(i + 1, cb);
      if (htid == kGroupThr - 32) {                   // lane 0 of the last helper warp: output tile i -> TMA store
        mbar_wait_relaxed(smem_u32(&sm.out_full[i & 1]), (uint32_t)((i >> 1) & 1));
        tma_store_3d(&map_out, smem_u32(sm.OUT[i & 1]), i * kT, d0, b);
        tma_store_commit();
        tma_store_wait_read<0>();
        mbar_arrive(smem_u32(&sm.out_free[i & 1]));
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) cb[r] = nb[r];
    }
    if (htid == kGroupThr - 32) tma_store_wait_all<0>();
    return;
  }

  // =========================================== recurrence warps ===========================================
  const int sq = lane & 3;                // which 4 states
  const int pr = lane >> 2;               // channel pair inside the warp
  const int r0 = (warp * 8 + pr) * 2;     // my rows (channels inside the CTA): r0, r0 + 1 -- the same swizzle key
  u64 A2p[2][2], x2[2][2];
  float* __restrict__ ck[2];
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const int d = d0 + r0 + c;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int n0 = sq * kLS + 2 * q;
      const float a0 = (n0 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)n0 * a.A_n_stride) * kLog2e : 0.f;
      const float a1 = (n0 + 1 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)(n0 + 1) * a.A_n_stride) * kLog2e : 0.f;
      A2p[c][q] = pk2(a0, a1);
      x2[c][q] = pk2(0.f, 0.f);
    }
    ck[c] = a.ckpt ? a.ckpt + ((int64_t)b * a.dim + d) * p.n_ckpt * kStatePad + sq * kLS : nullptr;
  }
  const bool hi1 = (sq & 2) != 0, hi0 = (sq & 1) != 0;
  const uint32_t key = (uint32_t)((r0 >> 1) & 3);
  const uint32_t row_off = (uint32_t)r0 * (kT * 4);        // second channel: + kT * 4

  for (int i = 0; i < n_tiles; ++i) {
    const int s = i % kStg;
    mbar_wait(smem_u32(&sm.prep_done[s]), (uint32_t)((i / kStg) & 1));
    if (i >= 2) mbar_wait(smem_u32(&sm.out_free[i & 1]), (uint32_t)(((i >> 1) - 1) & 1));   // the store of tile i-2 has read this buffer
    const uint32_t du_b = smem_u32(sm.U[s]) + row_off, dl_b = smem_u32(sm.DT[s]) + row_off, su_b = smem_u32(sm.SDU[s]) + row_off;
    const uint32_t bc_base = smem_u32(sm.BC[s]) + sq * (kLS * 4);
    const uint32_t out_b = smem_u32(sm.OUT[i & 1]) + row_off + sq * 4;
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const uint32_t qo = (((uint32_t)q) ^ key) << 4;
      float dl[2][4], du[2][4], y[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float4 t4 = lds_f4(dl_b + c * (kT * 4) + qo), v4 = lds_f4(su_b + c * (kT * 4) + qo);
        dl[c][0] = t4.x; dl[c][1] = t4.y; dl[c][2] = t4.z; dl[c][3] = t4.w;
        du[c][0] = v4.x; du[c][1] = v4.y; du[c][2] = v4.z; du[c][3] = v4.w;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t bc = bc_base + (uint32_t)(4 * q + j) * (kPitch * 4);
        u64 Bp[2], Cp[2];
        lds_2x64(bc, Bp[0], Bp[1]);
        lds_2x64(bc + 64, Cp[0], Cp[1]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const u64 dd = pk2(dl[c][j], dl[c][j]);
          const u64 duu = pk2(du[c][j], du[c][j]);
          float t0, t1, t2, t3;
          upk2(mul2(dd, A2p[c][0]), t0, t1);
          upk2(mul2(dd, A2p[c][1]), t2, t3);
          const u64 e0 = pk2(ex2(t0), ex2(t1)), e1 = pk2(ex2(t2), ex2(t3));
          x2[c][0] = fma2(e0, x2[c][0], mul2(duu, Bp[0]));
          x2[c][1] = fma2(e1, x2[c][1], mul2(duu, Bp[1]));
          y[c][j] = hsum2(fma2(Cp[1], x2[c][1], mul2(Cp[0], x2[c][0])));
        }
      }
      // reduce-scatter the 4 partial sums of each channel over its 4 lanes: lane sq ends with position 4q + sq
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const float s0 = hi1 ? y[c][0] : y[c][2], s1 = hi1 ? y[c][1] : y[c][3];
        float k0 = hi1 ? y[c][2] : y[c][0], k1 = hi1 ? y[c][3] : y[c][1];
        k0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        k1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        const float s2 = hi0 ? k0 : k1;
        float kk = hi0 ? k1 : k0;
        kk += __shfl_xor_sync(0xffffffffu, s2, 1);
        const float Du = lds_f1(du_b + c * (kT * 4) + qo + sq * 4);       // D * u of my position
        sts_f1(out_b + c * (kT * 4) + qo, kk + Du);
      }
      if ((q & 1) && ck[0] != nullptr) {   // position 16 i + 4q + 3 closes an interval of 8
        const int done = i * kT + 4 * q + 4;
        if (done < L) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float xs[4];
            upk2(x2[c][0], xs[0], xs[1]);
            upk2(x2[c][1], xs[2], xs[3]);
            *reinterpret_cast<float4*>(ck[c] + (int64_t)(done / kCkptInterval - 1) * kStatePad) = make_float4(xs[0], xs[1], xs[2], xs[3]);
          }
        }
      }
    }
    fence_proxy_async_smem();   // my OUT writes -> visible to the TMA store
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(smem_u32(&sm.out_full[i & 1]));
      mbar_arrive(smem_u32(&sm.stage_free[s]));
    }
  }
  if (a.last_state != nullptr) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float xs[4];
      upk2(x2[c][0], xs[0], xs[1]);
      upk2(x2[c][1], xs[2], xs[3]);
      const int64_t row = (int64_t)b * a.dim + d0 + r0 + c;
#pragma unroll
      for (int n = 0; n < kLS; ++n)
        if (sq * kLS + n < N) a.last_state[row * N + sq * kLS + n] = xs[n];
    }
  }
}

}  // namespace

bool fwd_ws_enabled() {
  static const bool on = [] {
    const char* e = getenv("SELSCAN_B200_FWD");
    return e != nullptr && e[0] == 'w';      // "ws": opt in
  }();
  return on;
}

cudaError_t launch_fwd_ws(const FwdLaunch& p, cudaStream_t stream) {
  const selscan_fwd_args& a = p.a;
  CUtensorMap mu, mdt, mout;
  if (!make_row_map(&mu, a.u, a.seqlen, a.dim, a.batch, a.u_d_stride, a.u_batch_stride, kT, kR) ||
      !make_row_map(&mdt, a.delta, a.seqlen, a.dim, a.batch, a.delta_d_stride, a.delta_batch_stride, kT, kR) ||
      !make_row_map(&mout, a.out, a.seqlen, a.dim, a.batch, a.out_d_stride, a.out_batch_stride, kT, kR))
    return cudaErrorInvalidValue;
  const int smem = (int)sizeof(FwsSmem) + 1024;
  static_assert(sizeof(FwsSmem) + 1024 + 1024 <= 116736, "two CTAs per SM");
  cudaError_t e = cudaFuncSetAttribute(selscan_fwd_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((int64_t)a.batch * a.ngroups * (p.dim_per_group / kR));
  selscan_fwd_ws_kernel<<<grid, kThr, smem, stream>>>(mu, mdt, mout, p);
  return cudaGetLastError();
}

}  // namespace selscan
