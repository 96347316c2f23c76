// Backward selective scan, TMEM-pipelined tiled path for sm_100a.  Replaces selective_scan_bwd_kernel
// (/root/reference/mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:75-489) for the aligned shapes Mamba-UNet produces
// (channels per group a multiple of 64, 16-byte aligned rows, no z); same arithmetic and the same saved-state scheme as
// selscan_bwd_ws.cu, which it supersedes as the default (SELSCAN_B200_BWD=ws / tma select the older kernels).
//
// What changed against the warp-specialised kernel.  There a compute thread owned 2 channels x 4 states and ran BOTH recurrences
// of a chunk of 8 positions: the forward one (restarted from the saved state) into 64 registers of states, then the reverse one,
// for which it had to evaluate every decay a second time (no registers left to keep them).  Four lanes per channel pair meant
// that every per-element scalar (delta, delta*u, dout), every B / C load and every partial sum was handled four times.
// Here the two recurrences run in DIFFERENT warps that hand the chunk's states and decays over through TENSOR MEMORY used as
// per-thread scratch (tcgen05.st / tcgen05.ld, 32 lanes x 32-bit: thread i of a warp owns TMEM lane 32*(warp%4)+i; warps w and
// w+4 address the same lanes).  TMEM traffic does not touch the shared-memory pipe that bound the old kernel, and 256 columns per
// thread per chunk hold what 64 registers could not:
//   warps 0-3   FORWARD (one per TMEM lane quarter): thread = 2 channels x 8 states (two lanes per channel pair).  Per chunk:
//               states from the saved checkpoint, 8 positions of x <- a*x + (delta*u)*B, each position's 16 states and 16 decays
//               stored to TMEM (two tcgen05.st.x16).  One MUFU.EX2 per (position, state) -- half of the old kernel.
//   warps 4-7   REVERSE: same thread <-> (channel pair, state half) map, reads the chunk back position by position (prefetched
//               one position ahead), runs dx <- C*dout + a*dx in registers and leaves the channel-pair products for dB / dC in
//               the P tile and its 8-state partial sums of dx*B and dx*a*x*A in the S12 tile (two lanes per channel instead of
//               four).  No transcendental, no shuffle, no global access.
//   warps 8-15  HELPERS (4 per unit), unchanged role: TMA loads two chunks ahead into a 3-stage ring, softplus / sigmoid once
//               per element one chunk ahead, B/C gather (any strides), contraction of P over the 32 channel pairs (ONE atomic
//               per (state, position) per 64 channels; reference: one per (channel, state, position), bwd_kernel.cuh:298-316),
//               finalisation of du / ddelta / dD / ddelta_bias, per-half TMA stores.
// A CTA carries TWO independent 64-channel units (unit k: forward warps 2k, 2k+1, reverse warps 4+2k, 5+2k, helper warps 8+4k..):
// one CTA per SM then owns all 512 TMEM columns, i.e. two chunk buffers per thread, so that the forward warp fills chunk i+1
// while the reverse warp drains chunk i.  Register budgets by setmaxnreg (launch 128 x 512 threads = the whole file).
// All hand-overs are mbarrier phases (TMEM ones bracketed by tcgen05.fence::before/after_thread_sync); waits are bounded.
#include <atomic>
#include <type_traits>

#include "selscan_common.cuh"
#include "selscan_kernels.h"
#include "selscan_ptx.cuh"
#include "selscan_tma_host.h"

namespace selscan {

// not declared in the product headers: this kernel is not part of the library (experiments/README.md)
bool bwd_tm_usable();
cudaError_t launch_bwd_tm(const BwdLaunch& p, cudaStream_t stream);

namespace {

constexpr int kR = 64;            // channels per unit
constexpr int kNP = kR / 2;       // channel pairs per unit
constexpr int kC = kCkptInterval; // positions per chunk (8)
constexpr int kHP = kC / 2;       // positions per half chunk
constexpr int kStg = 4;         // TMA runs three chunks ahead, the discretisation two: the forward warps never wait for the helpers
constexpr int kPitch = 36;        // B/C tile pitch (floats)
constexpr int kUnits = 2;         // units per CTA
constexpr int kGroupThr = 128;    // helper threads per unit
constexpr int kThr = 640;         // 20 warps: 4 forward, 8 reverse, 8 helper
constexpr int kRecThr = 384;      // the recurrence warps (forward + reverse), which use tensor memory
constexpr int kRevRegs = 120;     // launch 96 x 640; the helpers release 2 x 128 x 24 registers, the reverse warps take 2 x 128 x 24
constexpr int kHelpRegs = 72;
constexpr int kLaunchRegs = 96;   // what ptxas must report for the kernel (checked by the launcher: setmaxnreg.inc would block otherwise)
constexpr int kTmemCols = 512;
constexpr int kRowCols = 32;      // TMEM columns per (thread, position): 16 states | 16 decays

constexpr int kPP = kHP * 32 + 4;    // floats per channel pair in a P half: 4 positions x 32 values + 16 bytes: the 8 pairs of a
                                     // quarter-warp store to 8 distinct bank groups; readers walk 128 contiguous bytes
constexpr int kS12P = 36;            // floats per row in an S12 half: [state quarter][s1 x4 | s2 x4] + 16 bytes (conflict-free by row)

struct alignas(1024) UnitSmem {
  float CK[kStg][kR * kStatePad];   // [row][16 states]                                    (TMA)
  float U[kStg][kR * kC];           // [row][8 positions], 32-byte swizzle                  (TMA)
  float DT[kStg][kR * kC];          //   raw delta, overwritten IN PLACE by delta = softplus(raw + bias) (helper: each thread its own 4 elements)
  float DY[kStg][kR * kC];
  float SDU[kStg][kR * kC];         // delta*u [row][8], same swizzle                       (helper)
  float DU[2][kR * kC];             // output tiles [half][row][4] (one 64 x 4 TMA store per half), double-buffered by chunk parity (helper)
  float DDT[2][kR * kC];
  float P[2][kNP * kPP];            // per half chunk: [pair][position][dB 0..15 | dC 0..15]  (reverse)
  float S12[2][kR * kS12P];         // per half chunk: [row][state quarter][s1 x4 positions | s2 x4] (reverse)
  float BC[kStg][kC * kPitch];      // [position][B0..15 C0..15]                            (helper)
  u64 tma_full[kStg];               // TMA transaction bytes of a stage
  u64 prep_done[kStg];              // helper warps: delta / delta*u / B/C tiles of a stage written
  u64 stage_free[kStg];             // forward + reverse + helper warps: stage no longer read
  u64 half_full[2];                 // reverse warps: P / S12 of a half chunk written
  u64 half_free[2];                 // helper warps: P / S12 of a half chunk consumed
  u64 tm_full[2][2];                // [slab][buffer] forward warp: chunk stored to TMEM
  u64 tm_free[2][2];                // [slab][buffer] reverse warp: chunk read back
};
static_assert(sizeof(UnitSmem) % 1024 == 0, "unit tiles must keep the alignment of the swizzle atoms (32-byte swizzle: 256 B, 64-byte: 512 B)");

struct TmSmem {
  UnitSmem un[kUnits];
  uint32_t tmem_base;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void sts_2x64(uint32_t addr, u64 a, u64 b) {
  asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void sts_f1(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }

// ---- tensor memory as per-thread scratch (32x32b: thread i of the warp <-> its own TMEM lane, 16 consecutive columns) ----
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const u64 (&v)[8]) {
  asm volatile(
      "{\n\t.reg .b32 r<16>;\n\t"
      "mov.b64 {r0, r1}, %1;\n\tmov.b64 {r2, r3}, %2;\n\tmov.b64 {r4, r5}, %3;\n\tmov.b64 {r6, r7}, %4;\n\t"
      "mov.b64 {r8, r9}, %5;\n\tmov.b64 {r10, r11}, %6;\n\tmov.b64 {r12, r13}, %7;\n\tmov.b64 {r14, r15}, %8;\n\t"
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {r0, r1, r2, r3, r4, r5, r6, r7, r8, r9, r10, r11, r12, r13, r14, r15};\n\t}" ::"r"(taddr),
      "l"(v[0]), "l"(v[1]), "l"(v[2]), "l"(v[3]), "l"(v[4]), "l"(v[5]), "l"(v[6]), "l"(v[7])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// asynchronous: the destination registers are valid only after tmem_ld_wait* on the same registers
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                 "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr)
               : "memory");
}
// wait for every tcgen05.ld of this thread; the registers pass through the asm so that no use can be scheduled ahead of the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&a)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]),
                 "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9]),
                 "+r"(a[10]), "+r"(a[11]), "+r"(a[12]), "+r"(a[13]), "+r"(a[14]), "+r"(a[15]), "+r"(b[0]), "+r"(b[1]), "+r"(b[2]), "+r"(b[3]),
                 "+r"(b[4]), "+r"(b[5]), "+r"(b[6]), "+r"(b[7]), "+r"(b[8]), "+r"(b[9]), "+r"(b[10]), "+r"(b[11]), "+r"(b[12]), "+r"(b[13]),
                 "+r"(b[14]), "+r"(b[15])
               :
               : "memory");
}
__device__ __forceinline__ u64 pku(uint32_t lo, uint32_t hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}

struct UnitItem {
  int b, g, d0;
  bool active;
};
__device__ __forceinline__ UnitItem decode_unit(const BwdLaunch& p, int unit_idx, int n_units) {
  UnitItem it;
  it.active = unit_idx < n_units;
  int bid = it.active ? unit_idx : 0;
  const int tiles_per_group = p.dim_per_group / kR;
  const int tile_g = bid % tiles_per_group;
  bid /= tiles_per_group;
  it.g = bid % p.a.ngroups;
  it.b = bid / p.a.ngroups;
  it.d0 = it.g * p.dim_per_group + tile_g * kR;
  return it;
}

__global__ void __launch_bounds__(kThr, 1)
selscan_bwd_tm_kernel(const __grid_constant__ CUtensorMap map_u, const __grid_constant__ CUtensorMap map_dt,
                      const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_ck,
                      const __grid_constant__ CUtensorMap map_du, const __grid_constant__ CUtensorMap map_ddt, const BwdLaunch p,
                      const int n_units) {
  extern __shared__ unsigned char smem_raw[];
  TmSmem& sm = *reinterpret_cast<TmSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);   // swizzle atoms: 256 B (32-byte), 512 B (64-byte)
  const selscan_bwd_args& a = p.a;
  const int L = a.seqlen, N = a.dstate;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (L + kC - 1) / kC;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < kUnits; ++k) {
      UnitSmem& us = sm.un[k];
#pragma unroll
      for (int s = 0; s < kStg; ++s) {
        mbar_init(smem_u32(&us.tma_full[s]), 1);
        mbar_init(smem_u32(&us.prep_done[s]), 4);
        mbar_init(smem_u32(&us.stage_free[s]), 10);  // 2 forward + 4 reverse + 4 helper warps
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mbar_init(smem_u32(&us.half_full[h]), 4);    // 2 slabs x 2 state quarters
        mbar_init(smem_u32(&us.half_free[h]), 4);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          mbar_init(smem_u32(&us.tm_full[h][t]), 1);
          mbar_init(smem_u32(&us.tm_free[h][t]), 2);   // both reverse warps of the slab
        }
      }
    }
    mbar_fence_init();
    tma_prefetch_desc(&map_u);
    tma_prefetch_desc(&map_dt);
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_ck);
    tma_prefetch_desc(&map_du);
    tma_prefetch_desc(&map_ddt);
  }
  if (warp == 0) {   // one CTA per SM (shared memory): all 512 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // [row][8] tiles carry the TMA 32-byte swizzle (16-byte half index ^= bit 2 of the row): 8 consecutive rows x one half are then
  // 8 distinct bank groups for per-row 128-bit accesses
  constexpr uint32_t kCT = kNP * kC * 4;    // byte offset of a pair's second channel (row + 32) in such a tile

  if (warp >= 12) {
    // =========================================== helper warpgroups (one per unit) ===========================================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kHelpRegs));
    const int unit = (warp - 12) >> 2;
    const UnitItem it = decode_unit(p, blockIdx.x * kUnits + unit, n_units);
    if (!it.active) return;
    UnitSmem& us = sm.un[unit];
    const int b = it.b, g = it.g, d0 = it.d0;
    const int htid = (int)threadIdx.x - kRecThr - unit * kGroupThr;
    const int hw = (warp - 12) & 3;
    const int row = htid & (kR - 1);        // my channel inside the unit ...
    const int hf = htid >> 6;               // ... and my half of every chunk (warp-uniform)
    const int d = d0 + row;
    const float Dv = a.D ? __ldg(a.D + d) : 0.f;
    const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.f;
    const bool softplus = a.delta_softplus != 0;
    const uint32_t my16 = (uint32_t)row * (kC * 4) + (uint32_t)((hf ^ ((row >> 2) & 1)) << 4);   // my 4 elements in every [row][8] tile
    const uint32_t out16 = (uint32_t)hf * (kR * 16) + (uint32_t)row * 16;    // ... and in the [half][row][4] output tiles

    // ---- B/C gather: elements htid and htid + 128 of a chunk's [8 positions][32 values] tile; the pointers walk backwards ----
    // (B and C have the same position stride on this path: bwd_ws_eligible)
    const int64_t bc_step = (int64_t)kC * a.B_l_stride;
    const bool along_l = a.B_l_stride == 1;   // (.., N, L) layout: 8 consecutive threads read 8 consecutive positions of a state row;
                                              // l-major x_dbl layout: a warp reads the 16 B and 16 C values of one position
    const int bc_e1 = htid + kGroupThr;
    const int bc_pos0 = along_l ? (htid & 7) : (htid >> 5), bc_val0 = along_l ? (htid >> 3) : (htid & 31);
    const int bc_pos1 = along_l ? (bc_e1 & 7) : (bc_e1 >> 5), bc_val1 = along_l ? (bc_e1 >> 3) : (bc_e1 & 31);
    auto bc_src = [&](int pos, int val) -> const float* {
      const int n = val & 15;
      const float* base = (val >= 16) ? (a.C + (int64_t)b * a.C_batch_stride + (int64_t)g * a.C_group_stride + (int64_t)n * a.C_n_stride)
                                      : (a.B + (int64_t)b * a.B_batch_stride + (int64_t)g * a.B_group_stride + (int64_t)n * a.B_n_stride);
      return base + (int64_t)((n_tiles - 1) * kC + pos) * a.B_l_stride;   // first chunk processed = last of the sequence
    };
    const float* bcp0 = bc_src(bc_pos0, bc_val0);
    const float* bcp1 = bc_src(bc_pos1, bc_val1);
    const uint32_t bc_i0 = (uint32_t)(bc_pos0 * kPitch + bc_val0) * 4, bc_i1 = (uint32_t)(bc_pos1 * kPitch + bc_val1) * 4;   // byte offsets in a tile
    const bool bc_ok0 = (bc_val0 & 15) < N, bc_ok1 = (bc_val1 & 15) < N;
    // ---- contraction role: warp = position of the half chunk, lane = (pair quarter q, 16-byte chunk c of the 32 values) ----
    const int c_c = lane & 7, c_q = lane >> 3;
    const int c_n = (c_c & 3) * 4 + 2 * (c_q >> 1) + (c_q & 1);   // the state whose sum this lane ends up with
    const bool c_ok = c_n < N;
    // address of my (state, position hw of half 0) in the chunk being processed; walks backwards by one chunk per iteration
    float* dbc = ((c_c >= 4) ? a.dC : a.dB) + (((int64_t)b * a.ngroups + g) * N + (c_ok ? c_n : 0)) * (int64_t)L + (n_tiles - 1) * kC + hw;

    auto issue_tma = [&](int j) {            // chunk j (processing order) -> stage j % kStg
      const int t = n_tiles - 1 - j, s = j % kStg, l0 = t * kC;
      if (j >= kStg) mbar_wait(smem_u32(&us.stage_free[s]), (uint32_t)((j / kStg - 1) & 1));
      const uint32_t full = smem_u32(&us.tma_full[s]);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((uint32_t)(3 * kR * kC * 4 + kR * kStatePad * 4)) : "memory");
      tma_load_3d(smem_u32(us.U[s]), &map_u, l0, d0, b, full);
      tma_load_3d(smem_u32(us.DT[s]), &map_dt, l0, d0, b, full);
      tma_load_3d(smem_u32(us.DY[s]), &map_dy, l0, d0, b, full);
      // saved state t-1 = state before the chunk's first position; state "-1" is out of bounds -> zeros
      tma_load_2d(smem_u32(us.CK[s]), &map_ck, (t - 1) * kStatePad, b * a.dim + d0, full);
    };
    // discretise my 4 elements of chunk j, publish delta, delta*u, softplus' and the chunk's B/C values (loaded one iteration earlier)
    auto prep = [&](int j, float cb0, float cb1, float4& sg_out) {
      const int t = n_tiles - 1 - j, s = j % kStg, l0 = t * kC + hf * kHP;
      mbar_wait(smem_u32(&us.tma_full[s]), (uint32_t)((j / kStg) & 1));
      const float4 u4 = lds_f4(smem_u32(us.U[s]) + my16);
      const float4 t4 = lds_f4(smem_u32(us.DT[s]) + my16);
      const float uu[4] = {u4.x, u4.y, u4.z, u4.w}, tt[4] = {t4.x, t4.y, t4.z, t4.w};
      float v[4], vu[4], sg[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float xb = tt[e] + bias;
        float dv = xb, sgm = 1.f;
        if (softplus) {
          float wexp;
          dv = softplus_fast(xb, wexp);
          sgm = sigmoid_from_w(xb, wexp);   // softplus' (bwd_kernel.cuh:446-450; == 1 to rounding for x > 20)
        }
        dv = (l0 + e < L) ? dv : 0.f;       // past the end: a = 1, b = 0 (u and dout are TMA zero fill there)
        v[e] = dv;
        vu[e] = dv * uu[e];
        sg[e] = sgm;
      }
      sts_f4(smem_u32(us.DT[s]) + my16, make_float4(v[0], v[1], v[2], v[3]));
      sts_f4(smem_u32(us.SDU[s]) + my16, make_float4(vu[0], vu[1], vu[2], vu[3]));
      sg_out = make_float4(sg[0], sg[1], sg[2], sg[3]);
      sts_f1(smem_u32(us.BC[s]) + bc_i0, cb0);
      sts_f1(smem_u32(us.BC[s]) + bc_i1, cb1);
      fence_proxy_async_smem();   // the in-place delta: ordered before the TMA load that refills this stage three chunks later
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&us.prep_done[s]));
    };

    float dD_acc = 0.f, dbias_acc = 0.f;
    float cb0 = 0.f, cb1 = 0.f;
    // softplus' of my 4 elements: chunk being finalised (i), and the two prepared ahead (i+1, i+2)
    float4 sg_cur, sg_n1 = make_float4(0.f, 0.f, 0.f, 0.f), sg_n2 = make_float4(0.f, 0.f, 0.f, 0.f);
    {
      if (htid == 0) {
        issue_tma(0);
        if (n_tiles > 1) issue_tma(1);
        if (n_tiles > 2) issue_tma(2);
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
      if (n_tiles > 1) prep(1, cb0, cb1, sg_n1);
      cb0 = (bc_ok0 && n_tiles > 2) ? __ldg(bcp0) : 0.f;
      cb1 = (bc_ok1 && n_tiles > 2) ? __ldg(bcp1) : 0.f;
      bcp0 -= bc_step;
      bcp1 -= bc_step;
    }
    const int bar_id = 2 + 2 * unit;         // named barriers 2..5: (unit, half)
    for (int i = 0; i < n_tiles; ++i) {
      const int t = n_tiles - 1 - i, s = i % kStg, c0 = t * kC, ob = i & 1;
      float nb0 = 0.f, nb1 = 0.f;
      if (i + 3 < n_tiles) {                 // B/C of chunk i+3: in flight while chunk i is finalised
        if (bc_ok0) nb0 = __ldg(bcp0);
        if (bc_ok1) nb1 = __ldg(bcp1);
        bcp0 -= bc_step;
        bcp1 -= bc_step;
      }
      // first the discretisation of chunk i+2 (its TMA data landed an iteration ago; the forward warps work on it while the
      // reverse warps are still on chunk i+1), then the TMA loads of chunk i+3, whose stage must first be released by every
      // warp's work on chunk i-1
      if (i + 2 < n_tiles) prep(i + 2, cb0, cb1, sg_n2);
      if (htid == 0 && i + 3 < n_tiles) issue_tma(i + 3);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 1 - hh;
        mbar_wait_relaxed(smem_u32(&us.half_full[h]), (uint32_t)(i & 1));
        // ---------------- contraction of the pair products over the unit's 32 channel pairs ----------------
        {
          const uint32_t src = smem_u32(us.P[h]) + (uint32_t)(c_q * 8) * (kPP * 4) + (uint32_t)hw * 128 + (uint32_t)c_c * 16;
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
          if (c_ok && (i > 0 || c0 + h * kHP + hw < L)) atomicAdd(dbc + h * kHP, kk);
        }
        // ---------------- du / ddelta of my (row, half): sum the 4 state quarters' partials, finalise ----------------
        if (hf == h) {
          const uint32_t srow = smem_u32(us.S12[h]) + (uint32_t)row * (kS12P * 4);
          float4 s1 = lds_f4(srow), s2 = lds_f4(srow + 16);
#pragma unroll
          for (int q = 1; q < 4; ++q) {
            const float4 x1 = lds_f4(srow + q * 32), x2 = lds_f4(srow + q * 32 + 16);
            s1.x += x1.x; s1.y += x1.y; s1.z += x1.z; s1.w += x1.w;
            s2.x += x2.x; s2.y += x2.y; s2.z += x2.z; s2.w += x2.w;
          }
          const float4 u4 = lds_f4(smem_u32(us.U[s]) + my16);
          const float4 y4 = lds_f4(smem_u32(us.DY[s]) + my16);
          const float4 l4 = lds_f4(smem_u32(us.DT[s]) + my16);
          const float4 g4 = sg_cur;
          float4 o_du, o_dd;
          o_du.x = fmaf(l4.x, s1.x, Dv * y4.x);                               // bwd_kernel.cuh:211, :280
          o_du.y = fmaf(l4.y, s1.y, Dv * y4.y);
          o_du.z = fmaf(l4.z, s1.z, Dv * y4.z);
          o_du.w = fmaf(l4.w, s1.w, Dv * y4.w);
          o_dd.x = fmaf(u4.x, s1.x, s2.x * kLn2) * g4.x;                      // :281-284, :446-450 (s2 in units of log2 e)
          o_dd.y = fmaf(u4.y, s1.y, s2.y * kLn2) * g4.y;
          o_dd.z = fmaf(u4.z, s1.z, s2.z * kLn2) * g4.z;
          o_dd.w = fmaf(u4.w, s1.w, s2.w * kLn2) * g4.w;
          // positions past the end contribute exact zeros (delta forced to 0, u / dout / B / C zero fill), so no masking is needed
          dbias_acc += (o_dd.x + o_dd.y) + (o_dd.z + o_dd.w);
          dD_acc = fmaf(y4.x, u4.x, fmaf(y4.y, u4.y, fmaf(y4.z, u4.z, fmaf(y4.w, u4.w, dD_acc))));   // :213
          sts_f4(smem_u32(us.DU[ob]) + out16, o_du);
          sts_f4(smem_u32(us.DDT[ob]) + out16, o_dd);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&us.half_free[h]));
        if (hf == h) {                       // the two warps of this half hand their tiles to the TMA store
          if (hh == 1 && lane == 0) mbar_arrive(smem_u32(&us.stage_free[s]));   // (warps of half 0 are done with the stage here ...
          fence_proxy_async_smem();          // my du / ddelta writes -> visible to the async proxy
          if ((htid & 63) == 0) tma_store_wait_read<0>();   // my store of chunk i-1 has read its tiles: buffer ob^1 is free for chunk i+1
          named_bar_sync(bar_id + h, 64);
          if ((htid & 63) == 0) {
            tma_store_3d(&map_du, smem_u32(us.DU[ob]) + h * (kR * 16), c0 + kHP * h, d0, b);
            tma_store_3d(&map_ddt, smem_u32(us.DDT[ob]) + h * (kR * 16), c0 + kHP * h, d0, b);
            tma_store_commit();
          }
        } else if (hh == 1 && lane == 0) {
          mbar_arrive(smem_u32(&us.stage_free[s]));                             //  ... and the warps of half 1 after contracting half 0)
        }
      }
      dbc -= kC;
      sg_cur = sg_n1;
      sg_n1 = sg_n2;
      cb0 = nb0;
      cb1 = nb1;
    }
    if ((htid & 63) == 0) tma_store_wait_all<0>();
    if (a.dD != nullptr) atomicAdd(a.dD + d, dD_acc);                      // two threads (halves) per channel, summed over batch
    if (a.ddelta_bias != nullptr) atomicAdd(a.ddelta_bias + d, dbias_acc);
    return;
  }

  // =========================================== recurrence warps ===========================================
  const int q = warp & 3;                  // TMEM lane quarter
  const int unit = q >> 1, slab = q & 1;
  const int hs = lane >> 4;                // which 8 states (the 16 lanes of a half warp read the same B / C addresses)
  const int pp = slab * 16 + (lane & 15);  // channel pair inside the unit: rows pp and pp + 32
  const UnitItem it = decode_unit(p, blockIdx.x * kUnits + unit, n_units);
  UnitSmem& us = sm.un[unit];
  const uint32_t tm_row0 = sm.tmem_base + ((uint32_t)(32 * q) << 16);
  const uint32_t hoff0 = (uint32_t)((pp >> 2) & 1) << 4;   // byte offset of half 0 inside my rows of the swizzled [row][8] tiles (half 1: ^ 16)
  const uint32_t ck_sw = (uint32_t)((pp >> 1) & 3);        // 64-byte swizzle of the saved-state tile: 16-byte chunk index ^= (row >> 1) & 3
  const int d_lo = it.d0 + pp;

  if (warp < 4) {
    // ------------------------------------------- forward warps -------------------------------------------
    if (it.active) {
      u64 A2p[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int d = d_lo + c * kNP;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int n0 = hs * 8 + 2 * k;
          const float a0 = (n0 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)n0 * a.A_n_stride) * kLog2e : 0.f;
          const float a1 = (n0 + 1 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)(n0 + 1) * a.A_n_stride) * kLog2e : 0.f;
          A2p[c][k] = pk2(a0, a1);
        }
      }
      for (int i = 0; i < n_tiles; ++i) {
        const int s = i % kStg, tb = i & 1;
        const uint32_t par = (uint32_t)((i / kStg) & 1);
        mbar_wait(smem_u32(&us.prep_done[s]), par);
        mbar_wait(smem_u32(&us.tma_full[s]), par);     // completed long ago: makes the TMA-written state tile visible to me
        const uint32_t ck_row = smem_u32(us.CK[s]) + pp * (kStatePad * 4);
        const uint32_t bc_base = smem_u32(us.BC[s]) + hs * 32;
        const uint32_t sd_row = smem_u32(us.DT[s]) + pp * (kC * 4);
        const uint32_t sdu_row = smem_u32(us.SDU[s]) + pp * (kC * 4);
        u64 x[2][4];
#pragma unroll
        for (int c = 0; c < 2; ++c) {     // rows pp and pp + 32 have the same swizzle key
          lds_2x64(ck_row + c * (kNP * kStatePad * 4) + (((uint32_t)(2 * hs) ^ ck_sw) << 4), x[c][0], x[c][1]);
          lds_2x64(ck_row + c * (kNP * kStatePad * 4) + (((uint32_t)(2 * hs + 1) ^ ck_sw) << 4), x[c][2], x[c][3]);
        }
        if (i >= 2) {
          mbar_wait(smem_u32(&us.tm_free[slab][tb]), (uint32_t)(((i >> 1) - 1) & 1));   // the reverse warps have read chunk i-2 back
          tc_fence_after();
          __syncwarp();
        }
        const uint32_t tm_chunk = tm_row0 + (uint32_t)(tb * (kC * kRowCols));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float dl[2][4], duk[2][4];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float4 t4 = lds_f4(sd_row + (hoff0 ^ (h * 16u)) + c * kCT);
            const float4 v4 = lds_f4(sdu_row + (hoff0 ^ (h * 16u)) + c * kCT);
            dl[c][0] = t4.x; dl[c][1] = t4.y; dl[c][2] = t4.z; dl[c][3] = t4.w;
            duk[c][0] = v4.x; duk[c][1] = v4.y; duk[c][2] = v4.z; duk[c][3] = v4.w;
          }
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = h * 4 + jj;
            u64 Bq[4];
            lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bq[0], Bq[1]);
            lds_2x64(bc_base + (uint32_t)j * (kPitch * 4) + 16, Bq[2], Bq[3]);
            u64 e[2][4];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const u64 dd = pk2(dl[c][jj], dl[c][jj]);
              const u64 duu = pk2(duk[c][jj], duk[c][jj]);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                float t0, t1;
                upk2(mul2(dd, A2p[c][k]), t0, t1);
                e[c][k] = pk2(ex2(t0), ex2(t1));
                x[c][k] = fma2(e[c][k], x[c][k], mul2(duu, Bq[k]));
              }
            }
            // one 16-column group per reverse warp of the slab (state quarter sb): [x c0 | x c1 | a c0 | a c1], 2 pairs each
#pragma unroll
            for (int sb = 0; sb < 2; ++sb) {
              const u64 grp[8] = {x[0][2 * sb], x[0][2 * sb + 1], x[1][2 * sb], x[1][2 * sb + 1],
                                  e[0][2 * sb], e[0][2 * sb + 1], e[1][2 * sb], e[1][2 * sb + 1]};
              tmem_st16(tm_chunk + (uint32_t)(j * kRowCols + sb * 16), grp);
            }
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&us.tm_full[slab][tb]));
          mbar_arrive(smem_u32(&us.stage_free[s]));   // ordered after every lane's reads of the stage by the __syncwarp
        }
      }
    }
  } else {
    // ------------------------------------------- reverse warps -------------------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRevRegs));
    const int sb = (warp >> 2) - 1;          // which 4 of my lane half's 8 states (warps 4-7: 0, warps 8-11: 1)
    const int st0 = hs * 8 + sb * 4;         // my first state
    if (it.active) {
      u64 A2p[2][2], dA2[2][2], w2[2][2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int d = d_lo + c * kNP;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int n0 = st0 + 2 * k;
          const float a0 = (n0 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)n0 * a.A_n_stride) * kLog2e : 0.f;
          const float a1 = (n0 + 1 < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)(n0 + 1) * a.A_n_stride) * kLog2e : 0.f;
          A2p[c][k] = pk2(a0, a1);
          dA2[c][k] = pk2(0.f, 0.f);
          w2[c][k] = pk2(0.f, 0.f);   // a_{l+1} * dx_{l+1}: zero beyond the last position
        }
      }
      const uint32_t p_off = (uint32_t)pp * (kPP * 4) + (uint32_t)st0 * 4;                  // dB values st0..st0+3; dC at + 64
      const uint32_t s_off = (uint32_t)pp * (kS12P * 4) + (uint32_t)(hs * 2 + sb) * 32;    // s1 x4; s2 x4 at + 16; second channel at + kNP rows

      for (int i = 0; i < n_tiles; ++i) {
        const int s = i % kStg, tb = i & 1;
        const uint32_t par = (uint32_t)((i / kStg) & 1);
        mbar_wait(smem_u32(&us.prep_done[s]), par);
        mbar_wait(smem_u32(&us.tma_full[s]), par);     // makes the TMA-written dout / state tiles visible to me
        const uint32_t dy_row = smem_u32(us.DY[s]) + pp * (kC * 4);   // + hoff[half] + c * kCT
        const uint32_t ck_row = smem_u32(us.CK[s]) + pp * (kStatePad * 4) + (((uint32_t)(2 * hs + sb) ^ ck_sw) << 4);
        const uint32_t bc_base = smem_u32(us.BC[s]) + st0 * 4;
        const uint32_t sd_row = smem_u32(us.DT[s]) + pp * (kC * 4);
        const uint32_t sdu_row = smem_u32(us.SDU[s]) + pp * (kC * 4);
        const uint32_t tm_chunk = tm_row0 + (uint32_t)(tb * (kC * kRowCols) + sb * 16);

        mbar_wait(smem_u32(&us.tm_full[slab][tb]), (uint32_t)((i >> 1) & 1));   // the forward warp has stored this chunk
        tc_fence_after();
        __syncwarp();
        // G[j] = my 16 columns of position j: states after the position (x c0, x c1) and its decays (a c0, a c1), two packed pairs
        // each, as raw 32-bit registers; x0 = the saved state in front of the chunk
        uint32_t G[kC][16];
        u64 x0[2][2];
        tmem_ld16(tm_chunk + (uint32_t)((kC - 1) * kRowCols), G[kC - 1]);
        tmem_ld16(tm_chunk + (uint32_t)((kC - 2) * kRowCols), G[kC - 2]);
#pragma unroll
        for (int c = 0; c < 2; ++c) lds_2x64(ck_row + c * (kNP * kStatePad * 4), x0[c][0], x0[c][1]);
        tmem_ld_wait32(G[kC - 1], G[kC - 2]);

        auto rev_half = [&](auto HF) {
          constexpr int h = decltype(HF)::value;
          if (i > 0) mbar_wait(smem_u32(&us.half_free[h]), (uint32_t)((i - 1) & 1));   // the helpers have consumed this half of chunk i-1
          const uint32_t p_row = smem_u32(us.P[h]) + p_off;
          float dyh[2][4], dlh[2][4], duh[2][4], s1p[2][4], s2p[2][4];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float4 y4 = lds_f4(dy_row + (hoff0 ^ (h * 16u)) + c * kCT);
            const float4 t4 = lds_f4(sd_row + (hoff0 ^ (h * 16u)) + c * kCT);
            const float4 v4 = lds_f4(sdu_row + (hoff0 ^ (h * 16u)) + c * kCT);
            dyh[c][0] = y4.x; dyh[c][1] = y4.y; dyh[c][2] = y4.z; dyh[c][3] = y4.w;
            dlh[c][0] = t4.x; dlh[c][1] = t4.y; dlh[c][2] = t4.z; dlh[c][3] = t4.w;
            duh[c][0] = v4.x; duh[c][1] = v4.y; duh[c][2] = v4.z; duh[c][3] = v4.w;
          }
#pragma unroll
          for (int jj = 3; jj >= 0; --jj) {
            const int j = h * 4 + jj;
            // rows j and j-1 are in registers; fetch what position j-1 needs (row j-2) while j is computed
            if (j >= 2) tmem_ld16(tm_chunk + (uint32_t)((j - 2) * kRowCols), G[j >= 2 ? j - 2 : 0]);
            u64 Bq[2], Cq[2];
            lds_2x64(bc_base + (uint32_t)j * (kPitch * 4), Bq[0], Bq[1]);
            lds_2x64(bc_base + (uint32_t)j * (kPitch * 4) + 64, Cq[0], Cq[1]);
            u64 pB[2], pC[2];
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const u64 dyy = pk2(dyh[c][jj], dyh[c][jj]);
              const u64 dd = pk2(dlh[c][jj], dlh[c][jj]);
              const u64 duu = pk2(duh[c][jj], duh[c][jj]);
              u64 s1a, s2a;
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const int e = c * 2 + k, jm = j == 0 ? 0 : j - 1;
                const u64 xprev = (j == 0) ? x0[c][k] : pku(G[jm][2 * e], G[jm][2 * e + 1]);
                const u64 xcur = pku(G[j][2 * e], G[j][2 * e + 1]);
                const u64 acur = pku(G[j][8 + 2 * e], G[j][8 + 2 * e + 1]);
                const u64 dx = fma2(Cq[k], dyy, w2[c][k]);                       // dx_{l,n}
                s1a = (k == 0) ? mul2(dx, Bq[k]) : fma2(dx, Bq[k], s1a);         // sum_n dx * B          (bwd_kernel.cuh:280-281)
                w2[c][k] = mul2(acur, dx);                                       // a_l * dx_l: carried to position l-1 ...
                const u64 wg = mul2(w2[c][k], xprev);                            // ... and dx * a_l * x_{l-1}  (:283)
                s2a = (k == 0) ? mul2(wg, A2p[c][k]) : fma2(wg, A2p[c][k], s2a); // in units of log2(e)
                dA2[c][k] = fma2(wg, dd, dA2[c][k]);                             // :286
                if (c == 0) {                                                    // channel-pair products for dB / dC
                  pB[k] = mul2(duu, dx);
                  pC[k] = mul2(dyy, xcur);
                } else {
                  pB[k] = fma2(duu, dx, pB[k]);
                  pC[k] = fma2(dyy, xcur, pC[k]);
                }
              }
              s1p[c][jj] = hsum2(s1a);
              s2p[c][jj] = hsum2(s2a);
            }
            sts_2x64(p_row + (uint32_t)jj * 128, pB[0], pB[1]);
            sts_2x64(p_row + (uint32_t)jj * 128 + 64, pC[0], pC[1]);
            if (j >= 2) tmem_ld_wait16(G[j >= 2 ? j - 2 : 0]);
          }
          const uint32_t s_row = smem_u32(us.S12[h]) + s_off;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            sts_f4(s_row + c * (kNP * kS12P * 4), make_float4(s1p[c][0], s1p[c][1], s1p[c][2], s1p[c][3]));
            sts_f4(s_row + c * (kNP * kS12P * 4) + 16, make_float4(s2p[c][0], s2p[c][1], s2p[c][2], s2p[c][3]));
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&us.half_full[h]));
        };
        rev_half(std::integral_constant<int, 1>{});
        rev_half(std::integral_constant<int, 0>{});
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_u32(&us.tm_free[slab][tb]));
          mbar_arrive(smem_u32(&us.stage_free[s]));   // ordered after my reads by the __syncwarp
        }
      }

#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int d = d_lo + c * kNP;
        float da[4];
        upk2(dA2[c][0], da[0], da[1]);
        upk2(dA2[c][1], da[2], da[3]);
#pragma unroll
        for (int n = 0; n < 4; ++n)
          if (st0 + n < N) atomicAdd(a.dA + (int64_t)d * N + st0 + n, da[n]);   // sum over batch
      }
    }
  }
  // every TMEM access of the CTA is complete before the columns are returned
  tc_fence_before();
  named_bar_sync(1, kRecThr);
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(sm.tmem_base), "n"(kTmemCols) : "memory");
  }
}

inline bool make_ckpt_map(CUtensorMap* map, const float* base, int64_t rows, int n_ckpt) {
  auto enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)n_ckpt * kStatePad, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)n_ckpt * kStatePad * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kStatePad, (cuuint32_t)kR};
  const cuuint32_t estr[2] = {1, 1};
  // 64-byte rows with the 64-byte swizzle: the 8 rows a quarter-warp reads (one 16-byte chunk each) fall on 8 distinct bank groups
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// Usable when ptxas gave the kernel exactly the launch register count the setmaxnreg arithmetic assumes (otherwise
// setmaxnreg.inc could wait for registers that never come).
bool bwd_tm_usable() {
  static const bool ok = [] {
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, selscan_bwd_tm_kernel) != cudaSuccess) {
      (void)cudaGetLastError();
      return false;
    }
    return fa.numRegs == kLaunchRegs;
  }();
  return ok;
}

cudaError_t launch_bwd_tm(const BwdLaunch& p, cudaStream_t stream) {
  const selscan_bwd_args& a = p.a;
  CUtensorMap mu, mdt, mdy, mck, mdu, mddt;
  if (!make_row_map_sw(&mu, a.u, a.seqlen, a.dim, a.batch, a.u_d_stride, a.u_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B) ||
      !make_row_map_sw(&mdt, a.delta, a.seqlen, a.dim, a.batch, a.delta_d_stride, a.delta_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B) ||
      !make_row_map_sw(&mdy, a.dout, a.seqlen, a.dim, a.batch, a.dout_d_stride, a.dout_batch_stride, kC, kR, CU_TENSOR_MAP_SWIZZLE_32B) ||
      !make_row_map(&mdu, a.du, a.seqlen, a.dim, a.batch, a.du_d_stride, a.du_batch_stride, kHP, kR) ||
      !make_row_map(&mddt, a.ddelta, a.seqlen, a.dim, a.batch, a.ddelta_d_stride, a.ddelta_batch_stride, kHP, kR) ||
      !make_ckpt_map(&mck, a.ckpt, (int64_t)a.batch * a.dim, p.n_ckpt))
    return cudaErrorNotSupported;
  constexpr int smem = (int)sizeof(TmSmem) + 1024;
  static_assert(sizeof(TmSmem) + 1024 <= 232448, "one CTA per SM");
  static std::atomic<unsigned long long> configured{0};   // one cudaFuncSetAttribute per device, not per launch
  if (const cudaError_t e = set_smem_once(configured, selscan_bwd_tm_kernel, smem)) return e;
  const int n_units = (int)((int64_t)a.batch * a.ngroups * (p.dim_per_group / kR));
  const unsigned grid = (unsigned)((n_units + kUnits - 1) / kUnits);
  selscan_bwd_tm_kernel<<<grid, kThr, smem, stream>>>(mu, mdt, mdy, mck, mdu, mddt, p, n_units);
  return cudaGetLastError();
}

}  // namespace selscan
