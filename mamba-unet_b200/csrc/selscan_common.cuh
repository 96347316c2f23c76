// Device-side helpers shared by the forward and backward selective-scan kernels (sm_100a).
//
// Numerics follow the reference kernels so that parity holds to fp32 round-off:
//   softplus     x <= 20 ? log1pf(expf(x)) : x         selective_scan_fwd_kernel.cuh:153-156
//   decay        a = exp2f(delta * (A * log2(e)))      selective_scan_fwd_kernel.cuh:168-175,216
//   silu gate    z / (1 + expf(-z))                    selective_scan_fwd_kernel.cuh:293
//   softplus'    g / (1 + expf(-x))  for x <= 20       selective_scan_bwd_kernel.cuh:446-450
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace selscan {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kStatePad = 16;     // SELSCAN_B200_STATE_PAD
constexpr int kCkptInterval = 8;  // SELSCAN_B200_CKPT_INTERVAL
constexpr int kMaxFusedDtRank = 12;   // fused dt_proj: ranks above this cost more FMAs per element than a GEMM pass saves

// MUFU.EX2 (one SFU op, flush-to-zero): the only transcendental on the per-(position, state) path.
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float softplus20(float x) { return x <= 20.f ? log1pf(__expf(x)) : x; }

// Branch-free softplus for the tiled kernels, ~16 instructions and 2 MUFU (EX2, RCP):
//   softplus(x) = max(x, 0) + log1p(w),  w = exp(-|x|) in (0, 1],
//   log1p(w)    = 2 atanh(t),            t = w / (2 + w) in (0, 1/3]  (degree-4 minimax in t^2: rel. err of the whole function < 2.5e-7).
// Same function as the reference's `x <= 20 ? log1pf(expf(x)) : x` (fwd_kernel.cuh:155) to fp32 round-off: for
// x > 20 the correction term is < 2.1e-9, below half an ulp of x.  `w` is returned for the sigmoid in the backward.
// atanh(t) / t on t^2 in [0, 1/9]: degree-4 minimax (max error 4e-9; Remez, scripts not needed at run time)
constexpr float kAtanhC1 = 0.33333155512809753f, kAtanhC2 = 0.2001255750656128f, kAtanhC3 = 0.13978832960128784f, kAtanhC4 = 0.14095774292945862f;
__device__ __forceinline__ float softplus_fast(float x, float& w_out) {
  float w;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(-fabsf(x) * 1.4426950408889634f));
  w_out = w;
  const float t = __fdividef(w, 2.f + w);
  const float t2 = t * t;
  float p = fmaf(t2, kAtanhC4, kAtanhC3);
  p = fmaf(p, t2, kAtanhC2);
  p = fmaf(p, t2, kAtanhC1);
  p = fmaf(p, t2, 1.f);
  return fmaf(2.f * t, p, fmaxf(x, 0.f));
}
// sigmoid(x) from w = exp(-|x|):  x >= 0 ? 1/(1+w) : w/(1+w)
__device__ __forceinline__ float sigmoid_from_w(float x, float w) {
  const float r = __fdividef(1.f, 1.f + w);
  return x >= 0.f ? r : w * r;
}
__device__ __forceinline__ float sigmoidf_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// 4 consecutive sequence positions starting at l0 from a unit-stride row; positions >= L read as 0.
__device__ __forceinline__ void load_row4(const float* __restrict__ row, int l0, int L, bool vec, float (&v)[4]) {
  if (vec && l0 + 4 <= L) {
    const float4 t = ldg4(row + l0);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (l0 + j < L) ? __ldg(row + l0 + j) : 0.f;
  }
}

__device__ __forceinline__ void store_row4(float* __restrict__ row, int l0, int L, bool vec, const float (&v)[4]) {
  if (vec && l0 + 4 <= L) {
    *reinterpret_cast<float4*>(row + l0) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (l0 + j < L) row[l0 + j] = v[j];
  }
}

// 4 consecutive positions of one (group, state) row of B or C with arbitrary position stride.
__device__ __forceinline__ void load_bc4(const float* __restrict__ base, int64_t l_stride, int l0, int L, bool vec,
                                         float (&v)[4]) {
  if (vec && l0 + 4 <= L) {
    const float4 t = ldg4(base + l0);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (l0 + j < L) ? __ldg(base + (int64_t)(l0 + j) * l_stride) : 0.f;
  }
}

}  // namespace selscan
