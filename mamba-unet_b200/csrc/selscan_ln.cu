// LayerNorm over short channel rows (d_model = 96 .. 1536), the op on the caller side of every SS2D block
// (VSSBlock.ln_1, code/networks/mamba_sys.py:552,559; PatchMerging2D / PatchExpand norms :205,242).  sm_100a.
//
// One warp owns a row: lane l holds channels l, l + 32, ... (NV per lane) in registers, so the row is read once, both
// moments are exact two-pass sums over registers, and several short rows are in flight per warp to keep enough loads
// outstanding.  Backward keeps per-lane gamma / beta partial sums in registers across the rows of a CTA and writes one
// (2, dim) partial per CTA; the caller sums the partials.
#include <cuda_runtime.h>
#include <stdint.h>

#include "selscan_kernels.h"
#include "selscan_tma_host.h"

namespace selscan {

namespace {

constexpr int kLnThreads = 256;
constexpr int kLnWarps = kLnThreads / 32;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV, int ROWS>
__global__ void __launch_bounds__(kLnThreads)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float eps, float* __restrict__ y,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int64_t rows, int D) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gw[NV], gb[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int d = lane + 32 * j;
    gw[j] = d < D ? __ldg(w + d) : 0.f;
    gb[j] = d < D ? __ldg(b + d) : 0.f;
  }
  const float inv_d = 1.f / (float)D;
  const int64_t stride = (int64_t)gridDim.x * kLnWarps * ROWS;
  for (int64_t r0 = ((int64_t)blockIdx.x * kLnWarps + warp) * ROWS; r0 < rows; r0 += stride) {
    float v[ROWS][NV];
#pragma unroll
    for (int k = 0; k < ROWS; ++k)
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int d = lane + 32 * j;
        v[k][j] = (r0 + k < rows && d < D) ? __ldg(x + (r0 + k) * D + d) : 0.f;
      }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      if (r0 + k >= rows) break;
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) s += v[k][j];
      const float mean = wsum(s) * inv_d;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float c = (lane + 32 * j < D) ? v[k][j] - mean : 0.f;
        q = fmaf(c, c, q);
      }
      const float rs = rsqrtf(wsum(q) * inv_d + eps);
      if (mean_out != nullptr && lane == 0) {
        mean_out[r0 + k] = mean;
        rstd_out[r0 + k] = rs;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int d = lane + 32 * j;
        if (d < D) y[(r0 + k) * D + d] = fmaf((v[k][j] - mean) * rs, gw[j], gb[j]);
      }
    }
  }
}

template <int NV, int ROWS>
__global__ void __launch_bounds__(kLnThreads)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
              const float* __restrict__ w, float* __restrict__ dx, float* __restrict__ part, int64_t rows, int D) {
  __shared__ float red[kLnWarps][2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float gw[NV], ag[NV], ab[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int d = lane + 32 * j;
    gw[j] = d < D ? __ldg(w + d) : 0.f;
    ag[j] = 0.f;
    ab[j] = 0.f;
  }
  const float inv_d = 1.f / (float)D;
  const int64_t stride = (int64_t)gridDim.x * kLnWarps * ROWS;
  for (int64_t r0 = ((int64_t)blockIdx.x * kLnWarps + warp) * ROWS; r0 < rows; r0 += stride) {
    float g[ROWS][NV], v[ROWS][NV], mu[ROWS], rs[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      const bool ok = r0 + k < rows;
      mu[k] = ok ? __ldg(mean + r0 + k) : 0.f;
      rs[k] = ok ? __ldg(rstd + r0 + k) : 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int d = lane + 32 * j;
        g[k][j] = (ok && d < D) ? __ldg(dy + (r0 + k) * D + d) : 0.f;
        v[k][j] = (ok && d < D) ? __ldg(x + (r0 + k) * D + d) : 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) {
      if (r0 + k >= rows) break;
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float xh = (lane + 32 * j < D) ? (v[k][j] - mu[k]) * rs[k] : 0.f;
        v[k][j] = xh;
        ab[j] += g[k][j];
        ag[j] = fmaf(g[k][j], xh, ag[j]);
        const float t = g[k][j] * gw[j];
        g[k][j] = t;
        a1 += t;
        a2 = fmaf(t, xh, a2);
      }
      a1 = wsum(a1) * inv_d;
      a2 = wsum(a2) * inv_d;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const int d = lane + 32 * j;
        if (d < D) dx[(r0 + k) * D + d] = rs[k] * (g[k][j] - a1 - v[k][j] * a2);
      }
    }
  }
  // per-CTA (2, D) partial: sum the 8 warps' register partials channel block by channel block
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    red[warp][0][lane] = ag[j];
    red[warp][1][lane] = ab[j];
    __syncthreads();
    if (warp < 2) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < kLnWarps; ++k) s += red[k][warp][lane];
      const int d = lane + 32 * j;
      if (d < D) part[((int64_t)blockIdx.x * 2 + warp) * D + d] = s;
    }
    __syncthreads();
  }
}

int ln_grid(int64_t rows, int rows_per_cta) {
  const int64_t want = (rows + rows_per_cta - 1) / rows_per_cta;
  const int64_t cap = (int64_t)sm_count() * 8;   // same helper as every other grid heuristic: the partial-buffer query and the launcher agree
  return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

// rows in flight per warp: about 12 loads outstanding per lane
constexpr int ln_rows(int nv) { return nv >= 12 ? 1 : (nv >= 6 ? 2 : (nv >= 3 ? 4 : 8)); }

template <int NV>
cudaError_t run_fwd(const float* x, const float* w, const float* b, float eps, float* y, float* mean, float* rstd, int64_t rows, int D,
                    cudaStream_t s) {
  constexpr int R = ln_rows(NV);
  ln_fwd_kernel<NV, R><<<ln_grid(rows, kLnWarps * R), kLnThreads, 0, s>>>(x, w, b, eps, y, mean, rstd, rows, D);
  return cudaGetLastError();
}

template <int NV>
cudaError_t run_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* w, float* dx, float* part,
                    int64_t rows, int D, cudaStream_t s) {
  constexpr int R = NV >= 12 ? 1 : ln_rows(NV) / 2 > 0 ? ln_rows(NV) / 2 : 1;   // two tensors are loaded per row here
  ln_bwd_kernel<NV, R><<<ln_grid(rows, kLnWarps * R), kLnThreads, 0, s>>>(dy, x, mean, rstd, w, dx, part, rows, D);
  return cudaGetLastError();
}

int ln_bwd_rows_per_cta(int nv) {
  const int r = nv >= 12 ? 1 : (ln_rows(nv) / 2 > 0 ? ln_rows(nv) / 2 : 1);
  return kLnWarps * r;
}

}  // namespace

// values per lane the kernels are instantiated for (dim <= 32 * NV); 0 = unsupported
int ln_nv(int D) {
  static const int nvs[] = {1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48};
  for (int nv : nvs)
    if (D <= 32 * nv) return nv;
  return 0;
}

int64_t ln_bwd_ctas(int64_t rows, int D) {
  const int nv = ln_nv(D);
  if (nv == 0 || rows <= 0) return 0;
  return ln_grid(rows, ln_bwd_rows_per_cta(nv));
}

#define LN_DISPATCH(FN, ...)                     \
  switch (ln_nv(D)) {                            \
    case 1: return FN<1>(__VA_ARGS__);           \
    case 2: return FN<2>(__VA_ARGS__);           \
    case 3: return FN<3>(__VA_ARGS__);           \
    case 4: return FN<4>(__VA_ARGS__);           \
    case 6: return FN<6>(__VA_ARGS__);           \
    case 8: return FN<8>(__VA_ARGS__);           \
    case 12: return FN<12>(__VA_ARGS__);         \
    case 16: return FN<16>(__VA_ARGS__);         \
    case 24: return FN<24>(__VA_ARGS__);         \
    case 32: return FN<32>(__VA_ARGS__);         \
    case 48: return FN<48>(__VA_ARGS__);         \
    default: return cudaErrorInvalidValue;       \
  }

cudaError_t launch_ln_fwd(const float* x, const float* w, const float* b, float eps, float* y, float* mean, float* rstd, int64_t rows,
                          int D, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  LN_DISPATCH(run_fwd, x, w, b, eps, y, mean, rstd, rows, D, stream)
}

cudaError_t launch_ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* w, float* dx, float* part,
                          int64_t rows, int D, cudaStream_t stream) {
  if (rows == 0) return cudaSuccess;
  LN_DISPATCH(run_bwd, dy, x, mean, rstd, w, dx, part, rows, D, stream)
}

}  // namespace selscan
