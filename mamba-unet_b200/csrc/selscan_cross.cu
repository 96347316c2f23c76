// CrossScan / CrossMerge of SS2D as two single-pass kernels (sm_100a).
//
// Reference: code/networks/mamba_sys.py:403-404 builds the 4 scan orders with stack + transpose.contiguous + flip + cat
// (4 kernels, ~18 tensor passes), and :429-432 merges the 4 scan outputs with flip + 2 x transpose.contiguous + 3 adds.
// Here one CTA owns one (batch, channel) image plane, stages it in shared memory once and emits / consumes all four
// orders with coalesced global accesses:
//   scatter  x (B, D, H, W)        -> xs (B, 4, D, L):  k=0 row-major, k=1 column-major, k=2/3 the same reversed
//   gather   ys (B, 4, D, L)       -> y (B, D, H*W):    y[h,w] = ys0[hW+w] + ys2[L-1-(hW+w)] + ys1[wH+h] + ys3[L-1-(wH+h)]
// `gather` is also the backward of `scatter` and vice versa.  The (B, 4, D, L) tensors may have a padded row pitch
// (rows 16-byte aligned for uneven L), which is what the tiled scan kernels want.
#include <cuda_runtime.h>
#include <stdint.h>

namespace selscan {

namespace {

// shared plane, pitch W+1: column-major reads hit distinct banks
__global__ void cross_scatter_kernel(const float* __restrict__ x, float* __restrict__ xs, int D, int H, int W, int64_t pitch) {
  extern __shared__ float plane[];
  const int L = H * W, P = W + 1;
  const int64_t bd = blockIdx.x;                 // b * D + d
  const int64_t b = bd / D, d = bd - b * D;
  const float* src = x + bd * L;
  float* o0 = xs + ((b * 4 + 0) * D + d) * pitch;
  float* o1 = xs + ((b * 4 + 1) * D + d) * pitch;
  float* o2 = xs + ((b * 4 + 2) * D + d) * pitch;
  float* o3 = xs + ((b * 4 + 3) * D + d) * pitch;
  for (int l = threadIdx.x; l < L; l += blockDim.x) {
    const float v = __ldg(src + l);
    plane[(l / W) * P + (l % W)] = v;
    o0[l] = v;
    o2[L - 1 - l] = v;
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += blockDim.x) {   // l = w * H + h
    const float v = plane[(l % H) * P + (l / H)];
    o1[l] = v;
    o3[L - 1 - l] = v;
  }
}

__global__ void cross_gather_kernel(const float* __restrict__ ys, float* __restrict__ y, int D, int H, int W, int64_t pitch) {
  extern __shared__ float plane[];
  const int L = H * W, P = W + 1;
  const int64_t bd = blockIdx.x;
  const int64_t b = bd / D, d = bd - b * D;
  const float* i0 = ys + ((b * 4 + 0) * D + d) * pitch;
  const float* i1 = ys + ((b * 4 + 1) * D + d) * pitch;
  const float* i2 = ys + ((b * 4 + 2) * D + d) * pitch;
  const float* i3 = ys + ((b * 4 + 3) * D + d) * pitch;
  for (int l = threadIdx.x; l < L; l += blockDim.x)     // column-major pair summed, parked transposed
    plane[(l % H) * P + (l / H)] = __ldg(i1 + l) + __ldg(i3 + L - 1 - l);
  __syncthreads();
  float* dst = y + bd * L;
  for (int l = threadIdx.x; l < L; l += blockDim.x)
    dst[l] = (__ldg(i0 + l) + __ldg(i2 + L - 1 - l)) + plane[(l / W) * P + (l % W)];
}

}  // namespace

cudaError_t launch_cross(bool scatter, const float* in, float* out, int B, int D, int H, int W, int64_t pitch, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  const int L = H * W;
  int threads = ((L + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  const size_t smem = (size_t)H * (W + 1) * sizeof(float);
  const unsigned grid = (unsigned)((int64_t)B * D);
  cudaError_t e = cudaSuccess;
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(scatter ? cross_scatter_kernel : cross_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (scatter)
    cross_scatter_kernel<<<grid, threads, smem, stream>>>(in, out, D, H, W, pitch);
  else
    cross_gather_kernel<<<grid, threads, smem, stream>>>(in, out, D, H, W, pitch);
  return cudaGetLastError();
}

}  // namespace selscan
