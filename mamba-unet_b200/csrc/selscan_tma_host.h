// Host helper: build a 3-D TMA tensor map over a (batch, dim, seqlen) fp32 tensor with unit seqlen stride.
// cuTensorMapEncodeTiled is resolved through the runtime (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace selscan {

inline PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  }();
  return fn;
}

// dims (fastest first): {seqlen, dim, batch}; box {box_l, box_rows, 1}; rows of box_l*4 = 128 B use the 128-byte swizzle
// and rows of 64 B the 64-byte swizzle (unless swizzle = false); other boxes are dense.
inline bool make_row_map(CUtensorMap* map, const float* base, int seqlen, int dim, int batch, int64_t d_stride,
                         int64_t batch_stride, int box_l, int box_rows, bool swizzle = true) {
  auto enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)seqlen, (cuuint64_t)dim, (cuuint64_t)batch};
  const cuuint64_t gstr[2] = {(cuuint64_t)d_stride * 4, (cuuint64_t)(batch > 1 ? batch_stride : (int64_t)dim * d_stride) * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_l, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = !swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE
                                : (box_l * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_l * 4 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE));
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// same, with an explicit swizzle mode
inline bool make_row_map_sw(CUtensorMap* map, const float* base, int inner, int rows, int batch, int64_t row_stride, int64_t batch_stride,
                            int box_inner, int box_rows, CUtensorMapSwizzle sw) {
  auto enc = tensor_map_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
  const cuuint64_t gstr[2] = {(cuuint64_t)row_stride * 4, (cuuint64_t)(batch > 1 ? batch_stride : (int64_t)rows * row_stride) * 4};
  const cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// batch_stride: pass any positive multiple of 4 for batch == 1; a zero (expanded) batch stride cannot be a TMA stride
inline bool tma_row_ok(const void* base, int64_t d_stride, int64_t batch_stride) {
  return (reinterpret_cast<uintptr_t>(base) & 15u) == 0 && (d_stride & 3) == 0 && (batch_stride & 3) == 0 && d_stride > 0 &&
         batch_stride > 0 && d_stride < ((int64_t)1 << 36) && batch_stride < ((int64_t)1 << 36);
}

// SM count of the current device (cached per device); 148 (B200) when no device can be queried.  Every grid-size and
// segmenting heuristic derives its CTA-slot count from this one helper.
inline int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    (void)cudaGetLastError();
    return 148;
  }
  int v = cached[dev];
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) {
    (void)cudaGetLastError();
    v = 148;
  }
  cached[dev] = v;
  return v;
}

// Opt a kernel into its dynamic shared-memory size once per device (bit `device` of `done`) instead of on every launch.
template <typename Kernel>
inline cudaError_t set_smem_once(std::atomic<unsigned long long>& done, Kernel kernel, int smem_bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cudaGetLastError();
  const unsigned long long bit = 1ull << (dev & 63);
  if (done.load(std::memory_order_acquire) & bit) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e == cudaSuccess) done.fetch_or(bit, std::memory_order_release);
  return e;
}

}  // namespace selscan
