// Link scaffolding for oracle/build_ref.py -- OUR code, not the reference's.
// The reference host dispatcher (selective_scan.cpp:14-51,328-332,483-487) names every (input_t, weight_t)
// instantiation; only <float, float> is compiled from the reference sources.  The rest are defined here so the
// extension links, and refuse to run.
#include <c10/util/BFloat16.h>
#include <c10/util/Half.h>
#include <c10/util/complex.h>
#include <cuda_runtime.h>
#include <torch/extension.h>

#include "selective_scan.h"

using complex_t = c10::complex<float>;

template <typename input_t, typename weight_t>
void selective_scan_fwd_cuda(SSMParamsBase& params, cudaStream_t stream);
template <typename input_t, typename weight_t>
void selective_scan_bwd_cuda(SSMParamsBwd& params, cudaStream_t stream);

#define STUB_FWD(I, W)                                                                    \
  template <>                                                                             \
  void selective_scan_fwd_cuda<I, W>(SSMParamsBase&, cudaStream_t) {                      \
    TORCH_CHECK(false, "oracle/_ref build: this dtype combination was not compiled");     \
  }
#define STUB_BWD(I, W)                                                                    \
  template <>                                                                             \
  void selective_scan_bwd_cuda<I, W>(SSMParamsBwd&, cudaStream_t) {                       \
    TORCH_CHECK(false, "oracle/_ref build: this dtype combination was not compiled");     \
  }

STUB_FWD(at::Half, float)
STUB_FWD(at::Half, complex_t)
STUB_FWD(at::BFloat16, float)
STUB_FWD(at::BFloat16, complex_t)
STUB_BWD(float, complex_t)
STUB_BWD(at::Half, float)
STUB_BWD(at::Half, complex_t)
STUB_BWD(at::BFloat16, float)
STUB_BWD(at::BFloat16, complex_t)
