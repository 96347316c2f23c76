// Host-visible launch interface between the C ABI (selscan_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>

#include "../../include/selscan_b200.h"

namespace selscan {

struct FwdLaunch {
  selscan_fwd_args a;
  int dim_per_group;  // dim / ngroups
  int n_ckpt;         // saved states per row = ceil(seqlen / interval) - 1
  int vec_rows;       // u/delta/z/out rows are 16-byte aligned at every multiple-of-4 position
  int vec_bc;         // B and C rows likewise (and unit position stride)
  int state_block;    // generic kernels only: states [16*state_block, 16*state_block + 16) of dstate (<= 256) per launch;
  int n_state_blocks; // block 0 starts `out` (D*u), later blocks accumulate into it, the last one applies the z gate
  float* seg_ws;      // tiled forward only: workspace of the segmented (small-batch) path or nullptr
  int n_segs;         //   number of sequence segments and tiles of 32 positions per segment (fwd_plan_segments)
  int seg_tiles;
};

struct BwdLaunch {
  selscan_bwd_args a;
  int dim_per_group;
  int n_ckpt;
  int tiles_per_group;  // ceil(dim_per_group / rows per CTA)
  int vec_rows;         // u/delta/dout/z/out/du/ddelta/dz rows 16-byte aligned at multiples of 4
  int state_block;      // generic kernel only: see FwdLaunch; block 0 stores du / ddelta / dz / dD, later blocks accumulate
  int n_state_blocks;
};

cudaError_t launch_fwd(const FwdLaunch& p, cudaStream_t stream);
// tiled TMA path (selscan_fwd_tma.cu): aligned shapes with channels-per-group % 64 == 0 (z gate included);
// launch_fwd_tma returns cudaErrorNotSupported when a tensor map cannot be encoded (caller falls back to the generic kernel)
bool fwd_tma_eligible(const FwdLaunch& p);
void fwd_plan_segments(int batch, int dim, int seqlen, int ngroups, int* n_segs, int* seg_tiles);
cudaError_t launch_fwd_tma(const FwdLaunch& p, cudaStream_t stream);
// SELSCAN_B200_GENERIC=1 in the environment forces the generic kernels (debugging / A-B timing only)
bool force_generic();
cudaError_t launch_bwd(const BwdLaunch& p, cudaStream_t stream);
// tiled path (selscan_bwd_ws.cu, warp-specialised): aligned shapes with channels-per-group % 64 == 0, seqlen > 8, no z, B and C with
// the same position stride, and a device that grants the kernel's setmaxnreg budgets; launch_bwd_ws returns
// cudaErrorNotSupported when a tensor map cannot be encoded (caller falls back to the generic kernel)
bool bwd_ws_usable();
bool bwd_ws_eligible(const BwdLaunch& p);
cudaError_t launch_bwd_ws(const BwdLaunch& p, cudaStream_t stream);

// CrossScan (scatter = true) / CrossMerge (scatter = false) plane kernels (selscan_cross.cu)
cudaError_t launch_cross(bool scatter, const float* in, float* out, int B, int D, int H, int W, int64_t pitch, cudaStream_t stream);

// SS2D edge kernels (selscan_ss2d.cu): prologue = permute + depthwise 3x3 conv + SiLU + CrossScan, epilogue = CrossMerge +
// transpose + LayerNorm + silu(z) gate, and their backwards
bool ss2d_in_supported(int H, int W);
bool ss2d_out_supported(int D);
int64_t ss2d_out_ctas(int B, int D, int H, int W);
cudaError_t launch_ss2d_in_fwd(const float* x, int64_t ld, const float* cw, const float* cb, float* xs, int B, int D, int H, int W,
                               int64_t pitch, int n_planes, cudaStream_t stream);
cudaError_t launch_ss2d_in_bwd(const float* dxs, const float* x, int64_t ld, const float* cw, const float* cb, float* dx, int64_t dld,
                               float* wpart, int B, int D, int H, int W, int64_t pitch, int n_planes, cudaStream_t stream);
cudaError_t launch_ss2d_out_fwd(const float* ys, int64_t pitch, const float* z, int64_t zld, const float* gamma, const float* beta,
                                float eps, float* out, float* xhat, float* rstd, int B, int D, int H, int W, int n_planes, cudaStream_t stream);
cudaError_t launch_ss2d_out_bwd(const float* gout, const float* z, int64_t zld, const float* xhat, const float* rstd,
                                const float* gamma, const float* beta, float* dz, int64_t dzld, float* dys, int64_t pitch,
                                float* part, int B, int D, int H, int W, int n_planes, cudaStream_t stream);

// LayerNorm over short rows (selscan_ln.cu): ln_nv(dim) == 0 means the row length is not instantiated
int ln_nv(int D);
int64_t ln_bwd_ctas(int64_t rows, int D);
cudaError_t launch_ln_fwd(const float* x, const float* w, const float* b, float eps, float* y, float* mean, float* rstd, int64_t rows,
                          int D, cudaStream_t stream);
cudaError_t launch_ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* w, float* dx, float* part,
                          int64_t rows, int D, cudaStream_t stream);

// fp32 GEMM with the 3xTF32 split on tcgen05 tensor cores (selscan_tcgemm.cu)
bool tcgemm_operand_ok(const float* p, int64_t ld, int64_t batch_stride, int batch);
cudaError_t launch_tcgemm(const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn, float* C, int64_t ldc, int M,
                          int N, int K, int batch, int64_t strideA, int64_t strideB, int64_t strideC, int accumulate, int a_bmod,
                          int b_bmod, int c_bmod, cudaStream_t stream);

}  // namespace selscan
