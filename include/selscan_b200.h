/*
 * selscan_b200.h -- C ABI of the B200-native selective scan (libselscan_b200.so).
 *
 * This is the drop-in boundary for the one hot path of Grozta/Mamba-UNet:
 *   SS2D.forward_core            code/networks/mamba_sys.py:396-436
 *   -> selective_scan_fn         mamba/mamba_ssm/ops/selective_scan_interface.py:77-83
 *   -> selective_scan_cuda.fwd   mamba/csrc/selective_scan/selective_scan.cpp:226-336
 *   -> selective_scan_cuda.bwd   mamba/csrc/selective_scan/selective_scan.cpp:338-492
 * The entry points below are what the reference's pybind module (selective_scan.cpp:494-497) binds, with
 * ATen tensors replaced by raw device pointers + sizes + element strides (the reference's own
 * SSMParamsBase / SSMParamsBwd, selective_scan.h:26-101, minus the torch types).
 *
 * Contract
 *  - Every pointer is a DEVICE pointer on the current CUDA device.  The caller allocates every output
 *    and scratch buffer and owns it; the library never allocates, frees or caches device memory.
 *  - Launches go to `stream` (a cudaStream_t passed as void*); nothing synchronises, nothing changes the
 *    current device.  Re-entrant: safe from the main thread and autograd worker threads at once.
 *  - Return value: 0 ok; < 0 invalid argument (message via selscan_b200_last_error(), thread local);
 *    > 0 the cudaError_t of a failed launch.
 *  - dtype: fp32 real (u, delta, A, B, C, D, z, delta_bias).  B and C are "variable" (input dependent):
 *    (batch, ngroups, dstate, seqlen), any element strides; channel d uses group d / (dim / ngroups)
 *    (selective_scan_fwd_kernel.cuh:99).  dstate <= 256 as in the reference; <= 16 (Mamba-UNet: 16) is one launch on the tiled kernels, larger
 *    state counts run 16 states per launch on the generic kernels.  fp16/bf16 I/O and constant (dim, dstate)
 *    B/C are handled by the host shim (widening; zero-stride broadcast views); complex A is rejected with a message.
 *  - u, delta, z, out, dout, du, ddelta, dz have unit stride along seqlen (the reference requires the
 *    same, selective_scan.cpp:252-253); batch/channel strides are free.
 */
#ifndef SELSCAN_B200_H_
#define SELSCAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SELSCAN_B200_ABI_VERSION 7
/* distance (in sequence positions) between two saved scan states; also the backward's chunk length */
#define SELSCAN_B200_CKPT_INTERVAL 8
/* states are padded to this count inside the kernels and in the checkpoint buffer */
#define SELSCAN_B200_STATE_PAD 16

/* replaces selective_scan_cuda.fwd(u, delta, A, B, C, D_, z_, delta_bias_, delta_softplus)
 * (selective_scan.cpp:226-232) */
typedef struct selscan_fwd_args {
  int32_t batch, dim, seqlen, dstate, ngroups;
  int32_t delta_softplus;            /* 1: delta = softplus(delta + delta_bias), threshold 20 */
  const float* u;                    /* (batch, dim, seqlen) */
  const float* delta;                /* (batch, dim, seqlen) */
  const float* A;                    /* (dim, dstate) */
  const float* B;                    /* (batch, ngroups, dstate, seqlen) */
  const float* C;                    /* (batch, ngroups, dstate, seqlen) */
  const float* D;                    /* (dim) or NULL */
  const float* z;                    /* (batch, dim, seqlen) or NULL */
  const float* delta_bias;           /* (dim) or NULL */
  int64_t u_batch_stride, u_d_stride;
  int64_t delta_batch_stride, delta_d_stride;
  int64_t A_d_stride, A_n_stride;
  int64_t B_batch_stride, B_group_stride, B_n_stride, B_l_stride;
  int64_t C_batch_stride, C_group_stride, C_n_stride, C_l_stride;
  int64_t z_batch_stride, z_d_stride;
  float* out;                        /* (batch, dim, seqlen): y + D*u, NOT gated */
  int64_t out_batch_stride, out_d_stride;
  float* out_z;                      /* (batch, dim, seqlen): out * silu(z); required iff z != NULL */
  int64_t out_z_batch_stride, out_z_d_stride;
  float* last_state;                 /* (batch, dim, dstate) contiguous, or NULL */
  float* ckpt;                       /* selscan_b200_ckpt_elems() floats, or NULL (inference) */
  float* workspace;                  /* selscan_b200_fwd_workspace_elems() floats or NULL: lets small-batch calls split the
                                        sequence into segments that run concurrently (results identical to rounding) */
  /* Fused dt_proj (code/networks/mamba_sys.py:408-412).  When dt_w != NULL, `delta` is ignored (may be
   * NULL: no (batch, dim, seqlen) step tensor exists) and the kernels form the raw step themselves,
   *     delta[b, d, l] = sum_r dt_w[d, r] * dt_x[b, d / (dim / ngroups), r, l],
   * before the bias / softplus.  Only the tiled kernels do this: ask selscan_b200_dt_fusable() first. */
  const float* dt_w;                 /* (dim, dt_rank), unit stride along r, or NULL */
  const float* dt_x;                 /* (batch, ngroups, dt_rank, seqlen), unit stride along seqlen, 16-byte aligned rows */
  int64_t dt_w_d_stride;
  int64_t dt_x_batch_stride, dt_x_group_stride, dt_x_r_stride;
  int32_t dt_rank;
  /* Mirrored direction pairs (the reversed half of SS2D's CrossScan / CrossMerge inside the scan kernels; code/networks/mamba_sys.py:
   * 404, 429).  When mirror_pairs != 0 the groups come in pairs (2j, 2j+1): the odd group scans the SAME rows of `u` as the even
   * one, back to front (scan position l <-> source index seqlen-1-l).  `u` and `out` then have dim/2 rows per batch (row of channel c
   * of group g: (g / 2) * (dim / ngroups) + c), `out` must be ZERO-INITIALISED and receives the SUM of both groups of a pair, each in
   * source order (out_pair[l] = y_even[l] + y_odd[seqlen-1-l]); delta / dt_x, B and C stay per group, in SOURCE order (the kernels
   * read them back to front for odd groups).  Needs ngroups even, z == NULL, the tiled kernels: ask selscan_b200_mirror_ok(). */
  int32_t mirror_pairs;
} selscan_fwd_args;

/* replaces selective_scan_cuda.bwd(u, delta, A, B, C, D_, z_, delta_bias_, dout, x_, out_, dz_,
 * delta_softplus, recompute_out_z) (selective_scan.cpp:338-349) */
typedef struct selscan_bwd_args {
  int32_t batch, dim, seqlen, dstate, ngroups;
  int32_t delta_softplus;
  const float* u;
  const float* delta;
  const float* A;
  const float* B;
  const float* C;
  const float* D;                    /* or NULL */
  const float* z;                    /* or NULL */
  const float* delta_bias;           /* or NULL */
  const float* dout;                 /* (batch, dim, seqlen) */
  const float* out;                  /* ungated forward output; required iff z != NULL */
  const float* ckpt;                 /* written by selscan_b200_fwd on the same inputs */
  int64_t u_batch_stride, u_d_stride;
  int64_t delta_batch_stride, delta_d_stride;
  int64_t A_d_stride, A_n_stride;
  int64_t B_batch_stride, B_group_stride, B_n_stride, B_l_stride;
  int64_t C_batch_stride, C_group_stride, C_n_stride, C_l_stride;
  int64_t z_batch_stride, z_d_stride;
  int64_t dout_batch_stride, dout_d_stride;
  int64_t out_batch_stride, out_d_stride;
  int64_t du_batch_stride, du_d_stride;          /* unit stride along seqlen; a row pitch that is a multiple of 4 */
  int64_t ddelta_batch_stride, ddelta_d_stride;  /* floats keeps rows 16-byte aligned for uneven seqlen (e.g. 49 -> 52) */
  int64_t dz_batch_stride, dz_d_stride;
  float* du;                         /* (batch, dim, seqlen), fully written */
  float* ddelta;                     /* (batch, dim, seqlen), fully written */
  float* dz;                         /* (batch, dim, seqlen); required iff z != NULL */
  float* dA;                         /* (dim, dstate)                  contiguous, ZERO-INITIALISED by caller */
  float* dB;                         /* (batch, ngroups, dstate, seqlen) contiguous, ZERO-INITIALISED */
  float* dC;                         /* (batch, ngroups, dstate, seqlen) contiguous, ZERO-INITIALISED */
  float* dD;                         /* (dim) ZERO-INITIALISED; required iff D != NULL */
  float* ddelta_bias;                /* (dim) ZERO-INITIALISED; required iff delta_bias != NULL */
  /* Fused dt_proj (code/networks/mamba_sys.py:408-412).  When dt_w != NULL, `delta` is ignored (may be
   * NULL: no (batch, dim, seqlen) step tensor exists) and the kernels form the raw step themselves,
   *     delta[b, d, l] = sum_r dt_w[d, r] * dt_x[b, d / (dim / ngroups), r, l],
   * before the bias / softplus; `ddelta` is still written (the gradient w.r.t. that raw step).  Only the tiled kernels do this: ask selscan_b200_dt_fusable() first. */
  const float* dt_w;                 /* (dim, dt_rank), unit stride along r, or NULL */
  const float* dt_x;                 /* (batch, ngroups, dt_rank, seqlen), unit stride along seqlen, 16-byte aligned rows */
  int64_t dt_w_d_stride;
  int64_t dt_x_batch_stride, dt_x_group_stride, dt_x_r_stride;
  int32_t dt_rank;
  /* Mirrored direction pairs (the reversed half of SS2D's CrossScan / CrossMerge inside the scan kernels; code/networks/mamba_sys.py:
   * 404, 429).  When mirror_pairs != 0 the groups come in pairs (2j, 2j+1): the odd group scans the SAME rows of `u` as the even
   * one, back to front (scan position l <-> source index seqlen-1-l).  `u`, `dout` and `du` then have dim/2 rows per batch (row of channel c
   * of group g: (g / 2) * (dim / ngroups) + c), `du` must be ZERO-INITIALISED and receives the SUM of both groups' gradients
   * in source order; ddelta, dB, dC are per group, in source order; delta / dt_x, B and C stay per group, in SOURCE order (the kernels
   * read them back to front for odd groups).  Needs ngroups even, z == NULL, the tiled kernels: ask selscan_b200_mirror_ok(). */
  int32_t mirror_pairs;
} selscan_bwd_args;

int selscan_b200_abi_version(void);
/* name of the kernel aligned backward calls launch on the current device: "selscan_bwd_ws_kernel" (warp-specialised tiled kernel,
 * the default) or "selscan_bwd_chunk_kernel" (the generic kernel, when the device cannot grant the former's register split).
 * Needs a GPU. */
const char* selscan_b200_bwd_kernel(void);
const char* selscan_b200_last_error(void);

/* number of floats the `ckpt` scratch of one (batch, dim, seqlen, dstate) problem needs (may be 0):
 * ceil(dstate / 16) * batch * dim * (ceil(seqlen / 8) - 1) * 16 */
int64_t selscan_b200_ckpt_elems(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate);

/* 1 when a call of these sizes can take dt_w / dt_x (fused dt_proj): dt_rank <= 12, dstate <= 16, channels per group a multiple of
 * 64, seqlen > 8 and a multiple of 4, enough work to run unsegmented, and a device that runs the tiled kernels; else 0 */
int selscan_b200_dt_fusable(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate, int32_t ngroups, int32_t dt_rank);

/* 1 when a call of these sizes can use mirror_pairs: ngroups even, dstate <= 16, channels per group a multiple of 64, seqlen > 8 and
 * a multiple of 4, enough work to run unsegmented, and a device that runs the tiled kernels; else 0 */
int selscan_b200_mirror_ok(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate, int32_t ngroups);

/* floats of forward `workspace` that make the segmented small-batch path available; 0 when the call fills the chip anyway */
int64_t selscan_b200_fwd_workspace_elems(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate, int32_t ngroups);

int selscan_b200_fwd(const selscan_fwd_args* args, void* stream);
int selscan_b200_bwd(const selscan_bwd_args* args, void* stream);

/* The 4-direction CrossScan / CrossMerge of SS2D (code/networks/mamba_sys.py:403-404 and :429-432), one pass each.
 *   cross_scan : x  (batch, dim, H, W) contiguous -> xs (batch, 4, dim, H*W) with row pitch `row_pitch` floats
 *                k=0 row-major, k=1 column-major, k=2 / k=3 the same two reversed
 *   cross_merge: ys (batch, 4, dim, H*W), pitch `row_pitch` -> y (batch, dim, H*W) contiguous, row-major
 * Each is the other's backward.  row_pitch >= H*W; the (batch, 4, dim) dimensions are dense on top of it. */
int selscan_b200_cross_scan(const float* x, float* xs, int32_t batch, int32_t dim, int32_t H, int32_t W, int64_t row_pitch, void* stream);
int selscan_b200_cross_merge(const float* ys, float* y, int32_t batch, int32_t dim, int32_t H, int32_t W, int64_t row_pitch, void* stream);

/* The two edges of SS2D.forward around the scan (code/networks/mamba_sys.py:527-540), one kernel per direction of autograd.
 * "channels-last" tensors are (batch, H, W, channels) with a free POSITION stride (floats between two pixels), so the x and z
 * halves of in_proj's output (mamba_sys.py:530-531, `xz.chunk(2, dim=-1)`) are read, and their gradients written, in place.
 *
 * n_planes = 4: xs / ys / dxs / dys hold all four scan orders, (batch, 4, dim, H*W): row-major, column-major and both reversed.
 * n_planes = 2: only the row-major and the column-major plane, (batch, 2, dim, H*W) -- the reversed orders are then walked by the scan
 *              kernels themselves (mirror_pairs above): ys is the per-pair sum they accumulated, dys / xs are read by both groups.
 * ss2d_in_fwd : x (batch,H,W,dim) channels-last -> permute (:533) -> depthwise 3x3 conv, padding 1, + bias (:534; conv_w is
 *               (dim,1,3,3) contiguous, conv_b (dim) or NULL) -> SiLU (:534) -> CrossScan (:403-404) -> xs (batch,4,dim,H*W), row
 *               pitch `row_pitch` floats (k=0 row-major, 1 column-major, 2/3 reversed)
 * ss2d_in_bwd : dxs (layout of xs) -> dx (channels-last, stride dx_pos_stride); dconv_part (batch, dim, 10): per-image sums for
 *               d conv_w (9 taps) and d conv_b, summed over batch by the caller
 * ss2d_out_fwd: ys (batch,4,dim,H*W) scan outputs -> CrossMerge (:429-432) -> transpose (:433) -> LayerNorm over dim (:434,
 *               ln_weight, ln_bias, eps) -> * silu(z) (:536; z channels-last or NULL = no gate) -> out (batch,H,W,dim) contiguous.
 *               xhat (batch*H*W, dim) and rstd (batch*H*W) are saved for the backward when both are non-NULL.
 * ss2d_out_bwd: dout (batch,H,W,dim) contiguous -> dz (channels-last, stride dz_pos_stride; iff z), dys (layout of ys), and
 *               dln_part: selscan_b200_ss2d_out_partial_elems() floats = (n_tiles, 2, dim) per-tile sums for d ln_weight, d ln_bias */
int selscan_b200_ss2d_in_fwd(const float* x, int64_t x_pos_stride, const float* conv_w, const float* conv_b, float* xs,
                             int32_t batch, int32_t dim, int32_t H, int32_t W, int64_t row_pitch, int32_t n_planes, void* stream);
int selscan_b200_ss2d_in_bwd(const float* dxs, const float* x, int64_t x_pos_stride, const float* conv_w, const float* conv_b,
                             float* dx, int64_t dx_pos_stride, float* dconv_part, int32_t batch, int32_t dim, int32_t H, int32_t W,
                             int64_t row_pitch, int32_t n_planes, void* stream);
int64_t selscan_b200_ss2d_out_partial_elems(int32_t batch, int32_t dim, int32_t H, int32_t W);
int selscan_b200_ss2d_out_fwd(const float* ys, int64_t row_pitch, const float* z, int64_t z_pos_stride, const float* ln_weight,
                              const float* ln_bias, float eps, float* out, float* xhat, float* rstd, int32_t batch, int32_t dim,
                              int32_t H, int32_t W, int32_t n_planes, void* stream);
int selscan_b200_ss2d_out_bwd(const float* dout, const float* z, int64_t z_pos_stride, const float* xhat, const float* rstd,
                              const float* ln_weight, const float* ln_bias, float* dz, int64_t dz_pos_stride, float* dys,
                              int64_t row_pitch, float* dln_part, int32_t batch, int32_t dim, int32_t H, int32_t W, int32_t n_planes, void* stream);

/* LayerNorm over the last dimension of a contiguous (rows, dim) fp32 tensor: the op on the caller side of every SS2D block
 * (VSSBlock.ln_1, code/networks/mamba_sys.py:552,559; PatchMerging2D.norm :205; PatchExpand.norm :242; VSSM.norm / norm_up).
 * Biased variance, y = (x - mean) * rsqrt(var + eps) * weight + bias, as torch.nn.LayerNorm.  dim <= 1536
 * (selscan_b200_layernorm_supported).  mean / rstd (rows each) are saved for the backward when non-NULL.
 * bwd: dx fully written; dwb_part = selscan_b200_layernorm_partial_elems() floats = (n_ctas, 2, dim) partial sums of
 * d weight and d bias, summed over n_ctas by the caller. */
int selscan_b200_layernorm_supported(int32_t dim);
int64_t selscan_b200_layernorm_partial_elems(int64_t rows, int32_t dim);
int selscan_b200_layernorm_fwd(const float* x, const float* weight, const float* bias, float eps, float* y, float* mean, float* rstd,
                               int64_t rows, int32_t dim, void* stream);
int selscan_b200_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* weight, float* dx,
                               float* dwb_part, int64_t rows, int32_t dim, void* stream);

/* fp32 GEMM on the tcgen05 tensor cores with the 3xTF32 split (a = a_hi + a_lo in TF32; a_lo*b_hi + a_hi*b_lo + a_hi*b_hi with
 * fp32 accumulation): agrees with an fp32 SIMT GEMM to ~1e-6 relative.  OPT-IN replacement for the cuBLAS fp32 GEMMs around the
 * scan (SS2D.in_proj / out_proj / x_proj, code/networks/mamba_sys.py:299,336,406) and their dgrad / wgrad forms.
 *   C[b] (M x N, row stride ldc) (+)= A[b] (M x K) * B[b] (N x K)^T          b = 0 .. batch-1, strides strideA/B/C floats
 *   A: a_mn_major = 0: stored [M][K] (row stride lda, K contiguous); 1: stored [K][M] (row stride lda, M contiguous)
 *   B: b_mn_major = 0: stored [N][K];                                1: stored [K][N]
 *   accumulate = 1: C += ...   A, B, C 16-byte aligned, lda / ldb / ldc / batch strides multiples of 4 floats (TMA).
 *   x_batch_mod > 0: operand x uses batch entry (b % x_batch_mod) -- weights shared by all images of a batch (a/b), or one C
 *   that sums over images (c: the weight gradients of x_proj / dt_proj); 0: entry b. */
int selscan_b200_gemm_3xtf32(const float* A, int64_t lda, int32_t a_mn_major, const float* B, int64_t ldb, int32_t b_mn_major,
                             float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t batch, int64_t strideA,
                             int64_t strideB, int64_t strideC, int32_t accumulate, int32_t a_batch_mod, int32_t b_batch_mod,
                             int32_t c_batch_mod, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SELSCAN_B200_H_ */
