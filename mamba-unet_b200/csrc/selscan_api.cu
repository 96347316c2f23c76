// C ABI of libselscan_b200.so (declared in include/selscan_b200.h): argument validation + dispatch.
// Mirrors the host side of the reference extension,
// /root/reference/mamba/csrc/selective_scan/selective_scan.cpp:226-336 (fwd) and :338-492 (bwd):
// same shape / stride / presence rules, reported through return codes instead of TORCH_CHECK.
#include <cstdarg>
#include <cstdio>

#include "selscan_common.cuh"
#include "selscan_kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool m4(int64_t s) { return (s & 3) == 0; }

// shape rules shared by fwd and bwd (selective_scan.cpp:233-305, 351-447)
template <typename Args>
int check_common(const Args& a, const char* who) {
  if (a.batch < 0 || a.dim <= 0 || a.seqlen < 0) return fail("%s: bad sizes batch=%d dim=%d seqlen=%d", who, a.batch, a.dim, a.seqlen);
  if (a.dstate < 1) return fail("%s: dstate must be >= 1 (got %d)", who, a.dstate);
  if (a.dstate > 256) return fail("%s: selective_scan only supports state dimension <= 256 (got %d)", who, a.dstate);
  if (a.ngroups < 1 || a.dim % a.ngroups != 0) return fail("%s: dim (%d) must be divisible by ngroups (%d)", who, a.dim, a.ngroups);
  if (!a.u || (!a.delta && !a.dt_w) || !a.A || !a.B || !a.C) return fail("%s: u, delta (or dt_w / dt_x), A, B, C must not be NULL", who);
  if (a.mirror_pairs) {   // the odd group of every pair walks the even group's rows back to front (tiled kernels only)
    if (a.ngroups % 2 != 0) return fail("%s: mirror_pairs needs an even number of groups (got %d)", who, a.ngroups);
    if (a.z) return fail("%s: mirror_pairs is not available with z", who);
    if (!selscan_b200_mirror_ok(a.batch, a.dim, a.seqlen, a.dstate, a.ngroups))
      return fail("%s: mirror_pairs needs sizes accepted by selscan_b200_mirror_ok()", who);
  }
  if (a.dt_w) {   // fused dt_proj: the raw step is formed inside the tiled kernels
    if (!a.dt_x) return fail("%s: dt_x is required when dt_w is given", who);
    if (a.dt_rank < 1 || a.dt_rank > selscan::kMaxFusedDtRank) return fail("%s: fused dt_proj supports 1 <= dt_rank <= %d (got %d)", who, selscan::kMaxFusedDtRank, a.dt_rank);
    if (!al16(a.dt_x) || !m4(a.dt_x_batch_stride) || !m4(a.dt_x_group_stride) || !m4(a.dt_x_r_stride))
      return fail("%s: dt_x rows must be 16-byte aligned (pointer and strides multiples of 4 floats)", who);
    if (a.dt_w_d_stride < a.dt_rank) return fail("%s: dt_w_d_stride must be >= dt_rank", who);
  }
  return 0;
}

}  // namespace

extern "C" {

__attribute__((visibility("default"))) int selscan_b200_abi_version(void) { return SELSCAN_B200_ABI_VERSION; }
__attribute__((visibility("default"))) const char* selscan_b200_bwd_kernel(void) {
  return selscan::bwd_ws_usable() ? "selscan_bwd_ws_kernel" : "selscan_bwd_chunk_kernel";
}

__attribute__((visibility("default"))) const char* selscan_b200_last_error(void) { return g_err; }

__attribute__((visibility("default"))) int64_t selscan_b200_ckpt_elems(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate) {
  if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0) return 0;
  const int64_t n_ckpt = (seqlen + SELSCAN_B200_CKPT_INTERVAL - 1) / SELSCAN_B200_CKPT_INTERVAL - 1;
  const int64_t n_blocks = (dstate + SELSCAN_B200_STATE_PAD - 1) / SELSCAN_B200_STATE_PAD;   // states are processed 16 at a time
  return n_blocks * batch * dim * n_ckpt * SELSCAN_B200_STATE_PAD;
}

__attribute__((visibility("default"))) int64_t selscan_b200_fwd_workspace_elems(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate,
                                                                              int32_t ngroups) {
  if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0 || dstate > SELSCAN_B200_STATE_PAD || ngroups <= 0) return 0;
  int n_segs = 1, seg_tiles = 0;
  selscan::fwd_plan_segments(batch, dim, seqlen, ngroups, &n_segs, &seg_tiles);
  if (n_segs < 2) return 0;
  return (int64_t)batch * dim * n_segs * (SELSCAN_B200_STATE_PAD + 1);   // segment states + sums of delta
}

__attribute__((visibility("default"))) int selscan_b200_dt_fusable(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate, int32_t ngroups,
                                                                 int32_t dt_rank) {
  if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0 || ngroups <= 0 || dim % ngroups) return 0;
  if (dt_rank < 1 || dt_rank > selscan::kMaxFusedDtRank || dstate > SELSCAN_B200_STATE_PAD) return 0;
  if ((dim / ngroups) % 64 != 0 || seqlen <= SELSCAN_B200_CKPT_INTERVAL || (seqlen & 3)) return 0;
  if (selscan_b200_fwd_workspace_elems(batch, dim, seqlen, dstate, ngroups) > 0) return 0;   // small batches run segmented: not fused
  return selscan::bwd_ws_usable() && !selscan::force_generic() ? 1 : 0;
}

__attribute__((visibility("default"))) int selscan_b200_mirror_ok(int32_t batch, int32_t dim, int32_t seqlen, int32_t dstate, int32_t ngroups) {
  if (batch <= 0 || dim <= 0 || seqlen <= 0 || dstate <= 0 || ngroups <= 0 || (ngroups & 1) || dim % ngroups) return 0;
  if (dstate > SELSCAN_B200_STATE_PAD || (dim / ngroups) % 64 != 0 || seqlen <= SELSCAN_B200_CKPT_INTERVAL) return 0;
  if (seqlen & 3) return 0;   // the mirrored walk keeps its 128-bit accesses only when the partial tile is a whole number of quads
  if (selscan_b200_fwd_workspace_elems(batch, dim, seqlen, dstate, ngroups) > 0) return 0;   // small batches run segmented
  return selscan::bwd_ws_usable() && !selscan::force_generic() ? 1 : 0;
}

__attribute__((visibility("default"))) int selscan_b200_fwd(const selscan_fwd_args* args, void* stream) {
  if (!args) return fail("selscan_b200_fwd: args is NULL");
  const selscan_fwd_args& a = *args;
  if (int rc = check_common(a, "selscan_b200_fwd")) return rc;
  if (!a.out) return fail("selscan_b200_fwd: out must not be NULL");
  if (a.z && !a.out_z) return fail("selscan_b200_fwd: out_z is required when z is given");
  if (a.dt_w && (a.z || !selscan_b200_dt_fusable(a.batch, a.dim, a.seqlen, a.dstate, a.ngroups, a.dt_rank)))
    return fail("selscan_b200_fwd: dt_w / dt_x (fused dt_proj) need z == NULL and sizes accepted by selscan_b200_dt_fusable()");
  selscan::FwdLaunch p;
  p.a = a;
  p.dim_per_group = a.dim / a.ngroups;
  p.n_ckpt = (a.seqlen + selscan::kCkptInterval - 1) / selscan::kCkptInterval - 1;
  if (p.n_ckpt < 0) p.n_ckpt = 0;
  if (a.ckpt && !al16(a.ckpt)) return fail("selscan_b200_fwd: ckpt must be 16-byte aligned");
  p.vec_rows = al16(a.u) && (a.dt_w || al16(a.delta)) && al16(a.out) && m4(a.u_batch_stride) && m4(a.u_d_stride) &&
               m4(a.delta_batch_stride) && m4(a.delta_d_stride) && m4(a.out_batch_stride) && m4(a.out_d_stride) &&
               (!a.z || (al16(a.z) && al16(a.out_z) && m4(a.z_batch_stride) && m4(a.z_d_stride) &&
                         m4(a.out_z_batch_stride) && m4(a.out_z_d_stride)));
  p.vec_bc = al16(a.B) && al16(a.C) && a.B_l_stride == 1 && a.C_l_stride == 1 && m4(a.B_batch_stride) &&
             m4(a.B_group_stride) && m4(a.B_n_stride) && m4(a.C_batch_stride) && m4(a.C_group_stride) &&
             m4(a.C_n_stride);
  p.seg_ws = a.workspace;
  p.n_segs = 1;
  p.seg_tiles = 0;
  if (a.workspace && (reinterpret_cast<uintptr_t>(a.workspace) & 15u)) return fail("selscan_b200_fwd: workspace must be 16-byte aligned");
  p.n_state_blocks = (a.dstate + selscan::kStatePad - 1) / selscan::kStatePad;
  for (p.state_block = 0; p.state_block < p.n_state_blocks; ++p.state_block) {   // one launch for dstate <= 16 (Mamba-UNet)
    const cudaError_t e = selscan::launch_fwd(p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
      fail("selscan_b200_fwd: launch failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  return 0;
}

__attribute__((visibility("default"))) int selscan_b200_bwd(const selscan_bwd_args* args, void* stream) {
  if (!args) return fail("selscan_b200_bwd: args is NULL");
  const selscan_bwd_args& a = *args;
  if (int rc = check_common(a, "selscan_b200_bwd")) return rc;
  if (!a.dout || !a.du || !a.ddelta || !a.dA || !a.dB || !a.dC)
    return fail("selscan_b200_bwd: dout, du, ddelta, dA, dB, dC must not be NULL");
  if (a.z && (!a.out || !a.dz)) return fail("selscan_b200_bwd: out and dz are required when z is given");
  if (a.D && !a.dD) return fail("selscan_b200_bwd: dD is required when D is given");
  if (a.delta_bias && !a.ddelta_bias) return fail("selscan_b200_bwd: ddelta_bias is required when delta_bias is given");
  if (a.dt_w && (a.z || !selscan_b200_dt_fusable(a.batch, a.dim, a.seqlen, a.dstate, a.ngroups, a.dt_rank)))
    return fail("selscan_b200_bwd: dt_w / dt_x (fused dt_proj) need z == NULL and sizes accepted by selscan_b200_dt_fusable()");
  selscan::BwdLaunch p;
  p.a = a;
  if (!a.D) p.a.dD = nullptr;
  if (!a.delta_bias) p.a.ddelta_bias = nullptr;
  p.dim_per_group = a.dim / a.ngroups;
  p.n_ckpt = (a.seqlen + selscan::kCkptInterval - 1) / selscan::kCkptInterval - 1;
  if (p.n_ckpt < 0) p.n_ckpt = 0;
  if (p.n_ckpt > 0 && !a.ckpt) return fail("selscan_b200_bwd: ckpt (saved scan states from selscan_b200_fwd) is required when seqlen > %d", SELSCAN_B200_CKPT_INTERVAL);
  if (a.ckpt && !al16(a.ckpt)) return fail("selscan_b200_bwd: ckpt must be 16-byte aligned");
  p.tiles_per_group = (p.dim_per_group + 63) / 64;
  if (a.du_d_stride < a.seqlen || a.ddelta_d_stride < a.seqlen || (a.z && a.dz_d_stride < a.seqlen))
    return fail("selscan_b200_bwd: du / ddelta / dz channel strides must be >= seqlen");
  p.vec_rows = al16(a.u) && (a.dt_w || al16(a.delta)) && al16(a.dout) && al16(a.du) && al16(a.ddelta) &&
               m4(a.u_batch_stride) && m4(a.u_d_stride) && m4(a.delta_batch_stride) && m4(a.delta_d_stride) &&
               m4(a.dout_batch_stride) && m4(a.dout_d_stride) && m4(a.du_batch_stride) && m4(a.du_d_stride) &&
               m4(a.ddelta_batch_stride) && m4(a.ddelta_d_stride) &&
               (!a.z || (al16(a.z) && al16(a.out) && al16(a.dz) && m4(a.z_batch_stride) && m4(a.z_d_stride) &&
                         m4(a.out_batch_stride) && m4(a.out_d_stride) && m4(a.dz_batch_stride) && m4(a.dz_d_stride)));
  p.n_state_blocks = (a.dstate + selscan::kStatePad - 1) / selscan::kStatePad;
  for (p.state_block = 0; p.state_block < p.n_state_blocks; ++p.state_block) {
    const cudaError_t e = selscan::launch_bwd(p, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) {
      fail("selscan_b200_bwd: launch failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
  }
  return 0;
}

static int cross_common(bool scatter, const float* in, float* out, int32_t batch, int32_t dim, int32_t H, int32_t W, int64_t pitch, void* stream) {
  const char* who = scatter ? "selscan_b200_cross_scan" : "selscan_b200_cross_merge";
  if (!in || !out) return fail("%s: NULL pointer", who);
  if (batch < 0 || dim <= 0 || H <= 0 || W <= 0) return fail("%s: bad sizes batch=%d dim=%d H=%d W=%d", who, batch, dim, H, W);
  if (pitch < (int64_t)H * W) return fail("%s: row_pitch (%lld) must be >= H*W (%d)", who, (long long)pitch, H * W);
  if ((int64_t)H * (W + 1) * 4 > 200 * 1024) return fail("%s: image plane of %dx%d does not fit in shared memory", who, H, W);
  const cudaError_t e = selscan::launch_cross(scatter, in, out, batch, dim, H, W, pitch, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    fail("%s: launch failed: %s", who, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

__attribute__((visibility("default"))) int selscan_b200_cross_scan(const float* x, float* xs, int32_t batch, int32_t dim, int32_t H, int32_t W,
                                                                 int64_t row_pitch, void* stream) {
  return cross_common(true, x, xs, batch, dim, H, W, row_pitch, stream);
}

__attribute__((visibility("default"))) int selscan_b200_cross_merge(const float* ys, float* y, int32_t batch, int32_t dim, int32_t H, int32_t W,
                                                                  int64_t row_pitch, void* stream) {
  return cross_common(false, ys, y, batch, dim, H, W, row_pitch, stream);
}

static int edge_planes(const char* who, int32_t n_planes) {
  if (n_planes != 4 && n_planes != 2) return fail("%s: n_planes must be 4 (all scan orders) or 2 (row- and column-major only), got %d", who, n_planes);
  return 0;
}

static int edge_sizes(const char* who, int32_t batch, int32_t dim, int32_t H, int32_t W, int64_t pitch) {
  if (batch < 0 || dim <= 0 || H <= 0 || W <= 0) return fail("%s: bad sizes batch=%d dim=%d H=%d W=%d", who, batch, dim, H, W);
  if (pitch < (int64_t)H * W) return fail("%s: row_pitch (%lld) must be >= H*W (%d)", who, (long long)pitch, H * W);
  return 0;
}

static int edge_done(const char* who, cudaError_t e) {
  if (e == cudaSuccess) return 0;
  fail("%s: launch failed: %s", who, cudaGetErrorString(e));
  return (int)e;
}

#define SELSCAN_EXPORT __attribute__((visibility("default")))

SELSCAN_EXPORT int selscan_b200_ss2d_in_fwd(const float* x, int64_t x_pos_stride, const float* conv_w, const float* conv_b, float* xs,
                                            int32_t batch, int32_t dim, int32_t H, int32_t W, int64_t row_pitch, int32_t n_planes, void* stream) {
  const char* who = "selscan_b200_ss2d_in_fwd";
  if (!x || !conv_w || !xs) return fail("%s: x, conv_w, xs must not be NULL", who);
  if (edge_sizes(who, batch, dim, H, W, row_pitch) || edge_planes(who, n_planes)) return -1;
  if (x_pos_stride < dim) return fail("%s: x_pos_stride (%lld) must be >= dim (%d)", who, (long long)x_pos_stride, dim);
  if (!selscan::ss2d_in_supported(H, W)) return fail("%s: image plane of %dx%d does not fit in shared memory", who, H, W);
  return edge_done(who, selscan::launch_ss2d_in_fwd(x, x_pos_stride, conv_w, conv_b, xs, batch, dim, H, W, row_pitch,
                                                    n_planes, static_cast<cudaStream_t>(stream)));
}

SELSCAN_EXPORT int selscan_b200_ss2d_in_bwd(const float* dxs, const float* x, int64_t x_pos_stride, const float* conv_w,
                                            const float* conv_b, float* dx, int64_t dx_pos_stride, float* dconv_part, int32_t batch,
                                            int32_t dim, int32_t H, int32_t W, int64_t row_pitch, int32_t n_planes, void* stream) {
  const char* who = "selscan_b200_ss2d_in_bwd";
  if (!dxs || !x || !conv_w || !dx || !dconv_part) return fail("%s: dxs, x, conv_w, dx, dconv_part must not be NULL", who);
  if (edge_sizes(who, batch, dim, H, W, row_pitch) || edge_planes(who, n_planes)) return -1;
  if (x_pos_stride < dim || dx_pos_stride < dim) return fail("%s: position strides must be >= dim (%d)", who, dim);
  if (!selscan::ss2d_in_supported(H, W)) return fail("%s: image plane of %dx%d does not fit in shared memory", who, H, W);
  return edge_done(who, selscan::launch_ss2d_in_bwd(dxs, x, x_pos_stride, conv_w, conv_b, dx, dx_pos_stride, dconv_part, batch, dim, H,
                                                    W, row_pitch, n_planes, static_cast<cudaStream_t>(stream)));
}

SELSCAN_EXPORT int64_t selscan_b200_ss2d_out_partial_elems(int32_t batch, int32_t dim, int32_t H, int32_t W) {
  return selscan::ss2d_out_ctas(batch, dim, H, W) * 2 * (int64_t)(dim > 0 ? dim : 0);
}

SELSCAN_EXPORT int selscan_b200_ss2d_out_fwd(const float* ys, int64_t row_pitch, const float* z, int64_t z_pos_stride,
                                             const float* ln_weight, const float* ln_bias, float eps, float* out, float* xhat,
                                             float* rstd, int32_t batch, int32_t dim, int32_t H, int32_t W, int32_t n_planes, void* stream) {
  const char* who = "selscan_b200_ss2d_out_fwd";
  if (!ys || !ln_weight || !ln_bias || !out) return fail("%s: ys, ln_weight, ln_bias, out must not be NULL", who);
  if ((xhat == nullptr) != (rstd == nullptr)) return fail("%s: xhat and rstd are saved together (both or neither)", who);
  if (edge_sizes(who, batch, dim, H, W, row_pitch) || edge_planes(who, n_planes)) return -1;
  if (z && z_pos_stride < dim) return fail("%s: z_pos_stride must be >= dim (%d)", who, dim);
  if (!selscan::ss2d_out_supported(dim)) return fail("%s: dim=%d does not fit a shared-memory tile", who, dim);
  return edge_done(who, selscan::launch_ss2d_out_fwd(ys, row_pitch, z, z_pos_stride, ln_weight, ln_bias, eps, out, xhat, rstd, batch,
                                                     dim, H, W, n_planes, static_cast<cudaStream_t>(stream)));
}

SELSCAN_EXPORT int selscan_b200_ss2d_out_bwd(const float* dout, const float* z, int64_t z_pos_stride, const float* xhat,
                                             const float* rstd, const float* ln_weight, const float* ln_bias, float* dz,
                                             int64_t dz_pos_stride, float* dys, int64_t row_pitch, float* dln_part, int32_t batch,
                                             int32_t dim, int32_t H, int32_t W, int32_t n_planes, void* stream) {
  const char* who = "selscan_b200_ss2d_out_bwd";
  if (!dout || !xhat || !rstd || !ln_weight || !ln_bias || !dys || !dln_part)
    return fail("%s: dout, xhat, rstd, ln_weight, ln_bias, dys, dln_part must not be NULL", who);
  if (z && !dz) return fail("%s: dz is required when z is given", who);
  if (edge_sizes(who, batch, dim, H, W, row_pitch) || edge_planes(who, n_planes)) return -1;
  if (z && (z_pos_stride < dim || dz_pos_stride < dim)) return fail("%s: position strides must be >= dim (%d)", who, dim);
  if (!selscan::ss2d_out_supported(dim)) return fail("%s: dim=%d does not fit a shared-memory tile", who, dim);
  return edge_done(who, selscan::launch_ss2d_out_bwd(dout, z, z_pos_stride, xhat, rstd, ln_weight, ln_bias, dz, dz_pos_stride, dys,
                                                     row_pitch, dln_part, batch, dim, H, W, n_planes, static_cast<cudaStream_t>(stream)));
}

SELSCAN_EXPORT int selscan_b200_layernorm_supported(int32_t dim) { return dim > 0 && selscan::ln_nv(dim) != 0; }

SELSCAN_EXPORT int64_t selscan_b200_layernorm_partial_elems(int64_t rows, int32_t dim) {
  return dim > 0 ? selscan::ln_bwd_ctas(rows, dim) * 2 * (int64_t)dim : 0;
}

SELSCAN_EXPORT int selscan_b200_layernorm_fwd(const float* x, const float* weight, const float* bias, float eps, float* y, float* mean,
                                              float* rstd, int64_t rows, int32_t dim, void* stream) {
  const char* who = "selscan_b200_layernorm_fwd";
  if (!x || !weight || !bias || !y) return fail("%s: x, weight, bias, y must not be NULL", who);
  if ((mean == nullptr) != (rstd == nullptr)) return fail("%s: mean and rstd are saved together (both or neither)", who);
  if (rows < 0 || dim <= 0) return fail("%s: bad sizes rows=%lld dim=%d", who, (long long)rows, dim);
  if (selscan::ln_nv(dim) == 0) return fail("%s: dim=%d is larger than the kernels are instantiated for (1536)", who, dim);
  return edge_done(who, selscan::launch_ln_fwd(x, weight, bias, eps, y, mean, rstd, rows, dim, static_cast<cudaStream_t>(stream)));
}

SELSCAN_EXPORT int selscan_b200_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* weight,
                                              float* dx, float* dwb_part, int64_t rows, int32_t dim, void* stream) {
  const char* who = "selscan_b200_layernorm_bwd";
  if (!dy || !x || !mean || !rstd || !weight || !dx || !dwb_part)
    return fail("%s: dy, x, mean, rstd, weight, dx, dwb_part must not be NULL", who);
  if (rows < 0 || dim <= 0) return fail("%s: bad sizes rows=%lld dim=%d", who, (long long)rows, dim);
  if (selscan::ln_nv(dim) == 0) return fail("%s: dim=%d is larger than the kernels are instantiated for (1536)", who, dim);
  return edge_done(who, selscan::launch_ln_bwd(dy, x, mean, rstd, weight, dx, dwb_part, rows, dim, static_cast<cudaStream_t>(stream)));
}

SELSCAN_EXPORT int selscan_b200_gemm_3xtf32(const float* A, int64_t lda, int32_t a_mn_major, const float* B, int64_t ldb,
                                            int32_t b_mn_major, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t batch,
                                            int64_t strideA, int64_t strideB, int64_t strideC, int32_t accumulate,
                                            int32_t a_batch_mod, int32_t b_batch_mod, int32_t c_batch_mod, void* stream) {
  const char* who = "selscan_b200_gemm_3xtf32";
  if (!A || !B || !C) return fail("%s: A, B, C must not be NULL", who);
  if (M < 0 || N < 0 || K < 0 || batch < 0) return fail("%s: bad sizes M=%d N=%d K=%d batch=%d", who, M, N, K, batch);
  if (a_batch_mod < 0 || b_batch_mod < 0 || c_batch_mod < 0) return fail("%s: batch moduli must be >= 0", who);
  if (!selscan::tcgemm_operand_ok(A, lda, strideA, a_batch_mod ? a_batch_mod : batch) ||
      !selscan::tcgemm_operand_ok(B, ldb, strideB, b_batch_mod ? b_batch_mod : batch) ||
      !selscan::tcgemm_operand_ok(C, ldc, strideC, c_batch_mod ? c_batch_mod : batch))
    return fail("%s: A, B and C must be 16-byte aligned with row / batch strides that are multiples of 4 floats", who);
  if (lda < (a_mn_major ? M : K) || ldb < (b_mn_major ? N : K) || ldc < N)
    return fail("%s: a leading dimension is smaller than the row it strides over", who);
  return edge_done(who, selscan::launch_tcgemm(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, batch, strideA, strideB, strideC,
                                               accumulate, a_batch_mod, b_batch_mod, c_batch_mod, static_cast<cudaStream_t>(stream)));
}

}  // extern "C"
