// Forward selective scan for sm_100a -- replaces selective_scan_fwd_kernel
// (/root/reference/mamba/csrc/selective_scan/selective_scan_fwd_kernel.cuh:67-303) behind the C ABI.
//
// Mapping (v1, "row-serial"): one thread owns one (batch, channel) row and walks the sequence with all
// dstate (<= 16) states in registers, so the recurrence x = a*x + b needs no cross-thread scan at all:
// per (position, state) it costs FMUL (delta*A2), MUFU.EX2, FMUL (delta*u*B), FFMA (state), FFMA (y) --
// versus the reference's generic (a, b)-pair block scan (selective_scan_common.h:110-144) which also
// carries the running product.  The 32 lanes of a warp are 32 consecutive channels of one group, so the
// B/C loads are warp-uniform (one sector, broadcast).  Every kCkptInterval positions the state is saved
// for the backward kernel (the role of the reference's per-chunk `x` scratch, selective_scan.cpp:313).
#include <cstdlib>

#include "selscan_common.cuh"
#include "selscan_kernels.h"

namespace selscan {

constexpr int kFwdThreads = 128;

template <bool kHasZ>
__global__ void __launch_bounds__(kFwdThreads) selscan_fwd_rowserial_kernel(const FwdLaunch p) {
  const selscan_fwd_args& a = p.a;
  const int row = blockIdx.x * kFwdThreads + threadIdx.x;
  if (row >= a.batch * a.dim) return;
  const int b = row / a.dim;
  const int d = row - b * a.dim;
  const int g = d / p.dim_per_group;
  const int L = a.seqlen;
  const int N = a.dstate;
  const int n0 = p.state_block * kStatePad;   // first state of this launch's block
  const bool first_blk = p.state_block == 0, last_blk = p.state_block == p.n_state_blocks - 1;

  const float* __restrict__ u = a.u + (int64_t)b * a.u_batch_stride + (int64_t)d * a.u_d_stride;
  const float* __restrict__ dt = a.delta + (int64_t)b * a.delta_batch_stride + (int64_t)d * a.delta_d_stride;
  const float* __restrict__ Bg = a.B + (int64_t)b * a.B_batch_stride + (int64_t)g * a.B_group_stride;
  const float* __restrict__ Cg = a.C + (int64_t)b * a.C_batch_stride + (int64_t)g * a.C_group_stride;
  float* out = a.out + (int64_t)b * a.out_batch_stride + (int64_t)d * a.out_d_stride;
  const float* __restrict__ z = nullptr;
  float* __restrict__ out_z = nullptr;
  if (kHasZ) {
    z = a.z + (int64_t)b * a.z_batch_stride + (int64_t)d * a.z_d_stride;
    out_z = a.out_z + (int64_t)b * a.out_z_batch_stride + (int64_t)d * a.out_z_d_stride;
  }

  float A2[kStatePad], x[kStatePad];
#pragma unroll
  for (int n = 0; n < kStatePad; ++n) {
    A2[n] = (n0 + n < N) ? __ldg(a.A + (int64_t)d * a.A_d_stride + (int64_t)(n0 + n) * a.A_n_stride) * kLog2e : 0.f;
    x[n] = 0.f;
  }
  const float Dv = a.D ? __ldg(a.D + d) : 0.f;
  const float bias = a.delta_bias ? __ldg(a.delta_bias + d) : 0.f;
  const bool softplus = a.delta_softplus != 0;
  const bool vr = p.vec_rows != 0, vb = p.vec_bc != 0;
  float* __restrict__ ck = a.ckpt ? a.ckpt + ((int64_t)p.state_block * a.batch * a.dim + row) * p.n_ckpt * kStatePad : nullptr;

  for (int l0 = 0; l0 < L; l0 += 4) {
    float uv[4], dl[4], du[4], y[4], acc[4];
    load_row4(u, l0, L, vr, uv);
    load_row4(dt, l0, L, vr, dl);
    if (!first_blk) load_row4(out, l0, L, vr, acc);   // partial sum over the previous state blocks
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float v = dl[j] + bias;
      if (softplus) v = softplus20(v);
      v = (l0 + j < L) ? v : 0.f;  // past the end: a = 1, b = 0, the state is carried unchanged
      dl[j] = v;
      du[j] = v * uv[j];
      y[j] = first_blk ? Dv * uv[j] : acc[j];
    }
#pragma unroll
    for (int n = 0; n < kStatePad; ++n) {
      float Bv[4], Cv[4];
      if (n0 + n < N) {
        load_bc4(Bg + (int64_t)(n0 + n) * a.B_n_stride, a.B_l_stride, l0, L, vb, Bv);
        load_bc4(Cg + (int64_t)(n0 + n) * a.C_n_stride, a.C_l_stride, l0, L, vb, Cv);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) Bv[j] = Cv[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float e = ex2(dl[j] * A2[n]);
        x[n] = fmaf(e, x[n], du[j] * Bv[j]);
        y[j] = fmaf(Cv[j], x[n], y[j]);
      }
    }
    store_row4(out, l0, L, vr, y);
    if (kHasZ && last_blk) {
      float zv[4];
      load_row4(z, l0, L, vr, zv);
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = y[j] * zv[j] * sigmoidf_fast(zv[j]);
      store_row4(out_z, l0, L, vr, y);
    }
    // state after position l0+3; saved when that closes an interval that is not the last one
    const int done = l0 + 4;
    if (ck != nullptr && (done % kCkptInterval) == 0 && done < L) {
      float4* dst = reinterpret_cast<float4*>(ck + (int64_t)(done / kCkptInterval - 1) * kStatePad);
#pragma unroll
      for (int q = 0; q < kStatePad / 4; ++q) dst[q] = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
    }
  }
  if (a.last_state != nullptr) {
#pragma unroll
    for (int n = 0; n < kStatePad; ++n)
      if (n0 + n < N) a.last_state[(int64_t)row * N + n0 + n] = x[n];
  }
}

bool force_generic() {
  const char* e = getenv("SELSCAN_B200_GENERIC");
  return e != nullptr && e[0] == '1';
}

cudaError_t launch_fwd(const FwdLaunch& p, cudaStream_t stream) {
  const int64_t rows = (int64_t)p.a.batch * p.a.dim;
  if (rows == 0 || p.a.seqlen == 0) return cudaSuccess;
  if (!force_generic() && fwd_tma_eligible(p)) {
    const cudaError_t e = launch_fwd_tma(p, stream);
    if (e != cudaErrorNotSupported) return e;   // a tensor map could not be encoded for this layout: generic kernel below
  }
  if (p.a.dt_w != nullptr || p.a.mirror_pairs) return cudaErrorNotSupported;   // fused dt_proj / mirrored pairs: tiled kernels only
  const unsigned grid = (unsigned)((rows + kFwdThreads - 1) / kFwdThreads);
  if (p.a.z != nullptr)
    selscan_fwd_rowserial_kernel<true><<<grid, kFwdThreads, 0, stream>>>(p);
  else
    selscan_fwd_rowserial_kernel<false><<<grid, kFwdThreads, 0, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace selscan
