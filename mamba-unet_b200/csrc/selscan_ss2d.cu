// The two edges of SS2D around the scan, each as ONE kernel per direction of autograd (sm_100a).
//
// Reference (code/networks/mamba_sys.py), per SS2D.forward call:
//   prologue  :533-534 + :403-404   x.permute(0,3,1,2).contiguous() -> depthwise 3x3 conv + bias -> SiLU -> CrossScan
//                                   (stack + transpose.contiguous + flip + cat): ~8 ATen kernels, ~11 passes over (B, D, L)
//   epilogue  :429-434 + :536-537   CrossMerge (flip, 2 x transpose.contiguous, 3 adds) -> transpose(1,2).contiguous()
//                                   -> LayerNorm(d_inner) -> * silu(z): ~10 kernels, ~14 passes
// and roughly twice that in their backward.  Here:
//   ss2d_in_fwd   reads the x half of in_proj's output IN PLACE (channels-last, strided); one CTA per (image, group of
//                 4 / 8 channels) stages the zero-padded planes in shared memory, evaluates conv + SiLU once (two rows per
//                 thread), writes the row-major orders from registers and the column-major orders through a second plane.
//   ss2d_in_bwd   gathers the four incoming gradients into the plane, recomputes the pre-activation, applies SiLU', the
//                 transposed conv, and writes dx channels-last into its half of d(xz); per-image conv weight/bias partials.
//   ss2d_out_fwd  one CTA per (image, TH x TW tile of positions) x ALL channels: merges the four scan outputs through a
//                 shared tile [position][channel], LayerNorm per position (two-pass moments), gate, channels-last store.
//   ss2d_out_bwd  gate', LayerNorm backward, per-CTA gamma/beta partials, dz into its half of d(xz), and the four
//                 directional gradients written straight in scan layout.
// Tile runs along a scan direction are TW (or TH) floats = one 32-byte sector at 8, so the strided side of every
// transpose still moves whole sectors.  These kernels are latency-bound unless many loads are in flight: every global
// phase is unrolled so that a thread issues 4-8 independent loads before it uses the first.
#include <cuda_runtime.h>
#include <stdint.h>

#include "selscan_kernels.h"

namespace selscan {

namespace {

constexpr int kEdgeThreads = 256;
constexpr int kEdgeWarps = kEdgeThreads / 32;
constexpr int kInThreads = 512;                // prologue kernels: two big-plane CTAs per SM, so more threads per CTA
constexpr int kInWarps = kInThreads / 32;

// division by a runtime constant as one IMAD.HI (exact while p * d < 2^32); m == 0 encodes d == 1
struct FastDiv {
  int d;
  unsigned m;
};
FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  f.m = d > 1 ? (unsigned)(((1ull << 32) + (unsigned)d - 1) / (unsigned)d) : 0u;
  return f;
}
__device__ __forceinline__ int fdiv(int p, FastDiv f) { return f.m ? (int)__umulhi((unsigned)p, f.m) : p; }

// 1 / (1 + e^-v): ex2.approx + rcp.approx (rel. error ~2e-7, the level at which torch's own CPU and CUDA silu differ)
__device__ __forceinline__ float sigmoid_fast(float v) { return __fdividef(1.f, 1.f + __expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// prologue
// ------------------------------------------------------------------------------------------------------------------
// X: zero-bordered plane per channel, (H + 2) rows x RP floats, RP odd; channel pitch PP == 32 / kCT (mod 32) so that the
// channels-fastest global <-> shared transposes are bank-conflict free.  V: plain plane, row pitch VP odd (column walks
// hit distinct banks).
struct InGeom {
  int D, H, W, L, RP, PP, VP, VPP;
  FastDiv dW, dH, dL, dH2, dLq;  // dH2: by ceil(H / 2); dLq: by L / 4
  int64_t pitch;
  int vec;                        // channels-last rows can be moved as float4
  int vec4;                       // scan-order rows can be moved as float4 (H, W, pitch multiples of 4, aligned base)
  int ndir;                       // scan-order planes per image: 4 (row, column, and both reversed) or 2 (the reversed orders are
                                  // walked by the scan kernels themselves: selscan_b200.h, mirror_pairs)
};

template <int kCT>
__device__ __forceinline__ void load_planes(float* sm, const float* __restrict__ xb, int64_t ld, int c0, const InGeom& g) {
  constexpr int kParts = kCT / 4;
  const int nthr = blockDim.x;
  if (g.vec && c0 + kCT <= g.D) {
    const int n = g.L * kParts;
    for (int i0 = threadIdx.x; i0 < n; i0 += nthr * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < n) v[u] = __ldg(reinterpret_cast<const float4*>(xb + (int64_t)(i / kParts) * ld + (i % kParts) * 4));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < n) {
          const int p = i / kParts, c = (i % kParts) * 4;
          const int h = fdiv(p, g.dW), w = p - h * g.W;
          float* q = sm + c * g.PP + (h + 1) * g.RP + (w + 1);
          q[0] = v[u].x;
          q[g.PP] = v[u].y;
          q[2 * g.PP] = v[u].z;
          q[3 * g.PP] = v[u].w;
        }
      }
    }
    return;
  }
  const int n = g.L * kCT;
  for (int i0 = threadIdx.x; i0 < n; i0 += nthr * 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      v[u] = (i < n && c0 + i % kCT < g.D) ? __ldg(xb + (int64_t)(i / kCT) * ld + i % kCT) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * nthr;
      if (i < n) {
        const int p = i / kCT, c = i % kCT;
        const int h = fdiv(p, g.dW), w = p - h * g.W;
        sm[c * g.PP + (h + 1) * g.RP + (w + 1)] = v[u];
      }
    }
  }
}

template <int kCT>
__device__ __forceinline__ void load_taps(float* wsm, const float* __restrict__ cw, const float* __restrict__ cb, int c0, int D) {
  if (threadIdx.x < kCT * 10) {
    const int c = threadIdx.x / 10, j = threadIdx.x - c * 10;
    float v = 0.f;
    if (c0 + c < D) v = j < 9 ? __ldg(cw + (int64_t)(c0 + c) * 9 + j) : (cb ? __ldg(cb + c0 + c) : 0.f);
    wsm[threadIdx.x] = v;
  }
}

template <int kCT>
__global__ void __launch_bounds__(kInThreads, 2)
ss2d_in_fwd_kernel(const float* __restrict__ x, int64_t ld, const float* __restrict__ cw, const float* __restrict__ cb,
                   float* __restrict__ xs, InGeom g) {
  extern __shared__ __align__(16) float sm[];
  float* V = sm + kCT * g.PP;
  float* wsm = V + kCT * g.VPP;                // [kCT][10]: 9 taps + bias
  const int L = g.L, W = g.W, H = g.H, RP = g.RP;
  const int nthr = blockDim.x;
  const int ngrp = (g.D + kCT - 1) / kCT;
  const int b = blockIdx.x / ngrp, c0 = (blockIdx.x - b * ngrp) * kCT;
  for (int i = threadIdx.x; i < kCT * g.PP / 4; i += nthr) reinterpret_cast<float4*>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  load_taps<kCT>(wsm, cw, cb, c0, g.D);
  __syncthreads();
  load_planes<kCT>(sm, x + (int64_t)b * L * ld + c0, ld, c0, g);
  __syncthreads();
  const int nc = g.D - c0 < kCT ? g.D - c0 : kCT;
  const int H2 = (H + 1) / 2;
  // conv + SiLU once, two rows per thread (12 shared loads for 2 outputs); row-major orders k = 0 / 2 leave from registers
  for (int i = threadIdx.x; i < nc * H2 * W; i += nthr) {
    const int r = fdiv(i, g.dW), w = i - r * W;
    const int c = fdiv(r, g.dH2), h = (r - c * H2) * 2;
    const float* q = sm + c * g.PP + h * RP + w;
    const float* t = wsm + c * 10;
    float a0 = t[9], a1 = t[9];
    float r0[3], r1[3], r2[3], r3[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      r0[s] = q[s];
      r1[s] = q[RP + s];
      r2[s] = q[2 * RP + s];
      r3[s] = (h + 1 < H) ? q[3 * RP + s] : 0.f;
    }
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      a0 = fmaf(t[s], r0[s], a0);
      a0 = fmaf(t[3 + s], r1[s], a0);
      a0 = fmaf(t[6 + s], r2[s], a0);
      a1 = fmaf(t[s], r1[s], a1);
      a1 = fmaf(t[3 + s], r2[s], a1);
      a1 = fmaf(t[6 + s], r3[s], a1);
    }
    const float v0 = a0 * sigmoid_fast(a0), v1 = a1 * sigmoid_fast(a1);
    float* o = xs + ((int64_t)b * g.ndir * g.D + c0 + c) * g.pitch;
    float* o2 = o + 2 * g.D * g.pitch;
    const bool has_rev = g.ndir == 4;
    const int p = h * W + w;
    o[p] = v0;
    if (has_rev) o2[L - 1 - p] = v0;
    V[c * g.VPP + h * g.VP + w] = v0;
    if (h + 1 < H) {
      o[p + W] = v1;
      if (has_rev) o2[L - 1 - p - W] = v1;
      V[c * g.VPP + (h + 1) * g.VP + w] = v1;
    }
  }
  __syncthreads();
  // column-major orders k = 1 / 3: lanes walk h
  for (int i = threadIdx.x; i < nc * L; i += nthr) {
    const int c = fdiv(i, g.dL), p = i - c * L;
    const int w = fdiv(p, g.dH), h = p - w * H;
    const float v = V[c * g.VPP + h * g.VP + w];
    float* o = xs + (((int64_t)b * g.ndir + 1) * g.D + c0 + c) * g.pitch;
    o[p] = v;
    if (g.ndir == 4) o[2 * g.D * g.pitch + (L - 1 - p)] = v;
  }
}

template <int kCT>
__global__ void __launch_bounds__(kInThreads, 2)
ss2d_in_bwd_kernel(const float* __restrict__ dxs, const float* __restrict__ x, int64_t ld, const float* __restrict__ cw,
                   const float* __restrict__ cb, float* __restrict__ dx, int64_t dld, float* __restrict__ wpart, InGeom g) {
  extern __shared__ __align__(16) float sm[];
  float* X = sm;
  float* G = sm + kCT * g.PP;
  float* wsm = G + kCT * g.PP;                 // [kCT][10]
  float* red = wsm + kCT * 10;                 // [warps][10]
  const int L = g.L, W = g.W, H = g.H, RP = g.RP, PP = g.PP, D = g.D;
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  const int wpc = nwarps > kCT ? nwarps / kCT : 1;      // warps per channel in the d(pre-activation) pass
  const int cslots = nwarps / wpc;                       // channels in flight (divides kCT: powers of two)
  const int ngrp = (D + kCT - 1) / kCT;
  const int b = blockIdx.x / ngrp, c0 = (blockIdx.x - b * ngrp) * kCT;
  for (int i = threadIdx.x; i < 2 * kCT * PP / 4; i += nthr) reinterpret_cast<float4*>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  load_taps<kCT>(wsm, cw, cb, c0, D);
  __syncthreads();
  load_planes<kCT>(X, x + (int64_t)b * L * ld + c0, ld, c0, g);
  const int nvalid = (D - c0 < kCT ? D - c0 : kCT) * L;
  const int64_t dir2 = 2 * (int64_t)D * g.pitch;
  const bool has_rev = g.ndir == 4;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (g.vec4) {
    // CrossScan backward with 16-byte loads: a thread takes 4 consecutive positions of an order and of its reverse
    const int Lq = L >> 2, nq = nvalid >> 2;
    for (int q0 = threadIdx.x; q0 < nq; q0 += nthr * 2) {         // row-major pair
      float4 a[2], r[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + u * nthr;
        if (q < nq) {
          const int c = fdiv(q, g.dLq), p = (q - c * Lq) << 2;
          const float* g0 = dxs + ((int64_t)b * g.ndir * D + c0 + c) * g.pitch;
          a[u] = __ldg(reinterpret_cast<const float4*>(g0 + p));
          r[u] = has_rev ? __ldg(reinterpret_cast<const float4*>(g0 + dir2 + (L - 4 - p))) : zero4;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + u * nthr;
        if (q < nq) {
          const int c = fdiv(q, g.dLq), p = (q - c * Lq) << 2;
          const int h = fdiv(p, g.dW), w = p - h * W;
          float* d = G + c * PP + (h + 1) * RP + (w + 1);
          d[0] = a[u].x + r[u].w;
          d[1] = a[u].y + r[u].z;
          d[2] = a[u].z + r[u].y;
          d[3] = a[u].w + r[u].x;
        }
      }
    }
    __syncthreads();
    for (int q0 = threadIdx.x; q0 < nq; q0 += nthr * 2) {         // column-major pair: 4 consecutive rows of one column
      float4 a[2], r[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + u * nthr;
        if (q < nq) {
          const int c = fdiv(q, g.dLq), p = (q - c * Lq) << 2;
          const float* g1 = dxs + (((int64_t)b * g.ndir + 1) * D + c0 + c) * g.pitch;
          a[u] = __ldg(reinterpret_cast<const float4*>(g1 + p));
          r[u] = has_rev ? __ldg(reinterpret_cast<const float4*>(g1 + dir2 + (L - 4 - p))) : zero4;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = q0 + u * nthr;
        if (q < nq) {
          const int c = fdiv(q, g.dLq), p = (q - c * Lq) << 2;
          const int w = fdiv(p, g.dH), h = p - w * H;
          float* d = G + c * PP + (h + 1) * RP + (w + 1);
          d[0] += a[u].x + r[u].w;
          d[RP] += a[u].y + r[u].z;
          d[2 * RP] += a[u].z + r[u].y;
          d[3 * RP] += a[u].w + r[u].x;
        }
      }
    }
    __syncthreads();
  } else {
    for (int i0 = threadIdx.x; i0 < nvalid; i0 += nthr * 4) {   // CrossScan backward, row-major pair
      float v[4];
  #pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < nvalid) {
          const int c = fdiv(i, g.dL), p = i - c * L;
          const float* g0 = dxs + ((int64_t)b * g.ndir * D + c0 + c) * g.pitch;
          v[u] = __ldg(g0 + p) + (has_rev ? __ldg(g0 + dir2 + (L - 1 - p)) : 0.f);
        }
      }
  #pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < nvalid) {
          const int c = fdiv(i, g.dL), p = i - c * L;
          const int h = fdiv(p, g.dW), w = p - h * W;
          G[c * PP + (h + 1) * RP + (w + 1)] = v[u];
        }
      }
    }
    __syncthreads();
    for (int i0 = threadIdx.x; i0 < nvalid; i0 += nthr * 4) {   // column-major pair
      float v[4];
  #pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < nvalid) {
          const int c = fdiv(i, g.dL), p = i - c * L;
          const float* g1 = dxs + (((int64_t)b * g.ndir + 1) * D + c0 + c) * g.pitch;
          v[u] = __ldg(g1 + p) + (has_rev ? __ldg(g1 + dir2 + (L - 1 - p)) : 0.f);
        }
      }
  #pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * nthr;
        if (i < nvalid) {
          const int c = fdiv(i, g.dL), p = i - c * L;
          const int w = fdiv(p, g.dH), h = p - w * H;
          G[c * PP + (h + 1) * RP + (w + 1)] += v[u];
        }
      }
    }
    __syncthreads();
  }
  {  // d(pre-activation) in place, conv weight / bias partial sums; a warp stays on one channel at a time
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = warp % wpc;
    for (int c = warp / wpc; c < kCT; c += cslots) {
      float acc[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) acc[j] = 0.f;
      if (c0 + c < D) {
        const float* wc = wsm + c * 10;
        for (int p = sub * 32 + lane; p < L; p += wpc * 32) {
          const int h = fdiv(p, g.dW), w = p - h * W;
          const float* q = X + c * PP + h * RP + w;
          float xv[9];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) xv[r * 3 + s] = q[r * RP + s];
          float a = wc[9];
#pragma unroll
          for (int j = 0; j < 9; ++j) a = fmaf(wc[j], xv[j], a);
          const float sg = sigmoid_fast(a);
          float* gp = G + c * PP + (h + 1) * RP + (w + 1);
          const float dp = *gp * (sg * (1.f + a * (1.f - sg)));
          *gp = dp;
#pragma unroll
          for (int j = 0; j < 9; ++j) acc[j] = fmaf(dp, xv[j], acc[j]);
          acc[9] += dp;
        }
      }
#pragma unroll
      for (int j = 0; j < 10; ++j) acc[j] = warp_sum(acc[j]);
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 10; ++j) red[(c * wpc + sub) * 10 + j] = acc[j];
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < kCT * 10) {
    const int c = threadIdx.x / 10, j = threadIdx.x - c * 10;
    float s = 0.f;
    for (int k = 0; k < wpc; ++k) s += red[(c * wpc + k) * 10 + j];
    if (c0 + c < D) wpart[((int64_t)b * D + c0 + c) * 10 + j] = s;
  }
  // transposed conv, channels-fastest store.  padded coords: dpre[h' - r + 1][w' - s + 1] = q[(2 - r) * RP + (2 - s)]
  float* dxb = dx + (int64_t)b * L * dld + c0;
  constexpr int kParts = kCT / 4;
  const bool vec = g.vec && c0 + kCT <= D && (dld % 4 == 0) && ((reinterpret_cast<uintptr_t>(dxb) & 15) == 0);
  if (vec) {
    for (int i = threadIdx.x; i < L * kParts; i += nthr) {
      const int p = i / kParts, cq = (i % kParts) * 4;
      const int h = fdiv(p, g.dW), w = p - h * W;
      float o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* q = G + (cq + k) * PP + h * RP + w;
        const float* wc = wsm + (cq + k) * 10;
        float a = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s) a = fmaf(wc[r * 3 + s], q[(2 - r) * RP + (2 - s)], a);
        o[k] = a;
      }
      *reinterpret_cast<float4*>(dxb + (int64_t)p * dld + cq) = make_float4(o[0], o[1], o[2], o[3]);
    }
  } else {
    for (int i = threadIdx.x; i < L * kCT; i += nthr) {
      const int c = i % kCT, p = i / kCT;
      if (c0 + c >= D) continue;
      const int h = fdiv(p, g.dW), w = p - h * W;
      const float* q = G + c * PP + h * RP + w;
      const float* wc = wsm + c * 10;
      float a = 0.f;
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) a = fmaf(wc[r * 3 + s], q[(2 - r) * RP + (2 - s)], a);
      dxb[(int64_t)p * dld + c] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// epilogue
// ------------------------------------------------------------------------------------------------------------------
struct OutGeom {
  int D, H, W, L, tiles_h, tiles_w, DP;
  int ndir;                      // scan-order planes per image: 4, or 2 (no reversed orders)
  FastDiv dDb;                   // by the number of 32-channel blocks
  int64_t pitch;
};

// Walk every (channel, run) of the tile for one pair of scan orders.  kCol = false: runs along w (orders 0 / 2);
// kCol = true: runs along h (orders 1 / 3).  A lane moves VEC consecutive run elements (float2 when the plane sides are
// even), a warp instruction covers one run x 32 / (R / VEC) channels; the channel pattern keeps the accesses to the
// [position][channel] tile (pitch DP == 1 mod 32) at most 2-way bank conflicted.  Four steps are decoded at once:
// body(d, gpos, slot0, slot_step) issues its global loads for all four before `finish` consumes them.
template <int VEC>
struct Run {
  float a[VEC];
};

template <int VEC>
__device__ __forceinline__ Run<VEC> ld_run(const float* p) {          // elements p[0 .. VEC)
  Run<VEC> r;
  if constexpr (VEC == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    r.a[0] = t.x;
    r.a[1] = t.y;
  } else {
    r.a[0] = __ldg(p);
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ Run<VEC> ld_run_rev(const float* base, int L, int gp) {   // elements base[L-1-gp], base[L-2-gp]
  Run<VEC> r;
  if constexpr (VEC == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(base + (L - 2 - gp)));
    r.a[0] = t.y;
    r.a[1] = t.x;
  } else {
    r.a[0] = __ldg(base + (L - 1 - gp));
  }
  return r;
}

template <int VEC>
__device__ __forceinline__ void st_run(float* p, const Run<VEC>& r) {
  if constexpr (VEC == 2) *reinterpret_cast<float2*>(p) = make_float2(r.a[0], r.a[1]);
  else *p = r.a[0];
}

template <int VEC>
__device__ __forceinline__ void st_run_rev(float* base, int L, int gp, const Run<VEC>& r) {
  if constexpr (VEC == 2) *reinterpret_cast<float2*>(base + (L - 2 - gp)) = make_float2(r.a[1], r.a[0]);
  else base[L - 1 - gp] = r.a[0];
}

template <int TH, int TW, bool kCol, int VEC, typename FL, typename FS>
__device__ __forceinline__ void for_each_run(const OutGeom& g, int h0, int w0, FL load, FS finish) {
  constexpr int R = kCol ? TH : TW;            // run length
  constexpr int O = kCol ? TW : TH;            // runs per channel
  constexpr int RL = R / VEC;                  // lanes per run
  constexpr int DS = 32 / RL;                  // channels per warp instruction
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = (lane % RL) * VEC, dsub = lane / RL;
  const int dblocks = g.dDb.d;
  const int step = (kCol ? TW : 1) * g.DP;     // tile distance of two consecutive run elements
  for (int q = warp; q < O * dblocks; q += kEdgeWarps) {          // (run o, block of 32 channels)
    const int o = fdiv(q, g.dDb), dblk = q - o * dblocks;
    const int hh = kCol ? i : o, ww = kCol ? o : i;
    const int h = h0 + hh, w = w0 + ww;
    const bool inside = h < g.H && w < g.W;                       // VEC = 2 only with even H and W: both elements are inside
    const int gp = kCol ? w * g.H + h : h * g.W + w;
    const int slot0 = (hh * TW + ww) * g.DP + dblk * 32;
    Run<VEC> v[RL];
#pragma unroll
    for (int dlo = 0; dlo < RL; ++dlo) {                          // the RL x 2 global loads of a step are issued back to back
      const int dd = kCol ? dlo * DS + dsub : dsub * RL + dlo;
      if (inside && dblk * 32 + dd < g.D) v[dlo] = load(dblk * 32 + dd, gp, slot0 + dd, step);
    }
#pragma unroll
    for (int dlo = 0; dlo < RL; ++dlo) {
      const int dd = kCol ? dlo * DS + dsub : dsub * RL + dlo;
      if (inside && dblk * 32 + dd < g.D) finish(slot0 + dd, step, v[dlo]);
    }
  }
}

template <int TH, int TW, int VEC>
__global__ void __launch_bounds__(kEdgeThreads, 4)
ss2d_out_fwd_kernel(const float* __restrict__ ys, const float* __restrict__ z, int64_t zld, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, float* __restrict__ out, float* __restrict__ xhat,
                    float* __restrict__ rstd_out, OutGeom g) {
  extern __shared__ __align__(16) float T[];
  const int tiles = g.tiles_h * g.tiles_w;
  const int b = blockIdx.x / tiles, t = blockIdx.x - b * tiles;
  const int h0 = (t / g.tiles_w) * TH, w0 = (t % g.tiles_w) * TW;
  const float* yb = ys + (int64_t)b * g.ndir * g.D * g.pitch;
  const int64_t dir = (int64_t)g.D * g.pitch;
  const int L = g.L, D = g.D;
  const bool has_rev = g.ndir == 4;
  for_each_run<TH, TW, false, VEC>(
      g, h0, w0,
      [&](int d, int gp, int, int) {
        const float* r = yb + (int64_t)d * g.pitch;
        Run<VEC> a = ld_run<VEC>(r + gp);
        if (has_rev) {
          const Run<VEC> c = ld_run_rev<VEC>(r + 2 * dir, L, gp);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a.a[k] += c.a[k];
        }
        return a;
      },
      [&](int slot, int step, const Run<VEC>& v) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) T[slot + k * step] = v.a[k];
      });
  __syncthreads();
  for_each_run<TH, TW, true, VEC>(
      g, h0, w0,
      [&](int d, int gp, int, int) {
        const float* r = yb + dir + (int64_t)d * g.pitch;
        Run<VEC> a = ld_run<VEC>(r + gp);
        if (has_rev) {
          const Run<VEC> c = ld_run_rev<VEC>(r + 2 * dir, L, gp);
#pragma unroll
          for (int k = 0; k < VEC; ++k) a.a[k] += c.a[k];
        }
        return a;
      },
      [&](int slot, int step, const Run<VEC>& v) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) T[slot + k * step] += v.a[k];
      });
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_d = 1.f / (float)D;
  for (int tp = warp; tp < TH * TW; tp += kEdgeWarps) {
    const int h = h0 + tp / TW, w = w0 + tp % TW;
    if (h >= g.H || w >= g.W) continue;
    const float* row = T + tp * g.DP;
    const int64_t pos = (int64_t)b * L + h * g.W + w;
    float s = 0.f;
#pragma unroll 4
    for (int d = lane; d < D; d += 32) s += row[d];
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll 4
    for (int d = lane; d < D; d += 32) {
      const float c = row[d] - mean;
      q = fmaf(c, c, q);
    }
    const float rs = rsqrtf(warp_sum(q) * inv_d + eps);
    if (rstd_out != nullptr && lane == 0) rstd_out[pos] = rs;
    if (z != nullptr) {
      const float* zr = z + pos * zld;
#pragma unroll 4
      for (int d = lane; d < D; d += 32) {
        const float zz = __ldg(zr + d);
        const float xh = (row[d] - mean) * rs;
        out[pos * D + d] = fmaf(xh, __ldg(gamma + d), __ldg(beta + d)) * (zz * sigmoid_fast(zz));
        if (xhat != nullptr) xhat[pos * D + d] = xh;
      }
    } else {
#pragma unroll 4
      for (int d = lane; d < D; d += 32) {
        const float xh = (row[d] - mean) * rs;
        out[pos * D + d] = fmaf(xh, __ldg(gamma + d), __ldg(beta + d));
        if (xhat != nullptr) xhat[pos * D + d] = xh;
      }
    }
  }
}

template <int TH, int TW, int VEC>
__global__ void __launch_bounds__(kEdgeThreads, 4)
ss2d_out_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ z, int64_t zld, const float* __restrict__ xhat,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ dz, int64_t dzld, float* __restrict__ dys, float* __restrict__ part, OutGeom g) {
  extern __shared__ __align__(16) float T[];
  constexpr int P = TH * TW;
  float* m1 = T + P * g.DP;                    // [P] mean_d(dln * gamma)
  float* m2 = m1 + P;                          // [P] mean_d(dln * gamma * xhat)
  const int tiles = g.tiles_h * g.tiles_w;
  const int b = blockIdx.x / tiles, t = blockIdx.x - b * tiles;
  const int h0 = (t / g.tiles_w) * TH, w0 = (t % g.tiles_w) * TW;
  const int L = g.L, D = g.D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_d = 1.f / (float)D;
  for (int tp = warp; tp < P; tp += kEdgeWarps) {                  // gate backward, LayerNorm row moments
    const int h = h0 + tp / TW, w = w0 + tp % TW;
    float* row = T + tp * g.DP;
    if (h >= g.H || w >= g.W) {
      for (int d = lane; d < D; d += 32) row[d] = 0.f;
      continue;
    }
    const int64_t pos = (int64_t)b * L + h * g.W + w;
    const float* gr = gout + pos * D;
    const float* xr = xhat + pos * D;
    float a1 = 0.f, a2 = 0.f;
    if (z != nullptr) {
      const float* zr = z + pos * zld;
      float* dzr = dz + pos * dzld;
#pragma unroll 4
      for (int d = lane; d < D; d += 32) {
        const float go = __ldg(gr + d), xh = __ldg(xr + d), zz = __ldg(zr + d);
        const float gm = __ldg(gamma + d);
        const float sg = sigmoid_fast(zz);
        dzr[d] = go * fmaf(xh, gm, __ldg(beta + d)) * (sg * (1.f + zz * (1.f - sg)));
        const float dln = go * (zz * sg);
        row[d] = dln;
        const float dy = dln * gm;
        a1 += dy;
        a2 = fmaf(dy, xh, a2);
      }
    } else {
#pragma unroll 4
      for (int d = lane; d < D; d += 32) {
        const float dln = __ldg(gr + d), xh = __ldg(xr + d);
        row[d] = dln;
        const float dy = dln * __ldg(gamma + d);
        a1 += dy;
        a2 = fmaf(dy, xh, a2);
      }
    }
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
      m1[tp] = a1 * inv_d;
      m2[tp] = a2 * inv_d;
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += kEdgeThreads) {            // per-CTA gamma / beta partials (summed by the host)
    float sg = 0.f, sb = 0.f;
#pragma unroll 8
    for (int tp = 0; tp < P; ++tp) {
      const int h = h0 + tp / TW, w = w0 + tp % TW;
      if (h < g.H && w < g.W) {
        const float dln = T[tp * g.DP + d];
        sb += dln;
        sg = fmaf(dln, __ldg(xhat + ((int64_t)b * L + h * g.W + w) * D + d), sg);
      }
    }
    part[(int64_t)blockIdx.x * 2 * D + d] = sg;
    part[(int64_t)blockIdx.x * 2 * D + D + d] = sb;
  }
  __syncthreads();
  for (int tp = warp; tp < P; tp += kEdgeWarps) {                  // LayerNorm input gradient, in place
    const int h = h0 + tp / TW, w = w0 + tp % TW;
    if (h >= g.H || w >= g.W) continue;
    const int64_t pos = (int64_t)b * L + h * g.W + w;
    const float rs = __ldg(rstd + pos), c1 = m1[tp], c2 = m2[tp];
    float* row = T + tp * g.DP;
    const float* xr = xhat + pos * D;
#pragma unroll 4
    for (int d = lane; d < D; d += 32) row[d] = rs * (fmaf(row[d], __ldg(gamma + d), -c1) - __ldg(xr + d) * c2);
  }
  __syncthreads();
  float* db = dys + (int64_t)b * g.ndir * D * g.pitch;
  const int64_t dir = (int64_t)D * g.pitch;
  const bool has_rev = g.ndir == 4;
  // CrossMerge backward: every order receives the same value (the global stores need d and gp, so they sit in the first functor)
  for_each_run<TH, TW, false, VEC>(
      g, h0, w0,
      [&](int d, int gp, int slot, int step) {
        Run<VEC> v;
#pragma unroll
        for (int k = 0; k < VEC; ++k) v.a[k] = T[slot + k * step];
        float* r = db + (int64_t)d * g.pitch;
        st_run<VEC>(r + gp, v);
        if (has_rev) st_run_rev<VEC>(r + 2 * dir, L, gp, v);
        return v;
      },
      [&](int, int, const Run<VEC>&) {});
  for_each_run<TH, TW, true, VEC>(
      g, h0, w0,
      [&](int d, int gp, int slot, int step) {
        Run<VEC> v;
#pragma unroll
        for (int k = 0; k < VEC; ++k) v.a[k] = T[slot + k * step];
        float* r = db + dir + (int64_t)d * g.pitch;
        st_run<VEC>(r + gp, v);
        if (has_rev) st_run_rev<VEC>(r + 2 * dir, L, gp, v);
        return v;
      },
      [&](int, int, const Run<VEC>&) {});
}

struct InPlan {
  int ct;
  size_t smem;
  InGeom g;
};

// channels per CTA: 8 when the planes of 8 channels leave room for two CTAs per SM, else 4.
// planes: forward = padded X + plain V, backward = padded X + padded G
InPlan plan_in(int D, int H, int W, int64_t pitch, bool bwd) {
  InPlan p;
  InGeom& g = p.g;
  g.D = D; g.H = H; g.W = W; g.L = H * W; g.pitch = pitch; g.ndir = 4;
  g.RP = (W + 2) | 1;
  g.VP = W | 1;
  g.VPP = H * g.VP;
  g.dW = make_fastdiv(W); g.dH = make_fastdiv(H); g.dL = make_fastdiv(H * W); g.dH2 = make_fastdiv((H + 1) / 2);
  g.dLq = make_fastdiv(H * W / 4 > 0 ? H * W / 4 : 1);
  g.vec = 0;
  g.vec4 = 0;
  for (p.ct = 8;; p.ct = 4) {
    const int want = 32 / p.ct;
    int pp = (H + 2) * g.RP;
    pp += ((want - pp % 32) % 32 + 32) % 32;
    g.PP = pp;
    const size_t planes = bwd ? 2 * (size_t)p.ct * pp : (size_t)p.ct * (pp + g.VPP);
    p.smem = (planes + p.ct * 10 + 10 * kInWarps) * sizeof(float);
    if (p.ct == 4 || p.smem <= 112 * 1024) break;
  }
  return p;
}

template <int TH, int TW>
bool out_fits(int DP, size_t* smem) {
  *smem = ((size_t)TH * TW * DP + 2 * TH * TW) * sizeof(float);
  return *smem <= 52 * 1024;                    // four CTAs per SM
}

// tile shape index: 0 = 8x8, 1 = 8x4, 2 = 4x4, 3 = 4x2, 4 = 2x2, 5 = 1x1 (any D up to ~55K)
int plan_out(int D, int H, int W, int64_t pitch, OutGeom* g, size_t* smem) {
  g->D = D; g->H = H; g->W = W; g->L = H * W; g->pitch = pitch; g->ndir = 4;
  g->DP = D + ((1 - D % 32) + 32) % 32;
  g->dDb = make_fastdiv((D + 31) / 32);
  int shape, th, tw;
  if (out_fits<8, 8>(g->DP, smem)) { shape = 0; th = 8; tw = 8; }
  else if (out_fits<8, 4>(g->DP, smem)) { shape = 1; th = 8; tw = 4; }
  else if (out_fits<4, 4>(g->DP, smem)) { shape = 2; th = 4; tw = 4; }
  else if (out_fits<4, 2>(g->DP, smem)) { shape = 3; th = 4; tw = 2; }
  else if (out_fits<2, 2>(g->DP, smem)) { shape = 4; th = 2; tw = 2; }
  else {
    *smem = ((size_t)g->DP + 2) * sizeof(float);
    if (*smem > 220 * 1024) return -1;
    shape = 5; th = 1; tw = 1;
  }
  g->tiles_h = (H + th - 1) / th;
  g->tiles_w = (W + tw - 1) / tw;
  return shape;
}

// threads per prologue CTA: about 8 elements per thread, 128 .. 512 (>= 80 threads are needed to stage the taps)
int in_threads(int ct, int L) {
  const int want = (ct * L / 8 + 31) / 32 * 32;
  return want < 128 ? 128 : (want > kInThreads ? kInThreads : want);
}

template <typename K>
cudaError_t allow_smem(K kernel, size_t smem) {
  if (smem <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool ss2d_in_supported(int H, int W) {
  return plan_in(8, H, W, (int64_t)H * W, true).smem <= 220 * 1024 && plan_in(8, H, W, (int64_t)H * W, false).smem <= 220 * 1024 &&
         (int64_t)H * W * 8 < (1 << 24);
}

cudaError_t launch_ss2d_in_fwd(const float* x, int64_t ld, const float* cw, const float* cb, float* xs, int B, int D, int H, int W,
                               int64_t pitch, int n_planes, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  InPlan p = plan_in(D, H, W, pitch, false);
  p.g.ndir = n_planes;
  p.g.vec = al16(x) && ld % 4 == 0;
  const unsigned grid = (unsigned)((int64_t)B * ((D + p.ct - 1) / p.ct));
  cudaError_t e;
  if (p.ct == 8) {
    if ((e = allow_smem(ss2d_in_fwd_kernel<8>, p.smem)) != cudaSuccess) return e;
    ss2d_in_fwd_kernel<8><<<grid, in_threads(p.ct, H * W), p.smem, stream>>>(x, ld, cw, cb, xs, p.g);
  } else {
    if ((e = allow_smem(ss2d_in_fwd_kernel<4>, p.smem)) != cudaSuccess) return e;
    ss2d_in_fwd_kernel<4><<<grid, in_threads(p.ct, H * W), p.smem, stream>>>(x, ld, cw, cb, xs, p.g);
  }
  return cudaGetLastError();
}

cudaError_t launch_ss2d_in_bwd(const float* dxs, const float* x, int64_t ld, const float* cw, const float* cb, float* dx, int64_t dld,
                               float* wpart, int B, int D, int H, int W, int64_t pitch, int n_planes, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  InPlan p = plan_in(D, H, W, pitch, true);
  p.g.ndir = n_planes;
  p.g.vec = al16(x) && ld % 4 == 0;
  p.g.vec4 = al16(dxs) && H % 4 == 0 && W % 4 == 0 && pitch % 4 == 0;
  const unsigned grid = (unsigned)((int64_t)B * ((D + p.ct - 1) / p.ct));
  cudaError_t e;
  if (p.ct == 8) {
    if ((e = allow_smem(ss2d_in_bwd_kernel<8>, p.smem)) != cudaSuccess) return e;
    ss2d_in_bwd_kernel<8><<<grid, in_threads(p.ct, H * W), p.smem, stream>>>(dxs, x, ld, cw, cb, dx, dld, wpart, p.g);
  } else {
    if ((e = allow_smem(ss2d_in_bwd_kernel<4>, p.smem)) != cudaSuccess) return e;
    ss2d_in_bwd_kernel<4><<<grid, in_threads(p.ct, H * W), p.smem, stream>>>(dxs, x, ld, cw, cb, dx, dld, wpart, p.g);
  }
  return cudaGetLastError();
}

bool ss2d_out_supported(int D) {
  OutGeom g;
  size_t smem;
  return D > 0 && plan_out(D, 8, 8, 64, &g, &smem) >= 0;
}

int64_t ss2d_out_ctas(int B, int D, int H, int W) {
  OutGeom g;
  size_t smem;
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || plan_out(D, H, W, (int64_t)H * W, &g, &smem) < 0) return 0;
  return (int64_t)B * g.tiles_h * g.tiles_w;
}

#define SS2D_OUT_CASE(KERNEL, TH, TW, ...)                                                                        \
  if (vec2) {                                                                                                      \
    if ((e = allow_smem(KERNEL<TH, TW, 2>, smem)) == cudaSuccess) KERNEL<TH, TW, 2><<<grid, kEdgeThreads, smem, stream>>>(__VA_ARGS__); \
  } else {                                                                                                         \
    if ((e = allow_smem(KERNEL<TH, TW, 1>, smem)) == cudaSuccess) KERNEL<TH, TW, 1><<<grid, kEdgeThreads, smem, stream>>>(__VA_ARGS__); \
  }                                                                                                                \
  break;

#define SS2D_OUT_DISPATCH(KERNEL, ...)                                                                             \
  switch (shape) {                                                                                                 \
    case 0: SS2D_OUT_CASE(KERNEL, 8, 8, __VA_ARGS__)                                                               \
    case 1: SS2D_OUT_CASE(KERNEL, 8, 4, __VA_ARGS__)                                                               \
    case 2: SS2D_OUT_CASE(KERNEL, 4, 4, __VA_ARGS__)                                                               \
    case 3: SS2D_OUT_CASE(KERNEL, 4, 2, __VA_ARGS__)                                                               \
    case 4: SS2D_OUT_CASE(KERNEL, 2, 2, __VA_ARGS__)                                                               \
    default:                                                                                                       \
      if ((e = allow_smem(KERNEL<1, 1, 1>, smem)) == cudaSuccess) KERNEL<1, 1, 1><<<grid, kEdgeThreads, smem, stream>>>(__VA_ARGS__); \
      break;                                                                                                       \
  }

// float2 runs: even plane sides and row pitch, 8-byte aligned base
bool out_vec2(const float* p, int H, int W, int64_t pitch) {
  return H % 2 == 0 && W % 2 == 0 && pitch % 2 == 0 && (reinterpret_cast<uintptr_t>(p) & 7) == 0;
}

cudaError_t launch_ss2d_out_fwd(const float* ys, int64_t pitch, const float* z, int64_t zld, const float* gamma, const float* beta,
                                float eps, float* out, float* xhat, float* rstd, int B, int D, int H, int W, int n_planes, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  OutGeom g;
  size_t smem;
  const int shape = plan_out(D, H, W, pitch, &g, &smem);
  if (shape < 0) return cudaErrorInvalidValue;
  g.ndir = n_planes;
  const unsigned grid = (unsigned)((int64_t)B * g.tiles_h * g.tiles_w);
  cudaError_t e = cudaSuccess;
  const bool vec2 = out_vec2(ys, H, W, pitch);
  SS2D_OUT_DISPATCH(ss2d_out_fwd_kernel, ys, z, zld, gamma, beta, eps, out, xhat, rstd, g)
  return e != cudaSuccess ? e : cudaGetLastError();
}

cudaError_t launch_ss2d_out_bwd(const float* gout, const float* z, int64_t zld, const float* xhat, const float* rstd,
                                const float* gamma, const float* beta, float* dz, int64_t dzld, float* dys, int64_t pitch,
                                float* part, int B, int D, int H, int W, int n_planes, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  OutGeom g;
  size_t smem;
  const int shape = plan_out(D, H, W, pitch, &g, &smem);
  if (shape < 0) return cudaErrorInvalidValue;
  g.ndir = n_planes;
  const unsigned grid = (unsigned)((int64_t)B * g.tiles_h * g.tiles_w);
  cudaError_t e = cudaSuccess;
  const bool vec2 = out_vec2(dys, H, W, pitch);
  SS2D_OUT_DISPATCH(ss2d_out_bwd_kernel, gout, z, zld, xhat, rstd, gamma, beta, dz, dzld, dys, part, g)
  return e != cudaSuccess ? e : cudaGetLastError();
}

}  // namespace selscan
