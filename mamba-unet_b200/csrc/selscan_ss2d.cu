// The two edges of SS2D around the scan, each as ONE kernel per direction of autograd (sm_100a).
//
// Reference (code/networks/mamba_sys.py), per SS2D.forward call:
//   prologue  :533-534 + :403-404   x.permute(0,3,1,2).contiguous() -> depthwise 3x3 conv + bias -> SiLU -> CrossScan
//                                   (stack + transpose.contiguous + flip + cat): ~8 ATen kernels, ~11 passes over (B, D, L)
//   epilogue  :429-434 + :536-537   CrossMerge (flip, 2 x transpose.contiguous, 3 adds) -> transpose(1,2).contiguous()
//                                   -> LayerNorm(d_inner) -> * silu(z): ~10 kernels, ~14 passes
// and roughly twice that in their backward.  Here:
//   ss2d_in_fwd   reads the x half of in_proj's output IN PLACE (channels-last, strided), one CTA per (image, group of
//                 4 / 8 channels) stages the zero-padded planes in shared memory, and writes the four scan orders directly.
//   ss2d_in_bwd   gathers the four incoming gradients into the plane, recomputes the pre-activation, applies SiLU', the
//                 transposed conv, and writes dx channels-last into its half of d(xz); per-CTA conv weight/bias partials.
//   ss2d_out_fwd  one CTA per (image, TH x TW tile of positions) x ALL channels: merges the four scan outputs through a
//                 shared tile [position][channel], LayerNorm per position (two-pass moments), gate, channels-last store.
//   ss2d_out_bwd  gate', LayerNorm backward, per-CTA gamma/beta partials, dz into its half of d(xz), and the four
//                 directional gradients written straight in scan layout.
// Tile runs along a scan direction are TW (or TH) floats = one 32-byte sector at 8, so the strided side of every
// transpose still moves whole sectors.
#include <cuda_runtime.h>
#include <stdint.h>

#include "selscan_kernels.h"

namespace selscan {

namespace {

constexpr int kEdgeThreads = 256;

__device__ __forceinline__ float sigmoid_acc(float v) { return 1.f / (1.f + expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------------------------
// prologue
// ------------------------------------------------------------------------------------------------------------------
// Shared plane of one channel: (H + 2) rows x RP floats with a zero border, RP odd (column walks hit distinct banks);
// channel pitch PP == 32 / kCT (mod 32) so the channels-fastest global <-> shared transposes are conflict-free too.

template <int kCT>
__device__ __forceinline__ void load_planes(float* sm, const float* __restrict__ xb, int64_t ld, int c0, int D, int L, int W,
                                            int RP, int PP) {
  for (int i = threadIdx.x; i < L * kCT; i += kEdgeThreads) {
    const int c = i % kCT, p = i / kCT;
    const int h = p / W, w = p - h * W;
    if (c0 + c < D) sm[c * PP + (h + 1) * RP + (w + 1)] = __ldg(xb + (int64_t)p * ld + c);
  }
}

__device__ __forceinline__ float conv9(const float* q, int RP, const float* w, float bias) {
  float acc = bias;
  acc = fmaf(w[0], q[0], acc);
  acc = fmaf(w[1], q[1], acc);
  acc = fmaf(w[2], q[2], acc);
  acc = fmaf(w[3], q[RP], acc);
  acc = fmaf(w[4], q[RP + 1], acc);
  acc = fmaf(w[5], q[RP + 2], acc);
  acc = fmaf(w[6], q[2 * RP], acc);
  acc = fmaf(w[7], q[2 * RP + 1], acc);
  acc = fmaf(w[8], q[2 * RP + 2], acc);
  return acc;
}

template <int kCT>
__global__ void __launch_bounds__(kEdgeThreads)
ss2d_in_fwd_kernel(const float* __restrict__ x, int64_t ld, const float* __restrict__ cw, const float* __restrict__ cb,
                   float* __restrict__ xs, int D, int H, int W, int RP, int PP, int64_t pitch) {
  extern __shared__ __align__(16) float sm[];
  float* wsm = sm + kCT * PP;                  // [kCT][10]: 9 taps + bias
  const int L = H * W;
  const int ngrp = (D + kCT - 1) / kCT;
  const int b = blockIdx.x / ngrp, c0 = (blockIdx.x - b * ngrp) * kCT;
  for (int i = threadIdx.x; i < kCT * PP / 4; i += kEdgeThreads) reinterpret_cast<float4*>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x < kCT * 10) {
    const int c = threadIdx.x / 10, j = threadIdx.x - c * 10;
    float v = 0.f;
    if (c0 + c < D) v = j < 9 ? __ldg(cw + (int64_t)(c0 + c) * 9 + j) : (cb ? __ldg(cb + c0 + c) : 0.f);
    wsm[threadIdx.x] = v;
  }
  __syncthreads();
  load_planes<kCT>(sm, x + (int64_t)b * L * ld + c0, ld, c0, D, L, W, RP, PP);
  __syncthreads();
  const int nvalid = (D - c0 < kCT ? D - c0 : kCT) * L;
  // row-major orders (k = 0 and its reverse k = 2): lanes walk w
  for (int i = threadIdx.x; i < nvalid; i += kEdgeThreads) {
    const int c = i / L, p = i - c * L;
    const int h = p / W, w = p - h * W;
    const float a = conv9(sm + c * PP + h * RP + w, RP, wsm + c * 10, wsm[c * 10 + 9]);
    const float v = a * sigmoid_acc(a);
    float* o = xs + ((int64_t)b * 4 * D + c0 + c) * pitch;
    o[p] = v;
    o[2 * D * pitch + (L - 1 - p)] = v;
  }
  // column-major orders (k = 1, k = 3): lanes walk h; the conv is simply evaluated again (9 FMAs) instead of parking it
  for (int i = threadIdx.x; i < nvalid; i += kEdgeThreads) {
    const int c = i / L, p = i - c * L;
    const int w = p / H, h = p - w * H;
    const float a = conv9(sm + c * PP + h * RP + w, RP, wsm + c * 10, wsm[c * 10 + 9]);
    const float v = a * sigmoid_acc(a);
    float* o = xs + (((int64_t)b * 4 + 1) * D + c0 + c) * pitch;
    o[p] = v;
    o[2 * D * pitch + (L - 1 - p)] = v;
  }
}

template <int kCT>
__global__ void __launch_bounds__(kEdgeThreads)
ss2d_in_bwd_kernel(const float* __restrict__ dxs, const float* __restrict__ x, int64_t ld, const float* __restrict__ cw,
                   const float* __restrict__ cb, float* __restrict__ dx, int64_t dld, float* __restrict__ wpart, int D, int H,
                   int W, int RP, int PP, int64_t pitch) {
  extern __shared__ __align__(16) float sm[];
  float* X = sm;
  float* G = sm + kCT * PP;
  float* wsm = G + kCT * PP;                   // [kCT][10]
  float* red = wsm + kCT * 10;                 // [nwarps][10]
  constexpr int kWarps = kEdgeThreads / 32;
  constexpr int kWarpsPerC = kWarps / kCT;     // 1 (kCT = 8) or 2 (kCT = 4)
  const int L = H * W;
  const int ngrp = (D + kCT - 1) / kCT;
  const int b = blockIdx.x / ngrp, c0 = (blockIdx.x - b * ngrp) * kCT;
  for (int i = threadIdx.x; i < 2 * kCT * PP / 4; i += kEdgeThreads) reinterpret_cast<float4*>(sm)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x < kCT * 10) {
    const int c = threadIdx.x / 10, j = threadIdx.x - c * 10;
    float v = 0.f;
    if (c0 + c < D) v = j < 9 ? __ldg(cw + (int64_t)(c0 + c) * 9 + j) : (cb ? __ldg(cb + c0 + c) : 0.f);
    wsm[threadIdx.x] = v;
  }
  __syncthreads();
  load_planes<kCT>(X, x + (int64_t)b * L * ld + c0, ld, c0, D, L, W, RP, PP);
  const int nvalid = (D - c0 < kCT ? D - c0 : kCT) * L;
  for (int i = threadIdx.x; i < nvalid; i += kEdgeThreads) {          // CrossScan backward, row-major pair
    const int c = i / L, p = i - c * L;
    const int h = p / W, w = p - h * W;
    const float* g0 = dxs + ((int64_t)b * 4 * D + c0 + c) * pitch;
    G[c * PP + (h + 1) * RP + (w + 1)] = __ldg(g0 + p) + __ldg(g0 + 2 * D * pitch + (L - 1 - p));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nvalid; i += kEdgeThreads) {          // column-major pair
    const int c = i / L, p = i - c * L;
    const int w = p / H, h = p - w * H;
    const float* g1 = dxs + (((int64_t)b * 4 + 1) * D + c0 + c) * pitch;
    G[c * PP + (h + 1) * RP + (w + 1)] += __ldg(g1 + p) + __ldg(g1 + 2 * D * pitch + (L - 1 - p));
  }
  __syncthreads();
  {  // d(pre-activation) in place, conv weight / bias partial sums; a warp stays on one channel
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = warp / kWarpsPerC, sub = warp - c * kWarpsPerC;
    float acc[10];
#pragma unroll
    for (int j = 0; j < 10; ++j) acc[j] = 0.f;
    if (c0 + c < D) {
      const float* wc = wsm + c * 10;
      for (int p = sub * 32 + lane; p < L; p += kWarpsPerC * 32) {
        const int h = p / W, w = p - h * W;
        const float* q = X + c * PP + h * RP + w;
        float xv[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
          for (int s = 0; s < 3; ++s) xv[r * 3 + s] = q[r * RP + s];
        float a = wc[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) a = fmaf(wc[j], xv[j], a);
        const float sg = sigmoid_acc(a);
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
      for (int j = 0; j < 10; ++j) red[warp * 10 + j] = acc[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < kCT * 10) {
    const int c = threadIdx.x / 10, j = threadIdx.x - c * 10;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kWarpsPerC; ++k) s += red[(c * kWarpsPerC + k) * 10 + j];
    if (c0 + c < D) wpart[((int64_t)b * D + c0 + c) * 10 + j] = s;
  }
  float* dxb = dx + (int64_t)b * L * dld + c0;
  for (int i = threadIdx.x; i < L * kCT; i += kEdgeThreads) {         // transposed conv, channels-fastest store
    const int c = i % kCT, p = i / kCT;
    if (c0 + c >= D) continue;
    const int h = p / W, w = p - h * W;
    const float* q = G + c * PP + h * RP + w;     // padded coords: dpre[h' - i + 1][w' - j + 1] = q[(2 - i) * RP + (2 - j)]
    const float* wc = wsm + c * 10;
    float a = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) a = fmaf(wc[r * 3 + s], q[(2 - r) * RP + (2 - s)], a);
    dxb[(int64_t)p * dld + c] = a;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// epilogue
// ------------------------------------------------------------------------------------------------------------------
struct OutGeom {
  int D, H, W, L, TH, TW, tiles_h, tiles_w, DP;
  int64_t pitch;
};

// Walk every (channel, run) of the tile for one pair of scan orders.  kCol = false: runs along w (orders 0 / 2);
// kCol = true: runs along h (orders 1 / 3).  A warp instruction covers R consecutive run elements x 32 / R channels;
// the channel pattern is chosen so that the accesses to the [position][channel] tile (pitch DP == 1 mod 32) spread
// over all banks.  f(d, gpos, tpos) gets the channel, the position in THIS order's sequence, and the tile slot.
template <bool kCol, typename F>
__device__ __forceinline__ void for_each_run(const OutGeom& g, int h0, int w0, F f) {
  const int R = kCol ? g.TH : g.TW;            // run length (power of two <= 8)
  const int O = kCol ? g.TW : g.TH;            // runs per channel
  const int DS = 32 / R;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = lane % R, dsub = lane / R;
  const int dblocks = (g.D + 31) / 32;
  const int dlo_n = kCol ? R : R;              // d = dblk * 32 + (row-runs: dsub * R + dlo | column-runs: dlo * DS + dsub)
  const int total = dlo_n * dblocks * O;
  for (int m = warp; m < total; m += kEdgeThreads / 32) {
    const int dlo = m % dlo_n;
    const int r = m / dlo_n;
    const int dblk = r % dblocks, o = r / dblocks;
    const int d = dblk * 32 + (kCol ? dlo * DS + dsub : dsub * R + dlo);
    const int hh = kCol ? i : o, ww = kCol ? o : i;
    const int h = h0 + hh, w = w0 + ww;
    if (d < g.D && h < g.H && w < g.W) f(d, kCol ? w * g.H + h : h * g.W + w, hh * g.TW + ww);
  }
}

__global__ void __launch_bounds__(kEdgeThreads)
ss2d_out_fwd_kernel(const float* __restrict__ ys, const float* __restrict__ z, int64_t zld, const float* __restrict__ gamma,
                    const float* __restrict__ beta, float eps, float* __restrict__ out, float* __restrict__ xhat,
                    float* __restrict__ rstd_out, OutGeom g) {
  extern __shared__ __align__(16) float T[];
  const int tiles = g.tiles_h * g.tiles_w;
  const int b = blockIdx.x / tiles, t = blockIdx.x - b * tiles;
  const int h0 = (t / g.tiles_w) * g.TH, w0 = (t % g.tiles_w) * g.TW;
  const float* yb = ys + (int64_t)b * 4 * g.D * g.pitch;
  const int64_t dir = (int64_t)g.D * g.pitch;
  const int L = g.L;
  for_each_run<false>(g, h0, w0, [&](int d, int gp, int tp) {
    const float* r = yb + (int64_t)d * g.pitch;
    T[tp * g.DP + d] = __ldg(r + gp) + __ldg(r + 2 * dir + (L - 1 - gp));
  });
  __syncthreads();
  for_each_run<true>(g, h0, w0, [&](int d, int gp, int tp) {
    const float* r = yb + dir + (int64_t)d * g.pitch;
    T[tp * g.DP + d] += __ldg(r + gp) + __ldg(r + 2 * dir + (L - 1 - gp));
  });
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_d = 1.f / (float)g.D;
  for (int tp = warp; tp < g.TH * g.TW; tp += kEdgeThreads / 32) {
    const int h = h0 + tp / g.TW, w = w0 + tp % g.TW;
    if (h >= g.H || w >= g.W) continue;
    const float* row = T + tp * g.DP;
    float s = 0.f;
    for (int d = lane; d < g.D; d += 32) s += row[d];
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
    for (int d = lane; d < g.D; d += 32) {
      const float c = row[d] - mean;
      q = fmaf(c, c, q);
    }
    const float rs = rsqrtf(warp_sum(q) * inv_d + eps);
    const int64_t pos = (int64_t)b * L + h * g.W + w;
    if (rstd_out != nullptr && lane == 0) rstd_out[pos] = rs;
    for (int d = lane; d < g.D; d += 32) {
      const float xh = (row[d] - mean) * rs;
      float v = fmaf(xh, __ldg(gamma + d), __ldg(beta + d));
      if (z != nullptr) {
        const float zz = __ldg(z + pos * zld + d);
        v *= zz * sigmoid_acc(zz);
      }
      out[pos * g.D + d] = v;
      if (xhat != nullptr) xhat[pos * g.D + d] = xh;
    }
  }
}

__global__ void __launch_bounds__(kEdgeThreads)
ss2d_out_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ z, int64_t zld, const float* __restrict__ xhat,
                    const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                    float* __restrict__ dz, int64_t dzld, float* __restrict__ dys, float* __restrict__ part, OutGeom g) {
  extern __shared__ __align__(16) float T[];
  const int P = g.TH * g.TW;
  float* m1 = T + P * g.DP;                    // [P] mean_d(dln * gamma)
  float* m2 = m1 + P;                          // [P] mean_d(dln * gamma * xhat)
  const int tiles = g.tiles_h * g.tiles_w;
  const int b = blockIdx.x / tiles, t = blockIdx.x - b * tiles;
  const int h0 = (t / g.tiles_w) * g.TH, w0 = (t % g.tiles_w) * g.TW;
  const int L = g.L;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float inv_d = 1.f / (float)g.D;
  for (int tp = warp; tp < P; tp += kEdgeThreads / 32) {           // gate backward, LayerNorm row moments
    const int h = h0 + tp / g.TW, w = w0 + tp % g.TW;
    float* row = T + tp * g.DP;
    if (h >= g.H || w >= g.W) {
      for (int d = lane; d < g.D; d += 32) row[d] = 0.f;
      continue;
    }
    const int64_t pos = (int64_t)b * L + h * g.W + w;
    float a1 = 0.f, a2 = 0.f;
    for (int d = lane; d < g.D; d += 32) {
      const float go = __ldg(gout + pos * g.D + d);
      const float xh = __ldg(xhat + pos * g.D + d);
      const float gm = __ldg(gamma + d);
      float dln = go;
      if (z != nullptr) {
        const float zz = __ldg(z + pos * zld + d);
        const float sg = sigmoid_acc(zz);
        const float lnout = fmaf(xh, gm, __ldg(beta + d));
        dz[pos * dzld + d] = go * lnout * (sg * (1.f + zz * (1.f - sg)));
        dln = go * (zz * sg);
      }
      row[d] = dln;
      const float dy = dln * gm;
      a1 += dy;
      a2 = fmaf(dy, xh, a2);
    }
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
      m1[tp] = a1 * inv_d;
      m2[tp] = a2 * inv_d;
    }
  }
  __syncthreads();
  for (int d = threadIdx.x; d < g.D; d += kEdgeThreads) {         // per-CTA gamma / beta partials (summed by the host)
    float sg = 0.f, sb = 0.f;
    for (int tp = 0; tp < P; ++tp) {
      const int h = h0 + tp / g.TW, w = w0 + tp % g.TW;
      if (h >= g.H || w >= g.W) continue;
      const float dln = T[tp * g.DP + d];
      sb += dln;
      sg = fmaf(dln, __ldg(xhat + ((int64_t)b * L + h * g.W + w) * g.D + d), sg);
    }
    part[(int64_t)blockIdx.x * 2 * g.D + d] = sg;
    part[(int64_t)blockIdx.x * 2 * g.D + g.D + d] = sb;
  }
  __syncthreads();
  for (int tp = warp; tp < P; tp += kEdgeThreads / 32) {           // LayerNorm input gradient, in place
    const int h = h0 + tp / g.TW, w = w0 + tp % g.TW;
    if (h >= g.H || w >= g.W) continue;
    const int64_t pos = (int64_t)b * L + h * g.W + w;
    const float rs = __ldg(rstd + pos), c1 = m1[tp], c2 = m2[tp];
    float* row = T + tp * g.DP;
    for (int d = lane; d < g.D; d += 32) {
      const float xh = __ldg(xhat + pos * g.D + d);
      row[d] = rs * (fmaf(row[d], __ldg(gamma + d), -c1) - xh * c2);
    }
  }
  __syncthreads();
  float* db = dys + (int64_t)b * 4 * g.D * g.pitch;
  const int64_t dir = (int64_t)g.D * g.pitch;
  for_each_run<false>(g, h0, w0, [&](int d, int gp, int tp) {      // CrossMerge backward: every order gets the same value
    const float v = T[tp * g.DP + d];
    float* r = db + (int64_t)d * g.pitch;
    r[gp] = v;
    r[2 * dir + (L - 1 - gp)] = v;
  });
  for_each_run<true>(g, h0, w0, [&](int d, int gp, int tp) {
    const float v = T[tp * g.DP + d];
    float* r = db + dir + (int64_t)d * g.pitch;
    r[gp] = v;
    r[2 * dir + (L - 1 - gp)] = v;
  });
}

struct InPlan {
  int ct, RP, PP;
  size_t smem;
};

// planes per CTA: 8 channels when `nplanes` padded planes of 8 channels leave room for two CTAs per SM, else 4
InPlan plan_in(int H, int W, int nplanes) {
  InPlan p;
  p.RP = (W + 2) | 1;
  for (p.ct = 8;; p.ct = 4) {
    const int want = 32 / p.ct;
    int pp = (H + 2) * p.RP;
    pp += ((want - pp % 32) % 32 + 32) % 32;
    p.PP = pp;
    p.smem = ((size_t)nplanes * p.ct * pp + p.ct * 10 + 10 * (kEdgeThreads / 32)) * sizeof(float);
    if (p.ct == 4 || p.smem <= 112 * 1024) break;
  }
  return p;
}

bool plan_out(int D, int H, int W, int64_t pitch, OutGeom* g, size_t* smem) {
  g->D = D; g->H = H; g->W = W; g->L = H * W; g->pitch = pitch;
  g->DP = D + ((1 - D % 32) + 32) % 32;
  static const int shapes[][2] = {{8, 8}, {8, 4}, {4, 4}, {4, 2}, {2, 2}, {2, 1}, {1, 1}};
  for (const auto& s : shapes) {
    const size_t need = ((size_t)s[0] * s[1] * g->DP + 2 * s[0] * s[1]) * sizeof(float);
    if (need <= 100 * 1024 || (s[0] == 1 && s[1] == 1 && need <= 220 * 1024)) {
      g->TH = s[0]; g->TW = s[1];
      g->tiles_h = (H + g->TH - 1) / g->TH;
      g->tiles_w = (W + g->TW - 1) / g->TW;
      *smem = need;
      return true;
    }
  }
  return false;
}

template <typename K>
cudaError_t allow_smem(K kernel, size_t smem) {
  if (smem <= 48 * 1024) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

}  // namespace

bool ss2d_in_supported(int H, int W) { return plan_in(H, W, 2).smem <= 220 * 1024; }

cudaError_t launch_ss2d_in_fwd(const float* x, int64_t ld, const float* cw, const float* cb, float* xs, int B, int D, int H, int W,
                               int64_t pitch, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  const InPlan p = plan_in(H, W, 1);
  const unsigned grid = (unsigned)((int64_t)B * ((D + p.ct - 1) / p.ct));
  cudaError_t e;
  if (p.ct == 8) {
    if ((e = allow_smem(ss2d_in_fwd_kernel<8>, p.smem)) != cudaSuccess) return e;
    ss2d_in_fwd_kernel<8><<<grid, kEdgeThreads, p.smem, stream>>>(x, ld, cw, cb, xs, D, H, W, p.RP, p.PP, pitch);
  } else {
    if ((e = allow_smem(ss2d_in_fwd_kernel<4>, p.smem)) != cudaSuccess) return e;
    ss2d_in_fwd_kernel<4><<<grid, kEdgeThreads, p.smem, stream>>>(x, ld, cw, cb, xs, D, H, W, p.RP, p.PP, pitch);
  }
  return cudaGetLastError();
}

cudaError_t launch_ss2d_in_bwd(const float* dxs, const float* x, int64_t ld, const float* cw, const float* cb, float* dx, int64_t dld,
                               float* wpart, int B, int D, int H, int W, int64_t pitch, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  const InPlan p = plan_in(H, W, 2);
  const unsigned grid = (unsigned)((int64_t)B * ((D + p.ct - 1) / p.ct));
  cudaError_t e;
  if (p.ct == 8) {
    if ((e = allow_smem(ss2d_in_bwd_kernel<8>, p.smem)) != cudaSuccess) return e;
    ss2d_in_bwd_kernel<8><<<grid, kEdgeThreads, p.smem, stream>>>(dxs, x, ld, cw, cb, dx, dld, wpart, D, H, W, p.RP, p.PP, pitch);
  } else {
    if ((e = allow_smem(ss2d_in_bwd_kernel<4>, p.smem)) != cudaSuccess) return e;
    ss2d_in_bwd_kernel<4><<<grid, kEdgeThreads, p.smem, stream>>>(dxs, x, ld, cw, cb, dx, dld, wpart, D, H, W, p.RP, p.PP, pitch);
  }
  return cudaGetLastError();
}

bool ss2d_out_supported(int D) {
  OutGeom g;
  size_t smem;
  return plan_out(D, 8, 8, 64, &g, &smem);
}

int64_t ss2d_out_ctas(int B, int D, int H, int W) {
  OutGeom g;
  size_t smem;
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || !plan_out(D, H, W, (int64_t)H * W, &g, &smem)) return 0;
  return (int64_t)B * g.tiles_h * g.tiles_w;
}

cudaError_t launch_ss2d_out_fwd(const float* ys, int64_t pitch, const float* z, int64_t zld, const float* gamma, const float* beta,
                                float eps, float* out, float* xhat, float* rstd, int B, int D, int H, int W, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  OutGeom g;
  size_t smem;
  if (!plan_out(D, H, W, pitch, &g, &smem)) return cudaErrorInvalidValue;
  cudaError_t e = allow_smem(ss2d_out_fwd_kernel, smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((int64_t)B * g.tiles_h * g.tiles_w);
  ss2d_out_fwd_kernel<<<grid, kEdgeThreads, smem, stream>>>(ys, z, zld, gamma, beta, eps, out, xhat, rstd, g);
  return cudaGetLastError();
}

cudaError_t launch_ss2d_out_bwd(const float* gout, const float* z, int64_t zld, const float* xhat, const float* rstd,
                                const float* gamma, const float* beta, float* dz, int64_t dzld, float* dys, int64_t pitch,
                                float* part, int B, int D, int H, int W, cudaStream_t stream) {
  if (B == 0 || D == 0 || H * W == 0) return cudaSuccess;
  OutGeom g;
  size_t smem;
  if (!plan_out(D, H, W, pitch, &g, &smem)) return cudaErrorInvalidValue;
  cudaError_t e = allow_smem(ss2d_out_bwd_kernel, smem);
  if (e != cudaSuccess) return e;
  const unsigned grid = (unsigned)((int64_t)B * g.tiles_h * g.tiles_w);
  ss2d_out_bwd_kernel<<<grid, kEdgeThreads, smem, stream>>>(gout, z, zld, xhat, rstd, gamma, beta, dz, dzld, dys, part, g);
  return cudaGetLastError();
}

}  // namespace selscan
