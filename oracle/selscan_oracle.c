/*
 * selscan_oracle.c -- CPU restatement of the reference selective scan (forward + analytic backward).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the sm_100a kernels.  Nothing in the
 * product path (mamba-unet_b200/) may import, link or call it; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs do.
 *
 * What it restates (all file:line are into /root/reference):
 *   forward   mamba/mamba_ssm/ops/selective_scan_interface.py:86-152   (selective_scan_ref)
 *               :104-107  delta = softplus(delta + delta_bias)
 *               :121      deltaA = exp(delta * A)
 *               :125-129  deltaB_u = delta * B * u, group g = d / (dim / G)   (:128, kernel fwd_kernel.cuh:99)
 *               :133-146  x = deltaA*x + deltaB_u ; y = sum_n x*C
 *               :148      out = y + u*D
 *               :149-150  out = out * silu(z)
 *               :142-143  last_state = x after position L-1
 *   backward  the identities the reference CUDA kernel implements,
 *             mamba/csrc/selective_scan/selective_scan_bwd_kernel.cuh:278-296 (du, ddelta, dA, dB, dC),
 *             :186-191 (dz, gated dout), :439-452 (softplus'), :467-475 (dD, ddelta_bias);
 *             equivalently the autograd derivative of selective_scan_ref.
 *
 * Pinning: tests/golden/*.npz hold outputs + autograd gradients of the reference's own
 * selective_scan_ref (imported from /root/reference in the build container by
 * tests/golden/make_golden.py); tests/test_oracle.py checks this file against them.
 *
 * Layouts (all contiguous fp32):  u, delta, z, dout, out: (batch, dim, L);  A: (dim, N);
 * B, C: (batch, G, N, L);  D, delta_bias: (dim);  last_state: (batch, dim, N).
 * `real` selects the arithmetic: double (ground truth) or float (same rounding class as the kernels).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* F.softplus(x) with the default threshold 20 (selective_scan_interface.py:107; kernel: fwd_kernel.cuh:155) */
static inline double softplus_d(double x) { return x <= 20.0 ? log1p(exp(x)) : x; }
static inline float softplus_f(float x) { return x <= 20.f ? log1pf(expf(x)) : x; }
static inline double sigmoid_d(double x) { return 1.0 / (1.0 + exp(-x)); }
static inline float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

#define DEFINE_ORACLE(SUFFIX, real, EXP, SOFTPLUS, SIGMOID)                                              \
  /* forward: selective_scan_interface.py:101-152 */                                                      \
  static void fwd_##SUFFIX(const float* u, const float* delta, const float* A, const float* B,           \
                           const float* C, const float* D, const float* z, const float* delta_bias,      \
                           int delta_softplus, int batch, int dim, int L, int N, int G, float* out,      \
                           float* last_state) {                                                          \
    const int dpg = dim / G;                                                                             \
    _Pragma("omp parallel for collapse(2) schedule(static)")                                             \
    for (int b = 0; b < batch; ++b)                                                                      \
      for (int d = 0; d < dim; ++d) {                                                                    \
        const int g = d / dpg;                                                                           \
        const float* ur = u + ((size_t)b * dim + d) * L;                                                 \
        const float* dr = delta + ((size_t)b * dim + d) * L;                                             \
        const float* zr = z ? z + ((size_t)b * dim + d) * L : NULL;                                      \
        const float* Bg = B + ((size_t)b * G + g) * N * L;                                               \
        const float* Cg = C + ((size_t)b * G + g) * N * L;                                               \
        float* orow = out + ((size_t)b * dim + d) * L;                                                   \
        real x[256];                                                                                     \
        for (int n = 0; n < N; ++n) x[n] = 0;                                                            \
        const real bias = delta_bias ? (real)delta_bias[d] : (real)0;                                    \
        for (int l = 0; l < L; ++l) {                                                                    \
          real dl = (real)dr[l] + bias;                                      /* :105 */                  \
          if (delta_softplus) dl = SOFTPLUS(dl);                             /* :107 */                  \
          const real uu = (real)ur[l];                                                                   \
          real y = 0;                                                                                    \
          for (int n = 0; n < N; ++n) {                                                                  \
            const real a = EXP(dl * (real)A[(size_t)d * N + n]);             /* :121 */                  \
            x[n] = a * x[n] + dl * (real)Bg[(size_t)n * L + l] * uu;         /* :129,:134 */             \
            y += x[n] * (real)Cg[(size_t)n * L + l];                         /* :141 */                  \
          }                                                                                              \
          if (D) y += uu * (real)D[d];                                       /* :148 */                  \
          if (zr) { const real zz = (real)zr[l]; y = y * (zz * SIGMOID(zz)); } /* :150 */                \
          orow[l] = (float)y;                                                                            \
        }                                                                                                \
        if (last_state)                                                      /* :142-143 */              \
          for (int n = 0; n < N; ++n) last_state[((size_t)b * dim + d) * N + n] = (float)x[n];           \
      }                                                                                                  \
  }                                                                                                      \
  /* backward: bwd_kernel.cuh:186-191,211-213,278-296,439-452,467-475 */                                  \
  static void bwd_##SUFFIX(const float* u, const float* delta, const float* A, const float* B,           \
                           const float* C, const float* D, const float* z, const float* delta_bias,      \
                           const float* dout, int delta_softplus, int batch, int dim, int L, int N,      \
                           int G, float* du, float* ddelta, float* dA, float* dB, float* dC, float* dD,  \
                           float* dz, float* ddelta_bias) {                                              \
    const int dpg = dim / G;                                                                             \
    /* accumulate the reductions in `real`, one accumulator set per (b, g) for dB/dC and per */          \
    /* d for dA/dD/dbias; parallelise over (b, g) so that no two threads share an accumulator */         \
    real* dA_acc = (real*)calloc((size_t)batch * dim * N, sizeof(real));                                 \
    real* dD_acc = (real*)calloc((size_t)batch * dim, sizeof(real));                                     \
    real* db_acc = (real*)calloc((size_t)batch * dim, sizeof(real));                                     \
    _Pragma("omp parallel for collapse(2) schedule(dynamic)")                                            \
    for (int b = 0; b < batch; ++b)                                                                      \
      for (int g = 0; g < G; ++g) {                                                                      \
        const float* Bg = B + ((size_t)b * G + g) * N * L;                                               \
        const float* Cg = C + ((size_t)b * G + g) * N * L;                                               \
        real* dBg = (real*)calloc((size_t)N * L, sizeof(real));                                          \
        real* dCg = (real*)calloc((size_t)N * L, sizeof(real));                                          \
        real* xs = (real*)malloc((size_t)N * L * sizeof(real));    /* x_{l,n} of the current row */      \
        real* dls = (real*)malloc((size_t)L * sizeof(real));       /* discretised delta */               \
        real* dys = (real*)malloc((size_t)L * sizeof(real));       /* (gated) dout */                    \
        for (int dd = 0; dd < dpg; ++dd) {                                                               \
          const int d = g * dpg + dd;                                                                    \
          const size_t row = ((size_t)b * dim + d) * L;                                                  \
          const float* ur = u + row;                                                                     \
          const float* dr = delta + row;                                                                 \
          const float* zr = z ? z + row : NULL;                                                          \
          const float* gr = dout + row;                                                                  \
          const real bias = delta_bias ? (real)delta_bias[d] : (real)0;                                  \
          const real Dd = D ? (real)D[d] : (real)0;                                                      \
          real x[256];                                                                                   \
          for (int n = 0; n < N; ++n) x[n] = 0;                                                          \
          /* forward recompute, keeping every state */                                                   \
          for (int l = 0; l < L; ++l) {                                                                  \
            real dl = (real)dr[l] + bias;                                                                \
            if (delta_softplus) dl = SOFTPLUS(dl);                                                       \
            dls[l] = dl;                                                                                 \
            const real uu = (real)ur[l];                                                                 \
            real y = 0;                                                                                  \
            for (int n = 0; n < N; ++n) {                                                                \
              const real a = EXP(dl * (real)A[(size_t)d * N + n]);                                       \
              x[n] = a * x[n] + dl * (real)Bg[(size_t)n * L + l] * uu;                                   \
              xs[(size_t)n * L + l] = x[n];                                                              \
              y += x[n] * (real)Cg[(size_t)n * L + l];                                                   \
            }                                                                                            \
            real dy = (real)gr[l];                                                                       \
            if (zr) {                      /* bwd_kernel.cuh:186-191 */                                  \
              const real zz = (real)zr[l];                                                               \
              const real sg = SIGMOID(zz);                                                               \
              const real o = y + uu * Dd;  /* ungated out */                                             \
              dz[row + l] = (float)(dy * o * sg * ((real)1 + zz * ((real)1 - sg)));                      \
              dy = dy * zz * sg;                                                                         \
            }                                                                                            \
            dys[l] = dy;                                                                                 \
          }                                                                                              \
          /* reverse sweep: dx_l = C_l*dy_l + a_{l+1}*dx_{l+1}   (bwd_kernel.cuh:243-274) */             \
          real dx[256];                                                                                  \
          for (int n = 0; n < N; ++n) dx[n] = 0;                                                         \
          real dDr = 0, dbr = 0;                                                                         \
          for (int l = L - 1; l >= 0; --l) {                                                             \
            const real dl = dls[l], uu = (real)ur[l], dy = dys[l];                                       \
            real s1 = 0, s2 = 0;                                                                         \
            for (int n = 0; n < N; ++n) {                                                                \
              const real An = (real)A[(size_t)d * N + n];                                                \
              const real Bn = (real)Bg[(size_t)n * L + l], Cn = (real)Cg[(size_t)n * L + l];             \
              const real xl = xs[(size_t)n * L + l];                                                     \
              const real dxl = Cn * dy + dx[n];                                                          \
              const real gprev = xl - dl * Bn * uu;          /* a_l * x_{l-1}  (:283) */                 \
              dCg[(size_t)n * L + l] += dy * xl;             /* :296 */                                  \
              dBg[(size_t)n * L + l] += dxl * dl * uu;       /* :292 */                                  \
              s1 += dxl * Bn;                                                                            \
              s2 += dxl * An * gprev;                                                                    \
              dA_acc[((size_t)b * dim + d) * N + n] += dxl * dl * gprev; /* :286 */                      \
              dx[n] = EXP(dl * An) * dxl;                    /* carried to position l-1 */               \
            }                                                                                            \
            du[row + l] = (float)(Dd * dy + dl * s1);        /* :211,:280 */                             \
            real ddl = uu * s1 + s2;                         /* :281-284 */                              \
            if (delta_softplus) ddl = ddl * SIGMOID((real)dr[l] + bias);  /* :446-450 */                 \
            ddelta[row + l] = (float)ddl;                                                                \
            dbr += ddl;                                                                                  \
            dDr += dy * uu;                                  /* :213 */                                  \
          }                                                                                              \
          dD_acc[(size_t)b * dim + d] = dDr;                                                             \
          db_acc[(size_t)b * dim + d] = dbr;                                                             \
        }                                                                                                \
        for (size_t i = 0; i < (size_t)N * L; ++i) {                                                     \
          dB[((size_t)b * G + g) * N * L + i] = (float)dBg[i];                                           \
          dC[((size_t)b * G + g) * N * L + i] = (float)dCg[i];                                           \
        }                                                                                                \
        free(dBg); free(dCg); free(xs); free(dls); free(dys);                                            \
      }                                                                                                  \
    for (int d = 0; d < dim; ++d) {                                                                      \
      real sD = 0, sb = 0;                                                                               \
      for (int b = 0; b < batch; ++b) { sD += dD_acc[(size_t)b * dim + d]; sb += db_acc[(size_t)b * dim + d]; } \
      if (dD) dD[d] = (float)sD;                                                                         \
      if (ddelta_bias) ddelta_bias[d] = (float)sb;                                                       \
      for (int n = 0; n < N; ++n) {                                                                      \
        real s = 0;                                                                                      \
        for (int b = 0; b < batch; ++b) s += dA_acc[((size_t)b * dim + d) * N + n];                      \
        dA[(size_t)d * N + n] = (float)s;                                                                \
      }                                                                                                  \
    }                                                                                                    \
    free(dA_acc); free(dD_acc); free(db_acc);                                                            \
  }

DEFINE_ORACLE(f64, double, exp, softplus_d, sigmoid_d)
DEFINE_ORACLE(f32, float, expf, softplus_f, sigmoid_f)

/* precision: 64 = double arithmetic (ground truth), 32 = float arithmetic.  Returns 0, or -1 on bad args. */
ORACLE_API int selscan_oracle_fwd(const float* u, const float* delta, const float* A, const float* B,
                                  const float* C, const float* D, const float* z,
                                  const float* delta_bias, int delta_softplus, int batch, int dim, int L,
                                  int N, int G, float* out, float* last_state, int precision) {
  if (batch < 0 || dim <= 0 || L < 0 || N <= 0 || N > 256 || G <= 0 || dim % G) return -1;
  if (precision == 64)
    fwd_f64(u, delta, A, B, C, D, z, delta_bias, delta_softplus, batch, dim, L, N, G, out, last_state);
  else
    fwd_f32(u, delta, A, B, C, D, z, delta_bias, delta_softplus, batch, dim, L, N, G, out, last_state);
  return 0;
}

ORACLE_API int selscan_oracle_bwd(const float* u, const float* delta, const float* A, const float* B,
                                  const float* C, const float* D, const float* z,
                                  const float* delta_bias, const float* dout, int delta_softplus,
                                  int batch, int dim, int L, int N, int G, float* du, float* ddelta,
                                  float* dA, float* dB, float* dC, float* dD, float* dz,
                                  float* ddelta_bias, int precision) {
  if (batch < 0 || dim <= 0 || L < 0 || N <= 0 || N > 256 || G <= 0 || dim % G) return -1;
  if (z && !dz) return -1;
  if (precision == 64)
    bwd_f64(u, delta, A, B, C, D, z, delta_bias, dout, delta_softplus, batch, dim, L, N, G, du, ddelta,
            dA, dB, dC, dD, dz, ddelta_bias);
  else
    bwd_f32(u, delta, A, B, C, D, z, delta_bias, dout, delta_softplus, batch, dim, L, N, G, du, ddelta,
            dA, dB, dC, dD, dz, ddelta_bias);
  return 0;
}

ORACLE_API int selscan_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
