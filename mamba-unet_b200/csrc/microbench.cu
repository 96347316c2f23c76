// Pipe-throughput microbenchmarks for the selective-scan design (standalone binary, B200 only).
// Answers the questions DESIGN.md's issue-rate model depends on: FFMA vs packed FFMA2 (fma.rn.f32x2) lane
// throughput, MUFU.EX2 throughput, how well MUFU co-issues with FP32 work, and the shared-memory -> register return path
// (128-bit broadcast loads, shuffles: reported as bytes received per thread, x32 lanes = bytes/clk/SM) -- all per SM per clock,
// measured with clock64() inside the kernel so the result does not depend on the (unknown) SM clock.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu && ./microbench
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));          \
      exit(1);                                                                            \
    }                                                                                     \
  } while (0)

constexpr int kThreads = 256;
constexpr int kIters = 2048;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm volatile(
      "{ .reg .b64 ra, rb, rd;\n"
      "  mov.b64 ra, {%2, %3};\n  mov.b64 rb, {%4, %5};\n  mov.b64 rd, {%0, %1};\n"
      "  fma.rn.f32x2 rd, ra, rb, rd;\n"
      "  mov.b64 {%0, %1}, rd; }\n"
      : "+f"(d0), "+f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// MODE 0: FFMA x8 chains; 1: FFMA2 x8 packed chains; 2: MUFU x8; 3..: 1 MUFU + K FFMA; 20+: 1 MUFU + K FFMA2
template <int MODE, int K>
__global__ void __launch_bounds__(kThreads) bench(float* out, long long* cycles, float seed) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = seed + 0.001f * (threadIdx.x + i);
  const float a = 0.9999f + seed * 1e-6f, b = seed * 1e-7f;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) ffma2(v[i], v[i + 1], v[i], v[i + 1], a, a);
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = ex2(v[i]);
    } else if (MODE == 3) {  // 1 MUFU + K FFMA per "state step", 4 steps unrolled
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const float e = ex2(v[8 + s]);
        v[8 + s] = e * -0.5f;
#pragma unroll
        for (int i = 0; i < K - 1; ++i) v[(s * 2 + i) & 7] = fmaf(v[(s * 2 + i) & 7], e, b);
      }
    } else if (MODE == 4) {  // 2 MUFU + K FFMA2 per packed pair of state steps, 2 pairs unrolled
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const float e0 = ex2(v[8 + 2 * s]), e1 = ex2(v[9 + 2 * s]);
        v[8 + 2 * s] = e0 * -0.5f;
        v[9 + 2 * s] = e1 * -0.5f;
#pragma unroll
        for (int i = 0; i < K - 1; ++i) ffma2(v[(2 * i) & 7], v[(2 * i + 1) & 7], v[(2 * i) & 7], v[(2 * i + 1) & 7], e0, e1);
      }
    } else if (MODE >= 10 && MODE <= 15) {
      // The scan kernels' own access patterns.  10: LDS.128, lane reads chunk (lane & 3): 4 distinct 16-byte chunks per warp, each
      // shared by 8 lanes (the B / C loads: 4 state quads x 8 channel pairs); 11: LDS.128, lane reads chunk (lane >> 2): 8 distinct
      // chunks, each shared by 4 lanes (the per-channel row data); 12: STS.128, 32 distinct chunks (the product tile stores);
      // 13 / 14 / 15: other lane -> chunk maps (8 consecutive lanes share a chunk; chunk = lane & 7; 16 consecutive lanes share).
      extern __shared__ float4 sh4[];
      const int lane = threadIdx.x & 31;
      const unsigned base = (unsigned)__cvta_generic_to_shared(sh4);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int blk = ((it + i) & 7) * 32;
        if (MODE == 12) {
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base + (blk + lane) * 16), "f"(v[i]), "f"(v[i + 8]), "f"(v[i]), "f"(v[i + 8]) : "memory");
        } else {
          const int idx = blk + (MODE == 10 ? (lane & 3) : MODE == 11 ? (lane >> 2) : MODE == 13 ? (lane >> 3) : MODE == 14 ? (lane & 7) : (lane >> 4));
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w) : "r"(base + idx * 16) : "memory");
          v[i] += x.x + x.w;
          v[i + 8] += x.y + x.z;
        }
      }
    } else if (MODE >= 6 && MODE <= 8) {
      // Shared-memory -> register return path.  6: LDS.128, the whole warp reads the SAME 16 bytes (pure broadcast);
      // 7: LDS.128, 32 distinct 16-byte chunks (512 contiguous bytes); 8: LDS.32, the whole warp reads the same word.
      // Counted in bytes RECEIVED per thread: if the limit is the 128 B/clk/SM return path, 6 and 7 both give 128 B/clk/SM
      // (a broadcast 128-bit load is not cheaper than a distinct one) and 8 gives 128 B/clk/SM as well (1 instruction per clock).
      extern __shared__ float4 sh4[];
      const int lane = threadIdx.x & 31;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int idx = (MODE == 7) ? ((it + i) & 7) * 32 + lane : ((it + i) & 255);
        if (MODE == 8) {
          float x;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"((unsigned)__cvta_generic_to_shared(sh4) + idx * 4) : "memory");
          v[i] += x;
        } else {
          float4 x;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x.x), "=f"(x.y), "=f"(x.z), "=f"(x.w)
                       : "r"((unsigned)__cvta_generic_to_shared(sh4) + idx * 16)
                       : "memory");   // the clobber keeps the unrolled loop from merging identical loads of two iterations
          v[i] += x.x + x.w;
          v[i + 8] += x.y + x.z;
        }
      }
    } else if (MODE == 9) {   // SHFL: 4 bytes per thread per instruction through the same crossbar
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[(i + 1) & 7], 1 + (it & 3));
    } else if (MODE == 5) {  // the forward inner body, scalar: t=dl*A2; e=ex2(t); x=e*x+du*B; y+=C*x  (16 states)
      const float dl = v[15] * 1e-3f, du = v[14];
      float y = 0.f;
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        const float e = ex2(dl * (-1.f - n));
        v[n] = fmaf(e, v[n], du * a);
        y = fmaf(b, v[n], y);
      }
      v[14] = y;
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 12345.678f) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int K>
void run(const char* name, double ops_per_iter_per_thread, int ctas_per_sm) {
  int sms = 0;
  CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = sms * ctas_per_sm;
  float* out;
  long long* cyc;
  CHECK(cudaMalloc(&out, 4));
  CHECK(cudaMalloc(&cyc, grid * sizeof(long long)));
  const int smem = ((MODE >= 6 && MODE <= 8) || (MODE >= 10 && MODE <= 15)) ? 8 * 32 * 16 : 0;
  bench<MODE, K><<<grid, kThreads, smem>>>(out, cyc, 0.5f);
  CHECK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  CHECK(cudaEventCreate(&e0));
  CHECK(cudaEventCreate(&e1));
  CHECK(cudaEventRecord(e0));
  bench<MODE, K><<<grid, kThreads, smem>>>(out, cyc, 0.5f);
  CHECK(cudaEventRecord(e1));
  CHECK(cudaDeviceSynchronize());
  float ms = 0;
  CHECK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<long long> h(grid);
  CHECK(cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
  double mean = 0;
  for (long long c : h) mean += (double)c;
  mean /= grid;
  const double per_sm_clk = ops_per_iter_per_thread * kIters * kThreads * ctas_per_sm / mean;
  const double ghz = mean / (ms * 1e6);
  // second figure: from the event time of the whole launch and the SM clock under load (max clock; the CTA-local clock64 window
  // over-counts when several CTAs per SM do not run fully concurrently)
  int khz = 0;
  CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double per_sm_clk_t = ops_per_iter_per_thread * kIters * kThreads * ctas_per_sm / (ms * 1e-3 * khz * 1e3);
  printf("%-34s ctas/sm=%d  %8.2f ops/clk/SM (CTA clock64)  %8.2f (launch time @ %.2f GHz)   (%.0f cycles, %.3f ms, ~%.2f GHz)\n", name,
         ctas_per_sm, per_sm_clk, per_sm_clk_t, khz * 1e-6, mean, ms, ghz);
  CHECK(cudaFree(out));
  CHECK(cudaFree(cyc));
}

int main() {
  cudaDeviceProp prop;
  CHECK(cudaGetDeviceProperties(&prop, 0));
  printf("device: %s, %d SMs, cc %d.%d\n", prop.name, prop.multiProcessorCount, prop.major, prop.minor);
  for (int occ : {1, 2, 4}) {
    run<0, 0>("FFMA (3-reg), thread-FMAs", 16, occ);
    run<1, 0>("FFMA2 packed, thread-FMAs (x2)", 16, occ);
    run<2, 0>("MUFU.EX2", 8, occ);
    run<3, 3>("1 MUFU + 3 FP (count: steps)", 4, occ);
    run<3, 5>("1 MUFU + 5 FP (count: steps)", 4, occ);
    run<3, 7>("1 MUFU + 7 FP (count: steps)", 4, occ);
    run<3, 9>("1 MUFU + 9 FP (count: steps)", 4, occ);
    run<4, 3>("2 MUFU + 3 FP2 (count: steps)", 4, occ);
    run<4, 5>("2 MUFU + 5 FP2 (count: steps)", 4, occ);
    run<5, 0>("fwd body scalar (count: steps)", 8, occ);
    run<6, 0>("LDS.128 broadcast, bytes/thread", 8 * 16, occ);
    run<7, 0>("LDS.128 distinct, bytes/thread", 8 * 16, occ);
    run<8, 0>("LDS.32 broadcast, bytes/thread", 8 * 4, occ);
    run<9, 0>("SHFL.BFLY, bytes/thread", 8 * 4, occ);
    run<10, 0>("LDS.128 4 chunks x8 lanes, B/thr", 8 * 16, occ);
    run<11, 0>("LDS.128 8 chunks x4 lanes, B/thr", 8 * 16, occ);
    run<12, 0>("STS.128 distinct, bytes/thread", 8 * 16, occ);
    run<13, 0>("LDS.128 chunk = lane>>3 (4 chunks)", 8 * 16, occ);
    run<14, 0>("LDS.128 chunk = lane&7  (8 chunks)", 8 * 16, occ);
    run<15, 0>("LDS.128 chunk = lane>>4 (2 chunks)", 8 * 16, occ);
  }
  return 0;
}
