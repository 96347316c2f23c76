// TMEM as per-thread scratch: round-trip bandwidth of tcgen05.st / tcgen05.ld (32x32b.x32: thread i of warp w <-> TMEM lane 32*(w%4)+i,
// 32 consecutive columns).  Question: can the backward scan keep its per-chunk saved states / decays in TMEM instead of registers?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu && ./tmem_bw
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define R32(p) p##0, p##1, p##2, p##3, p##4, p##5, p##6, p##7, p##8, p##9, p##10, p##11, p##12, p##13, p##14, p##15, p##16, p##17, p##18, \
               p##19, p##20, p##21, p##22, p##23, p##24, p##25, p##26, p##27, p##28, p##29, p##30, p##31

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
      "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

// MODE 0: st + wait + ld + wait per iteration (round trip).  1: ld only (4 loads in flight per wait).  2: st only.
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, long long* cyc, int iters, int cols_per_cta) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(cols_per_cta));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  // warps w and w+4 share a lane quarter: give them different column ranges
  const uint32_t taddr = tbase + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * (cols_per_cta / 2));
  uint32_t v[32], acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 33 + i;
  tmem_st32(taddr, v);
  tmem_st32(taddr + 32, v);
  asm volatile("tcgen05.wait::st.sync.aligned;");
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      tmem_st32(taddr + (it & 1) * 32, v);
      asm volatile("tcgen05.wait::st.sync.aligned;");
      tmem_ld32(taddr + (it & 1) * 32, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      v[0] += it;
    } else if (MODE == 1) {
      uint32_t a[32], b[32];
      tmem_ld32(taddr, a);
      tmem_ld32(taddr + 32, b);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
      for (int i = 0; i < 32; ++i) acc += a[i] ^ b[i];
    } else {
      tmem_st32(taddr, v);
      tmem_st32(taddr + 32, v);
      asm volatile("tcgen05.wait::st.sync.aligned;");
      v[1] += it;
    }
  }
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(cols_per_cta));
}

template <int MODE>
void run(const char* name, int threads, int ctas_per_sm, double bytes_per_thread_iter) {
  const int iters = 4096, sms = 148;
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, sizeof(uint32_t) * sms * ctas_per_sm * threads);
  cudaMalloc(&cyc, sizeof(long long) * sms * ctas_per_sm);
  const int cols = ctas_per_sm == 1 ? 256 : 128;
  k<MODE><<<sms * ctas_per_sm, threads>>>(out, cyc, iters, cols);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[148 * 4];
  cudaMemcpy(h, cyc, sizeof(long long) * sms * ctas_per_sm, cudaMemcpyDeviceToHost);
  double mx = 0;
  for (int i = 0; i < sms * ctas_per_sm; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-28s threads=%3d ctas/sm=%d : %8.1f cycles/iter, %7.1f B/clk/SM (per direction where both)\n", name, threads, ctas_per_sm, mx / iters,
         bytes_per_thread_iter * threads * ctas_per_sm * iters / mx);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int thr : {128, 256}) {
    run<0>("st+wait+ld+wait x32", thr, 1, 128.0);
    run<1>("ld x32 x2 + wait", thr, 1, 256.0);
    run<2>("st x32 x2 + wait", thr, 1, 256.0);
  }
  run<0>("st+wait+ld+wait x32", 128, 2, 128.0);
  run<1>("ld x32 x2 + wait", 128, 2, 256.0);
  run<2>("st x32 x2 + wait", 128, 2, 256.0);
  return 0;
}
