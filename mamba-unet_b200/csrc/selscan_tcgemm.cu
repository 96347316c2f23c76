// fp32 GEMM on the 5th-generation tensor cores (tcgen05 / TMEM / TMA, sm_100a) with the 3xTF32 split:
//   C (M x N) (+)= A (M x K) * B (N x K)^T,   a = a_hi + a_lo,  a_hi = tf32(a),  a_lo = a - a_hi (exact),
//   a * b ~= a_lo * b_hi + a_hi * b_lo + a_hi * b_hi   (the dropped a_lo * b_lo term is 2^-22 relative), fp32 accumulation in TMEM.
// This is the arithmetic cuBLAS does NOT offer for fp32 (its tensor-core mode is a single TF32 product, 2^-11 relative): the
// result agrees with an fp32 SIMT GEMM to ~1e-6 relative, at several times its speed.  Used, opt-in, for the GEMMs around the
// scan (SS2D.in_proj / out_proj / x_proj: code/networks/mamba_sys.py:299,336,406) and their dgrad / wgrad forms.
//
// Both operands may be stored K-major (row = M or N index, K contiguous) or MN-major (row = K index, M or N contiguous), so
// that y = x W^T, dx = dy W and dW = dy^T x all run without a transpose pass.
//
// Persistent CTAs (one per SM) walk the 128 x BN tiles of C (BN <= 128); 10 warps, every hand-over is an mbarrier:
//   warp 0      TMA producer: 128-byte-swizzled boxes of 32 k per stage into a 3-stage ring (OOB rows / k are zero fill)
//   warps 2-5   split a stage in place into hi (tf32-exact) and lo tiles (generic proxy -> fence.proxy.async)
//   warp 1      one lane issues 12 tcgen05.mma.kind::tf32 per stage (4 k-steps x 3 products) into one of TWO TMEM
//               accumulators, commits to the stage's empty barrier and, after a tile's last k-block, to acc_full
//   warps 6-9   epilogue of tile j while tile j+1 is being multiplied: tcgen05.ld 32 lanes x 32 columns -> swizzled staging
//               tile -> TMA store, or TMA reduce-add for C += and for split-K (weight gradients: K = batch * L)
#include <atomic>

#include <cuda_runtime.h>
#include <stdint.h>

#include "selscan_kernels.h"
#include "selscan_ptx.cuh"
#include "selscan_tma_host.h"

namespace selscan {

namespace {

constexpr int kBM = 128;                 // UMMA M
constexpr int kBK = 32;                  // k per stage: 128 bytes of fp32 = one swizzle row
constexpr int kBNMax = 128;
constexpr int kTcMaxStages = 8;          // the ring is as deep as shared memory allows: 3 stages at BN = 128, 5 at BN = 16
constexpr int kTcThreads = 320;          // warp 0 TMA, warp 1 MMA, warps 2-5 split, warps 6-9 epilogue
constexpr int kTileBytes = kBM * kBK * 4;     // 16 KB: the A tile, and the largest B tile

struct TcParams {
  int M, N, K, BN;
  int a_mn, b_mn;                         // operand storage: 0 = K-major, 1 = MN-major
  int m_tiles, n_tiles, batch;
  int a_bmod, b_bmod, c_bmod;             // batch coordinate of an operand = bz % mod (0: bz): weights shared across images, sums over images
  int kb_total, kb_per_split, splits;     // k-blocks of kBK
  int stages;                             // depth of the operand ring
  uint32_t stage_bytes, b_off;            // one stage: [a_hi 16K][a_lo 16K][b_hi][b_lo]; b_off = bytes between b_hi and b_lo
  int reduce;                             // 0: C = tile (TMA store), 1: C += tile (TMA reduce-add: accumulate and / or split-K)
  uint32_t idesc;
};

struct TcSmem {                           // followed by the operand ring (runtime depth)
  float out[4][2][32 * 32];               // per epilogue warp: two 32 x 32 staging tiles for the TMA stores
  u64 full[kTcMaxStages], split[kTcMaxStages], empty[kTcMaxStages], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};
constexpr uint32_t kTcHeader = (sizeof(TcSmem) + 1023) / 1024 * 1024;
constexpr uint32_t kTcSmemMax = 232448 - 1024;    // opt-in dynamic shared memory minus the alignment slack

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), Blackwell version bits.
// layout 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B (the only layout tcgen05 accepts for MN-major tf32 operands)
__device__ __forceinline__ u64 smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (u64)((addr & 0x3FFFF) >> 4) | ((u64)(lbo_bytes >> 4) << 16) | ((u64)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((u64)layout << 61);
}

// round to the 10-bit TF32 mantissa (nearest, ties away) with two full-rate integer ops; cvt.rna.tf32.f32 runs on the
// quarter-rate conversion pipe and made the split pass, not the tensor core, the bottleneck of a stage
__device__ __forceinline__ float rna_tf32(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }

__device__ __forceinline__ void tc_mma(uint32_t tmem_c, u64 da, u64 db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// work item w (n tile fastest, so that CTAs running side by side share the A tile in L2) -> tile coordinates
struct TcTile {
  int m0, n0, bz, kb_begin, n_kb;
};
__device__ __forceinline__ TcTile tc_tile(const TcParams& p, int w) {
  TcTile t;
  const int nt = w % p.n_tiles;
  w /= p.n_tiles;
  const int split = w % p.splits;
  w /= p.splits;
  const int mt = w % p.m_tiles;
  t.bz = w / p.m_tiles;
  t.m0 = mt * kBM;
  t.n0 = nt * p.BN;
  t.kb_begin = split * p.kb_per_split;
  t.n_kb = min(p.kb_total, t.kb_begin + p.kb_per_split) - t.kb_begin;
  return t;
}

__global__ void __launch_bounds__(kTcThreads, 1)
tcgemm_3xtf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ CUtensorMap map_c, const TcParams p) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem_al = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_al);
  const uint32_t ring = smem_u32(smem_al) + kTcHeader;      // stage s: ring + s * stage_bytes
  unsigned char* ring_ptr = smem_al + kTcHeader;
  const int kTcStages = p.stages;
  auto a_hi_addr = [&](int st) { return ring + (uint32_t)st * p.stage_bytes; };
  auto a_lo_addr = [&](int st) { return ring + (uint32_t)st * p.stage_bytes + kTileBytes; };
  auto b_hi_addr = [&](int st) { return ring + (uint32_t)st * p.stage_bytes + 2 * kTileBytes; };
  auto b_lo_addr = [&](int st) { return ring + (uint32_t)st * p.stage_bytes + 2 * kTileBytes + p.b_off; };
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_work = p.m_tiles * p.n_tiles * p.batch * p.splits;
  const int b_chunks = (p.BN + 31) / 32;
  const uint32_t b_bytes = p.b_mn ? (uint32_t)b_chunks * 4096u : (uint32_t)p.BN * 128u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcMaxStages; ++s) {
      mbar_init(smem_u32(&sm.full[s]), 1);
      mbar_init(smem_u32(&sm.split[s]), 4);       // lane 0 of each of the four split warps
      mbar_init(smem_u32(&sm.empty[s]), 1);       // tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&sm.acc_full[a]), 1);    // tcgen05.commit after a tile's last k-block
      mbar_init(smem_u32(&sm.acc_empty[a]), 4);   // lane 0 of each epilogue warp once its TMEM reads are done
    }
    mbar_fence_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_c);
  }
  if (warp == 1) {                                // TMEM: two accumulator tiles of 128 lanes x 128 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "n"(2 * kBNMax) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ------------------------------------------------ TMA producer ------------------------------------------------
    if (lane == 0) {
      int it = 0;
      for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
        const TcTile t = tc_tile(p, w);
        for (int i = 0; i < t.n_kb; ++i, ++it) {
          const int s = it % kTcStages, k = it / kTcStages;
          if (k > 0) mbar_wait(smem_u32(&sm.empty[s]), (k - 1) & 1);
          const uint32_t full = smem_u32(&sm.full[s]);
          const int k0 = (t.kb_begin + i) * kBK;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((uint32_t)kTileBytes + b_bytes) : "memory");
          const int za = p.a_bmod ? t.bz % p.a_bmod : t.bz, zb = p.b_bmod ? t.bz % p.b_bmod : t.bz;
          if (p.a_mn) {
#pragma unroll
            for (int c = 0; c < kBM / 32; ++c) tma_load_3d(a_hi_addr(s) + c * 4096, &map_a, t.m0 + 32 * c, k0, za, full);
          } else {
            tma_load_3d(a_hi_addr(s), &map_a, k0, t.m0, za, full);
          }
          if (p.b_mn) {
            for (int c = 0; c < b_chunks; ++c) tma_load_3d(b_hi_addr(s) + c * 4096, &map_b, t.n0 + 32 * c, k0, zb, full);
          } else {
            tma_load_3d(b_hi_addr(s), &map_b, k0, t.n0, zb, full);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------ MMA issuer --------------------------------------------------
    // K-major (128B swizzle): atoms of 8 rows x 128 B (SBO 1024); a k-step of 8 is 32 B inside the swizzled row.
    // MN-major (128B swizzle on a 32 B base): [chunk of 32 mn][k][32 mn], atoms of 4 k x 128 B: LBO 4096 between
    // chunks, SBO 512 between groups of 4 k; a k-step of 8 is 1024 B.
    const uint32_t a_step = p.a_mn ? 1024u : 32u, b_step = p.b_mn ? 1024u : 32u;
    const uint32_t a_lbo = p.a_mn ? 4096u : 16u, b_lbo = p.b_mn ? 4096u : 16u;
    const uint32_t a_sbo = p.a_mn ? 512u : 1024u, b_sbo = p.b_mn ? 512u : 1024u;
    const uint32_t a_lt = p.a_mn ? 1u : 2u, b_lt = p.b_mn ? 1u : 2u;
    int it = 0, j = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++j) {
      const TcTile t = tc_tile(p, w);
      const int ab = j & 1, use = j >> 1;
      if (use > 0) mbar_wait(smem_u32(&sm.acc_empty[ab]), (use - 1) & 1);   // the epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t acc = tmem + (uint32_t)ab * kBNMax;
      for (int i = 0; i < t.n_kb; ++i, ++it) {
        const int s = it % kTcStages, k = it / kTcStages;
        mbar_wait(smem_u32(&sm.split[s]), k & 1);
        tc_fence_after();
        if (lane == 0) {
#pragma unroll
          for (int ks = 0; ks < kBK / 8; ++ks) {
            const u64 ah = smem_desc(a_hi_addr(s) + ks * a_step, a_lbo, a_sbo, a_lt);
            const u64 al = smem_desc(a_lo_addr(s) + ks * a_step, a_lbo, a_sbo, a_lt);
            const u64 bh = smem_desc(b_hi_addr(s) + ks * b_step, b_lbo, b_sbo, b_lt);
            const u64 bl = smem_desc(b_lo_addr(s) + ks * b_step, b_lbo, b_sbo, b_lt);
            tc_mma(acc, al, bh, p.idesc, (i > 0 || ks > 0) ? 1u : 0u);
            tc_mma(acc, ah, bl, p.idesc, 1u);
            tc_mma(acc, ah, bh, p.idesc, 1u);
          }
          tc_commit(smem_u32(&sm.empty[s]));      // the stage may be refilled once these MMAs have read it
          if (i + 1 == t.n_kb) tc_commit(smem_u32(&sm.acc_full[ab]));
        }
        __syncwarp();
      }
    }
  } else if (warp < 6) {
    // ------------------------------------------------ split warps -------------------------------------------------
    const int tix = threadIdx.x - 64;             // 0..127
    const int b_vec = (int)(b_bytes >> 4);
    int it = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
      const TcTile t = tc_tile(p, w);
      for (int i = 0; i < t.n_kb; ++i, ++it) {
        const int s = it % kTcStages, k = it / kTcStages;
        mbar_wait(smem_u32(&sm.full[s]), k & 1);
        unsigned char* st = ring_ptr + (size_t)s * p.stage_bytes;
        float4* ah = reinterpret_cast<float4*>(st);
        float4* al = reinterpret_cast<float4*>(st + kTileBytes);
        float4* bh = reinterpret_cast<float4*>(st + 2 * kTileBytes);
        float4* bl = reinterpret_cast<float4*>(st + 2 * kTileBytes + p.b_off);
        auto split4 = [](float4* hi, float4* lo, int q) {   // hi = tf32(x) round-to-nearest (exact operand), lo = x - hi (exact
          float4 v = hi[q], h, l;                             // difference; the tensor core drops its bits below 2^-21 |x|)
          h.x = rna_tf32(v.x); h.y = rna_tf32(v.y); h.z = rna_tf32(v.z); h.w = rna_tf32(v.w);
          l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
          hi[q] = h;
          lo[q] = l;
        };
#pragma unroll
        for (int q = 0; q < kTileBytes / 16 / 128; ++q) split4(ah, al, tix + 128 * q);
        for (int q = tix; q < b_vec; q += 128) split4(bh, bl, q);
        fence_proxy_async_smem();                 // generic-proxy writes -> visible to the tensor core's async-proxy reads
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&sm.split[s]));
      }
    }
  } else {
    // ------------------------------------------------ epilogue warps ----------------------------------------------
    // TMEM lane = tile row; a warp may only touch lanes 32 * (warp % 4) .. + 31.  32 x 32 chunks go through a swizzled
    // staging tile and leave as TMA stores (or reduce-adds), clipped at the edges of C by the tensor map.
    const int q = warp & 3;
    int j = 0, n_store = 0;
    for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++j) {
      const TcTile t = tc_tile(p, w);
      const int ab = j & 1, use = j >> 1;
      mbar_wait(smem_u32(&sm.acc_full[ab]), use & 1);
      tc_fence_after();
      const uint32_t acc = tmem + (uint32_t)ab * kBNMax + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.BN; c0 += 32, ++n_store) {
        uint32_t r[32];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,"
            "%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(acc + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (lane == 0) tma_store_wait_read<1>();  // the staging tile used two stores ago has been read
        __syncwarp();
        const uint32_t stage = smem_u32(sm.out[q][n_store & 1]);
        const uint32_t rowp = stage + (uint32_t)lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)               // 128-byte swizzle of the C tensor map: 16-byte chunk c of row r sits at c ^ (r & 7)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowp + (uint32_t)((c ^ (lane & 7)) << 4)), "r"(r[4 * c]),
                       "r"(r[4 * c + 1]), "r"(r[4 * c + 2]), "r"(r[4 * c + 3])
                       : "memory");
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int zc = p.c_bmod ? t.bz % p.c_bmod : t.bz;
          if (p.reduce) tma_reduce_add_3d(&map_c, stage, t.n0 + c0, t.m0 + q * 32, zc);
          else tma_store_3d(&map_c, stage, t.n0 + c0, t.m0 + q * 32, zc);
          tma_store_commit();
        }
      }
      tc_fence_before();                          // this warp's TMEM reads of the tile are complete
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm.acc_empty[ab]));
    }
    if (lane == 0) tma_store_wait_all<0>();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * kBNMax) : "memory");
  }
}

}  // namespace

bool tcgemm_operand_ok(const float* p, int64_t ld, int64_t batch_stride, int batch) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && ld % 4 == 0 && ld > 0 && (batch <= 1 || batch_stride % 4 == 0);
}

// C (M x N) (+)= A (M x K) * B (N x K)^T; see selscan_b200_gemm_3xtf32 in include/selscan_b200.h
cudaError_t launch_tcgemm(const float* A, int64_t lda, int a_mn, const float* B, int64_t ldb, int b_mn, float* C, int64_t ldc, int M,
                          int N, int K, int batch, int64_t strideA, int64_t strideB, int64_t strideC, int accumulate, int a_bmod,
                          int b_bmod, int c_bmod, cudaStream_t stream) {
  if (M == 0 || N == 0 || batch == 0) return cudaSuccess;
  TcParams p;
  p.a_bmod = a_bmod; p.b_bmod = b_bmod; p.c_bmod = c_bmod;
  p.M = M; p.N = N; p.K = K;
  p.a_mn = a_mn; p.b_mn = b_mn;
  p.batch = batch;
  // tile width: as few tiles of at most 128 columns as cover N; several tiles are multiples of 32 wide (their 32-column
  // stores must not spill into the next tile), a single tile only needs the UMMA granularity of 16
  p.n_tiles = (N + kBNMax - 1) / kBNMax;
  const int gran = p.n_tiles > 1 ? 32 : 16;
  p.BN = (((N + p.n_tiles - 1) / p.n_tiles) + gran - 1) / gran * gran;
  p.n_tiles = (N + p.BN - 1) / p.BN;
  p.m_tiles = (M + kBM - 1) / kBM;
  p.kb_total = (K + kBK - 1) / kBK;
  if (p.kb_total == 0) p.kb_total = 1;
  // split K when the tiles alone cannot fill the GPU and the reduction is long (weight gradients: K = batch * L)
  int splits = 1;
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles * batch;
  const int n_sm = sm_count();
  if (tiles < n_sm && p.kb_total >= 16) {
    splits = (int)((2 * n_sm + tiles - 1) / tiles);
    if (splits > p.kb_total / 4) splits = p.kb_total / 4;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  const bool shared_c = c_bmod > 0 && c_bmod < batch;     // several batch entries sum into one C
  p.reduce = (p.splits > 1 || accumulate || shared_c) ? 1 : 0;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a_mn ? 1 : 0) << 15) | ((uint32_t)(b_mn ? 1 : 0) << 16) |
            ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
  CUtensorMap ma, mb, mc;
  // K-major: rows = M (or N), inner = K; MN-major: rows = K, inner = M (or N).  Boxes are 32 floats (128 B, swizzled) wide.
  const int na = a_bmod ? a_bmod : batch, nb = b_bmod ? b_bmod : batch, nc = c_bmod ? c_bmod : batch;
  const bool ok_a = a_mn ? make_row_map_sw(&ma, A, M, K, na, lda, strideA, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                         : make_row_map_sw(&ma, A, K, M, na, lda, strideA, 32, kBM, CU_TENSOR_MAP_SWIZZLE_128B);
  const bool ok_b = b_mn ? make_row_map_sw(&mb, B, N, K, nb, ldb, strideB, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                         : make_row_map_sw(&mb, B, K, N, nb, ldb, strideB, 32, p.BN, CU_TENSOR_MAP_SWIZZLE_128B);
  const bool ok_c = make_row_map_sw(&mc, C, N, M, nc, ldc, strideC, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (!ok_a || !ok_b || !ok_c) return cudaErrorInvalidValue;
  if (p.reduce && !accumulate) {                  // split-K / shared C accumulate with reduce-adds: start from zero
    for (int b = 0; b < nc; ++b) {
      const cudaError_t e = cudaMemset2DAsync(C + (int64_t)b * strideC, (size_t)ldc * 4, 0, (size_t)N * 4, (size_t)M, stream);
      if (e != cudaSuccess) return e;
    }
  }
  const uint32_t b_tile = b_mn ? (uint32_t)((p.BN + 31) / 32) * 4096u : (uint32_t)p.BN * 128u;
  p.b_off = (b_tile + 1023u) / 1024u * 1024u;
  p.stage_bytes = 2u * kTileBytes + 2u * p.b_off;
  p.stages = (int)((kTcSmemMax - kTcHeader) / p.stage_bytes);
  if (p.stages > kTcMaxStages) p.stages = kTcMaxStages;
  const int smem = (int)(kTcHeader + (uint32_t)p.stages * p.stage_bytes + 1024u);
  {   // opt in to the largest dynamic shared-memory size seen so far, once per size increase and device instead of per launch
    static std::atomic<int> configured[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (configured[dev].load(std::memory_order_acquire) < smem) {
      const cudaError_t e = cudaFuncSetAttribute(tcgemm_3xtf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (e != cudaSuccess) return e;
      configured[dev].store(smem, std::memory_order_release);
    }
  }
  const int64_t n_work = tiles * p.splits;
  const unsigned grid = (unsigned)(n_work < n_sm ? n_work : n_sm);
  tcgemm_3xtf32_kernel<<<grid, kTcThreads, smem, stream>>>(ma, mb, mc, p);
  return cudaGetLastError();
}

}  // namespace selscan
