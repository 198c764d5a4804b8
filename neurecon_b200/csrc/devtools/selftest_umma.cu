// Single-CTA tcgen05 self-test: D[128, N] = A[128, K] * B[K, N] with exactly the operand layouts,
// descriptors, TMEM addressing and barrier protocol the fused MLP kernel uses.  Exercised by
// tests/test_gpu_umma.py against a bf16-rounded CPU matmul.
#include "../common.cuh"
#include "../umma.cuh"

namespace {

__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(const uint8_t* __restrict__ a_image, const float* __restrict__ B, int K, int N,
                     float* __restrict__ D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kchunks = (K + 63) / 64;
  uint8_t* sA = smem;                       // kchunks x 16 KB
  uint8_t* sB = smem + kchunks * 16384;     // K x N bf16, MN-major SW128
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t lbo_b = (uint32_t)((K + 7) / 8) * 1024;  // 64-column blocks are K/8 atoms apart

  if (warp == 0) {
    umma::tmem_alloc(&tmem_base_s, 128);
    umma::tmem_relinquish();
  }
  if (tid == 0) {
    umma::mbar_init(&bar_load, 1);
    umma::mbar_init(&bar_mma, 1);
    umma::fence_barrier_init();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    umma::mbar_arrive_expect_tx(&bar_load, (uint32_t)kchunks * 16384u);
    for (int c = 0; c < kchunks; ++c) umma::bulk_g2s(sA + c * 16384, a_image + (size_t)c * 16384, 16384u, &bar_load);
  }
  // B: thread owns k-rows tid, tid+128, ... and writes 16-byte chunks of 8 columns
  for (int k = tid; k < K; k += 128) {
    for (int n8 = 0; n8 < N / 8; ++n8) {
      const float* src = B + (size_t)k * N + n8 * 8;
      uint4 v;
      v.x = umma::pack_bf16(src[0], src[1]);
      v.y = umma::pack_bf16(src[2], src[3]);
      v.z = umma::pack_bf16(src[4], src[5]);
      v.w = umma::pack_bf16(src[6], src[7]);
      *reinterpret_cast<uint4*>(sB + umma::b_chunk_offset(k, n8, lbo_b)) = v;
    }
  }
  umma::fence_proxy_async_smem();
  __syncthreads();

  if (tid == 0) {
    umma::mbar_wait(&bar_load, 0);
    umma::tc_fence_after();
    const uint32_t idesc = umma::make_idesc_bf16(128, N, 0, 1);
    const uint32_t a_lbo = (variant & 1) ? 0u : 16u;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a_addr = umma::smem_u32(sA) + (ks >> 2) * 16384 + (ks & 3) * 32;
      const uint32_t b_addr = umma::smem_u32(sB) + ks * 2048;
      const uint64_t da = umma::make_smem_desc(a_addr, a_lbo, 1024);
      const uint64_t db = (variant & 2) ? umma::make_smem_desc(b_addr, 1024, lbo_b)
                                        : umma::make_smem_desc(b_addr, lbo_b, 1024);
      umma::mma_bf16_ss(tmem_base, da, db, idesc, ks > 0 ? 1u : 0u);
    }
    umma::mma_commit(&bar_mma);
  }
  __syncwarp();
  umma::mbar_wait(&bar_mma, 0);
  umma::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    umma::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tmem_ld_wait();
    const int row = warp * 32 + (tid & 31);
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 128);
}

// CTA-pair self-test: D[256, N] = A[256, K] * B[K, N] with tcgen05.mma.cta_group::2.  CTA r of the pair holds rows
// [128r, 128r+128) of A and of D and columns [N/2 r, N/2 (r+1)) of B.  variant bit 0: each CTA writes the PEER's half
// of B through distributed shared memory (the path the fused MLP's epilogue would use for its activations).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
selftest_umma2_kernel(const uint8_t* __restrict__ a_image, const float* __restrict__ B, int K, int N,
                      float* __restrict__ D, int variant) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int kchunks = (K + 63) / 64;
  const int nh = N / 2;                      // columns of B per CTA
  uint8_t* sA = smem;                       // kchunks x 16 KB: this CTA's 128 rows
  uint8_t* sB = smem + kchunks * 16384;     // K x nh, MN-major SW128
  __shared__ uint64_t bar_load, bar_mma, bar_ready;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = umma::cluster_ctarank();
  const uint32_t lbo_b = (uint32_t)((K + 7) / 8) * 1024;

  if (tid == 0) {
    umma::mbar_init(&bar_load, 1);
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_init(&bar_ready, 8);           // 4 warps x 2 CTAs
    umma::fence_barrier_init();
  }
  __syncthreads();
  umma::cluster_sync_all();                 // barriers of both CTAs initialised before anything can arrive on them
  if (warp == 0) {
    umma::tmem_alloc2(&tmem_base_s, 256);
    umma::tmem_relinquish2();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    umma::mbar_arrive_expect_tx(&bar_load, (uint32_t)kchunks * 16384u);
    // image layout: [k-chunk][M-tile 0 | M-tile 1] 16 KB each
    for (int c = 0; c < kchunks; ++c)
      umma::bulk_g2s(sA + c * 16384, a_image + ((size_t)c * 2 + rank) * 16384, 16384u, &bar_load);
  }
  // B: this CTA fills the half of `dst` = itself, or its peer (variant & 1) -- with the columns that half owns
  const uint32_t dst = (variant & 1) ? (rank ^ 1u) : rank;
  const uint32_t sB_dst = umma::map_to_cta(umma::smem_u32(sB), dst);
  for (int k = tid; k < K; k += 128) {
    for (int n8 = 0; n8 < nh / 8; ++n8) {
      const float* src = B + (size_t)k * N + dst * nh + n8 * 8;
      umma::st_cluster_v4(sB_dst + umma::b_chunk_offset(k, n8, lbo_b), umma::pack_f16(src[0], src[1]), umma::pack_f16(src[2], src[3]),
                          umma::pack_f16(src[4], src[5]), umma::pack_f16(src[6], src[7]));
    }
  }
  umma::mbar_wait(&bar_load, 0);
  if (variant & 2) {
    // the fused kernel's hand-off: every warp of both CTAs arrives (release.cluster) on the LEADER's barrier after a
    // full async-proxy fence; the issuing thread acquires at cluster scope
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncwarp();
    if ((tid & 31) == 0) umma::mbar_arrive_cluster(umma::map_to_cta(umma::smem_u32(&bar_ready), 0));
    if (rank == 0 && tid == 0) umma::mbar_wait_cluster(&bar_ready, 0);
  } else {
    umma::fence_proxy_async_smem();
    __syncthreads();
    umma::cluster_sync_all();                 // both halves of A and B in place and visible
  }

  if (rank == 0 && tid == 0) {
    umma::tc_fence_after();
    const uint32_t idesc = umma::make_idesc_f16(256, N, 0, 1);
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t a_addr = umma::smem_u32(sA) + (ks >> 2) * 16384 + (ks & 3) * 32;
      const uint32_t b_addr = umma::smem_u32(sB) + ks * 2048;
      umma::mma2_f16_ss(tmem_base, umma::make_smem_desc(a_addr, 16, 1024), umma::make_smem_desc(b_addr, lbo_b, 1024), idesc,
                        ks > 0 ? 1u : 0u);
    }
    umma::mma2_commit_mc(&bar_mma, 3);
  }
  __syncwarp();
  umma::mbar_wait(&bar_mma, 0);
  umma::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    umma::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    umma::tmem_ld_wait();
    const int row = (int)rank * 128 + warp * 32 + (tid & 31);
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();
  if (warp == 0) umma::tmem_dealloc2(tmem_base, 256);
}

// Tensor-pipe rate probe: one lane issues `n_mmas` back-to-back M=128 x N x 16 MMAs on operands resident in
// shared memory (K = 256 cycled), optionally while `store_warps` other warps hammer shared memory with 16-byte
// stores and/or a producer keeps 16 KB bulk copies in flight -- the traffic mix of the fused MLP kernel.
// out[block] = cycles from the first issue to the completion of the last MMA.
__global__ void __launch_bounds__(640, 1)
bench_umma_kernel(int N, int n_mmas, int store_warps, int bulk_copies, const uint8_t* __restrict__ gsrc,
                  long long* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 64 KB: 4 chunks of 128 x 64
  uint8_t* sB = smem + 65536;         // 64 KB (N <= 128) or 128 KB (N = 256): 256 k-rows, MN-major SW128
  uint8_t* sX = smem + 65536 + 131072;  // 16 KB scratch for the store / copy traffic
  __shared__ uint64_t bar_mma, bar_copy;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (65536 + 131072 + 16384) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) { umma::tmem_alloc(&tmem_base_s, 512); umma::tmem_relinquish(); }
  if (tid == 0) {
    umma::mbar_init(&bar_mma, 1);
    umma::mbar_init(&bar_copy, 1);
    umma::fence_barrier_init();
    stop_flag = 0;
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp == 1) {
    const uint32_t idesc = umma::make_idesc_f16(128, N, 0, 1);
    const uint32_t a_hi = umma::smem_desc_hi(1024), b_hi = umma::smem_desc_hi(1024);
    const uint32_t a_lo0 = umma::smem_desc_lo(umma::smem_u32(sA), 16), b_lo0 = umma::smem_desc_lo(umma::smem_u32(sB), 32768);
    const long long t0 = clock64();
    if (n_mmas < 0) { while (clock64() - t0 < -(long long)n_mmas) {} }   // no MMAs: just hold the window open
    for (int i = 0; i < n_mmas; i += 4) {
      const uint32_t kc = (uint32_t)(i >> 2) & 3u;
      const uint32_t a_lo = a_lo0 + kc * 1024, b_lo = b_lo0 + kc * 512;
      const uint32_t d = tmem_base + (((uint32_t)i >> 4) & 1u) * (uint32_t)N;
      if (umma::elect_one()) {
        umma::mma_bf16_ss(d, umma::desc64(a_lo, a_hi), umma::desc64(b_lo, b_hi), idesc, 1u);
        umma::mma_bf16_ss(d, umma::desc64(a_lo + 2, a_hi), umma::desc64(b_lo + 128, b_hi), idesc, 1u);
        umma::mma_bf16_ss(d, umma::desc64(a_lo + 4, a_hi), umma::desc64(b_lo + 256, b_hi), idesc, 1u);
        umma::mma_bf16_ss(d, umma::desc64(a_lo + 6, a_hi), umma::desc64(b_lo + 384, b_hi), idesc, 1u);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (umma::elect_one()) umma::mma_commit(&bar_mma);
    __syncwarp();
    umma::mbar_wait(&bar_mma, 0);
    const long long t2 = clock64();
    if (lane == 0) { out[2 * blockIdx.x] = t2 - t0; out[2 * blockIdx.x + 1] = t1 - t0; stop_flag = 1; }
  } else if (warp == 2 && bulk_copies) {
    uint32_t ph = 0;
    while (!stop_flag) {
      if (lane == 0) {
        umma::mbar_arrive_expect_tx(&bar_copy, 16384u);
        umma::bulk_g2s(sX, gsrc + (size_t)((blockIdx.x * 7 + ph) % 64) * 16384, 16384u, &bar_copy);
      }
      __syncwarp();
      umma::mbar_wait(&bar_copy, ph & 1);
      ++ph;
    }
  } else if (warp >= 4 && warp < 4 + store_warps) {
    const uint32_t base = umma::smem_u32(sX) + (uint32_t)((warp - 4) & 7) * 2048 + lane * 16;
    uint32_t it = 0;
    while (!stop_flag) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + r * 512), "r"(it) : "memory");
      ++it;
    }
    if (lane == 0 && warp == 4) out[2 * gridDim.x + blockIdx.x] = (long long)it * 4;   // 512-byte stores of one warp
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, 512);
}

// TMEM read-rate probe: `warps` warps each issue `iters` x 4 tcgen05.ld.32x32b.x16 (2 KB per instruction) on their
// lane quarter, waiting once per group of 4.  out[0] = cycles of warp 0.
__global__ void __launch_bounds__(1024, 1) bench_ldtm_kernel(int iters, long long* __restrict__ out, float* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { umma::tmem_alloc(&tmem_base_s, 512); umma::tmem_relinquish(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t taddr = tmem_base_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) & 3) * 128u;
  uint32_t a0[16], a1[16], a2[16], a3[16];
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    umma::tmem_ld16(taddr, a0);
    umma::tmem_ld16(taddr + 16, a1);
    umma::tmem_ld16(taddr + 32, a2);
    umma::tmem_ld16(taddr + 48, a3);
    umma::tmem_ld_wait();
    acc += __uint_as_float(a0[0] ^ a1[1] ^ a2[2] ^ a3[3]);
  }
  const long long t1 = clock64();
  if (acc == 123.456f) sink[0] = acc;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base_s, 512);
}

}  // namespace

extern "C" int nr_selftest_umma(const void* a_image, const float* B, int32_t K, int32_t N, float* D, int32_t variant,
                                void* stream) {
  NR_CHECK_ARG(a_image && B && D, "nr_selftest_umma: null pointer");
  NR_CHECK_ARG(K >= 16 && K <= 256 && K % 16 == 0, "nr_selftest_umma: K must be a multiple of 16 in [16,256]");
  NR_CHECK_ARG(N >= 32 && N <= 128 && N % 32 == 0, "nr_selftest_umma: N must be 32, 64, 96 or 128");
  const int kchunks = (K + 63) / 64;
  const size_t smem = 1024 + (size_t)kchunks * 16384 + (size_t)((K + 7) / 8) * 1024 * ((N + 63) / 64);
  NR_CHECK_CUDA(cudaFuncSetAttribute(selftest_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_umma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)a_image, B, K, N, D, variant);
  NR_CHECK_LAUNCH("selftest_umma_kernel");
  return NR_OK;
}

extern "C" int nr_bench_umma(int32_t N, int32_t n_mmas, int32_t store_warps, int32_t bulk_copies, const void* gsrc,
                             int32_t grid, long long* out, void* stream) {
  NR_CHECK_ARG(out && gsrc, "nr_bench_umma: null pointer");
  NR_CHECK_ARG(N == 32 || N == 64 || N == 128 || N == 256, "nr_bench_umma: N");
  NR_CHECK_ARG((n_mmas < 0 || (n_mmas >= 4 && n_mmas % 4 == 0)) && store_warps >= 0 && store_warps <= 16 && grid >= 1, "nr_bench_umma: sizes");
  const size_t smem = 1024 + 65536 + 131072 + 16384;
  NR_CHECK_CUDA(cudaFuncSetAttribute(bench_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  bench_umma_kernel<<<grid, 640, smem, (cudaStream_t)stream>>>(N, n_mmas, store_warps, bulk_copies, (const uint8_t*)gsrc, out);
  NR_CHECK_LAUNCH("bench_umma_kernel");
  return NR_OK;
}

extern "C" int nr_bench_ldtm(int32_t warps, int32_t iters, int32_t grid, long long* out, float* sink, void* stream) {
  NR_CHECK_ARG(out && sink && warps >= 1 && warps <= 32 && iters > 0 && grid > 0, "nr_bench_ldtm: args");
  bench_ldtm_kernel<<<grid, warps * 32, 0, (cudaStream_t)stream>>>(iters, out, sink);
  NR_CHECK_LAUNCH("bench_ldtm_kernel");
  return NR_OK;
}

extern "C" int nr_selftest_umma2(const void* a_image, const float* B, int32_t K, int32_t N, float* D, int32_t variant,
                                 void* stream) {
  NR_CHECK_ARG(a_image && B && D, "nr_selftest_umma2: null pointer");
  NR_CHECK_ARG(K >= 16 && K <= 256 && K % 16 == 0, "nr_selftest_umma2: K must be a multiple of 16 in [16,256]");
  NR_CHECK_ARG(N == 64 || N == 128 || N == 256, "nr_selftest_umma2: N must be 64, 128 or 256");
  const int kchunks = (K + 63) / 64;
  const size_t smem = 1024 + (size_t)kchunks * 16384 + (size_t)((K + 7) / 8) * 1024 * ((N / 2 + 63) / 64);
  NR_CHECK_CUDA(cudaFuncSetAttribute(selftest_umma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  selftest_umma2_kernel<<<2, 128, smem, (cudaStream_t)stream>>>((const uint8_t*)a_image, B, K, N, D, variant);
  NR_CHECK_LAUNCH("selftest_umma2_kernel");
  return NR_OK;
}
