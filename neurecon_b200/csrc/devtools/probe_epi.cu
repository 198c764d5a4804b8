// The forward epilogue of mlp_rev_kernel (softplus + softplus' codes of one 16-column chunk per thread) in isolation: no
// MMAs, no TMEM, the accumulators come from shared memory.  tools/probe_epi.py compares cycles per chunk of a few
// formulations with what the chunk costs inside the kernel (tools/trace_rev.py), to tell instruction-bound from
// interference-bound.
#include "../mlp_epilogue.cuh"

namespace {

__device__ __forceinline__ uint32_t code_pack4(f32x2 a, f32x2 b, float scale) {
  float a0, a1, b0, b1;
  upk2(fma2(a, splat2(scale), splat2(12583040.0f)), a0, a1);
  upk2(fma2(b, splat2(scale), splat2(12583040.0f)), b0, b1);
  const uint32_t p0 = __byte_perm(__float_as_uint(a0), __float_as_uint(a1), 0x0040);
  const uint32_t p1 = __byte_perm(__float_as_uint(b0), __float_as_uint(b1), 0x0040);
  return __byte_perm(p0, p1, 0x5410);
}

// variant 2: log1p as a quadratic, 1/(1+u) as a cubic
__device__ __forceinline__ void softplus_lo2(float a0, float a1, f32x2 b144, float& sp0, float& sp1, f32x2& sg) {
  const f32x2 t2 = fma2(pk2(a0, a1), splat2(144.26950408889634f), b144);
  float t0, t1;
  upk2(t2, t0, t1);
  const f32x2 u2 = pk2(ex2_approx(-fabsf(t0)), ex2_approx(-fabsf(t1)));
  f32x2 p = fma2(u2, splat2(-0.2914e-2f), splat2(0.9829e-2f));
  const f32x2 q = mul2(p, u2);
  upk2(fma2(pk2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), splat2(0.006931471805599453f), q), sp0, sp1);
  f32x2 r = fma2(u2, splat2(-0.2355f), splat2(0.6863f));
  r = fma2(r, u2, splat2(-0.9508f));
  r = fma2(r, u2, splat2(0.99874f - 0.5f));
  float s0, s1;
  upk2(r, s0, s1);
  s0 = __uint_as_float(__float_as_uint(s0) | (__float_as_uint(t0) & 0x80000000u));
  s1 = __uint_as_float(__float_as_uint(s1) | (__float_as_uint(t1) & 0x80000000u));
  sg = pk2(s0, s1);
}
// variant 3: no bias / scale FMA (t = accumulator: weights pre-scaled, bias in the accumulator)
__device__ __forceinline__ void softplus_nb2(float t0, float t1, float& sp0, float& sp1, f32x2& sg) {
  const f32x2 u2 = pk2(ex2_approx(-fabsf(t0)), ex2_approx(-fabsf(t1)));
  f32x2 p = fma2(u2, splat2(-0.05875718221068382e-2f), splat2(0.22568579018115997e-2f));
  p = fma2(p, u2, splat2(-0.4713013470172882e-2f));
  p = fma2(p, u2, splat2(0.9974489808082581e-2f));
  const f32x2 q = mul2(p, u2);
  upk2(fma2(pk2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), splat2(0.006931471805599453f), q), sp0, sp1);
  f32x2 r = fma2(u2, splat2(0.16162029f), splat2(-0.55180615f));
  r = fma2(r, u2, splat2(0.87791824f));
  r = fma2(r, u2, splat2(-0.9872991f));
  r = fma2(r, u2, splat2(0.99978334f - 0.5f));
  float s0, s1;
  upk2(r, s0, s1);
  s0 = __uint_as_float(__float_as_uint(s0) | (__float_as_uint(t0) & 0x80000000u));
  s1 = __uint_as_float(__float_as_uint(s1) | (__float_as_uint(t1) & 0x80000000u));
  sg = pk2(s0, s1);
}
// variant 4: the shipped math on scalar FFMA (one value per instruction)
__device__ __forceinline__ void softplus_sc1(float a, float b144, float& sp, float& sg) {
  const float t = fmaf(a, 144.26950408889634f, b144);
  const float u = ex2_approx(-fabsf(t));
  float p = fmaf(u, -0.05875718221068382e-2f, 0.22568579018115997e-2f);
  p = fmaf(p, u, -0.4713013470172882e-2f);
  p = fmaf(p, u, 0.9974489808082581e-2f);
  sp = fmaf(fmaxf(t, 0.0f), 0.006931471805599453f, p * u);
  float r = fmaf(u, 0.16162029f, -0.55180615f);
  r = fmaf(r, u, 0.87791824f);
  r = fmaf(r, u, -0.9872991f);
  r = fmaf(r, u, 0.99978334f - 0.5f);
  sg = __uint_as_float(__float_as_uint(r) | (__float_as_uint(t) & 0x80000000u));
}

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

// kV: 0 shipped (math, codes, operand stores, code stores), 1 no operand stores, 2 no code stores, 3 no stores, 4 scalar FFMA,
// 5 math only (no codes, no conversions, no stores), 6 no activation math (codes, conversions, stores), 7 lower-degree
// polynomials, 8 no bias FMA, 9 code stores with the default cache policy, 10 one transcendental (tanh) for both outputs
template <int kV>
__global__ void __launch_bounds__(640, 1) probe_epi_kernel(int iters, const float* bias, uint8_t* scratch, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t done_bar;
  if (threadIdx.x == 0) { umma::mbar_init(&done_bar, 1); umma::fence_barrier_init(); }
  uint8_t* act = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // 64 KB operand buffer
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 16) {
    umma::tmem_alloc(&tmem_base_s, 512);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp < 16) {
    const int q = warp & 3, mo = (warp >> 2) & 1, g = warp >> 3;
    const int F = mo * 128 + 32 * q + (tid & 31);
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(mo * 128) + 64u * g;
    for (int c = 0; c < 4; ++c) {   // accumulators: a mix of saturated and unsaturated pre-activations
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(0.004f * (float)(((F * 64 + c * 16 + j) * 37) % 121 - 60));
      tmem_st16(taddr + 16 * c, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const RowAddr ra(umma::smem_u32(act), F);
    const float b = bias[F];
    const f32x2 b144 = splat2(b * 144.26950408889634f);
    uint8_t* dst = scratch + ((size_t)blockIdx.x * 512 + tid) * 16;
    uint32_t sink = 0;
    asm volatile("bar.sync 1, 512;");
    const long long t0 = clock64();
    auto values = [&](const uint32_t (&r)[16], int c, int it) {
      float vv[16];
      f32x2 d2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float a0 = __uint_as_float(r[2 * j]), a1 = __uint_as_float(r[2 * j + 1]);
        if (kV == 7) softplus_lo2(a0, a1, b144, vv[2 * j], vv[2 * j + 1], d2[j]);
        else if (kV == 8) softplus_nb2(a0, a1, vv[2 * j], vv[2 * j + 1], d2[j]);
        else if (kV == 10) softplus_th2<false>(a0, a1, b144, vv[2 * j], vv[2 * j + 1], d2[j]);
        else if (kV == 11) softplus_th2<true>(a0, a1, b144, vv[2 * j], vv[2 * j + 1], d2[j]);
        else if (kV == 4) {
          float s0, s1;
          softplus_sc1(a0, b * 144.26950408889634f, vv[2 * j], s0);
          softplus_sc1(a1, b * 144.26950408889634f, vv[2 * j + 1], s1);
          d2[j] = pk2(s0, s1);
        } else if (kV == 6) { vv[2 * j] = a0; vv[2 * j + 1] = a1; d2[j] = pk2(a1, a0); }
        else softplus_sigq2(a0, a1, b144, vv[2 * j], vv[2 * j + 1], d2[j]);
      }
      if (kV == 5) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { float s0, s1; upk2(d2[j], s0, s1); sink ^= __float_as_uint(vv[2 * j] + s0) ^ __float_as_uint(vv[2 * j + 1] + s1); }
        return;
      }
      const float cs = kV >= 10 ? 127.0f : 254.0f;
      const uint4 w = make_uint4(code_pack4(d2[0], d2[1], cs), code_pack4(d2[2], d2[3], cs), code_pack4(d2[4], d2[5], cs),
                                 code_pack4(d2[6], d2[7], cs));
      if (kV == 1 || kV == 3) {
#pragma unroll
        for (int j = 0; j < 8; ++j) sink ^= umma::pack2<true>(vv[2 * j], vv[2 * j + 1]);
      } else store_row16<true>(ra, 16 * (4 * g + c), vv);
      uint4* cdst = reinterpret_cast<uint4*>(dst + (size_t)(c + 4 * (it & 7)) * 8192 * 148);
      if (kV == 2 || kV == 3) sink ^= w.x ^ w.y ^ w.z ^ w.w;
      else if (kV == 9) *cdst = w;
      else __stcg(cdst, w);
    };
#pragma unroll 1
    for (int it = 0; it < (iters & 0xfffff); ++it) {
      uint32_t raw[16], rawB[16];
      umma::tmem_ld16(taddr, raw);
#pragma unroll
      for (int k = 0; k < 4; k += 2) {
        umma::tmem_ld_wait();
        umma::tmem_ld16(taddr + 16 * (k + 1), rawB);
        values(raw, k, it);
        umma::tmem_ld_wait();
        if (k + 2 < 4) umma::tmem_ld16(taddr + 16 * (k + 2), raw);
        values(rawB, k + 1, it);
      }
    }
    const long long t1 = clock64();
    if (sink == 0x12345u) dst[0] = 1;
    asm volatile("bar.sync 1, 512;");
    if (tid == 0) { cycles[blockIdx.x] = t1 - t0; umma::mbar_arrive(&done_bar); }
  } else if (warp - 16 < (iters >> 20)) {
    // iters bits 20+: this many of the four idle warps poll an mbarrier the way the kernel's producer / MMA issuers do
    umma::mbar_wait(&done_bar, 0);
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 16) umma::tmem_dealloc(tmem_base, 512);
}

// The same chunk on EIGHT fat epilogue warps (384 threads: 168 registers each, two warps per scheduler) instead of sixteen:
// kWide = 0: 16-column chunks as above; 1: two 16-column chunks in flight per warp (16 value pairs of ILP)
template <int kWide>
__global__ void __launch_bounds__(384, 1) probe_epi8_kernel(int iters, const float* bias, uint8_t* scratch, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t tmem_base_s;
  uint8_t* act = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 8) {
    umma::tmem_alloc(&tmem_base_s, 512);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (warp < 8) {
    const int q = warp & 3, mo = warp >> 2;
    const int F = mo * 128 + 32 * q + (tid & 31);
    const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(mo * 128);
    for (int c = 0; c < 8; ++c) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(0.004f * (float)(((F * 64 + c * 16 + j) * 37) % 121 - 60));
      tmem_st16(taddr + 16 * c, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const RowAddr ra(umma::smem_u32(act), F);
    const f32x2 b50 = splat2(bias[F] * 50.0f);
    uint8_t* dst = scratch + ((size_t)blockIdx.x * 512 + tid) * 16;
    asm volatile("bar.sync 1, 256;");
    const long long t0 = clock64();
    auto values = [&](const uint32_t (&r)[16], int c, int it) {
      float vv[16];
      f32x2 d2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        softplus_th2<true>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), b50, vv[2 * j], vv[2 * j + 1], d2[j]);
      const uint4 w = make_uint4(code_pack4(d2[0], d2[1], 127.0f), code_pack4(d2[2], d2[3], 127.0f), code_pack4(d2[4], d2[5], 127.0f),
                                 code_pack4(d2[6], d2[7], 127.0f));
      store_row16<true>(ra, 16 * c, vv);
      __stcg(reinterpret_cast<uint4*>(dst + (size_t)(c + 8 * (it & 3)) * 8192 * 148), w);
    };
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
      if (kWide == 0) {
        uint32_t raw[16], rawB[16];
        umma::tmem_ld16(taddr, raw);
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          umma::tmem_ld_wait();
          umma::tmem_ld16(taddr + 16 * (k + 1), rawB);
          values(raw, k, it);
          umma::tmem_ld_wait();
          if (k + 2 < 8) umma::tmem_ld16(taddr + 16 * (k + 2), raw);
          values(rawB, k + 1, it);
        }
      } else {
        uint32_t r0[16], r1[16], r2[16], r3[16];
        umma::tmem_ld16(taddr, r0);
        umma::tmem_ld16(taddr + 16, r1);
#pragma unroll
        for (int k = 0; k < 8; k += 4) {
          umma::tmem_ld_wait();
          umma::tmem_ld16(taddr + 16 * (k + 2), r2);
          umma::tmem_ld16(taddr + 16 * (k + 3), r3);
          values(r0, k, it);
          values(r1, k + 1, it);
          umma::tmem_ld_wait();
          if (k + 4 < 8) { umma::tmem_ld16(taddr + 16 * (k + 4), r0); umma::tmem_ld16(taddr + 16 * (k + 5), r1); }
          values(r2, k + 2, it);
          values(r3, k + 3, it);
        }
      }
    }
    const long long t1 = clock64();
    asm volatile("bar.sync 1, 256;");
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 8) umma::tmem_dealloc(tmem_base, 512);
}

}  // namespace

// scratch: 32 * 8192 * grid bytes; cycles: [grid]; iters = tile steps of 4 chunks per warp
extern "C" int nr_probe_epi(int32_t variant, int32_t iters, int32_t grid, const float* bias, void* scratch, long long* cycles,
                            void* stream) {
  NR_CHECK_ARG(bias && scratch && cycles && iters > 0 && grid > 0 && variant >= 0 && variant <= 21, "nr_probe_epi: args");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = 65536 + 1024;
#define NR_PE(V)                                                                                                        \
  case V:                                                                                                               \
    NR_CHECK_CUDA(cudaFuncSetAttribute(probe_epi_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    probe_epi_kernel<V><<<grid, 640, smem, st>>>(iters, bias, (uint8_t*)scratch, cycles);                                \
    break;
  if (variant >= 20) {
    NR_CHECK_CUDA(cudaFuncSetAttribute(probe_epi8_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(probe_epi8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (variant == 20) probe_epi8_kernel<0><<<grid, 384, smem, st>>>(iters, bias, (uint8_t*)scratch, cycles);
    else probe_epi8_kernel<1><<<grid, 384, smem, st>>>(iters, bias, (uint8_t*)scratch, cycles);
    NR_CHECK_LAUNCH("probe_epi8_kernel");
    return NR_OK;
  }
  switch (variant) { NR_PE(0) NR_PE(1) NR_PE(2) NR_PE(3) NR_PE(4) NR_PE(5) NR_PE(6) NR_PE(7) NR_PE(8) NR_PE(9) NR_PE(10) NR_PE(11) }
#undef NR_PE
  NR_CHECK_LAUNCH("probe_epi_kernel");
  return NR_OK;
}
