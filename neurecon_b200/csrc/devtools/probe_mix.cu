// Which of the epilogue's instruction classes overlap?  Every thread runs iters x (8 x op A interleaved with 8 x op B) on
// independent registers; tools/probe_mix.py prints the cycles per scheduler for A alone, B alone and the mix.
// ops: 0 none, 1 fma.f32, 2 fma.f32x2, 3 ex2.approx, 4 lop3, 5 max.f32, 6 prmt, 7 cvt.f16x2.f32, 8 fma.f16x2, 9 add.s32,
// 10 fma.f32x2 with two register-pair operands and an immediate-like third (the epilogue's form)
#include "../common.cuh"

namespace {

template <int kOp>
__device__ __forceinline__ void op(unsigned long long& r, uint32_t k) {
  uint32_t lo = (uint32_t)r, hi = (uint32_t)(r >> 32);
  float& f = reinterpret_cast<float&>(lo);
  if (kOp == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f));
  else if (kOp == 2) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(r)); return; }
  else if (kOp == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
  else if (kOp == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0xf8;" : "+r"(lo) : "r"(k), "r"(hi));
  else if (kOp == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(f) : "f"(__uint_as_float(k)));
  else if (kOp == 6) asm volatile("prmt.b32 %0, %0, %1, 0x0040;" : "+r"(lo) : "r"(hi));
  else if (kOp == 7) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "+r"(lo) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(k)));
  else if (kOp == 8) asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(lo));
  else if (kOp == 9) asm volatile("add.s32 %0, %0, %1;" : "+r"(lo) : "r"(k));
  else if (kOp == 10) {
    const unsigned long long c = 0x3f0000003f000000ull;
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r) : "l"(r ^ 1ull), "l"(c));
    return;
  }
  r = ((unsigned long long)hi << 32) | lo;
}

template <int kA, int kB>
__global__ void __launch_bounds__(640, 1) probe_mix_kernel(int iters, uint32_t seed, uint32_t* out, long long* cycles) {
  unsigned long long a[8], b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = ((unsigned long long)__float_as_uint(0.5f + 0.001f * (float)(threadIdx.x + j)) << 32) | __float_as_uint(0.25f + 0.002f * j);
    b[j] = a[j] + seed;
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (kA) op<kA>(a[j], seed);
      if (kB) op<kB>(b[j], seed);
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  unsigned long long s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= a[j] ^ b[j];
  if (s == 0x123456789ull) out[0] = (uint32_t)s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kA>
int launch_b(int b, int threads, int iters, int grid, uint32_t* out, long long* cycles, cudaStream_t st) {
#define NR_PM(B) case B: probe_mix_kernel<kA, B><<<grid, threads, 0, st>>>(iters, 1u, out, cycles); break;
  switch (b) { NR_PM(0) NR_PM(1) NR_PM(2) NR_PM(3) NR_PM(4) NR_PM(5) NR_PM(6) NR_PM(7) NR_PM(8) NR_PM(9) NR_PM(10) default: return -1; }
#undef NR_PM
  return 0;
}

}  // namespace

extern "C" int nr_probe_mix(int32_t op_a, int32_t op_b, int32_t threads, int32_t iters, int32_t grid, uint32_t* out,
                            long long* cycles, void* stream) {
  NR_CHECK_ARG(out && cycles && threads >= 32 && threads <= 640 && iters > 0 && grid > 0, "nr_probe_mix: args");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = -1;
#define NR_PA(A) case A: rc = launch_b<A>(op_b, threads, iters, grid, out, cycles, st); break;
  switch (op_a) { NR_PA(0) NR_PA(1) NR_PA(2) NR_PA(3) NR_PA(4) NR_PA(5) NR_PA(6) NR_PA(7) NR_PA(8) NR_PA(9) NR_PA(10) default: break; }
#undef NR_PA
  NR_CHECK_ARG(rc == 0, "nr_probe_mix: ops %d, %d", op_a, op_b);
  NR_CHECK_LAUNCH("probe_mix_kernel");
  return NR_OK;
}
