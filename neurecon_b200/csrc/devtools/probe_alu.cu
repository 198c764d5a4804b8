// Issue-rate probes for the epilogue's instruction mix (tools/probe_alu.py): per-SM throughput of MUFU.EX2,
// MUFU.RCP, the fp32 -> 16-bit pack conversion and FFMA, measured with clock64 over a long unrolled loop.
#include "../common.cuh"

namespace {

template <int kOp>
__global__ void __launch_bounds__(1024, 1) probe_alu_kernel(int iters, float seed, float* out, long long* cycles) {
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = seed + 0.001f * (float)(threadIdx.x + j);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (kOp == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      else if (kOp == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      else if (kOp == 2) { uint32_t p; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v[j]), "f"(v[(j + 1) & 7])); acc ^= p; }
      else if (kOp == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[j]));
      else if (kOp == 4) { uint32_t p; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(v[j]), "f"(v[(j + 1) & 7])); acc ^= p; }
      else if (kOp == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      else if (kOp == 7) { uint32_t& w = reinterpret_cast<uint32_t&>(v[j]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(w)); }
      else if (kOp == 8) { uint32_t& w = reinterpret_cast<uint32_t&>(v[j]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(w)); }
      else if (kOp == 9) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[j]));
      else if (kOp == 10) { uint32_t& w = reinterpret_cast<uint32_t&>(v[j]); asm volatile("fma.rn.f16x2 %0, %0, %0, %0;" : "+r"(w)); }
      else if (kOp == 11) { uint32_t& w = reinterpret_cast<uint32_t&>(v[j]); asm volatile("max.f16x2 %0, %0, %1;" : "+r"(w) : "r"(acc)); }
      else if (kOp == 6) {
        if (j & 1) continue;
        uint64_t p = ((uint64_t)__float_as_uint(v[j + 1]) << 32) | __float_as_uint(v[j]);
        asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(p));
        v[j] = __uint_as_float((uint32_t)p); v[j + 1] = __uint_as_float((uint32_t)(p >> 32));
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  float s = __uint_as_float(acc);
#pragma unroll
  for (int j = 0; j < 8; ++j) s += v[j];
  if (s == 123.456f) out[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

}  // namespace

extern "C" int nr_probe_alu(int32_t op, int32_t threads, int32_t iters, int32_t grid, float* out, long long* cycles,
                            void* stream) {
  NR_CHECK_ARG(out && cycles && threads >= 32 && threads <= 1024 && iters > 0 && grid > 0 && op >= 0 && op <= 11, "nr_probe_alu: args");
  cudaStream_t st = (cudaStream_t)stream;
  switch (op) {
    case 0: probe_alu_kernel<0><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 1: probe_alu_kernel<1><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 2: probe_alu_kernel<2><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 3: probe_alu_kernel<3><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 4: probe_alu_kernel<4><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 5: probe_alu_kernel<5><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 6: probe_alu_kernel<6><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 7: probe_alu_kernel<7><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 8: probe_alu_kernel<8><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 9: probe_alu_kernel<9><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    case 10: probe_alu_kernel<10><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
    default: probe_alu_kernel<11><<<grid, threads, 0, st>>>(iters, 0.5f, out, cycles); break;
  }
  NR_CHECK_LAUNCH("probe_alu_kernel");
  return NR_OK;
}
