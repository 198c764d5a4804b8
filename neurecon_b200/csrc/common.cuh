// Shared helpers for the neurecon_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/neurecon_b200.h"

void nr_set_error(const char* fmt, ...);

#define NR_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      nr_set_error(__VA_ARGS__);           \
      return NR_ERR_INVALID;               \
    }                                      \
  } while (0)

void nr_count_launch();

#define NR_CHECK_LAUNCH(name)                                                      \
  do {                                                                             \
    nr_count_launch();                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      nr_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__));   \
      return NR_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

#define NR_CHECK_CUDA(call)                                                        \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      nr_set_error("%s failed: %s", #call, cudaGetErrorString(e__));               \
      return NR_ERR_CUDA;                                                          \
    }                                                                              \
  } while (0)

static inline int64_t nr_cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t nr_align(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int nr_pad4(int k) { return (k + 3) & ~3; }

// nn.Softplus(beta=100, threshold=20): x if 100x > 20 else log1p(exp(100x))/100.
__device__ __forceinline__ float nr_softplus100(float z) {
  float t = z * 100.0f;
  return t > 20.0f ? z : log1pf(expf(t)) / 100.0f;
}
// d/dz softplus100(z) = sigmoid(100 z) (1 above the threshold, as autograd gives).
__device__ __forceinline__ float nr_softplus100_grad(float z) {
  float t = z * 100.0f;
  return t > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-t));
}
__device__ __forceinline__ float nr_sigmoid(float x) { return 1.0f / (1.0f + expf(-x)); }

// torch.linspace(0, 1, n)[i] in fp32, bit-exact (symmetric FMA formula, see oracle/sampling.py).
__device__ __forceinline__ float nr_linspace01(int i, int n) {
  if (n <= 1) return 0.0f;
  float step = __fdiv_rn(1.0f, (float)(n - 1));
  return (i < n / 2) ? __fmaf_rn(step, (float)i, 0.0f) : __fmaf_rn(-step, (float)(n - 1 - i), 1.0f);
}
