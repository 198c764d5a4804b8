// Weight normalisation of every layer of a network in ONE launch, and its backward in one more (training glue:
// models/base.py:118-129, 226-227 of the reference re-normalise W = g v / ||v||_row in a pre-forward hook of every layer --
// three torch kernels per layer forward and half a dozen backward, ~130 launches per training step for the two networks).
// The layers are described by a table passed BY VALUE as a kernel argument (<= 16 entries, 1.5 KB): no host-to-device copy,
// so the launch can be captured into a CUDA graph; one warp per weight row.
#include <string.h>
#include "common.cuh"

namespace {

struct WnEntry {          // 96 bytes = twelve 8-byte words, packed by the caller
  const float* v;         // [rows, cols]
  const float* g;         // [rows]
  float* W;               // [rows, ldw] effective weight (pad columns zeroed)
  const float* dW;        // [rows, lddw] gradient w.r.t. the effective weight (backward only)
  float* dv;              // [rows, cols]
  float* dg;              // [rows]
  int64_t rows, cols, ldw, lddw;
  double scale;           // W = scale * g v / ||v||  (1/sqrt 2 on the skip layer)
  int64_t row0;           // first global row of this entry
};

constexpr int kMaxEntries = 16;
struct WnTable { WnEntry e[kMaxEntries]; };

constexpr unsigned kFull = 0xffffffffu;
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
  return x;
}
__device__ __forceinline__ const WnEntry* find_entry(const WnTable& tab, int n, int64_t row) {
  int e = 0;
  while (e + 1 < n && tab.e[e + 1].row0 <= row) ++e;
  return &tab.e[e];
}

__global__ void weight_norm_fwd_kernel(const __grid_constant__ WnTable tab, int n, int64_t total_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= total_rows) return;
  const WnEntry* E = find_entry(tab, n, row);
  const int64_t o = row - E->row0;
  const float* v = E->v + o * E->cols;
  float ss = 0.0f;
  for (int64_t i = lane; i < E->cols; i += 32) ss += v[i] * v[i];
  ss = warp_sum(ss);
  const float k = (float)E->scale * E->g[o] / sqrtf(ss);
  float* w = E->W + o * E->ldw;
  for (int64_t i = lane; i < E->ldw; i += 32) w[i] = i < E->cols ? k * v[i] : 0.0f;
}

__global__ void weight_norm_bwd_kernel(const __grid_constant__ WnTable tab, int n, int64_t total_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= total_rows) return;
  const WnEntry* E = find_entry(tab, n, row);
  const int64_t o = row - E->row0;
  const float* v = E->v + o * E->cols;
  const float* dW = E->dW + o * E->lddw;
  float ss = 0.0f, dot = 0.0f;
  for (int64_t i = lane; i < E->cols; i += 32) { ss += v[i] * v[i]; dot += dW[i] * v[i]; }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float inv = rsqrtf(ss), s = (float)E->scale;
  const float dgo = s * dot * inv;                 // dL/dg = scale * <dW, v / |v|>
  const float k = s * E->g[o] * inv;               // dL/dv = scale g / |v| (dW - vhat <dW, vhat>)
  float* dv = E->dv + o * E->cols;
  for (int64_t i = lane; i < E->cols; i += 32) dv[i] = k * (dW[i] - v[i] * (dot * inv * inv));
  if (lane == 0) E->dg[o] = dgo;
}

}  // namespace

extern "C" int nr_weight_norm(const void* table, int32_t n_entries, int64_t total_rows, int32_t backward, void* stream) {
  NR_CHECK_ARG(table && n_entries >= 1 && n_entries <= kMaxEntries && total_rows >= 0, "nr_weight_norm: 1..%d entries", kMaxEntries);
  if (total_rows == 0) return NR_OK;
  WnTable tab = {};
  memcpy(tab.e, table, sizeof(WnEntry) * (size_t)n_entries);       // `table` is HOST memory
  const unsigned blocks = (unsigned)nr_cdiv(total_rows * 32, 256);
  if (backward)
    weight_norm_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tab, n_entries, total_rows);
  else
    weight_norm_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tab, n_entries, total_rows);
  NR_CHECK_LAUNCH(backward ? "weight_norm_bwd_kernel" : "weight_norm_fwd_kernel");
  return NR_OK;
}
