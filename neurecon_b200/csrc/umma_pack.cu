// A-operand (weight) images of the fused MLP kernels, packed on the device: [128 rows x 64 k] K-major SWIZZLE_128B tiles
// (layout: umma.cuh) ordered (k-chunk, M-tile[, hi / lo]).  The images are re-packed after every optimiser step (the
// up-sampler of a training iteration runs on the inference kernels); the torch composition this replaces cost ~20 (fp16 tier)
// to ~60 (fp16x2 tier) small launches per iteration.
#include "common.cuh"
#include "umma.cuh"

namespace {
__global__ void umma_pack_a_kernel(const float* __restrict__ W, int64_t row_stride, int64_t col_stride, int rows, int K, int n_mt,
                                   int n_kc, int mode, uint8_t* __restrict__ img) {
  const int parts = mode == 2 ? 2 : 1;
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)n_kc * n_mt * parts * 128 * 8) return;
  const int c8 = (int)(idx & 7), r = (int)((idx >> 3) & 127);
  const int chunk = (int)(idx >> 10);
  const int part = parts == 2 ? (chunk & 1) : 0, cm = parts == 2 ? (chunk >> 1) : chunk;
  const int mt = cm % n_mt, kc = cm / n_mt;
  const int row = mt * 128 + r, k0 = kc * 64 + c8 * 8;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = (row < rows && k0 + j < K) ? W[row * row_stride + (k0 + j) * col_stride] : 0.0f;
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (mode == 1) {
      o[j] = umma::pack_bf16(v[2 * j], v[2 * j + 1]);
    } else {
      const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
      if (part == 0) {
        o[j] = *reinterpret_cast<const uint32_t*>(&h);
      } else {   // lo part of the split-precision image, scaled by 2^12 (csrc/mlp_rev_split.cu)
        const float2 hf = __half22float2(h);
        o[j] = umma::pack_f16((v[2 * j] - hf.x) * 4096.0f, (v[2 * j + 1] - hf.y) * 4096.0f);
      }
    }
  }
  *reinterpret_cast<uint4*>(img + (size_t)chunk * 16384 + r * 128 + ((c8 ^ (r & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
}
}  // namespace

extern "C" size_t nr_umma_pack_a_bytes(int32_t n_mt, int32_t k_pad, int32_t mode) {
  return (size_t)((k_pad + 63) / 64) * n_mt * (mode == 2 ? 2 : 1) * 16384;
}

extern "C" int nr_umma_pack_a(const float* W, int64_t row_stride, int64_t col_stride, int32_t rows, int32_t K, int32_t n_mt,
                              int32_t k_pad, int32_t mode, void* img, void* stream) {
  NR_CHECK_ARG(W && img && rows >= 1 && K >= 1 && n_mt >= 1 && rows <= n_mt * 128 && k_pad >= K && mode >= 0 && mode <= 2,
               "nr_umma_pack_a: bad arguments");
  NR_CHECK_ARG(((uintptr_t)img & 15) == 0, "nr_umma_pack_a: img must be 16-byte aligned");
  const int n_kc = (k_pad + 63) / 64;
  const int64_t total = (int64_t)n_kc * n_mt * (mode == 2 ? 2 : 1) * 128 * 8;
  umma_pack_a_kernel<<<(unsigned)nr_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(W, row_stride, col_stride, rows, K, n_mt, n_kc, mode,
                                                                                   (uint8_t*)img);
  NR_CHECK_LAUNCH("umma_pack_a_kernel");
  return NR_OK;
}
