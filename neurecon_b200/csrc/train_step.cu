// what follows the path in a NeuS training step -- the loss terms of Trainer.forward
// (neus.py:443-478) with their gradients, the gradient norm (train_util.py:5-16) and the Adam update
// (train.py:204-210) -- as HBM-bound kernels that keep every scalar on the device (the reference syncs with .item()).
#include "common.cuh"

namespace {

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float s = 0.f;
  if (w == 0) {
    s = l < (blockDim.x >> 5) ? red[l] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  }
  return s;   // valid in thread 0
}

// sums[0] = sum |rgb - t| * m (m = pixel weight: target_mask & mask_ignore, or 1), sums[1] = sum m,
// sums[2] = sum (|nabla| - 1)^2, sums[3] = sum bce(clamp(acc), target_mask)
__global__ void neus_loss_sums_kernel(const float* __restrict__ rgb, const float* __restrict__ target,
                                      const float* __restrict__ nablas, const float* __restrict__ acc,
                                      const uint8_t* __restrict__ target_mask, const uint8_t* __restrict__ mask_ignore,
                                      int64_t R, int64_t P, int with_mask, float* __restrict__ sums) {
  __shared__ float red[32];
  float s_img = 0.f, s_m = 0.f, s_eik = 0.f, s_bce = 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t r = i0; r < R; r += stride) {
    bool m = true;
    if (with_mask) m = target_mask[r] != 0;
    if (mask_ignore) m = m && mask_ignore[r] != 0;
    const float mf = (with_mask || mask_ignore) ? (m ? 1.f : 0.f) : 1.f;
    s_img += mf * (fabsf(rgb[3 * r] - target[3 * r]) + fabsf(rgb[3 * r + 1] - target[3 * r + 1]) + fabsf(rgb[3 * r + 2] - target[3 * r + 2]));
    s_m += mf;
    if (with_mask) {
      const float a = fminf(fmaxf(acc[r], 1e-3f), 1.0f - 1e-3f), t = target_mask[r] ? 1.f : 0.f;
      s_bce -= t * fmaxf(logf(a), -100.f) + (1.f - t) * fmaxf(logf(1.f - a), -100.f);   // F.binary_cross_entropy clamps log at -100
    }
  }
  for (int64_t i = i0; i < R * P; i += stride) {
    const float x = nablas[3 * i], y = nablas[3 * i + 1], z = nablas[3 * i + 2];
    const float e = sqrtf(x * x + y * y + z * z) - 1.f;
    s_eik += e * e;
  }
  float v;
  v = block_sum(s_img, red); if (threadIdx.x == 0) atomicAdd(&sums[0], v);
  v = block_sum(s_m, red);   if (threadIdx.x == 0) atomicAdd(&sums[1], v);
  v = block_sum(s_eik, red); if (threadIdx.x == 0) atomicAdd(&sums[2], v);
  v = block_sum(s_bce, red); if (threadIdx.x == 0) atomicAdd(&sums[3], v);
}

// losses[0..3] = loss_img, loss_eikonal, loss_mask, total; gradients of `total` w.r.t. rgb, nablas, acc
__global__ void neus_loss_grads_kernel(const float* __restrict__ rgb, const float* __restrict__ target,
                                       const float* __restrict__ nablas, const float* __restrict__ acc,
                                       const uint8_t* __restrict__ target_mask, const uint8_t* __restrict__ mask_ignore,
                                       int64_t R, int64_t P, int with_mask, float w_eik, float w_mask,
                                       const float* __restrict__ sums, float* __restrict__ losses,
                                       float* __restrict__ g_rgb, float* __restrict__ g_nablas, float* __restrict__ g_acc) {
  const bool weighted = with_mask || mask_ignore;
  const float inv_img = weighted ? 1.0f / (sums[1] + 1e-10f) : 1.0f / (3.0f * (float)R);
  const float inv_eik = w_eik / (float)(R * P), inv_bce = w_mask / (float)R;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i0 == 0) {
    const float li = sums[0] * inv_img, le = sums[2] * inv_eik, lm = with_mask ? sums[3] * inv_bce : 0.f;
    losses[0] = li; losses[1] = le; losses[2] = lm; losses[3] = li + le + lm;
  }
  for (int64_t r = i0; r < R; r += stride) {
    bool m = true;
    if (with_mask) m = target_mask[r] != 0;
    if (mask_ignore) m = m && mask_ignore[r] != 0;
    const float mf = weighted ? (m ? 1.f : 0.f) : 1.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float d = rgb[3 * r + c] - target[3 * r + c];
      g_rgb[3 * r + c] = mf * inv_img * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    float ga = 0.f;
    if (with_mask) {
      const float a0 = acc[r];
      if (a0 >= 1e-3f && a0 <= 1.0f - 1e-3f) {      // torch.clamp passes the gradient inside (and on) the bounds
        const float t = target_mask[r] ? 1.f : 0.f;
        ga = inv_bce * (-t / a0 + (1.f - t) / (1.f - a0));
      }
    }
    if (g_acc) g_acc[r] = ga;
  }
  for (int64_t i = i0; i < R * P; i += stride) {
    const float x = nablas[3 * i], y = nablas[3 * i + 1], z = nablas[3 * i + 2];
    const float nrm = sqrtf(x * x + y * y + z * z);
    const float k = nrm > 0.f ? 2.f * inv_eik * (nrm - 1.f) / nrm : 0.f;     // d|v|/dv = v/|v| (torch.norm: 0 at 0)
    g_nablas[3 * i] = k * x; g_nablas[3 * i + 1] = k * y; g_nablas[3 * i + 2] = k * z;
  }
}

struct TensorRef { float* p; const float* g; float* m; float* v; int64_t n; };

__global__ void sqsum_multi_kernel(const TensorRef* __restrict__ tab, int n_tensors, float* __restrict__ out) {
  __shared__ float red[32];
  float s = 0.f;
  for (int t = blockIdx.y; t < n_tensors; t += gridDim.y) {
    const TensorRef T = tab[t];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) s += T.g[i] * T.g[i];
  }
  const float v = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, v);
}

// torch.optim.Adam (no weight decay, no amsgrad): m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
__global__ void adam_multi_kernel(const TensorRef* __restrict__ tab, int n_tensors, float lr, float b1, float b2, float eps,
                                  float bc1, float bc2_sqrt) {
  const float step_size = lr / bc1;
  for (int t = blockIdx.y; t < n_tensors; t += gridDim.y) {
    const TensorRef T = tab[t];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) {
      const float g = T.g[i];
      const float m = T.m[i] + (g - T.m[i]) * (1.f - b1);          // lerp, as torch's _single_tensor_adam
      const float v = b2 * T.v[i] + (1.f - b2) * g * g;
      T.m[i] = m; T.v[i] = v;
      const float denom = sqrtf(v) / bc2_sqrt + eps;
      T.p[i] = T.p[i] - step_size * (m / denom);
    }
  }
}

// The same update with the step count and the learning rate in device memory, so that a CUDA graph of the whole
// training step replays with the right bias correction and a scheduler can change lr between replays.
__global__ void adam_tick_kernel(int64_t* step) { *step += 1; }

__global__ void adam_multi_dev_kernel(const TensorRef* __restrict__ tab, int n_tensors, const float* __restrict__ lr_dev,
                                      float b1, float b2, float eps, const int64_t* __restrict__ step_dev) {
  const double step = (double)*step_dev;
  const float bc1 = (float)(1.0 - pow((double)b1, step)), bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, step));
  const float step_size = *lr_dev / bc1;
  for (int t = blockIdx.y; t < n_tensors; t += gridDim.y) {
    const TensorRef T = tab[t];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < T.n; i += (int64_t)gridDim.x * blockDim.x) {
      const float g = T.g[i];
      const float m = T.m[i] + (g - T.m[i]) * (1.f - b1);
      const float v = b2 * T.v[i] + (1.f - b2) * g * g;
      T.m[i] = m; T.v[i] = v;
      const float denom = sqrtf(v) / bc2_sqrt + eps;
      T.p[i] = T.p[i] - step_size * (m / denom);
    }
  }
}

}  // namespace

extern "C" int nr_neus_loss(const float* rgb, const float* target_rgb, const float* nablas, const float* mask_volume,
                            const uint8_t* target_mask, const uint8_t* mask_ignore, int64_t R, int64_t P, float w_eikonal,
                            float w_mask, float* sums4, float* losses4, float* g_rgb, float* g_nablas, float* g_mask_volume,
                            void* stream) {
  NR_CHECK_ARG(R > 0 && P > 0, "nr_neus_loss: bad sizes");
  NR_CHECK_ARG(rgb && target_rgb && nablas && sums4 && losses4 && g_rgb && g_nablas, "nr_neus_loss: null pointer");
  const int with_mask = target_mask != nullptr;
  NR_CHECK_ARG(!with_mask || (mask_volume && g_mask_volume), "nr_neus_loss: the mask loss needs mask_volume and its gradient buffer");
  cudaStream_t st = (cudaStream_t)stream;
  NR_CHECK_CUDA(cudaMemsetAsync(sums4, 0, 4 * sizeof(float), st));
  const int blocks = (int)(nr_cdiv(R * P, 256) < 1184 ? nr_cdiv(R * P, 256) : 1184);
  neus_loss_sums_kernel<<<blocks, 256, 0, st>>>(rgb, target_rgb, nablas, mask_volume, target_mask, mask_ignore, R, P, with_mask, sums4);
  NR_CHECK_LAUNCH("neus_loss_sums_kernel");
  neus_loss_grads_kernel<<<blocks, 256, 0, st>>>(rgb, target_rgb, nablas, mask_volume, target_mask, mask_ignore, R, P, with_mask,
                                                 w_eikonal, w_mask, sums4, losses4, g_rgb, g_nablas, g_mask_volume);
  NR_CHECK_LAUNCH("neus_loss_grads_kernel");
  return NR_OK;
}

// table: n_tensors records of 5 x 8 bytes {param*, grad*, exp_avg*, exp_avg_sq*, numel} in device memory
extern "C" int nr_grad_sqsum(const void* table, int32_t n_tensors, float* out, void* stream) {
  NR_CHECK_ARG(table && out && n_tensors > 0, "nr_grad_sqsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  NR_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  dim3 grid(32, n_tensors < 64 ? n_tensors : 64);
  sqsum_multi_kernel<<<grid, 256, 0, st>>>((const TensorRef*)table, n_tensors, out);
  NR_CHECK_LAUNCH("sqsum_multi_kernel");
  return NR_OK;
}

extern "C" int nr_adam_step(const void* table, int32_t n_tensors, float lr, float beta1, float beta2, float eps, int64_t step,
                            void* stream) {
  NR_CHECK_ARG(table && n_tensors > 0 && step >= 1, "nr_adam_step: bad arguments");
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  dim3 grid(32, n_tensors < 64 ? n_tensors : 64);
  adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const TensorRef*)table, n_tensors, lr, beta1, beta2, eps, (float)bc1,
                                                            (float)sqrt(bc2));
  NR_CHECK_LAUNCH("adam_multi_kernel");
  return NR_OK;
}

extern "C" int nr_adam_step_dev(const void* table, int32_t n_tensors, const float* lr_dev, float beta1, float beta2, float eps,
                                int64_t* step_dev, void* stream) {
  NR_CHECK_ARG(table && n_tensors > 0 && lr_dev && step_dev, "nr_adam_step_dev: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  adam_tick_kernel<<<1, 1, 0, st>>>(step_dev);
  NR_CHECK_LAUNCH("adam_tick_kernel");
  dim3 grid(32, n_tensors < 64 ? n_tensors : 64);
  adam_multi_dev_kernel<<<grid, 256, 0, st>>>((const TensorRef*)table, n_tensors, lr_dev, beta1, beta2, eps, step_dev);
  NR_CHECK_LAUNCH("adam_multi_dev_kernel");
  return NR_OK;
}
