// Ray / sphere helpers of the NeRF++ background (utils/rend_util.py:188-234) and the inverted-sphere sample generator
// of VolSDF's background branch (volsdf.py:456-467).  Thread per (ray, sample); un-fused fp32 ops in the order the
// reference's separate elementwise kernels apply them.
#include "common.cuh"

namespace {

struct RayDots { float o2, od; };
// torch.sum(o**2, -1) and torch.sum(o*d, -1): the three products summed left to right
__device__ __forceinline__ RayDots ray_dots(const float* __restrict__ o, const float* __restrict__ d, int64_t ray) {
  const float ox = o[3 * ray], oy = o[3 * ray + 1], oz = o[3 * ray + 2];
  const float dx = d[3 * ray], dy = d[3 * ray + 1], dz = d[3 * ray + 2];
  RayDots r;
  r.o2 = __fadd_rn(__fadd_rn(__fmul_rn(ox, ox), __fmul_rn(oy, oy)), __fmul_rn(oz, oz));
  r.od = __fadd_rn(__fadd_rn(__fmul_rn(ox, dx), __fmul_rn(oy, dy)), __fmul_rn(oz, dz));
  return r;
}

// rend_util.py:188-210
__global__ void sphere_intersection_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t R, float r2,
                                           float* __restrict__ near, float* __restrict__ far, uint8_t* __restrict__ mask) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R) return;
  const RayDots q = ray_dots(o, d, i);
  // under_sqrt = ray_cam_dot ** 2 + r ** 2 - rayso_norm_square; r ** 2 is a Python float (a double), rounded to fp32 once
  // when it meets the tensor: r2 is that value
  const float under = __fsub_rn(__fadd_rn(__fmul_rn(q.od, q.od), r2), q.o2);
  const bool hit = under > 0.0f;
  float n = 0.0f, f = 0.0f;
  if (hit) {
    const float sq = __fsqrt_rn(under);
    n = __fsub_rn(-sq, q.od);
    f = __fsub_rn(sq, q.od);
  }
  near[i] = fmaxf(n, 0.0f);
  far[i] = fmaxf(f, 0.0f);
  if (mask) mask[i] = hit ? 1 : 0;
}

// rend_util.py:213-234; d_vals [R,N] for radii rs [R,N].  The reference asserts under_sqrt > 0 on the host; here the
// violations are counted on the device (*bad += 1 per offending entry) and the result there is NaN like torch.sqrt's.
__device__ __forceinline__ float dval_from_radius(const RayDots& q, float rs, bool far_end, int* bad) {
  const float under = __fsub_rn(__fmul_rn(rs, rs), __fsub_rn(q.o2, __fmul_rn(q.od, q.od)));
  if (!(under > 0.0f) && bad) atomicAdd(bad, 1);
  const float sq = __fsqrt_rn(under);
  return far_end ? __fadd_rn(-q.od, sq) : fmaxf(__fsub_rn(-q.od, sq), 0.0f);
}
__global__ void dvals_from_radius_kernel(const float* __restrict__ o, const float* __restrict__ d,
                                         const float* __restrict__ rs, int64_t R, int N, int far_end,
                                         float* __restrict__ d_vals, int* __restrict__ bad) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= R * N) return;
  const int64_t ray = idx / N;
  d_vals[idx] = dval_from_radius(ray_dots(o, d, ray), rs[idx], far_end != 0, bad);
}

// volsdf.py:456-467: radii r / flip(linspace(0,1,n+2)[1:-1]) (stratified jitter with u [R,n] if given), their far-end
// depths, and the inverted-sphere inputs x_out = [p / rs, 1 / rs] of NeRF.forward.
__global__ void volsdf_outside_points_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t R,
                                             float radius, int n_out, const float* __restrict__ u,
                                             float* __restrict__ d_out, float* __restrict__ x_out, int* __restrict__ bad) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= R * n_out) return;
  const int64_t ray = idx / n_out;
  const int j = (int)(idx - ray * n_out);
  auto radius_at = [&](int k) { return __fdiv_rn(radius, nr_linspace01(n_out - k, n_out + 2)); };   // flipped linspace
  float rs = radius_at(j);
  if (u) {
    const float lo = j == 0 ? rs : __fmul_rn(0.5f, __fadd_rn(rs, radius_at(j - 1)));
    const float hi = j == n_out - 1 ? rs : __fmul_rn(0.5f, __fadd_rn(radius_at(j + 1), rs));
    rs = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u[idx]));
  }
  const float dv = dval_from_radius(ray_dots(o, d, ray), rs, true, bad);
  d_out[idx] = dv;
  float* x = x_out + idx * 4;
#pragma unroll
  for (int c = 0; c < 3; ++c) x[c] = __fdiv_rn(__fadd_rn(o[3 * ray + c], __fmul_rn(d[3 * ray + c], dv)), rs);
  x[3] = __fdiv_rn(1.0f, rs);
}

}  // namespace

extern "C" int nr_sphere_intersection(const float* rays_o, const float* rays_d, int64_t R, double r, float* near,
                                      float* far, uint8_t* mask, void* stream) {
  NR_CHECK_ARG(R >= 0 && rays_o && rays_d && near && far, "nr_sphere_intersection: bad arguments");
  if (R == 0) return NR_OK;
  sphere_intersection_kernel<<<(unsigned)nr_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, R, (float)(r * r), near, far,
                                                                                            mask);
  NR_CHECK_LAUNCH("sphere_intersection_kernel");
  return NR_OK;
}

extern "C" int nr_dvals_from_radius(const float* rays_o, const float* rays_d, const float* rs, int64_t R, int32_t N,
                                    int32_t far_end, float* d_vals, int32_t* bad_count, void* stream) {
  NR_CHECK_ARG(R >= 0 && N >= 1 && rays_o && rays_d && rs && d_vals, "nr_dvals_from_radius: bad arguments");
  if (R == 0) return NR_OK;
  dvals_from_radius_kernel<<<(unsigned)nr_cdiv(R * N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, rs, R, N, far_end,
                                                                                          d_vals, bad_count);
  NR_CHECK_LAUNCH("dvals_from_radius_kernel");
  return NR_OK;
}

extern "C" int nr_volsdf_outside_points(const float* rays_o, const float* dirs, int64_t R, float radius, int32_t n_out,
                                        const float* u, float* d_out, float* x_out, int32_t* bad_count, void* stream) {
  NR_CHECK_ARG(R >= 0 && n_out >= 1 && rays_o && dirs && d_out && x_out, "nr_volsdf_outside_points: bad arguments");
  if (R == 0) return NR_OK;
  volsdf_outside_points_kernel<<<(unsigned)nr_cdiv(R * n_out, 256), 256, 0, (cudaStream_t)stream>>>(
      rays_o, dirs, R, radius, n_out, u, d_out, x_out, bad_count);
  NR_CHECK_LAUNCH("volsdf_outside_points_kernel");
  return NR_OK;
}
