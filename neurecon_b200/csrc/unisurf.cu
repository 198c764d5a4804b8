// UNISURF: root finding along the ray (uniform proposals -> first sign change -> secant), the
// interval / free-space sampler and occupancy compositing.
//
// Reference semantics: models/ray_casting.py:11-30 (secant), :35-160 (root_finding_surface_points),
// models/frameworks/unisurf.py:124-131 (near/far), :147-203 (samplers), :216-240 (compositing).
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// dirs, near/far (near_far_from_sphere, keepdim=False) and the n_steps uniform proposal points
__global__ void unisurf_ray_setup_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t R,
                                         float radius, float near_bypass, float far_bypass, int n_steps,
                                         float* __restrict__ dirs, float* __restrict__ near_out,
                                         float* __restrict__ far_out, float* __restrict__ pts) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)4 + warp;
  if (ray >= R) return;
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
  const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-12f);
  dx = __fdiv_rn(dx, nrm); dy = __fdiv_rn(dy, nrm); dz = __fdiv_rn(dz, nrm);
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(ox, dx), __fmul_rn(oy, dy)), __fmul_rn(oz, dz));
  float nr_ = fmaxf(-dot - radius, 0.0f), fr_ = fmaxf(-dot + radius, radius);
  if (!isnan(near_bypass)) nr_ = near_bypass;
  if (!isnan(far_bypass)) fr_ = far_bypass;
  if (lane == 0) {
    dirs[3 * ray] = dx; dirs[3 * ray + 1] = dy; dirs[3 * ray + 2] = dz;
    near_out[ray] = nr_; far_out[ray] = fr_;
  }
  for (int i = lane; i < n_steps; i += 32) {
    const float t = nr_linspace01(i, n_steps);
    const float d = __fadd_rn(__fmul_rn(nr_, __fsub_rn(1.0f, t)), __fmul_rn(fr_, t));
    float* p = pts + (ray * (int64_t)n_steps + i) * 3;
    p[0] = __fadd_rn(ox, __fmul_rn(d, dx)); p[1] = __fadd_rn(oy, __fmul_rn(d, dy)); p[2] = __fadd_rn(oz, __fmul_rn(d, dz));
  }
}

// ray_casting.py:86-137: first sign change of (val - tau) along the proposals, masks, secant bracket
// and the first secant estimate with its point.  state: [5][R] = d_low, f_low, d_high, f_high, d_pred.
__global__ void unisurf_crossing_kernel(const float* __restrict__ val, const float* __restrict__ rays_o,
                                        const float* __restrict__ dirs, const float* __restrict__ near,
                                        const float* __restrict__ far, int64_t R, int n_steps, float tau,
                                        float* __restrict__ state, uint8_t* __restrict__ mask,
                                        uint8_t* __restrict__ mask_sign_change, uint8_t* __restrict__ mask_0_free,
                                        float* __restrict__ pts_pred) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)4 + warp;
  if (ray >= R) return;
  const float* v = val + ray * (int64_t)n_steps;
  // cost_i = sign(v_i v_{i+1}) * (n_steps - i), last entry sign = +1; argmin with first-index tie break
  float best = INFINITY;
  int best_i = 0;
  for (int i = lane; i < n_steps; i += 32) {
    float sgn = 1.0f;
    if (i < n_steps - 1) {
      const float prod = (v[i] - tau) * (v[i + 1] - tau);
      sgn = prod > 0.0f ? 1.0f : (prod < 0.0f ? -1.0f : 0.0f);
    }
    const float c = sgn * (float)(n_steps - i);
    if (c < best) { best = c; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(kFull, best, o);
    const int oi = __shfl_xor_sync(kFull, best_i, o);
    if (ob < best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
  }
  if (lane == 0) {
    const float nr_ = near[ray], fr_ = far[ray];
    auto dprop = [&](int i) {
      const float t = nr_linspace01(i, n_steps);
      return __fadd_rn(__fmul_rn(nr_, __fsub_rn(1.0f, t)), __fmul_rn(fr_, t));
    };
    const bool sign_change = best < 0.0f;
    const bool free0 = (v[0] - tau) > 0.0f;
    const bool pos_to_neg = (v[best_i] - tau) > 0.0f;
    const bool m = sign_change && pos_to_neg && free0;
    const int i2 = min(best_i + 1, n_steps - 1);
    const float d_high = dprop(best_i), f_high = v[best_i] - tau;
    const float d_low = dprop(i2), f_low = v[i2] - tau;
    float d_pred = m ? (-f_low * (d_high - d_low) / (f_high - f_low) + d_low) : nr_;  // ray_casting.py:15
    state[0 * R + ray] = d_low; state[1 * R + ray] = f_low; state[2 * R + ray] = d_high; state[3 * R + ray] = f_high;
    state[4 * R + ray] = d_pred;
    mask[ray] = m; mask_sign_change[ray] = sign_change; mask_0_free[ray] = free0;
    pts_pred[3 * ray] = rays_o[3 * ray] + d_pred * dirs[3 * ray];
    pts_pred[3 * ray + 1] = rays_o[3 * ray + 1] + d_pred * dirs[3 * ray + 1];
    pts_pred[3 * ray + 2] = rays_o[3 * ray + 2] + d_pred * dirs[3 * ray + 2];
  }
}

// One secant update (ray_casting.py:16-29) given f_mid = sdf(p_mid); rays outside the mask are left alone.
__global__ void unisurf_secant_kernel(const float* __restrict__ f_mid_raw, float tau, const float* __restrict__ rays_o,
                                      const float* __restrict__ dirs, const uint8_t* __restrict__ mask, int64_t R,
                                      float* __restrict__ state, float* __restrict__ pts_pred) {
  const int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (ray >= R || !mask[ray]) return;
  float d_low = state[0 * R + ray], f_low = state[1 * R + ray], d_high = state[2 * R + ray], f_high = state[3 * R + ray];
  const float d_pred = state[4 * R + ray], f_mid = f_mid_raw[ray] - tau;
  if (f_mid < 0.0f) { d_low = d_pred; f_low = f_mid; } else { d_high = d_pred; f_high = f_mid; }
  const float dn = -f_low * (d_high - d_low) / (f_high - f_low) + d_low;
  state[0 * R + ray] = d_low; state[1 * R + ray] = f_low; state[2 * R + ray] = d_high; state[3 * R + ray] = f_high;
  state[4 * R + ray] = dn;
  pts_pred[3 * ray] = rays_o[3 * ray] + dn * dirs[3 * ray];
  pts_pred[3 * ray + 1] = rays_o[3 * ray + 1] + dn * dirs[3 * ray + 1];
  pts_pred[3 * ray + 2] = rays_o[3 * ray + 2] + dn * dirs[3 * ray + 2];
}

// Outputs of root finding (ray_casting.py:140-151) + the interval / free-space sampler and merge
// (unisurf.py:147-207).  u_int [R,n_query] / u_free [R,n_free]: stratified jitter uniforms or NULL.
__global__ void unisurf_sample_kernel(const float* __restrict__ rays_o, const float* __restrict__ dirs,
                                      const float* __restrict__ near, const float* __restrict__ far,
                                      const float* __restrict__ state, const uint8_t* __restrict__ mask,
                                      const uint8_t* __restrict__ mask_sign_change,
                                      const uint8_t* __restrict__ mask_0_free, int64_t R, float interval,
                                      float too_close, int n_query, int n_free, const float* __restrict__ u_int,
                                      const float* __restrict__ u_free, float* __restrict__ depth_surface,
                                      float* __restrict__ surface_pts, float* __restrict__ d_all,
                                      float* __restrict__ pts) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)4 + warp;
  if (ray >= R) return;
  const int M = n_query + n_free;
  float* sf = smem + (size_t)warp * 2 * M;  // free | interval
  float* si = sf + n_free;
  float* so = sf + M;
  const float nr_ = near[ray], fr_ = far[ray];
  const bool m = mask[ray];
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  const float dx = dirs[3 * ray], dy = dirs[3 * ray + 1], dz = dirs[3 * ray + 2];
  float d_out = m ? state[4 * R + ray] : fr_;          // fill_inf=False: far when no surface is hit
  if (!mask_0_free[ray]) d_out = 0.0f;                 // origin already occupied
  const float d_pred = fmaxf(fminf(d_out, fr_), nr_);  // unisurf.py:150
  const float d_upper = fminf(d_pred + interval, fr_);
  float d_lower = fmaxf(d_pred - interval, nr_);
  if (lane == 0) {
    depth_surface[ray] = d_pred;
    const float ds = state[4 * R + ray];
    surface_pts[3 * ray] = m ? ox + ds * dx : 1.0f;     // pt_pred: ones where no surface (ray_casting.py:139-140)
    surface_pts[3 * ray + 1] = m ? oy + ds * dy : 1.0f;
    surface_pts[3 * ray + 2] = m ? oz + ds * dz : 1.0f;
  }
  auto lerp_t = [](float a, float b, float t) { return __fadd_rn(__fmul_rn(a, __fsub_rn(1.0f, t)), __fmul_rn(b, t)); };
  for (int i = lane; i < n_query; i += 32) {
    if (u_int) {
      const float lo = lerp_t(d_lower, d_upper, nr_linspace01(i, n_query + 1));
      const float hi = lerp_t(d_lower, d_upper, nr_linspace01(i + 1, n_query + 1));
      si[i] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u_int[ray * (int64_t)n_query + i]));
    } else {
      si[i] = lerp_t(d_lower, d_upper, nr_linspace01(i, n_query));
    }
  }
  d_lower = fmaxf(d_lower, nr_ + (fr_ - nr_) * too_close);   // unisurf.py:176
  if (!mask_sign_change[ray]) d_lower = fr_;                 // no intersection: sample the whole ray
  if (d_lower < 1e-10f) d_lower = fr_;
  for (int i = lane; i < n_free; i += 32) {
    if (u_free) {
      const float lo = lerp_t(nr_, d_lower, nr_linspace01(i, n_free + 1));
      const float hi = lerp_t(nr_, d_lower, nr_linspace01(i + 1, n_free + 1));
      sf[i] = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u_free[ray * (int64_t)n_free + i]));
    } else {
      sf[i] = lerp_t(nr_, d_lower, nr_linspace01(i, n_free));
    }
  }
  __syncwarp();
  // both lists are non-decreasing: stable merge (free first on ties) == sort(cat(free, interval))
  for (int i = lane; i < n_free; i += 32) {
    const float v = sf[i];
    int lo = 0, hi = n_query;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (si[mid] < v) lo = mid + 1; else hi = mid; }
    so[i + lo] = v;
  }
  for (int j = lane; j < n_query; j += 32) {
    const float v = si[j];
    int lo = 0, hi = n_free;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sf[mid] <= v) lo = mid + 1; else hi = mid; }
    so[j + lo] = v;
  }
  __syncwarp();
  for (int i = lane; i < M; i += 32) {
    const float d = so[i];
    d_all[ray * (int64_t)M + i] = d;
    float* p = pts + (ray * (int64_t)M + i) * 3;
    p[0] = __fadd_rn(ox, __fmul_rn(dx, d)); p[1] = __fadd_rn(oy, __fmul_rn(dy, d)); p[2] = __fadd_rn(oz, __fmul_rn(dz, d));
  }
}

// unisurf.py:216-240: alpha = e^{-x} / (1 + e^{-x}) (NaN below ~ -88.7 like the reference),
// w = alpha * exclusive cumprod(1 - alpha + 1e-10), weights sit ON the samples.
__global__ void unisurf_composite_kernel(const float* __restrict__ logits, const float* __restrict__ nablas,
                                         const float* __restrict__ radiance, const float* __restrict__ d_all,
                                         int64_t R, int M, int white_bkgd, float* __restrict__ rgb,
                                         float* __restrict__ depth, float* __restrict__ acc, float* __restrict__ normals,
                                         float* __restrict__ alpha_out, float* __restrict__ w_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)4 + warp;
  if (ray >= R) return;
  float carry = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  for (int base = 0; base < M; base += 32) {
    const int i = base + lane;
    const bool ok = i < M;
    float alpha = 0.0f;
    if (ok) {
      const float odds = expf(-logits[ray * (int64_t)M + i]);
      alpha = __fdiv_rn(odds, __fadd_rn(1.0f, odds));
    }
    const float f = ok ? __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f) : 1.0f;
    float incl = f;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl *= t;
    }
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float w = alpha * (carry * excl);
    carry *= __shfl_sync(kFull, incl, 31);
    if (ok) {
      if (alpha_out) alpha_out[ray * (int64_t)M + i] = alpha;
      if (w_out) w_out[ray * (int64_t)M + i] = w;
      const float* c = radiance + (ray * (int64_t)M + i) * 3;
      ar += w * c[0]; ag += w * c[1]; ab += w * c[2];
      ad += w * d_all[ray * (int64_t)M + i];
      aw += w;
      if (nablas) {
        const float* nb = nablas + (ray * (int64_t)M + i) * 3;
        const float x = nb[0], y = nb[1], z = nb[2];
        const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
        nx += w * x * inv; ny += w * y * inv; nz += w * z * inv;
      }
    }
  }
  ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); ad = warp_sum(ad); aw = warp_sum(aw);
  if (nablas) { nx = warp_sum(nx); ny = warp_sum(ny); nz = warp_sum(nz); }
  if (lane == 0) {
    if (white_bkgd) { ar += 1.0f - aw; ag += 1.0f - aw; ab += 1.0f - aw; }
    rgb[3 * ray] = ar; rgb[3 * ray + 1] = ag; rgb[3 * ray + 2] = ab;
    depth[ray] = ad / (aw + 1e-10f);
    acc[ray] = aw;
    if (normals) { normals[3 * ray] = nx; normals[3 * ray + 1] = ny; normals[3 * ray + 2] = nz; }
  }
}

// Staged variant (M <= kUsMaxM), the structure of the NeuS / VolSDF compositing kernels: the block's four rays arrive by
// bulk copies issued by one thread (or 16-byte cp.async when a pointer is unaligned / the block is ragged), each lane owns
// ceil(M/32) CONSECUTIVE samples, one multiplicative warp scan per ray.  These kernels are bound by instruction issue.
// alpha keeps the IEEE division: inf / inf = NaN for logits below ~ -88.7 is the reference's behaviour.
constexpr int kUsMaxM = 256;
template <int kSeg>
__global__ void unisurf_composite_staged_kernel(const float* __restrict__ logits, const float* __restrict__ nablas,
                                                const float* __restrict__ radiance, const float* __restrict__ d_all,
                                                int64_t R, int M, int white_bkgd, float* __restrict__ rgb,
                                                float* __restrict__ depth, float* __restrict__ acc,
                                                float* __restrict__ normals, float* __restrict__ alpha_out,
                                                float* __restrict__ w_out, int vec16) {
  extern __shared__ __align__(16) float ustage[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray0 = blockIdx.x * (int64_t)4, ray = ray0 + warp;
  const int nrays = (int)((R - ray0) < 4 ? (R - ray0) : 4);
  float* s_lg = ustage;                 // [4][M]
  float* s_d = s_lg + 4 * M;            // [4][M]
  float* s_rad = s_d + 4 * M;           // [4][3M]
  float* s_nb = s_rad + 12 * M;         // [4][3M]
  if (vec16 && nrays == 4) {
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
      umma::mbar_init(&bar, 1);
      umma::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t b1 = 16u * M, b3 = 48u * M;
      umma::mbar_arrive_expect_tx(&bar, 2 * b1 + b3 + (nablas ? b3 : 0u));
      umma::bulk_g2s(s_lg, logits + ray0 * (int64_t)M, b1, &bar);
      umma::bulk_g2s(s_d, d_all + ray0 * (int64_t)M, b1, &bar);
      umma::bulk_g2s(s_rad, radiance + ray0 * (int64_t)(3 * M), b3, &bar);
      if (nablas) umma::bulk_g2s(s_nb, nablas + ray0 * (int64_t)(3 * M), b3, &bar);
    }
    umma::mbar_wait(&bar, 0);
  } else {
    auto stage_in = [&](float* dst, const float* src, int row) {
      const float* g = src + ray0 * (int64_t)row;
      const int n = nrays * row;
      const int n4 = vec16 ? n >> 2 : 0;
      for (int i = threadIdx.x; i < n4; i += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 4 * i)), "l"(g + 4 * i) : "memory");
      for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + i)), "l"(g + i) : "memory");
    };
    stage_in(s_lg, logits, M);
    stage_in(s_d, d_all, M);
    stage_in(s_rad, radiance, 3 * M);
    if (nablas) stage_in(s_nb, nablas, 3 * M);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
  if (ray >= R) return;
  const float* lg = s_lg + warp * M;
  const float* dd = s_d + warp * M;
  const float* rad = s_rad + warp * 3 * M;
  const float* nb = s_nb + warp * 3 * M;
  const int seg = (M + 31) >> 5;
  const int i0 = lane * seg;
  float al[kSeg], tr[kSeg];
  float prod = 1.0f;
#pragma unroll
  for (int k = 0; k < kSeg; ++k) {
    const int i = i0 + k;
    float alpha = 0.0f;
    if (k < seg && i < M) {
      const float odds = __expf(-lg[i]);
      alpha = __fdiv_rn(odds, 1.0f + odds);
    }
    al[k] = alpha;
    tr[k] = prod;
    if (k < seg && i < M) prod *= (1.0f - alpha) + 1e-10f;
  }
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl *= t;
  }
  float before = __shfl_up_sync(kFull, incl, 1);
  if (lane == 0) before = 1.0f;
  float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
#pragma unroll
  for (int k = 0; k < kSeg; ++k) {
    const int i = i0 + k;
    if (k < seg && i < M) {
      const float w = al[k] * (before * tr[k]);
      if (alpha_out) alpha_out[ray * (int64_t)M + i] = al[k];
      if (w_out) w_out[ray * (int64_t)M + i] = w;
      ar += w * rad[3 * i]; ag += w * rad[3 * i + 1]; ab += w * rad[3 * i + 2];
      ad += w * dd[i];
      aw += w;
      if (nablas) {
        const float x = nb[3 * i], y = nb[3 * i + 1], z = nb[3 * i + 2];
        const float inv = rsqrtf(fmaxf(x * x + y * y + z * z, 1e-24f));   // = 1 / max(|v|, 1e-12)
        nx += w * x * inv; ny += w * y * inv; nz += w * z * inv;
      }
    }
  }
  ar = warp_sum(ar); ag = warp_sum(ag); ab = warp_sum(ab); ad = warp_sum(ad); aw = warp_sum(aw);
  if (nablas) { nx = warp_sum(nx); ny = warp_sum(ny); nz = warp_sum(nz); }
  if (lane == 0) {
    if (white_bkgd) { ar += 1.0f - aw; ag += 1.0f - aw; ab += 1.0f - aw; }
    rgb[3 * ray] = ar; rgb[3 * ray + 1] = ag; rgb[3 * ray + 2] = ab;
    depth[ray] = ad / (aw + 1e-10f);
    acc[ray] = aw;
    if (normals) { normals[3 * ray] = nx; normals[3 * ray + 1] = ny; normals[3 * ray + 2] = nz; }
  }
}

// One sphere-tracing update (ray_casting.py:178-183): d[mask] += sdf[mask]; rays leaving [0, far] are dropped;
// emits the next query points.
__global__ void sphere_trace_step_kernel(const float* __restrict__ val, const float* __restrict__ rays_o,
                                         const float* __restrict__ dirs, float far, int64_t R, float* __restrict__ d,
                                         uint8_t* __restrict__ mask, float* __restrict__ pts) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R) return;
  float di = d[i];
  uint8_t m = mask[i];
  if (val && m) di = __fadd_rn(di, val[i]);
  if (di > far || di < 0.0f) m = 0;
  d[i] = di; mask[i] = m;
  pts[3 * i] = __fadd_rn(rays_o[3 * i], __fmul_rn(dirs[3 * i], di));
  pts[3 * i + 1] = __fadd_rn(rays_o[3 * i + 1], __fmul_rn(dirs[3 * i + 1], di));
  pts[3 * i + 2] = __fadd_rn(rays_o[3 * i + 2], __fmul_rn(dirs[3 * i + 2], di));
}

}  // namespace

extern "C" int nr_unisurf_ray_setup(const float* rays_o, const float* rays_d, int64_t R, float radius,
                                    float near_bypass, float far_bypass, int32_t n_steps, float* dirs, float* near,
                                    float* far, float* pts, void* stream) {
  NR_CHECK_ARG(R >= 0 && n_steps >= 2, "nr_unisurf_ray_setup: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(rays_o && rays_d && dirs && near && far && pts, "nr_unisurf_ray_setup: null pointer");
  unisurf_ray_setup_kernel<<<(unsigned)nr_cdiv(R, 4), 128, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, R, radius, near_bypass, far_bypass, n_steps, dirs, near, far, pts);
  NR_CHECK_LAUNCH("unisurf_ray_setup_kernel");
  return NR_OK;
}

extern "C" int nr_unisurf_first_crossing(const float* val, const float* rays_o, const float* dirs, const float* near,
                                         const float* far, int64_t R, int32_t n_steps, float logit_tau, float* state,
                                         uint8_t* mask, uint8_t* mask_sign_change, uint8_t* mask_0_free,
                                         float* pts_pred, void* stream) {
  NR_CHECK_ARG(R >= 0 && n_steps >= 2, "nr_unisurf_first_crossing: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(val && rays_o && dirs && near && far && state && mask && mask_sign_change && mask_0_free && pts_pred,
               "nr_unisurf_first_crossing: null pointer");
  unisurf_crossing_kernel<<<(unsigned)nr_cdiv(R, 4), 128, 0, (cudaStream_t)stream>>>(
      val, rays_o, dirs, near, far, R, n_steps, logit_tau, state, mask, mask_sign_change, mask_0_free, pts_pred);
  NR_CHECK_LAUNCH("unisurf_crossing_kernel");
  return NR_OK;
}

extern "C" int nr_unisurf_secant_step(const float* f_mid, float logit_tau, const float* rays_o, const float* dirs,
                                      const uint8_t* mask, int64_t R, float* state, float* pts_pred, void* stream) {
  NR_CHECK_ARG(R >= 0, "nr_unisurf_secant_step: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(f_mid && rays_o && dirs && mask && state && pts_pred, "nr_unisurf_secant_step: null pointer");
  unisurf_secant_kernel<<<(unsigned)nr_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(f_mid, logit_tau, rays_o, dirs, mask,
                                                                                     R, state, pts_pred);
  NR_CHECK_LAUNCH("unisurf_secant_kernel");
  return NR_OK;
}

extern "C" int nr_unisurf_sample(const float* rays_o, const float* dirs, const float* near, const float* far,
                                 const float* state, const uint8_t* mask, const uint8_t* mask_sign_change,
                                 const uint8_t* mask_0_free, int64_t R, float interval, float too_close,
                                 int32_t n_query, int32_t n_free, const float* u_int, const float* u_free,
                                 float* depth_surface, float* surface_pts, float* d_all, float* pts, void* stream) {
  NR_CHECK_ARG(R >= 0 && n_query >= 2 && n_free >= 2, "nr_unisurf_sample: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(rays_o && dirs && near && far && state && mask && mask_sign_change && mask_0_free && depth_surface &&
                   surface_pts && d_all && pts, "nr_unisurf_sample: null pointer");
  NR_CHECK_ARG((u_int != nullptr) == (u_free != nullptr), "nr_unisurf_sample: pass both uniform sets or none");
  const size_t smem = (size_t)4 * 2 * (n_query + n_free) * sizeof(float);
  NR_CHECK_ARG(smem <= 48 * 1024, "nr_unisurf_sample: too many samples per ray");
  unisurf_sample_kernel<<<(unsigned)nr_cdiv(R, 4), 128, smem, (cudaStream_t)stream>>>(
      rays_o, dirs, near, far, state, mask, mask_sign_change, mask_0_free, R, interval, too_close, n_query, n_free, u_int,
      u_free, depth_surface, surface_pts, d_all, pts);
  NR_CHECK_LAUNCH("unisurf_sample_kernel");
  return NR_OK;
}

extern "C" int nr_unisurf_composite(const float* logits, const float* nablas, const float* radiance,
                                    const float* d_all, int64_t R, int32_t M, int32_t white_bkgd, float* rgb,
                                    float* depth, float* acc, float* normals, float* alpha_out, float* weights_out,
                                    void* stream) {
  NR_CHECK_ARG(R >= 0 && M >= 1, "nr_unisurf_composite: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(logits && radiance && d_all && rgb && depth && acc, "nr_unisurf_composite: null pointer");
  NR_CHECK_ARG((nablas != nullptr) == (normals != nullptr), "nr_unisurf_composite: nablas and normals go together");
  if (M <= kUsMaxM) {
    const size_t smem = (size_t)4 * 8 * M * sizeof(float);
    const int vec16 = ((((uintptr_t)logits | (uintptr_t)nablas | (uintptr_t)radiance | (uintptr_t)d_all) & 15) == 0) ? 1 : 0;
    const unsigned grid = (unsigned)nr_cdiv(R, 4);
    if (M <= 96) {           // 64 interval + 32 free-space samples (configs/unisurf.yaml)
      unisurf_composite_staged_kernel<3><<<grid, 128, smem, (cudaStream_t)stream>>>(
          logits, nablas, radiance, d_all, R, M, white_bkgd, rgb, depth, acc, normals, alpha_out, weights_out, vec16);
    } else {
      unisurf_composite_staged_kernel<kUsMaxM / 32><<<grid, 128, smem, (cudaStream_t)stream>>>(
          logits, nablas, radiance, d_all, R, M, white_bkgd, rgb, depth, acc, normals, alpha_out, weights_out, vec16);
    }
    NR_CHECK_LAUNCH("unisurf_composite_staged_kernel");
    return NR_OK;
  }
  unisurf_composite_kernel<<<(unsigned)nr_cdiv(R, 4), 128, 0, (cudaStream_t)stream>>>(
      logits, nablas, radiance, d_all, R, M, white_bkgd, rgb, depth, acc, normals, alpha_out, weights_out);
  NR_CHECK_LAUNCH("unisurf_composite_kernel");
  return NR_OK;
}

extern "C" int nr_sphere_trace_step(const float* val, const float* rays_o, const float* dirs, float far, int64_t R,
                                    float* d, uint8_t* mask, float* pts, void* stream) {
  NR_CHECK_ARG(R >= 0, "nr_sphere_trace_step: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(rays_o && dirs && d && mask && pts, "nr_sphere_trace_step: null pointer");
  sphere_trace_step_kernel<<<(unsigned)nr_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(val, rays_o, dirs, far, R, d, mask, pts);
  NR_CHECK_LAUNCH("sphere_trace_step_kernel");
  return NR_OK;
}
