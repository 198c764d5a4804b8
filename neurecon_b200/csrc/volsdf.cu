// VolSDF: Laplace density, error bound, the error-bounded beta iteration of fine_sample and
// compositing.  One warp per ray; a ray's sample buffers (up to 3584 depths + sdf values) are staged
// in shared memory, scans are lane-segment + warp-shuffle scans, the whole per-ray state machine of
// the reference (boolean-mask gathers + >= 4 host syncs per iteration) runs on the device: a ray
// carries a status word and finished rays skip the work.
//
// Reference semantics: models/frameworks/volsdf.py:16-35 (sdf_to_sigma), :38-74 (error_bound),
// :77-272 (fine_sample), :402-417,436-443 (ray setup / merge), :452-503 (compositing).
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarps = 2;  // rays per block in the sampling kernels (3 x cap floats of smem each)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_excl_scan_add(float v, int lane, float* total) {
  float incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl += t;
  }
  *total = __shfl_sync(kFull, incl, 31);
  const float ex = __shfl_up_sync(kFull, incl, 1);
  return lane == 0 ? 0.0f : ex;
}

// volsdf.py:16-35
__device__ __forceinline__ float laplace_sigma(float sdf, float alpha, float beta) {
  const float e = 0.5f * expf(-fabsf(sdf) / beta);
  return alpha * (sdf >= 0.0f ? e : 1.0f - e);
}

struct Seg { int lo, hi; };
__device__ __forceinline__ Seg lane_segment(int n, int lane) {
  // an ODD segment length: lane l reads arr[l * seg + i], and with an even seg (20 at M = 640) the 32 lanes fall on 8 banks --
  // ncu: shared-memory wavefronts at 95 % of peak, 14 short-scoreboard stalls per issue, the kernel 4 x slower than its math
  const int seg = ((n + 31) >> 5) | 1;
  Seg s;
  s.lo = min(lane * seg, n);
  s.hi = min(s.lo + seg, n);
  return s;
}

// max_i bound_i of error_bound(d, sdf, alpha, beta) (volsdf.py:38-74); optionally stores the bounds
// (clamped to [0, clamp_hi] when clamp_hi > 0).  d, sdf: smem, M entries.  NaN -> inf like :73.
__device__ float error_bound_max(const float* d, const float* sdf, int M, float alpha, float beta, int lane,
                                 float* bounds_out, float clamp_hi) {
  const int n = M - 1;
  const Seg sg = lane_segment(n, lane);
  const float k = alpha / (4.0f * beta);
  // (d[i], sdf[i]) of the next interval are this one's (d[i + 1], sdf[i + 1]): carried in registers, two shared-memory
  // loads per interval instead of four
  float sR = 0.0f, sE = 0.0f;
  float d0 = 0.0f, s0 = 0.0f;
  if (sg.lo < sg.hi) { d0 = d[sg.lo]; s0 = sdf[sg.lo]; }
  for (int i = sg.lo; i < sg.hi; ++i) {
    const float d1 = d[i + 1], s1 = sdf[i + 1];
    const float delta = d1 - d0;
    sR += laplace_sigma(s0, alpha, beta) * delta;
    const float dstar = fmaxf(0.5f * (fabsf(s0) + fabsf(s1) - delta), 0.0f);
    sE += k * (delta * delta) * expf(-dstar / beta);
    d0 = d1; s0 = s1;
  }
  float tR, tE;
  float R = warp_excl_scan_add(sR, lane, &tR);
  float E = warp_excl_scan_add(sE, lane, &tE);
  float mx = -INFINITY;
  if (sg.lo < sg.hi) { d0 = d[sg.lo]; s0 = sdf[sg.lo]; }
  for (int i = sg.lo; i < sg.hi; ++i) {
    const float d1 = d[i + 1], s1 = sdf[i + 1];
    const float delta = d1 - d0;
    const float dstar = fmaxf(0.5f * (fabsf(s0) + fabsf(s1) - delta), 0.0f);
    E += k * (delta * delta) * expf(-dstar / beta);            // inclusive: E(t_{i+1})
    float b = expf(-R) * (expf(E) - 1.0f);                     // R exclusive: R(t_i)
    if (isnan(b)) b = INFINITY;
    mx = fmaxf(mx, b);
    if (bounds_out) bounds_out[i] = clamp_hi > 0.0f ? fminf(fmaxf(b, 0.0f), clamp_hi) : b;
    R += laplace_sigma(s0, alpha, beta) * delta;
    d0 = d1; s0 = s1;
  }
  mx = warp_max(mx);
  __syncwarp();
  return mx;
}

// In-place inclusive cumsum of arr[0..n) (smem) by one warp.
__device__ void warp_cumsum(float* arr, int n, int lane) {
  const Seg sg = lane_segment(n, lane);
  float s = 0.0f;
  for (int i = sg.lo; i < sg.hi; ++i) s += arr[i];
  float tot;
  float run = warp_excl_scan_add(s, lane, &tot);
  for (int i = sg.lo; i < sg.hi; ++i) { run += arr[i]; arr[i] = run; }
  __syncwarp();
}

__device__ __forceinline__ int lower_bound(const float* cdf, int M, float u) {
  int lo = 0, hi = M;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int M, float u, float eps) {
  const int ind = lower_bound(cdf, M, u);
  const int below = max(ind - 1, 0), above = min(ind, M - 1);
  const float cb = cdf[below], ca = cdf[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < eps) denom = 1.0f;
  const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  return __fadd_rn(bins[below], __fmul_rn(t, __fsub_rn(bins[above], bins[below])));
}

// opacity_invert_cdf_sample (volsdf.py:102-116): R_t[i] = sum_{j<i} sigma_j delta_j (i = 0..M-2),
// cdf = [0, 1 - exp(-R_t)] (M entries), then sample_cdf(N).  So cdf[k] = 1 - exp(-sum_{j<=k-2} sigma_j delta_j).
__device__ void opacity_sample(const float* d, const float* sdf, int M, float alpha, float beta, float* cdf, int N,
                               const float* u, float* out, int lane) {
  for (int k = lane; k < M; k += 32)
    cdf[k] = k >= 2 ? laplace_sigma(sdf[k - 2], alpha, beta) * (d[k - 1] - d[k - 2]) : 0.0f;
  __syncwarp();
  warp_cumsum(cdf, M, lane);
  for (int k = lane; k < M; k += 32) cdf[k] = 1.0f - expf(-cdf[k]);
  __syncwarp();
  for (int j = lane; j < N; j += 32) {
    const float uu = u ? u[j] : nr_linspace01(j, N);
    out[j] = invert_cdf(cdf, d, M, uu, 1e-5f);
  }
  __syncwarp();
}

__global__ void volsdf_error_bound_kernel(const float* __restrict__ d_vals, const float* __restrict__ sdf, int64_t R,
                                          int M, const float* __restrict__ alpha, int alpha_stride,
                                          const float* __restrict__ beta, int beta_stride,
                                          float* __restrict__ bounds, float* __restrict__ bound_max) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarps + warp;
  if (ray >= R) return;
  float* sd = smem + (size_t)warp * 3 * M;
  float* ss = sd + M;
  float* sb = ss + M;
  for (int i = lane; i < M; i += 32) { sd[i] = d_vals[ray * M + i]; ss[i] = sdf[ray * M + i]; }
  __syncwarp();
  const float mx = error_bound_max(sd, ss, M, alpha[ray * alpha_stride], beta[ray * beta_stride], lane, sb, -1.0f);
  if (bounds) for (int i = lane; i < M - 1; i += 32) bounds[ray * (int64_t)(M - 1) + i] = sb[i];
  if (bound_max && lane == 0) bound_max[ray] = mx;
}

// ---------------------------------------------------------------------------------------------
// Ray prologue (volsdf.py:169-172,402-427): dirs, far (constant or exact sphere exit for NeRF++),
// the dense initial depths d_init = linspace(near, far, n_init) and their points.
// ---------------------------------------------------------------------------------------------
__global__ void volsdf_ray_setup_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t R,
                                        float near, float far, float sphere_radius, int n_init,
                                        float* __restrict__ dirs, float* __restrict__ fars, int* __restrict__ miss_count,
                                        float* __restrict__ d_buf, int cap, float* __restrict__ pts_new) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)4 + warp;
  if (ray >= R) return;
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
  const float nrm = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);
  dx /= nrm; dy /= nrm; dz /= nrm;
  float fr = far;
  if (sphere_radius > 0.0f) {  // get_sphere_intersection (rend_util.py:188-210): far end, clamped at 0
    const float o2 = ox * ox + oy * oy + oz * oz, od = ox * dx + oy * dy + oz * dz;
    const float under = od * od + sphere_radius * sphere_radius - o2;
    if (under > 0.0f) fr = fmaxf(sqrtf(under) - od, 0.0f);
    else { fr = 0.0f; if (lane == 0) atomicAdd(miss_count, 1); }
  }
  if (lane == 0) { dirs[3 * ray] = dx; dirs[3 * ray + 1] = dy; dirs[3 * ray + 2] = dz; fars[ray] = fr; }
  for (int i = lane; i < n_init; i += 32) {
    const float t = nr_linspace01(i, n_init);
    const float d = __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(fr, t));
    d_buf[ray * (int64_t)cap + i] = d;
    float* p = pts_new + (ray * (int64_t)n_init + i) * 3;
    p[0] = ox + d * dx; p[1] = oy + d * dy; p[2] = oz + d * dz;
  }
}

// sdf = min(sdf, r - |x|)  (VolSDF.forward_surface / forward_surface_with_nablas, volsdf.py:310-325)
__global__ void sphere_min_kernel(const float* __restrict__ pts, float* __restrict__ sdf, int64_t n, float radius) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
  sdf[i] = fminf(sdf[i], radius - sqrtf(x * x + y * y + z * z));
}

// ---------------------------------------------------------------------------------------------
// One iteration of fine_sample (volsdf.py:129-270).  it = 0: first bound check on the initial samples;
// it >= 1: merge the n_up new samples, re-check, bisect beta.  status: 0 active, 1 done.
// ---------------------------------------------------------------------------------------------
struct FineArgs {
  const float* rays_o; const float* dirs; const float* fars; int64_t R;
  float* d_buf; float* sdf_buf; int cap; int m_cur;        // sorted state, m_cur valid entries (before merge)
  const float* sdf_new; int n_new;                          // it = 0: sdf of the n_new = m0 initial samples
  const float* d_new_in;                                    // it >= 1: the n_up depths proposed last iteration
  const float* alpha_net; const float* beta_net;            // device scalars
  float eps; int it; int max_iter; int max_bisection; int n_up; int n_final;
  const float* u_final;                                     // [R, n_final] uniforms or null (det)
  float* beta; int* status; float* iter_usage; float* beta_map; float* d_fine;
  float* d_new_out; float* pts_new;                         // [R, n_up], [R, n_up, 3]
  int m0;                                                   // number of initial samples (for beta_0)
  int ld;                                                   // floats per shared-memory array of a ray: this launch's M, not cap
};

__global__ void volsdf_fine_iter_kernel(const FineArgs g) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarps + warp;
  if (ray >= g.R) return;
  // three arrays of THIS iteration's sample count per ray: sized by the buffer capacity (128 + 5 x 512 samples = 32 KB per ray)
  // the kernel ran 4 warps per SM -- ncu: 4 - 6 % of the warps active, 0.39 instructions per cycle
  float* sd = smem + (size_t)warp * 3 * g.ld;
  float* ss = sd + g.ld;
  float* sc = ss + g.ld;  // scratch: bounds, then cdf
  float* dn_out = g.d_new_out + ray * (int64_t)g.n_up;
  float* pn_out = g.pts_new + ray * (int64_t)g.n_up * 3;
  const float ox = g.rays_o[3 * ray], oy = g.rays_o[3 * ray + 1], oz = g.rays_o[3 * ray + 2];
  const float dx = g.dirs[3 * ray], dy = g.dirs[3 * ray + 1], dz = g.dirs[3 * ray + 2];
  auto emit_points = [&](bool valid) {
    for (int j = lane; j < g.n_up; j += 32) {
      const float d = valid ? dn_out[j] : 0.0f;
      if (!valid) dn_out[j] = 0.0f;
      pn_out[3 * j] = ox + d * dx; pn_out[3 * j + 1] = oy + d * dy; pn_out[3 * j + 2] = oz + d * dz;
    }
  };
  if (g.it > 0 && g.status[ray] != 0) {  // finished earlier: keep the (ignored) proposals well defined
    if (g.it < g.max_iter) emit_points(false);
    return;
  }
  const float alpha_net = *g.alpha_net, beta_net = *g.beta_net;
  float* drow = g.d_buf + ray * (int64_t)g.cap;
  float* srow = g.sdf_buf + ray * (int64_t)g.cap;
  int M;
  if (g.it == 0) {
    M = g.n_new;
    for (int i = lane; i < M; i += 32) { sd[i] = drow[i]; ss[i] = g.sdf_new[ray * (int64_t)g.n_new + i]; }
    __syncwarp();
    for (int i = lane; i < M; i += 32) srow[i] = ss[i];
  } else {
    // merge: both lists are sorted (the proposals come from a deterministic inverse CDF); stable, old first
    M = g.m_cur + g.n_new;
    float* od = sc;              // old d
    float* nd = sc + g.m_cur;    // new d  (m_cur + n_new <= cap)
    for (int i = lane; i < g.m_cur; i += 32) od[i] = drow[i];
    for (int j = lane; j < g.n_new; j += 32) nd[j] = g.d_new_in[ray * (int64_t)g.n_new + j];
    __syncwarp();
    for (int i = lane; i < g.m_cur; i += 32) {
      const float v = od[i];
      int lo = 0, hi = g.n_new;  // count(new < v)
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (nd[mid] < v) lo = mid + 1; else hi = mid; }
      sd[i + lo] = v; ss[i + lo] = srow[i];
    }
    for (int j = lane; j < g.n_new; j += 32) {
      const float v = nd[j];
      int lo = 0, hi = g.m_cur;  // count(old <= v)
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (od[mid] <= v) lo = mid + 1; else hi = mid; }
      sd[j + lo] = v; ss[j + lo] = g.sdf_new[ray * (int64_t)g.n_new + j];
    }
    __syncwarp();
    for (int i = lane; i < M; i += 32) { drow[i] = sd[i]; srow[i] = ss[i]; }
  }
  __syncwarp();
  const float* uf = g.u_final ? g.u_final + ray * (int64_t)g.n_final : nullptr;
  float* dfine = g.d_fine + ray * (int64_t)g.n_final;

  // bound with the network's own beta (volsdf.py:138-141 / :210-212)
  const float net_max = error_bound_max(sd, ss, M, alpha_net, beta_net, lane, nullptr, -1.0f);
  if (!(net_max > g.eps)) {
    opacity_sample(sd, ss, M, alpha_net, beta_net, sc, g.n_final, uf, dfine, lane);
    if (lane == 0) { g.status[ray] = 1; g.iter_usage[ray] = (float)g.it; g.beta_map[ray] = beta_net; }
    if (g.it < g.max_iter) emit_points(false);
    return;
  }
  float beta;
  if (g.it == 0) {
    const float far = g.fars[ray];
    beta = sqrtf((far * far) / (4.0f * (float)(g.m0 - 1) * logf(1.0f + g.eps)));   // volsdf.py:129
    error_bound_max(sd, ss, M, 1.0f / beta, beta, lane, sc, -1.0f);                  // bounds (not clamped, :146)
  } else {
    // bisection for beta+ with B(beta+) == eps (volsdf.py:228-252)
    float right = g.beta[ray], left = beta_net;
    for (int b = 0; b < g.max_bisection; ++b) {
      const float mid = 0.5f * (left + right);
      const float mx = error_bound_max(sd, ss, M, 1.0f / mid, mid, lane, nullptr, -1.0f);
      if (mx <= g.eps) right = mid;
      if (mx > g.eps) left = mid;
    }
    beta = right;
    if (g.it < g.max_iter) error_bound_max(sd, ss, M, 1.0f / beta, beta, lane, sc, 1e5f);  // clamped (:252)
  }
  if (lane == 0) g.beta[ray] = beta;
  if (g.it < g.max_iter) {
    // sample_pdf(d, bounds, n_up + 2, det=True)[1:-1] (volsdf.py:173; rend_util.py:255-292), cdf built in place:
    // bounds sc[0..n) -> pdf shifted right by one -> cumsum -> cdf sc[0..M)
    const int n = M - 1;
    float part = 0.0f;
    for (int i = lane; i < n; i += 32) part += sc[i] + 1e-5f;
    const float tot = warp_sum(part);
    __syncwarp();
    // shift right by one while normalising: process from the back in 32-wide groups
    for (int base = ((n + 31) / 32) * 32; base > 0; base -= 32) {
      const int i = base - 32 + lane;            // source index
      float v = 0.0f;
      if (i < n) v = __fdiv_rn(sc[i] + 1e-5f, tot);
      __syncwarp();
      if (i < n) sc[i + 1] = v;
      __syncwarp();
    }
    if (lane == 0) sc[0] = 0.0f;
    __syncwarp();
    warp_cumsum(sc + 1, n, lane);
    const int N = g.n_up + 2;
    for (int j = lane; j < g.n_up; j += 32) dn_out[j] = invert_cdf(sc, sd, M, nr_linspace01(j + 1, N), 1e-5f);
    __syncwarp();
    emit_points(true);
  } else {
    // not converged after max_iter: sample with the last beta+ (volsdf.py:264-270)
    opacity_sample(sd, ss, M, 1.0f / beta, beta, sc, g.n_final, uf, dfine, lane);
    if (lane == 0) { g.status[ray] = 1; g.iter_usage[ray] = -1.0f; g.beta_map[ray] = beta; }
  }
}

// ---------------------------------------------------------------------------------------------
// d_all = sort(cat(d_coarse, d_fine)) and its points (volsdf.py:436-444)
// ---------------------------------------------------------------------------------------------
__global__ void volsdf_merge_kernel(const float* __restrict__ rays_o, const float* __restrict__ dirs,
                                    const float* __restrict__ fars, int64_t R, float near, int n_coarse,
                                    const float* __restrict__ d_fine, int n_fine, float* __restrict__ d_all,
                                    float* __restrict__ pts) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)4 + warp;
  if (ray >= R) return;
  const int M = n_coarse + n_fine;
  float* sc = smem + (size_t)warp * 2 * M;  // coarse | fine
  float* sf = sc + n_coarse;
  float* so = sc + M;                        // merged
  const float fr = fars[ray];
  for (int i = lane; i < n_coarse; i += 32) {
    const float t = nr_linspace01(i, n_coarse);
    sc[i] = __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(fr, t));
  }
  for (int j = lane; j < n_fine; j += 32) sf[j] = d_fine[ray * (int64_t)n_fine + j];
  __syncwarp();
  for (int i = lane; i < n_coarse; i += 32) {
    const float v = sc[i];
    int c = 0;
    for (int k = 0; k < n_fine; ++k) c += (sf[k] < v);
    so[i + c] = v;
  }
  for (int j = lane; j < n_fine; j += 32) {
    const float v = sf[j];
    int lo = 0, hi = n_coarse;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sc[mid] <= v) lo = mid + 1; else hi = mid; }
    int c = lo;
    for (int k = 0; k < n_fine; ++k) c += (sf[k] < v) || (sf[k] == v && k < j);
    so[c] = v;
  }
  __syncwarp();
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  const float dx = dirs[3 * ray], dy = dirs[3 * ray + 1], dz = dirs[3 * ray + 2];
  for (int i = lane; i < M; i += 32) {
    const float d = so[i];
    d_all[ray * (int64_t)M + i] = d;
    float* p = pts + (ray * (int64_t)M + i) * 3;
    p[0] = __fadd_rn(ox, __fmul_rn(dx, d)); p[1] = __fadd_rn(oy, __fmul_rn(dy, d)); p[2] = __fadd_rn(oz, __fmul_rn(dz, d));
  }
}

// ---------------------------------------------------------------------------------------------
// Compositing (volsdf.py:452-503): sigma = Laplace(sdf) for the M_in inside samples, optional M_out
// NeRF++ samples appended (raw sigma_out, radiance_out, d_out); p = exp(-relu(sigma delta));
// tau = (1 - p + 1e-10) * exclusive cumprod(p); weights sit on the LEFT end of each interval.
// ---------------------------------------------------------------------------------------------
__global__ void volsdf_composite_kernel(const float* __restrict__ sdf, const float* __restrict__ nablas,
                                        const float* __restrict__ radiance, const float* __restrict__ d_in,
                                        const float* __restrict__ alpha_dev, const float* __restrict__ beta_dev,
                                        int64_t R, int M_in, const float* __restrict__ sigma_out,
                                        const float* __restrict__ radiance_out, const float* __restrict__ d_out,
                                        int M_out, int white_bkgd, float* __restrict__ rgb, float* __restrict__ depth,
                                        float* __restrict__ acc, float* __restrict__ normals,
                                        float* __restrict__ sigma_all, float* __restrict__ p_out,
                                        float* __restrict__ tau_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)4 + warp;
  if (ray >= R) return;
  const float alpha = *alpha_dev, beta = *beta_dev;
  const int M = M_in + M_out;
  auto dval = [&](int i) { return i < M_in ? d_in[ray * (int64_t)M_in + i] : d_out[ray * (int64_t)M_out + (i - M_in)]; };
  auto sig = [&](int i) {
    return i < M_in ? laplace_sigma(sdf[ray * (int64_t)M_in + i], alpha, beta) : sigma_out[ray * (int64_t)M_out + (i - M_in)];
  };
  float carry = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  for (int base = 0; base < M; base += 32) {
    const int i = base + lane;
    const bool in_m = i < M, ok = i < M - 1;
    float s = 0.0f, di = 0.0f;
    if (in_m) { s = sig(i); di = dval(i); if (sigma_all) sigma_all[ray * (int64_t)M + i] = s; }
    float p = 1.0f;
    if (ok) p = expf(-fmaxf(s * (dval(i + 1) - di), 0.0f));
    float incl = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl *= t;
    }
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    const float tau = (1.0f - p + 1e-10f) * (carry * excl);
    carry *= __shfl_sync(kFull, incl, 31);
    if (ok) {
      if (p_out) p_out[ray * (int64_t)(M - 1) + i] = p;
      if (tau_out) tau_out[ray * (int64_t)(M - 1) + i] = tau;
      const float* c = i < M_in ? radiance + (ray * (int64_t)M_in + i) * 3 : radiance_out + (ray * (int64_t)M_out + (i - M_in)) * 3;
      ar += tau * c[0]; ag += tau * c[1]; ab += tau * c[2];
      ad += tau * di;
      aw += tau;
      if (nablas && i < M_in) {
        const float* nb = nablas + (ray * (int64_t)M_in + i) * 3;
        const float x = nb[0], y = nb[1], z = nb[2];
        const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-12f);
        nx += tau * x * inv; ny += tau * y * inv; nz += tau * z * inv;
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

// ---------------------------------------------------------------------------------------------
// Staged variant (M_in <= kVsMaxIn, M_out <= kVsMaxOut): the block's four rays are contiguous in every input, so whole
// blocks move into shared memory with 16-byte cp.async copies, all in flight at once; each lane then owns ceil(M/32)
// CONSECUTIVE samples (one exponential per sample, a serial product inside the lane) and the warp does ONE multiplicative
// scan per ray instead of one per 32 samples.  The first structure above was bound by instruction issue and load latency
// (33 % of the copy bandwidth).  Fast exponential (2 ulp): compositing is compared at 1e-4, sample positions are not
// computed here.
// ---------------------------------------------------------------------------------------------
constexpr int kVsMaxIn = 256, kVsMaxOut = 64;
constexpr int kVsSegMax = (kVsMaxIn + kVsMaxOut + 31) / 32;
template <int kVsSeg>   // samples per lane: ceil(M / 32) <= kVsSeg (every unrolled iteration costs issue slots, used or not)
__global__ void volsdf_composite_staged_kernel(const float* __restrict__ sdf, const float* __restrict__ nablas,
                                               const float* __restrict__ radiance, const float* __restrict__ d_in,
                                               const float* __restrict__ alpha_dev, const float* __restrict__ beta_dev,
                                               int64_t R, int M_in, const float* __restrict__ sigma_out,
                                               const float* __restrict__ radiance_out, const float* __restrict__ d_out,
                                               int M_out, int white_bkgd, float* __restrict__ rgb,
                                               float* __restrict__ depth, float* __restrict__ acc,
                                               float* __restrict__ normals, float* __restrict__ sigma_all,
                                               float* __restrict__ p_out, float* __restrict__ tau_out, int vec16) {
  extern __shared__ __align__(16) float vstage[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray0 = blockIdx.x * (int64_t)4, ray = ray0 + warp;
  const int nrays = (int)((R - ray0) < 4 ? (R - ray0) : 4);
  const int M = M_in + M_out;
  float* s_sd = vstage;                    // [4][M_in]
  float* s_d = s_sd + 4 * M_in;            // [4][M_in]
  float* s_rad = s_d + 4 * M_in;           // [4][3 M_in]
  float* s_nb = s_rad + 12 * M_in;         // [4][3 M_in]
  float* s_so = s_nb + 12 * M_in;          // [4][M_out]
  float* s_do = s_so + 4 * M_out;          // [4][M_out]
  float* s_ro = s_do + 4 * M_out;          // [4][3 M_out]
  auto stage_in = [&](float* dst, const float* src, int row) {
    const float* g = src + ray0 * (int64_t)row;
    const int n = nrays * row;
    const int n4 = vec16 ? n >> 2 : 0;
    for (int i = threadIdx.x; i < n4; i += blockDim.x)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 4 * i)), "l"(g + 4 * i) : "memory");
    for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + i)), "l"(g + i) : "memory");
  };
  if (vec16 && nrays == 4) {
    // full block, aligned inputs: one thread issues the bulk copies (TMA unit); spread over the block as 16-byte
    // cp.async the staging alone was ~5 % of the instructions of an issue-bound kernel
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
      umma::mbar_init(&bar, 1);
      umma::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t b1 = 16u * M_in, b3 = 48u * M_in, o1 = 16u * M_out, o3 = 48u * M_out;
      umma::mbar_arrive_expect_tx(&bar, 2 * b1 + b3 + (nablas ? b3 : 0u) + (M_out > 0 ? 2 * o1 + o3 : 0u));
      umma::bulk_g2s(s_sd, sdf + ray0 * (int64_t)M_in, b1, &bar);
      umma::bulk_g2s(s_d, d_in + ray0 * (int64_t)M_in, b1, &bar);
      umma::bulk_g2s(s_rad, radiance + ray0 * (int64_t)(3 * M_in), b3, &bar);
      if (nablas) umma::bulk_g2s(s_nb, nablas + ray0 * (int64_t)(3 * M_in), b3, &bar);
      if (M_out > 0) {
        umma::bulk_g2s(s_so, sigma_out + ray0 * (int64_t)M_out, o1, &bar);
        umma::bulk_g2s(s_do, d_out + ray0 * (int64_t)M_out, o1, &bar);
        umma::bulk_g2s(s_ro, radiance_out + ray0 * (int64_t)(3 * M_out), o3, &bar);
      }
    }
    umma::mbar_wait(&bar, 0);
  } else {
    stage_in(s_sd, sdf, M_in);
    stage_in(s_d, d_in, M_in);
    stage_in(s_rad, radiance, 3 * M_in);
    if (nablas) stage_in(s_nb, nablas, 3 * M_in);
    if (M_out > 0) {
      stage_in(s_so, sigma_out, M_out);
      stage_in(s_do, d_out, M_out);
      stage_in(s_ro, radiance_out, 3 * M_out);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
  if (ray >= R) return;
  const float* sd = s_sd + warp * M_in;
  const float* dd = s_d + warp * M_in;
  const float* rad = s_rad + warp * 3 * M_in;
  const float* nb = s_nb + warp * 3 * M_in;
  const float* so = s_so + warp * M_out;
  const float* dO = s_do + warp * M_out;
  const float* ro = s_ro + warp * 3 * M_out;
  const float alpha = *alpha_dev, inv_beta = 1.0f / *beta_dev;
  auto dval = [&](int i) { return i < M_in ? dd[i] : dO[i - M_in]; };
  auto sig = [&](int i) {
    if (i >= M_in) return so[i - M_in];
    const float v = sd[i];
    const float e = 0.5f * __expf(-fabsf(v) * inv_beta);
    return alpha * (v >= 0.0f ? e : 1.0f - e);
  };
  const int seg = (M + 31) >> 5;
  const int i0 = lane * seg;
  float pk[kVsSeg], tr[kVsSeg];
  float prod = 1.0f;
  float d_cur = dval(i0 < M ? i0 : M - 1);
#pragma unroll
  for (int k = 0; k < kVsSeg; ++k) {
    const int i = i0 + k;
    float p = 1.0f;
    if (k < seg && i < M) {
      const float sg = sig(i);
      if (sigma_all) sigma_all[ray * (int64_t)M + i] = sg;
      if (i < M - 1) {
        const float d_next = dval(i + 1);
        p = __expf(-fmaxf(sg * (d_next - d_cur), 0.0f));
        d_cur = d_next;
      }
    }
    pk[k] = p;
    tr[k] = prod;
    prod *= p;
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
  for (int k = 0; k < kVsSeg; ++k) {
    const int i = i0 + k;
    if (k < seg && i < M - 1) {
      const float tau = (1.0f - pk[k] + 1e-10f) * (before * tr[k]);
      if (p_out) p_out[ray * (int64_t)(M - 1) + i] = pk[k];
      if (tau_out) tau_out[ray * (int64_t)(M - 1) + i] = tau;
      const float* c = i < M_in ? rad + 3 * i : ro + 3 * (i - M_in);
      ar += tau * c[0]; ag += tau * c[1]; ab += tau * c[2];
      ad += tau * dval(i);
      aw += tau;
      if (nablas && i < M_in) {
        const float x = nb[3 * i], y = nb[3 * i + 1], z = nb[3 * i + 2];
        const float inv = rsqrtf(fmaxf(x * x + y * y + z * z, 1e-24f));   // = 1 / max(|v|, 1e-12)
        nx += tau * x * inv; ny += tau * y * inv; nz += tau * z * inv;
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

int set_smem(const void* fn, size_t bytes, const char* name) {
  if (bytes > 220 * 1024) { nr_set_error("%s: %zu bytes of shared memory needed (sample count too large)", name, bytes); return NR_ERR_INVALID; }
  if (bytes > 48 * 1024) NR_CHECK_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return NR_OK;
}

}  // namespace

extern "C" int nr_volsdf_error_bound(const float* d_vals, const float* sdf, int64_t R, int32_t M, const float* alpha,
                                     int32_t alpha_stride, const float* beta, int32_t beta_stride, float* bounds,
                                     float* bound_max, void* stream) {
  NR_CHECK_ARG(R >= 0 && M >= 2, "nr_volsdf_error_bound: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(d_vals && sdf && alpha && beta && (bounds || bound_max), "nr_volsdf_error_bound: null pointer");
  const size_t smem = (size_t)kWarps * 3 * M * sizeof(float);
  int rc = set_smem((const void*)volsdf_error_bound_kernel, smem, "nr_volsdf_error_bound");
  if (rc) return rc;
  volsdf_error_bound_kernel<<<(unsigned)nr_cdiv(R, kWarps), kWarps * 32, smem, (cudaStream_t)stream>>>(
      d_vals, sdf, R, M, alpha, alpha_stride, beta, beta_stride, bounds, bound_max);
  NR_CHECK_LAUNCH("volsdf_error_bound_kernel");
  return NR_OK;
}

extern "C" int nr_volsdf_ray_setup(const float* rays_o, const float* rays_d, int64_t R, float near, float far,
                                   float sphere_radius, int32_t n_init, float* dirs, float* fars, int32_t* miss_count,
                                   float* d_buf, int32_t cap, float* pts_new, void* stream) {
  NR_CHECK_ARG(R >= 0 && n_init >= 2 && cap >= n_init, "nr_volsdf_ray_setup: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(rays_o && rays_d && dirs && fars && d_buf && pts_new && miss_count, "nr_volsdf_ray_setup: null pointer");
  volsdf_ray_setup_kernel<<<(unsigned)nr_cdiv(R, 4), 128, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, R, near, far, sphere_radius, n_init, dirs, fars, miss_count, d_buf, cap, pts_new);
  NR_CHECK_LAUNCH("volsdf_ray_setup_kernel");
  return NR_OK;
}

extern "C" int nr_sphere_min(const float* pts, float* sdf, int64_t n, float radius, void* stream) {
  NR_CHECK_ARG(n >= 0, "nr_sphere_min: n < 0");
  if (n == 0) return NR_OK;
  NR_CHECK_ARG(pts && sdf, "nr_sphere_min: null pointer");
  sphere_min_kernel<<<(unsigned)nr_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(pts, sdf, n, radius);
  NR_CHECK_LAUNCH("sphere_min_kernel");
  return NR_OK;
}

extern "C" int nr_volsdf_fine_iter(const float* rays_o, const float* dirs, const float* fars, int64_t R, float* d_buf,
                                   float* sdf_buf, int32_t cap, int32_t m_cur, const float* sdf_new, int32_t n_new,
                                   const float* d_new_in, const float* alpha_net, const float* beta_net, float eps,
                                   int32_t it, int32_t max_iter, int32_t max_bisection, int32_t n_up, int32_t n_final,
                                   const float* u_final, int32_t m0, float* beta, int32_t* status, float* iter_usage,
                                   float* beta_map, float* d_fine, float* d_new_out, float* pts_new, void* stream) {
  NR_CHECK_ARG(R >= 0 && it >= 0 && it <= max_iter, "nr_volsdf_fine_iter: bad iteration");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(rays_o && dirs && fars && d_buf && sdf_buf && sdf_new && alpha_net && beta_net && beta && status &&
                   iter_usage && beta_map && d_fine && d_new_out && pts_new,
               "nr_volsdf_fine_iter: null pointer");
  NR_CHECK_ARG((it == 0 ? n_new : m_cur + n_new) <= cap && n_new >= 1, "nr_volsdf_fine_iter: m_cur=%d n_new=%d cap=%d",
               m_cur, n_new, cap);
  NR_CHECK_ARG(it == 0 || d_new_in, "nr_volsdf_fine_iter: d_new_in required for it >= 1");
  NR_CHECK_ARG(n_up >= 1 && n_final >= 1 && m0 >= 2, "nr_volsdf_fine_iter: bad sample counts");
  const int ld = (((it == 0 ? n_new : m_cur + n_new) + 31) / 32) * 32;
  const size_t smem = (size_t)kWarps * 3 * ld * sizeof(float);
  int rc = set_smem((const void*)volsdf_fine_iter_kernel, smem, "nr_volsdf_fine_iter");
  if (rc) return rc;
  FineArgs g{rays_o, dirs, fars, R, d_buf, sdf_buf, cap, m_cur, sdf_new, n_new, d_new_in, alpha_net, beta_net, eps, it,
             max_iter, max_bisection, n_up, n_final, u_final, beta, status, iter_usage, beta_map, d_fine, d_new_out,
             pts_new, m0, ld};
  volsdf_fine_iter_kernel<<<(unsigned)nr_cdiv(R, kWarps), kWarps * 32, smem, (cudaStream_t)stream>>>(g);
  NR_CHECK_LAUNCH("volsdf_fine_iter_kernel");
  return NR_OK;
}

extern "C" int nr_volsdf_merge(const float* rays_o, const float* dirs, const float* fars, int64_t R, float near,
                               int32_t n_coarse, const float* d_fine, int32_t n_fine, float* d_all, float* pts,
                               void* stream) {
  NR_CHECK_ARG(R >= 0 && n_coarse >= 2 && n_fine >= 1, "nr_volsdf_merge: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(rays_o && dirs && fars && d_fine && d_all && pts, "nr_volsdf_merge: null pointer");
  const size_t smem = (size_t)4 * 2 * (n_coarse + n_fine) * sizeof(float);
  int rc = set_smem((const void*)volsdf_merge_kernel, smem, "nr_volsdf_merge");
  if (rc) return rc;
  volsdf_merge_kernel<<<(unsigned)nr_cdiv(R, 4), 128, smem, (cudaStream_t)stream>>>(rays_o, dirs, fars, R, near,
                                                                                     n_coarse, d_fine, n_fine, d_all, pts);
  NR_CHECK_LAUNCH("volsdf_merge_kernel");
  return NR_OK;
}

extern "C" int nr_volsdf_composite(const float* sdf, const float* nablas, const float* radiance, const float* d_in,
                                   const float* alpha_dev, const float* beta_dev, int64_t R, int32_t M_in,
                                   const float* sigma_out, const float* radiance_out, const float* d_out, int32_t M_out,
                                   int32_t white_bkgd, float* rgb, float* depth, float* acc, float* normals,
                                   float* sigma_all, float* p_out, float* tau_out, void* stream) {
  NR_CHECK_ARG(R >= 0 && M_in >= 2 && M_out >= 0, "nr_volsdf_composite: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(sdf && radiance && d_in && alpha_dev && beta_dev && rgb && depth && acc, "nr_volsdf_composite: null pointer");
  NR_CHECK_ARG((nablas != nullptr) == (normals != nullptr), "nr_volsdf_composite: nablas and normals go together");
  NR_CHECK_ARG(M_out == 0 || (sigma_out && radiance_out && d_out), "nr_volsdf_composite: outside samples missing");
  if (M_in <= kVsMaxIn && M_out <= kVsMaxOut) {
    const size_t smem = (size_t)4 * (8 * M_in + 5 * M_out) * sizeof(float);
    uintptr_t al = (uintptr_t)sdf | (uintptr_t)radiance | (uintptr_t)d_in | (uintptr_t)nablas;
    if (M_out > 0) al |= (uintptr_t)sigma_out | (uintptr_t)radiance_out | (uintptr_t)d_out;
    const int vec16 = (al & 15) == 0 ? 1 : 0;
    const int seg = (M_in + M_out + 31) / 32;
#define NR_VS_LAUNCH(SEG)                                                                                              \
  do {                                                                                                                 \
    int rc = set_smem((const void*)volsdf_composite_staged_kernel<SEG>, smem, "nr_volsdf_composite");                   \
    if (rc) return rc;                                                                                                 \
    volsdf_composite_staged_kernel<SEG><<<(unsigned)nr_cdiv(R, 4), 128, smem, (cudaStream_t)stream>>>(                  \
        sdf, nablas, radiance, d_in, alpha_dev, beta_dev, R, M_in, sigma_out, radiance_out, d_out, M_out, white_bkgd,  \
        rgb, depth, acc, normals, sigma_all, p_out, tau_out, vec16);                                                   \
  } while (0)
    if (seg <= 6) NR_VS_LAUNCH(6);            // 128 + 64 samples (configs/volsdf.yaml)
    else if (seg <= 7) NR_VS_LAUNCH(7);       // + 32 NeRF++ samples
    else NR_VS_LAUNCH(kVsSegMax);
#undef NR_VS_LAUNCH
    NR_CHECK_LAUNCH("volsdf_composite_staged_kernel");
    return NR_OK;
  }
  volsdf_composite_kernel<<<(unsigned)nr_cdiv(R, 4), 128, 0, (cudaStream_t)stream>>>(
      sdf, nablas, radiance, d_in, alpha_dev, beta_dev, R, M_in, sigma_out, radiance_out, d_out, M_out, white_bkgd, rgb,
      depth, acc, normals, sigma_all, p_out, tau_out);
  NR_CHECK_LAUNCH("volsdf_composite_kernel");
  return NR_OK;
}
