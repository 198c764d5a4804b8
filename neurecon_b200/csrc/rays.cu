// Ray geometry, inverse-CDF sampling, the NeuS coarse-to-fine up-sampler and the NeuS
// alpha / transmittance / compositing pass.  All kernels are warp-per-ray: a ray's sample
// buffers (<= a few KB) are staged once in shared memory, scans are warp-shuffle scans over
// lane-contiguous segments, the inverse-CDF search is a per-lane binary search in smem.
// These kernels are HBM-bound by design (SURVEY.md section 8d gives the bytes per ray).
//
// Reference semantics: utils/rend_util.py:167-185,255-327 and
// models/frameworks/neus.py:21-70,184-210,249-288,296,346-381.
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// In-place exclusive scan (sum or product) of arr[0..n) in shared memory by one warp.
// Each lane owns a contiguous segment; returns the total.  `excl_out` may alias arr.
template <bool kMul>
__device__ __forceinline__ float warp_scan_excl(const float* arr, float* excl_out, float* incl_out, int n, int lane) {
  const int seg = (n + 31) >> 5;
  const int lo = min(lane * seg, n), hi = min(lo + seg, n);
  float tot = kMul ? 1.0f : 0.0f;
  for (int i = lo; i < hi; ++i) tot = kMul ? tot * arr[i] : tot + arr[i];
  float incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl = kMul ? incl * t : incl + t;
  }
  float run = __shfl_up_sync(kFull, incl, 1);
  if (lane == 0) run = kMul ? 1.0f : 0.0f;
  const float total = __shfl_sync(kFull, incl, 31);
  __syncwarp();
  for (int i = lo; i < hi; ++i) {
    const float v = arr[i];
    if (excl_out) excl_out[i] = run;
    run = kMul ? run * v : run + v;
    if (incl_out) incl_out[i] = run;
  }
  __syncwarp();
  return total;
}

// torch.searchsorted(cdf, u, right=False): first index i in [0, M] with cdf[i] >= u.
__device__ __forceinline__ int lower_bound(const float* cdf, int M, float u) {
  int lo = 0, hi = M;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cdf[mid] < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Tail of sample_pdf / sample_cdf (rend_util.py:275-292), un-fused fp32 ops so that the
// result is bit-identical to the separate torch elementwise kernels given the same cdf and u.
__device__ __forceinline__ float invert_cdf(const float* cdf, const float* bins, int M, float u, float eps,
                                            int* below_out, int* above_out) {
  const int ind = lower_bound(cdf, M, u);
  const int below = max(ind - 1, 0), above = min(ind, M - 1);
  const float cb = cdf[below], ca = cdf[above];
  float denom = __fsub_rn(ca, cb);
  if (denom < eps) denom = 1.0f;
  const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
  const float bb = bins[below], ba = bins[above];
  if (below_out) *below_out = below;
  if (above_out) *above_out = above;
  return __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
}

__global__ void near_far_kernel(const float* __restrict__ o, const float* __restrict__ d, int64_t R, float r,
                                float* __restrict__ near, float* __restrict__ far) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= R) return;
  // torch.sum over the 3 products, left to right
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(o[3 * i], d[3 * i]), __fmul_rn(o[3 * i + 1], d[3 * i + 1])),
                              __fmul_rn(o[3 * i + 2], d[3 * i + 2]));
  const float mid = -dot;
  near[i] = fmaxf(mid - r, 0.0f);
  far[i] = fmaxf(mid + r, r);
}

// ---------------------------------------------------------------------------------------------
// sample_pdf / sample_cdf: one warp per ray, CDF [M] staged in dynamic smem.
// ---------------------------------------------------------------------------------------------
__global__ void sample_pdf_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                  const float* __restrict__ u_in, int64_t R, int M, int N, int cdf_given, float eps,
                                  float* __restrict__ samples, int* __restrict__ below, int* __restrict__ above,
                                  float* __restrict__ cdf_out) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= R) return;
  float* cdf = smem + (size_t)warp * 2 * M;  // [M]
  float* sb = cdf + M;                       // bins [M]
  const float* w = weights + ray * (int64_t)(M - 1);
  const float* b = bins + ray * (int64_t)M;
  for (int i = lane; i < M; i += 32) sb[i] = b[i];
  if (cdf_given) {
    for (int i = lane; i < M - 1; i += 32) cdf[i + 1] = w[i];
    if (lane == 0) cdf[0] = 0.0f;
    __syncwarp();
  } else {
    float part = 0.0f;
    for (int i = lane; i < M - 1; i += 32) {
      const float v = w[i] + 1e-5f;
      cdf[i + 1] = v;
      part += v;
    }
    const float tot = warp_sum(part);
    __syncwarp();
    for (int i = lane; i < M - 1; i += 32) cdf[i + 1] = __fdiv_rn(cdf[i + 1], tot);
    if (lane == 0) cdf[0] = 0.0f;
    __syncwarp();
    warp_scan_excl<false>(cdf + 1, nullptr, cdf + 1, M - 1, lane);
  }
  if (cdf_out)
    for (int i = lane; i < M; i += 32) cdf_out[ray * (int64_t)M + i] = cdf[i];
  for (int j = lane; j < N; j += 32) {
    const float u = u_in ? u_in[ray * (int64_t)N + j] : nr_linspace01(j, N);
    int bl, ab;
    const float s = invert_cdf(cdf, sb, M, u, eps, &bl, &ab);
    samples[ray * (int64_t)N + j] = s;
    if (below) below[ray * (int64_t)N + j] = bl;
    if (above) above[ray * (int64_t)N + j] = ab;
  }
}

// ---------------------------------------------------------------------------------------------
// sample_pdf / sample_cdf for short rows (even M <= 128, the NeuS up-sampling shapes): one THREAD per ray.
// The warp-per-ray kernel above issues ~550 warp instructions per 64-bin ray (ncu: 93 % of the issue slots busy at
// 16 % of the copy bandwidth); here a warp serves 32 rays with the same instruction stream.  A block's rows are
// contiguous in HBM, so they are staged as FLAT copies (16-byte loads, no per-row index math): the weights row of
// thread t starts at t*(M-1) floats (odd stride -> conflict-free lock-step access), the bins row at t*M (accessed at
// data-dependent indices only).  The arithmetic reproduces the warp kernel's association order exactly -- the
// strided partial sums and the butterfly of warp_sum, the lane segments and the Hillis-Steele pass of
// warp_scan_excl -- so both kernels return the same bits.
// ---------------------------------------------------------------------------------------------
constexpr int kRowThreads = 64;

__device__ __forceinline__ void flat_copy(float* dst, const float* __restrict__ src, int count, int tid) {
  if (((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0) {
    const int c4 = count >> 2;
    for (int i = tid; i < c4; i += kRowThreads) reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
    for (int i = (c4 << 2) + tid; i < count; i += kRowThreads) dst[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < count; i += kRowThreads) dst[i] = __ldg(src + i);
  }
}

// n > 32 (kSeg - 1), so only the last stride / the last lanes' segments can run past the row: the other bounds checks
// fold away at compile time
template <int kSeg>
__device__ __forceinline__ void row_weights_to_cdf(float* row, int n) {
  float part[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) {          // warp_sum's input: lane l adds the elements l, l + 32, ...
    float p = 0.0f;
#pragma unroll
    for (int k = 0; k < kSeg; ++k) {
      const int i = l + 32 * k;
      if (k < kSeg - 1 || i < n) { const float v = row[i] + 1e-5f; row[i] = v; p += v; }
    }
    part[l] = p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)        // the xor butterfly as lane 0 sees it
#pragma unroll
    for (int j = 0; j < o; ++j) part[j] = part[j] + part[j + o];
  const float tot = part[0];
  float incl[32];
#pragma unroll
  for (int l = 0; l < 32; ++l) {          // warp_scan_excl: lane l owns [l * kSeg, (l + 1) * kSeg)
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < kSeg; ++k) {
      const int i = l * kSeg + k;
      if (i <= 32 * (kSeg - 1) || i < n) { const float c = __fdiv_rn(row[i], tot); row[i] = c; t = t + c; }
    }
    incl[l] = t;
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
#pragma unroll
    for (int l = 31; l >= o; --l) incl[l] = incl[l] + incl[l - o];
#pragma unroll
  for (int l = 0; l < 32; ++l) {
    float run = l == 0 ? 0.0f : incl[l - 1];
#pragma unroll
    for (int k = 0; k < kSeg; ++k) {
      const int i = l * kSeg + k;
      if (i <= 32 * (kSeg - 1) || i < n) { run = run + row[i]; row[i] = run; }
    }
  }
}

template <int kSeg>
__global__ void __launch_bounds__(kRowThreads) sample_pdf_rows_kernel(const float* __restrict__ bins, const float* __restrict__ weights,
                                                                       const float* __restrict__ u_in, int64_t R, int M, int N,
                                                                       int cdf_given, float eps, float* __restrict__ samples) {
  extern __shared__ __align__(16) float smem[];
  const int n = M - 1, tid = threadIdx.x;
  const int64_t ray0 = blockIdx.x * (int64_t)kRowThreads;
  const int rows = (int)(R - ray0 < kRowThreads ? R - ray0 : kRowThreads);
  float* sw = smem;                                   // [rows][n]  weights -> cdf[1..M)
  float* sb = sw + ((kRowThreads * n + 3) & ~3);      // [rows][M]  bins
  float* su = sb + kRowThreads * M;                   // [rows][N]  uniforms in, samples out
  flat_copy(sw, weights + ray0 * n, rows * n, tid);
  flat_copy(sb, bins + ray0 * M, rows * M, tid);
  if (u_in) flat_copy(su, u_in + ray0 * N, rows * N, tid);
  __syncthreads();
  if (tid < rows) {
    float* row = sw + tid * n;
    const float* b = sb + tid * M;
    if (!cdf_given) row_weights_to_cdf<kSeg>(row, n);
    // nr_linspace01 with its division hoisted; the search as a fixed number of predicated rounds (floor(log2 M) + 1 cover
    // [0, M]) so that the unrolled samples interleave instead of waiting out 7 dependent shared-memory loads each
    const float step = N > 1 ? __fdiv_rn(1.0f, (float)(N - 1)) : 0.0f;
    const int rounds = 32 - __clz(M);
    float* my_u = su + tid * N;
    int j = tid % N;                                  // rotate the sample order per thread: su's row stride N is a power of two
#pragma unroll 4
    for (int jj = 0; jj < N; ++jj) {
      float u;
      if (u_in) u = my_u[j];
      else u = N <= 1 ? 0.0f : ((j < N / 2) ? __fmaf_rn(step, (float)j, 0.0f) : __fmaf_rn(-step, (float)(N - 1 - j), 1.0f));
      int lo = 0, hi = M;                             // lower_bound over cdf = (0, row[0..n)), the warp kernel's probe sequence
      for (int r = 0; r < rounds; ++r) {
        const int mid = (lo + hi) >> 1;
        const bool act = lo < hi;
        const float c = (act && mid) ? row[mid - 1] : 0.0f;
        const bool lt = c < u;
        lo = (act && lt) ? mid + 1 : lo;
        hi = (act && !lt) ? mid : hi;
      }
      const int below = max(lo - 1, 0), above = min(lo, M - 1);
      const float cb = below ? row[below - 1] : 0.0f, ca = above ? row[above - 1] : 0.0f;
      float denom = __fsub_rn(ca, cb);
      if (denom < eps) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(u, cb), denom);
      const float bb = b[below], ba = b[above];
      my_u[j] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
      j = j + 1 == N ? 0 : j + 1;
    }
  }
  __syncthreads();
  float* out = samples + ray0 * N;
  const int count = rows * N;
  if ((((uintptr_t)out) & 15) == 0 && (count & 3) == 0) {
    for (int i = tid; i < (count >> 2); i += kRowThreads) reinterpret_cast<float4*>(out)[i] = reinterpret_cast<const float4*>(su)[i];
  } else {
    for (int i = tid; i < count; i += kRowThreads) out[i] = su[i];
  }
}

// ---------------------------------------------------------------------------------------------
// NeuS ray prologue (neus.py:169-172,184-210)
// ---------------------------------------------------------------------------------------------
__global__ void neus_ray_setup_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d, int64_t R,
                                      float radius, float near_bypass, float far_bypass, int n_samples,
                                      float* __restrict__ dirs, float* __restrict__ near_out,
                                      float* __restrict__ far_out, float* __restrict__ d_new,
                                      float* __restrict__ pts_new) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= R) return;
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  float dx = rays_d[3 * ray], dy = rays_d[3 * ray + 1], dz = rays_d[3 * ray + 2];
  // F.normalize(dim=-1): x / max(||x||, 1e-12)
  const float nrm = fmaxf(sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz))), 1e-12f);
  dx = __fdiv_rn(dx, nrm); dy = __fdiv_rn(dy, nrm); dz = __fdiv_rn(dz, nrm);
  const float dot = __fadd_rn(__fadd_rn(__fmul_rn(ox, dx), __fmul_rn(oy, dy)), __fmul_rn(oz, dz));
  const float mid = -dot;
  float nr_ = fmaxf(mid - radius, 0.0f), fr_ = fmaxf(mid + radius, radius);
  if (!isnan(near_bypass)) nr_ = near_bypass;
  if (!isnan(far_bypass)) fr_ = far_bypass;
  if (lane == 0) {
    dirs[3 * ray] = dx; dirs[3 * ray + 1] = dy; dirs[3 * ray + 2] = dz;
    near_out[ray] = nr_; far_out[ray] = fr_;
  }
  for (int i = lane; i < n_samples; i += 32) {
    const float t = nr_linspace01(i, n_samples);
    // near * (1 - t) + far * t  (separate torch ops)
    const float d = __fadd_rn(__fmul_rn(nr_, __fsub_rn(1.0f, t)), __fmul_rn(fr_, t));
    d_new[ray * (int64_t)n_samples + i] = d;
    float* p = pts_new + (ray * (int64_t)n_samples + i) * 3;
    p[0] = __fadd_rn(ox, __fmul_rn(d, dx));
    p[1] = __fadd_rn(oy, __fmul_rn(d, dy));
    p[2] = __fadd_rn(oz, __fmul_rn(d, dz));
  }
}

// ---------------------------------------------------------------------------------------------
// NeuS up-sampling iteration (neus.py:249-277): merge + weights + inverse-CDF sampling.
// smem per warp: d[cap], sdf[cap], a[cap], b[cap], dn[nmax], sn[nmax]
// ---------------------------------------------------------------------------------------------
__global__ void neus_upsample_kernel(const float* __restrict__ rays_o, const float* __restrict__ dirs, int64_t R,
                                     float* __restrict__ d_buf, float* __restrict__ sdf_buf, int cap, int m_cur,
                                     const float* __restrict__ d_new, const float* __restrict__ sdf_new, int n_new,
                                     int iter, int n_next, const float* __restrict__ u_next,
                                     float* __restrict__ d_next, float* __restrict__ pts_next,
                                     float* __restrict__ pts_all, float* __restrict__ d_mid_out,
                                     float* __restrict__ pts_mid, float* __restrict__ nab_buf,
                                     const float* __restrict__ nab_new) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= R) return;
  float* sd = smem + (size_t)warp * (nab_buf ? 9 * cap : 6 * cap);
  float* ss = sd + cap;
  float* sa = ss + cap;   // scratch: alpha / weights / cdf
  float* sc = sa + cap;   // scratch: old d (during merge) / cdf
  float* dn = sc + cap;   // new d   [<= cap]
  float* sn = dn + cap;   // new sdf [<= cap]
  float* on = sn + cap;   // old normals [<= 3 cap], only with nab_buf
  const int M = m_cur + n_new;

  // ---- 1. merge (neus.py:272-276: cat, sort, gather) ----
  float* od = sc;  // old d
  float* os = sa;  // old sdf
  for (int i = lane; i < m_cur; i += 32) {
    od[i] = d_buf[ray * (int64_t)cap + i];
    os[i] = sdf_buf[ray * (int64_t)cap + i];
  }
  float* nrow = nab_buf ? nab_buf + ray * (int64_t)cap * 3 : nullptr;   // the samples' normals ride along with the merge
  if (nab_buf)
    for (int i = lane; i < 3 * m_cur; i += 32) on[i] = nrow[i];
  for (int j = lane; j < n_new; j += 32) {
    dn[j] = d_new[ray * (int64_t)n_new + j];
    sn[j] = sdf_new[ray * (int64_t)n_new + j];
  }
  __syncwarp();
  for (int i = lane; i < m_cur; i += 32) {  // old element: stable, old before equal new
    const float v = od[i];
    int c = 0;
    for (int k = 0; k < n_new; ++k) c += (dn[k] < v);
    sd[i + c] = v;
    ss[i + c] = os[i];
    if (nab_buf) {
      nrow[3 * (i + c)] = on[3 * i];
      nrow[3 * (i + c) + 1] = on[3 * i + 1];
      nrow[3 * (i + c) + 2] = on[3 * i + 2];
    }
  }
  for (int j = lane; j < n_new; j += 32) {
    const float v = dn[j];
    int lo = 0, hi = m_cur;  // upper_bound in old: count(old <= v)
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (od[mid] <= v) lo = mid + 1; else hi = mid;
    }
    int c = lo;
    for (int k = 0; k < n_new; ++k) c += (dn[k] < v) || (dn[k] == v && k < j);
    sd[c] = v;
    ss[c] = sn[j];
    if (nab_buf) {
      const float* src = nab_new + (ray * (int64_t)n_new + j) * 3;
      nrow[3 * c] = src[0];
      nrow[3 * c + 1] = src[1];
      nrow[3 * c + 2] = src[2];
    }
  }
  __syncwarp();
  for (int i = lane; i < M; i += 32) {
    d_buf[ray * (int64_t)cap + i] = sd[i];
    sdf_buf[ray * (int64_t)cap + i] = ss[i];
  }
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  const float dx = dirs[3 * ray], dy = dirs[3 * ray + 1], dz = dirs[3 * ray + 2];

  if (n_next > 0) {
    // ---- 2. slopes -> logistic CDF -> alpha (neus.py:253-268) ----
    const float s = 64.0f * (float)(1 << iter);
    for (int i = lane; i < M - 1; i += 32) {
      const float ps = ss[i], ns = ss[i + 1], pz = sd[i], nz = sd[i + 1];
      const float mid_sdf = __fmul_rn(__fadd_rn(ps, ns), 0.5f);
      const float dot = __fdiv_rn(__fsub_rn(ns, ps), __fadd_rn(__fsub_rn(nz, pz), 1e-5f));
      float prev_dot = 0.0f;
      if (i > 0) prev_dot = __fdiv_rn(__fsub_rn(ps, ss[i - 1]), __fadd_rn(__fsub_rn(pz, sd[i - 1]), 1e-5f));
      const float dv = fminf(fmaxf(fminf(prev_dot, dot), -10.0f), 0.0f);
      const float dist = __fsub_rn(nz, pz);
      const float half = __fmul_rn(__fmul_rn(dv, dist), 0.5f);
      const float prev_cdf = nr_sigmoid(__fmul_rn(__fsub_rn(mid_sdf, half), s));
      const float next_cdf = nr_sigmoid(__fmul_rn(__fadd_rn(mid_sdf, half), s));
      const float alpha = __fdiv_rn(__fadd_rn(__fsub_rn(prev_cdf, next_cdf), 1e-5f), __fadd_rn(prev_cdf, 1e-5f));
      sa[i] = alpha;
      sc[i] = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
    }
    __syncwarp();
    // w = alpha * exclusive cumprod(1 - alpha + 1e-10)   (neus.py:57-70)
    warp_scan_excl<true>(sc, sc, nullptr, M - 1, lane);
    // sample_pdf (rend_util.py:255-292): w + 1e-5 -> pdf -> cdf
    float part = 0.0f;
    for (int i = lane; i < M - 1; i += 32) {
      const float w = __fadd_rn(__fmul_rn(sa[i], sc[i]), 1e-5f);
      sa[i] = w;
      part += w;
    }
    const float tot = warp_sum(part);
    __syncwarp();
    for (int i = lane; i < M - 1; i += 32) sc[i + 1] = __fdiv_rn(sa[i], tot);
    if (lane == 0) sc[0] = 0.0f;
    __syncwarp();
    warp_scan_excl<false>(sc + 1, nullptr, sc + 1, M - 1, lane);
    for (int j = lane; j < n_next; j += 32) {
      const float u = u_next ? u_next[ray * (int64_t)n_next + j] : nr_linspace01(j, n_next);
      const float d = invert_cdf(sc, sd, M, u, 1e-5f, nullptr, nullptr);
      d_next[ray * (int64_t)n_next + j] = d;
      float* p = pts_next + (ray * (int64_t)n_next + j) * 3;
      p[0] = __fadd_rn(ox, __fmul_rn(d, dx));
      p[1] = __fadd_rn(oy, __fmul_rn(d, dy));
      p[2] = __fadd_rn(oz, __fmul_rn(d, dz));
    }
  } else {
    // ---- 3. final: points, mid depths and mid points (neus.py:284-288) ----
    for (int i = lane; i < M; i += 32) {
      const float d = sd[i];
      float* p = pts_all + (ray * (int64_t)M + i) * 3;
      p[0] = __fadd_rn(ox, __fmul_rn(dx, d));
      p[1] = __fadd_rn(oy, __fmul_rn(dy, d));
      p[2] = __fadd_rn(oz, __fmul_rn(dz, d));
      if (i < M - 1) {
        const float dm = __fmul_rn(0.5f, __fadd_rn(sd[i + 1], d));
        d_mid_out[ray * (int64_t)(M - 1) + i] = dm;
        float* q = pts_mid + (ray * (int64_t)(M - 1) + i) * 3;
        q[0] = __fadd_rn(ox, __fmul_rn(dx, dm));
        q[1] = __fadd_rn(oy, __fmul_rn(dy, dm));
        q[2] = __fadd_rn(oz, __fmul_rn(dz, dm));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// NeuS compositing, staged variant (M <= kStagedMaxM): the ray's 8M-4 input floats are first pulled into the warp's
// shared-memory slice with coalesced 4-byte cp.async copies, all in flight at once (32 per lane at M = 128), and the
// scan below reads them from there: the memory system sees independent, fully coalesced 16-byte requests instead of
// 4 dependent rounds of strided loads.
// ---------------------------------------------------------------------------------------------
constexpr int kStagedMaxM = 160;
// The staged kernel is bound by instruction issue, not by HBM (ncu: issue slots 91 % busy with IEEE division, expf
// and sqrt): it uses the 2-ulp hardware forms -- well inside the 1e-4 compositing tolerance; sample positions, which
// must be bit-exact, are not computed here.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
template <int kSeg>   // intervals per lane: ceil((M - 1) / 32) <= kSeg (every unrolled iteration costs issue slots, used or not)
__global__ void neus_composite_staged_kernel(const float* __restrict__ sdf, const float* __restrict__ nablas,
                                             const float* __restrict__ radiance, const float* __restrict__ d_mid,
                                             const float* __restrict__ s_dev, int64_t R, int M, int white_bkgd,
                                             float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ acc,
                                             float* __restrict__ normals, float* __restrict__ cdf_out,
                                             float* __restrict__ alpha_out, float* __restrict__ w_out, int vec16) {
  extern __shared__ __align__(16) float stage[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray0 = blockIdx.x * (int64_t)kWarpsPerBlock, ray = ray0 + warp;
  const int M1 = M - 1;
  const int nrays = (int)((R - ray0) < kWarpsPerBlock ? (R - ray0) : kWarpsPerBlock);
  // block-level staging: the kWarpsPerBlock rays of a block are contiguous in every input, and 4 rows of any length are
  // a multiple of 16 bytes, so whole blocks move with 16-byte copies when the base pointers allow it
  float* s_sd = stage;                               // [4][M]
  float* s_nb = s_sd + kWarpsPerBlock * M;           // [4][3M]
  float* s_rad = s_nb + kWarpsPerBlock * 3 * M;      // [4][3(M-1)]
  float* s_dm = s_rad + kWarpsPerBlock * 3 * M1;     // [4][M-1]
  auto stage_in = [&](float* dst, const float* src, int row) {   // rows of `row` floats for the block's rays
    const float* g = src + ray0 * (int64_t)row;
    const int n = nrays * row;
    if (vec16) {
      const int n4 = n >> 2;
      for (int i = threadIdx.x; i < n4; i += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + 4 * i)), "l"(g + 4 * i) : "memory");
      for (int i = 4 * n4 + threadIdx.x; i < n; i += blockDim.x) cp_async4(dst + i, g + i);
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async4(dst + i, g + i);
    }
  };
  if (vec16 && nrays == kWarpsPerBlock) {
    // full block, aligned inputs: four bulk copies (TMA unit) issued by one thread instead of 1024 16-byte cp.async
    // spread over the block -- the kernel is bound by instruction issue, and this is ~5 % of its instructions
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
      umma::mbar_init(&bar, 1);
      umma::fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t b_sd = 4u * M * 4u, b_nb = nablas ? 12u * M * 4u : 0u, b_rad = 12u * M1 * 4u, b_dm = 4u * M1 * 4u;
      umma::mbar_arrive_expect_tx(&bar, b_sd + b_nb + b_rad + b_dm);
      umma::bulk_g2s(s_sd, sdf + ray0 * (int64_t)M, b_sd, &bar);
      if (nablas) umma::bulk_g2s(s_nb, nablas + ray0 * (int64_t)(3 * M), b_nb, &bar);
      umma::bulk_g2s(s_rad, radiance + ray0 * (int64_t)(3 * M1), b_rad, &bar);
      umma::bulk_g2s(s_dm, d_mid + ray0 * (int64_t)M1, b_dm, &bar);
    }
    umma::mbar_wait(&bar, 0);
  } else {
    stage_in(s_sd, sdf, M);
    if (nablas) stage_in(s_nb, nablas, 3 * M);
    stage_in(s_rad, radiance, 3 * M1);
    stage_in(s_dm, d_mid, M1);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
  }
  if (ray >= R) return;
  const float* sd = s_sd + warp * M;
  const float* nb = s_nb + warp * 3 * M;
  const float* rad = s_rad + warp * 3 * M1;
  const float* dm = s_dm + warp * M1;
  const float s = *s_dev;
  float ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  // lane l owns the kSeg consecutive intervals [kSeg l, kSeg l + kSeg): one logistic per sample, a serial product
  // inside the lane and ONE multiplicative warp scan over the lane products (5 shuffles per ray instead of 5 per
  // 32 intervals)
  const int seg = (M1 + 31) >> 5;                     // intervals per lane (<= kSeg)
  const int i0 = lane * seg;
  float alpha[kSeg], tr[kSeg];                        // alpha_i and the transmittance before interval i inside the lane
  float prod = 1.0f;
  float c_prev = fast_sigmoid(sd[i0 < M ? i0 : M - 1] * s);
#pragma unroll
  for (int k = 0; k < kSeg; ++k) {
    const int i = i0 + k;
    const bool ok = k < seg && i < M1;
    const float c1 = fast_sigmoid(sd[i + 1 < M ? i + 1 : M - 1] * s);
    alpha[k] = ok ? fmaxf(__fdividef(c_prev - c1, c_prev + 1e-10f), 0.0f) : 0.0f;
    if (ok && cdf_out) {
      cdf_out[ray * (int64_t)M + i] = c_prev;
      if (i == M - 2) cdf_out[ray * (int64_t)M + i + 1] = c1;
    }
    tr[k] = prod;
    if (ok) prod *= (1.0f - alpha[k]) + 1e-10f;
    c_prev = c1;
  }
  float incl = prod;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(kFull, incl, o);
    if (lane >= o) incl *= t;
  }
  float before = __shfl_up_sync(kFull, incl, 1);      // transmittance before the lane's first interval
  if (lane == 0) before = 1.0f;
#pragma unroll
  for (int k = 0; k < kSeg; ++k) {
    const int i = i0 + k;
    if (k < seg && i < M1) {
      const float w = alpha[k] * (before * tr[k]);
      if (alpha_out) alpha_out[ray * (int64_t)M1 + i] = alpha[k];
      if (w_out) w_out[ray * (int64_t)M1 + i] = w;
      ar += w * rad[3 * i]; ag += w * rad[3 * i + 1]; ab += w * rad[3 * i + 2];
      ad += w * dm[i];
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

// ---------------------------------------------------------------------------------------------
// NeuS compositing (neus.py:28-35,57-70,346-381): one warp per ray, 32 intervals per step,
// multiplicative warp scan with a running carry for the exclusive transmittance.
// ---------------------------------------------------------------------------------------------
__global__ void neus_composite_kernel(const float* __restrict__ sdf, const float* __restrict__ nablas,
                                      const float* __restrict__ radiance, const float* __restrict__ d_mid,
                                      const float* __restrict__ s_dev, int64_t R, int M, int white_bkgd,
                                      float* __restrict__ rgb, float* __restrict__ depth, float* __restrict__ acc,
                                      float* __restrict__ normals, float* __restrict__ cdf_out,
                                      float* __restrict__ alpha_out, float* __restrict__ w_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= R) return;
  const float s = *s_dev;
  const float* sd = sdf + ray * (int64_t)M;
  const float* rad = radiance + ray * (int64_t)(M - 1) * 3;
  const float* dm = d_mid + ray * (int64_t)(M - 1);
  float carry = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  for (int base = 0; base < M - 1; base += 32) {
    const int i = base + lane;
    const bool ok = i < M - 1;
    float alpha = 0.0f, c0 = 0.0f;
    if (ok) {
      c0 = nr_sigmoid(__fmul_rn(sd[i], s));
      const float c1 = nr_sigmoid(__fmul_rn(sd[i + 1], s));
      alpha = fmaxf(__fdiv_rn(__fsub_rn(c0, c1), __fadd_rn(c0, 1e-10f)), 0.0f);
      if (cdf_out) {
        cdf_out[ray * (int64_t)M + i] = c0;
        if (i == M - 2) cdf_out[ray * (int64_t)M + i + 1] = c1;
      }
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
      if (alpha_out) alpha_out[ray * (int64_t)(M - 1) + i] = alpha;
      if (w_out) w_out[ray * (int64_t)(M - 1) + i] = w;
      ar += w * rad[3 * i]; ag += w * rad[3 * i + 1]; ab += w * rad[3 * i + 2];
      ad += w * dm[i];
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

// sdf_to_w with a fixed slope (neus.py:28-70 as used by the 'direct_use' / 'direct_more' up-samplers, :216-243):
// cdf = sigmoid(sdf s), alpha_i = max((cdf_i - cdf_{i+1}) / (cdf_i + 1e-10), 0), w_i = alpha_i prod_{j<i} (1 - alpha_j + 1e-10).
// One warp per ray, 32 intervals per round, the running product carried between rounds.
__global__ void neus_sdf_to_w_kernel(const float* __restrict__ sdf, float s, int64_t R, int M, float* __restrict__ w) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= R) return;
  const float* row = sdf + ray * (int64_t)M;
  float* out = w + ray * (int64_t)(M - 1);
  float carry = 1.0f;
  for (int base = 0; base < M - 1; base += 32) {
    const int i = base + lane;
    float alpha = 0.0f, q = 1.0f;
    if (i < M - 1) {
      const float c0 = nr_sigmoid(__fmul_rn(row[i], s)), c1 = nr_sigmoid(__fmul_rn(row[i + 1], s));
      alpha = fmaxf(__fdiv_rn(__fsub_rn(c0, c1), __fadd_rn(c0, 1e-10f)), 0.0f);
      q = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
    }
    float incl = q;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl *= t;
    }
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    if (i < M - 1) out[i] = alpha * (carry * excl);
    carry *= __shfl_sync(kFull, incl, 31);
  }
}

// ---------------------------------------------------------------------------------------------
// NeuS + NeRF++ background (neus.py:303-343).
// ---------------------------------------------------------------------------------------------
// Outside samples: d_out = far / flip(linspace(0,1,N_out+2)[1:-1]) (optionally jittered), appended to
// d_mid; inverted-sphere coordinates x_out = [p/|p|, 1/|p|] for all M1 + N_out depths.
__global__ void neus_outside_points_kernel(const float* __restrict__ rays_o, const float* __restrict__ dirs,
                                           const float* __restrict__ far, const float* __restrict__ d_mid, int64_t R,
                                           int M1, int n_out, const float* __restrict__ u, float* __restrict__ d_vals,
                                           float* __restrict__ x_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= R) return;
  const int T = M1 + n_out;
  const float fr = far[ray];
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  const float dx = dirs[3 * ray], dy = dirs[3 * ray + 1], dz = dirs[3 * ray + 2];
  auto dout = [&](int j) {  // j in [0, n_out): far / t, t = linspace(0,1,n_out+2)[n_out - j]
    return __fdiv_rn(fr, nr_linspace01(n_out - j, n_out + 2));
  };
  for (int i = lane; i < T; i += 32) {
    float d;
    if (i < M1) {
      d = d_mid[ray * (int64_t)M1 + i];
    } else {
      const int j = i - M1;
      d = dout(j);
      if (u) {  // stratified jitter between the mid-points of neighbouring depths (neus.py:306-311)
        const float lo = j == 0 ? d : __fmul_rn(0.5f, __fadd_rn(d, dout(j - 1)));
        const float hi = j == n_out - 1 ? d : __fmul_rn(0.5f, __fadd_rn(dout(j + 1), d));
        d = __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u[ray * (int64_t)n_out + j]));
      }
    }
    d_vals[ray * (int64_t)T + i] = d;
    const float px = ox + d * dx, py = oy + d * dy, pz = oz + d * dz;
    const float r = sqrtf(px * px + py * py + pz * pz);
    float* x = x_out + (ray * (int64_t)T + i) * 4;
    x[0] = px / r; x[1] = py / r; x[2] = pz / r; x[3] = 1.0f / r;
  }
}

// Compositing with the background blend (neus.py:320-352): inside alpha from the sdf where the mid
// point is inside the bounding sphere, NeRF++ alpha = 1 - exp(-softplus(sigma) * dist) elsewhere and for
// the n_out appended samples (last dist = 1e10).
__global__ void neus_composite_bg_kernel(const float* __restrict__ sdf, const float* __restrict__ nablas,
                                         const float* __restrict__ radiance, const float* __restrict__ rays_o,
                                         const float* __restrict__ dirs, const float* __restrict__ d_vals,
                                         const float* __restrict__ sigma_out, const float* __restrict__ radiance_out,
                                         const float* __restrict__ s_dev, float radius, int64_t R, int M, int n_out,
                                         int white_bkgd, float* __restrict__ rgb, float* __restrict__ depth,
                                         float* __restrict__ acc, float* __restrict__ normals, float* __restrict__ cdf_out,
                                         float* __restrict__ alpha_out, float* __restrict__ w_out,
                                         float* __restrict__ rad_blend_out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarpsPerBlock + warp;
  if (ray >= R) return;
  const float s = *s_dev;
  const int M1 = M - 1, T = M1 + n_out;
  const float* sd = sdf + ray * (int64_t)M;
  const float* dv = d_vals + ray * (int64_t)T;
  const float ox = rays_o[3 * ray], oy = rays_o[3 * ray + 1], oz = rays_o[3 * ray + 2];
  const float dx = dirs[3 * ray], dy = dirs[3 * ray + 1], dz = dirs[3 * ray + 2];
  float carry = 1.0f, ar = 0.f, ag = 0.f, ab = 0.f, ad = 0.f, aw = 0.f, nx = 0.f, ny = 0.f, nz = 0.f;
  for (int base = 0; base < T; base += 32) {
    const int i = base + lane;
    const bool ok = i < T;
    float alpha = 0.0f, c[3] = {0.f, 0.f, 0.f}, d = 0.0f;
    if (ok) {
      d = dv[i];
      const float dist = i + 1 < T ? dv[i + 1] - d : 1e10f;
      const float so = sigma_out[ray * (int64_t)T + i];
      const float sp = so > 20.0f ? so : log1pf(expf(so));              // F.softplus (beta=1, threshold=20)
      const float a_bg = 1.0f - expf(-sp * dist);
      const float* cb = radiance_out + (ray * (int64_t)T + i) * 3;
      bool inside = false;
      if (i < M1) {
        const float px = ox + d * dx, py = oy + d * dy, pz = oz + d * dz;
        inside = sqrtf(px * px + py * py + pz * pz) <= radius;
        const float c0 = nr_sigmoid(__fmul_rn(sd[i], s)), c1 = nr_sigmoid(__fmul_rn(sd[i + 1], s));
        if (cdf_out) { cdf_out[ray * (int64_t)M + i] = c0; if (i == M1 - 1) cdf_out[ray * (int64_t)M + i + 1] = c1; }
        if (inside) alpha = fmaxf(__fdiv_rn(__fsub_rn(c0, c1), __fadd_rn(c0, 1e-10f)), 0.0f);
      }
      if (!inside) alpha = a_bg;
      const float* ci = radiance + (ray * (int64_t)M1 + min(i, M1 - 1)) * 3;
#pragma unroll
      for (int k = 0; k < 3; ++k) c[k] = inside ? ci[k] : cb[k];
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
      if (alpha_out) alpha_out[ray * (int64_t)T + i] = alpha;
      if (w_out) w_out[ray * (int64_t)T + i] = w;
      if (rad_blend_out) {
        float* ro = rad_blend_out + (ray * (int64_t)T + i) * 3;
        ro[0] = c[0]; ro[1] = c[1]; ro[2] = c[2];
      }
      ar += w * c[0]; ag += w * c[1]; ab += w * c[2];
      ad += w * d;
      aw += w;
      if (nablas && i < M) {  // N_pts = min(len(w), len(normals)) = M (neus.py:366)
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

// ---------------------------------------------------------------------------------------------
// get_rays (rend_util.py:95-164): pinhole rays for selected pixels of B cameras.  pose [B,4,4]
// camera-to-world, intr [B,5] = fx, fy, cx, cy, skew; select_inds [B,N] (pixel = h*W + w) or NULL = all.
// ---------------------------------------------------------------------------------------------
__global__ void get_rays_kernel(const float* __restrict__ pose, const float* __restrict__ intr,
                                const int64_t* __restrict__ select_inds, int B, int W, int64_t N,
                                float* __restrict__ rays_o, float* __restrict__ rays_d) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= (int64_t)B * N) return;
  const int b = (int)(idx / N);
  const int64_t pix = select_inds ? select_inds[idx] : idx % N;
  const float x = (float)(pix % W), y = (float)(pix / W);
  const float fx = intr[5 * b], fy = intr[5 * b + 1], cx = intr[5 * b + 2], cy = intr[5 * b + 3], sk = intr[5 * b + 4];
  // lift (rend_util.py:95-109), z = 1
  const float xl = __fdiv_rn(__fsub_rn(__fadd_rn(__fsub_rn(x, cx), __fdiv_rn(__fmul_rn(cy, sk), fy)),
                                       __fdiv_rn(__fmul_rn(sk, y), fy)), fx);
  const float yl = __fdiv_rn(__fsub_rn(y, cy), fy);
  const float* p = pose + 16 * b;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    // world = p @ [xl, yl, 1, 1]; rays_d = world - cam_loc   (same two roundings as the reference)
    const float t = p[4 * r + 3];
    const float w = p[4 * r] * xl + p[4 * r + 1] * yl + p[4 * r + 2] + t;
    rays_d[3 * idx + r] = __fsub_rn(w, t);
    rays_o[3 * idx + r] = t;
  }
}

}  // namespace

extern "C" int nr_near_far_from_sphere(const float* rays_o, const float* rays_d, int64_t R, float r, float* near,
                                       float* far, void* stream) {
  NR_CHECK_ARG(R >= 0 && rays_o && rays_d && near && far, "nr_near_far_from_sphere: bad arguments");
  if (R == 0) return NR_OK;
  near_far_kernel<<<(unsigned)nr_cdiv(R, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, R, r, near, far);
  NR_CHECK_LAUNCH("near_far_kernel");
  return NR_OK;
}

extern "C" int nr_sample_pdf(const float* bins, const float* weights, const float* u, int64_t R, int32_t M, int32_t N,
                             int32_t cdf_is_given, float eps, float* samples, int32_t* below, int32_t* above,
                             float* cdf_out, void* stream) {
  NR_CHECK_ARG(R >= 0 && M >= 2 && N >= 1, "nr_sample_pdf: need R>=0, M>=2, N>=1 (got R=%lld M=%d N=%d)", (long long)R, M, N);
  if (R == 0) return NR_OK;  // empty batches carry null data pointers
  NR_CHECK_ARG(bins && weights && samples, "nr_sample_pdf: null pointer");
  if (!below && !above && !cdf_out && (M & 1) == 0 && M <= 128 && N <= 128) {   // short rows: one thread per ray
    const size_t smem_rows = ((size_t)((kRowThreads * (M - 1) + 3) & ~3) + (size_t)kRowThreads * (M + N)) * sizeof(float);
    const int seg = (M - 1 + 31) >> 5;
    auto kern = seg == 1 ? sample_pdf_rows_kernel<1> : seg == 2 ? sample_pdf_rows_kernel<2>
              : seg == 3 ? sample_pdf_rows_kernel<3> : sample_pdf_rows_kernel<4>;
    if (smem_rows > 48 * 1024)
      NR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
    kern<<<(unsigned)nr_cdiv(R, kRowThreads), kRowThreads, smem_rows, (cudaStream_t)stream>>>(bins, weights, u, R, M, N,
                                                                                             cdf_is_given, eps, samples);
    NR_CHECK_LAUNCH("sample_pdf_rows_kernel");
    return NR_OK;
  }
  const size_t smem = (size_t)kWarpsPerBlock * 2 * M * sizeof(float);
  NR_CHECK_ARG(smem <= 200 * 1024, "nr_sample_pdf: M=%d too large for shared memory staging", M);
  if (smem > 48 * 1024)
    NR_CHECK_CUDA(cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sample_pdf_kernel<<<(unsigned)nr_cdiv(R, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      bins, weights, u, R, M, N, cdf_is_given, eps, samples, below, above, cdf_out);
  NR_CHECK_LAUNCH("sample_pdf_kernel");
  return NR_OK;
}

extern "C" int nr_neus_ray_setup(const float* rays_o, const float* rays_d, int64_t R, float radius, float near_bypass,
                                 float far_bypass, int32_t n_samples, float* dirs, float* near, float* far,
                                 float* d_new, float* pts_new, void* stream) {
  NR_CHECK_ARG(rays_o && rays_d && dirs && near && far && d_new && pts_new, "nr_neus_ray_setup: null pointer");
  NR_CHECK_ARG(R >= 0 && n_samples >= 2, "nr_neus_ray_setup: bad sizes");
  if (R == 0) return NR_OK;
  neus_ray_setup_kernel<<<(unsigned)nr_cdiv(R, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      rays_o, rays_d, R, radius, near_bypass, far_bypass, n_samples, dirs, near, far, d_new, pts_new);
  NR_CHECK_LAUNCH("neus_ray_setup_kernel");
  return NR_OK;
}

extern "C" int nr_neus_upsample_step(const float* rays_o, const float* dirs, int64_t R, float* d_buf, float* sdf_buf,
                                     int32_t cap, int32_t m_cur, const float* d_new, const float* sdf_new,
                                     int32_t n_new, int32_t iter, int32_t n_next, const float* u_next, float* d_next,
                                     float* pts_next, float* pts_all, float* d_mid, float* pts_mid, float* nab_buf,
                                     const float* nab_new, void* stream) {
  NR_CHECK_ARG(rays_o && dirs && d_buf && sdf_buf && d_new && sdf_new, "nr_neus_upsample_step: null pointer");
  NR_CHECK_ARG((nab_buf != nullptr) == (nab_new != nullptr), "nr_neus_upsample_step: nab_buf and nab_new go together");
  NR_CHECK_ARG(m_cur >= 0 && n_new >= 1 && m_cur + n_new <= cap && m_cur + n_new >= 2,
               "nr_neus_upsample_step: m_cur=%d n_new=%d cap=%d", m_cur, n_new, cap);
  NR_CHECK_ARG(n_next >= 0 && iter >= 0 && iter < 24, "nr_neus_upsample_step: bad iter/n_next");
  if (n_next > 0) NR_CHECK_ARG(d_next && pts_next, "nr_neus_upsample_step: d_next/pts_next required");
  else NR_CHECK_ARG(pts_all && d_mid && pts_mid, "nr_neus_upsample_step: final outputs required");
  if (R == 0) return NR_OK;
  const size_t smem = (size_t)kWarpsPerBlock * (nab_buf ? 9 : 6) * cap * sizeof(float);
  NR_CHECK_ARG(smem <= 200 * 1024, "nr_neus_upsample_step: cap=%d too large", cap);
  if (smem > 48 * 1024)
    NR_CHECK_CUDA(cudaFuncSetAttribute(neus_upsample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  neus_upsample_kernel<<<(unsigned)nr_cdiv(R, kWarpsPerBlock), kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
      rays_o, dirs, R, d_buf, sdf_buf, cap, m_cur, d_new, sdf_new, n_new, iter, n_next, u_next, d_next, pts_next,
      pts_all, d_mid, pts_mid, nab_buf, nab_new);
  NR_CHECK_LAUNCH("neus_upsample_kernel");
  return NR_OK;
}

extern "C" int nr_neus_composite(const float* sdf, const float* nablas, const float* radiance, const float* d_mid,
                                 const float* s_dev, int64_t R, int32_t M, int32_t white_bkgd, float* rgb,
                                 float* depth, float* acc, float* normals, float* cdf_out, float* alpha_out,
                                 float* weights_out, void* stream) {
  NR_CHECK_ARG(sdf && radiance && d_mid && s_dev && rgb && depth && acc, "nr_neus_composite: null pointer");
  NR_CHECK_ARG(R >= 0 && M >= 2, "nr_neus_composite: bad sizes");
  NR_CHECK_ARG((nablas != nullptr) == (normals != nullptr), "nr_neus_composite: nablas and normals go together");
  if (R == 0) return NR_OK;
  if (M <= kStagedMaxM) {
    const size_t smem = (size_t)kWarpsPerBlock * 8 * M * sizeof(float);
    static bool carve = false;   // 16 KB per 4-warp block: ask for the large shared-memory carve-out so 13 blocks fit per SM
    if (!carve) {
      cudaFuncSetAttribute(neus_composite_staged_kernel<4>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      cudaFuncSetAttribute(neus_composite_staged_kernel<(kStagedMaxM + 31) / 32>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      carve = true;
    }
    const int vec16 = ((((uintptr_t)sdf | (uintptr_t)nablas | (uintptr_t)radiance | (uintptr_t)d_mid) & 15) == 0) ? 1 : 0;
    const unsigned grid = (unsigned)nr_cdiv(R, kWarpsPerBlock);
    if (M - 1 <= 4 * 32)   // 64 + 64 samples (every shipped NeuS config): 4 intervals per lane
      neus_composite_staged_kernel<4><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
          sdf, nablas, radiance, d_mid, s_dev, R, M, white_bkgd, rgb, depth, acc, normals, cdf_out, alpha_out, weights_out, vec16);
    else
      neus_composite_staged_kernel<(kStagedMaxM + 31) / 32><<<grid, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
          sdf, nablas, radiance, d_mid, s_dev, R, M, white_bkgd, rgb, depth, acc, normals, cdf_out, alpha_out, weights_out, vec16);
    NR_CHECK_LAUNCH("neus_composite_staged_kernel");
    return NR_OK;
  }
  neus_composite_kernel<<<(unsigned)nr_cdiv(R, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      sdf, nablas, radiance, d_mid, s_dev, R, M, white_bkgd, rgb, depth, acc, normals, cdf_out, alpha_out,
      weights_out);
  NR_CHECK_LAUNCH("neus_composite_kernel");
  return NR_OK;
}

extern "C" int nr_neus_sdf_to_w(const float* sdf, float s, int64_t R, int32_t M, float* w, void* stream) {
  NR_CHECK_ARG(R >= 0 && M >= 2, "nr_neus_sdf_to_w: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(sdf && w, "nr_neus_sdf_to_w: null pointer");
  neus_sdf_to_w_kernel<<<(unsigned)nr_cdiv(R, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(sdf, s, R, M, w);
  NR_CHECK_LAUNCH("neus_sdf_to_w_kernel");
  return NR_OK;
}

extern "C" int nr_neus_outside_points(const float* rays_o, const float* dirs, const float* far, const float* d_mid,
                                      int64_t R, int32_t M1, int32_t n_out, const float* u, float* d_vals, float* x_out,
                                      void* stream) {
  NR_CHECK_ARG(R >= 0 && M1 >= 1 && n_out >= 1, "nr_neus_outside_points: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(rays_o && dirs && far && d_mid && d_vals && x_out, "nr_neus_outside_points: null pointer");
  neus_outside_points_kernel<<<(unsigned)nr_cdiv(R, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      rays_o, dirs, far, d_mid, R, M1, n_out, u, d_vals, x_out);
  NR_CHECK_LAUNCH("neus_outside_points_kernel");
  return NR_OK;
}

extern "C" int nr_neus_composite_bg(const float* sdf, const float* nablas, const float* radiance, const float* rays_o,
                                    const float* dirs, const float* d_vals, const float* sigma_out,
                                    const float* radiance_out, const float* s_dev, float radius, int64_t R, int32_t M,
                                    int32_t n_out, int32_t white_bkgd, float* rgb, float* depth, float* acc,
                                    float* normals, float* cdf_out, float* alpha_out, float* weights_out,
                                    float* radiance_blend_out, void* stream) {
  NR_CHECK_ARG(R >= 0 && M >= 2 && n_out >= 1, "nr_neus_composite_bg: bad sizes");
  if (R == 0) return NR_OK;
  NR_CHECK_ARG(sdf && radiance && rays_o && dirs && d_vals && sigma_out && radiance_out && s_dev && rgb && depth && acc,
               "nr_neus_composite_bg: null pointer");
  NR_CHECK_ARG((nablas != nullptr) == (normals != nullptr), "nr_neus_composite_bg: nablas and normals go together");
  neus_composite_bg_kernel<<<(unsigned)nr_cdiv(R, kWarpsPerBlock), kWarpsPerBlock * 32, 0, (cudaStream_t)stream>>>(
      sdf, nablas, radiance, rays_o, dirs, d_vals, sigma_out, radiance_out, s_dev, radius, R, M, n_out, white_bkgd, rgb,
      depth, acc, normals, cdf_out, alpha_out, weights_out, radiance_blend_out);
  NR_CHECK_LAUNCH("neus_composite_bg_kernel");
  return NR_OK;
}

extern "C" int nr_get_rays(const float* pose, const float* intr, const int64_t* select_inds, int32_t B, int32_t W,
                           int64_t N, float* rays_o, float* rays_d, void* stream) {
  NR_CHECK_ARG(B >= 0 && N >= 0 && W >= 1, "nr_get_rays: bad sizes");
  if ((int64_t)B * N == 0) return NR_OK;
  NR_CHECK_ARG(pose && intr && rays_o && rays_d, "nr_get_rays: null pointer");
  get_rays_kernel<<<(unsigned)nr_cdiv((int64_t)B * N, 256), 256, 0, (cudaStream_t)stream>>>(pose, intr, select_inds, B, W,
                                                                                            N, rays_o, rays_d);
  NR_CHECK_LAUNCH("get_rays_kernel");
  return NR_OK;
}
