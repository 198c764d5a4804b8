// Backward of the alpha / transmittance / compositing pass of the three frameworks (training):
//   NeuS     neus.py:28-35,57-70,296,320-352      sdf -> Phi = sigmoid(s sdf) -> alpha -> w = alpha * excl-cumprod(1 - alpha + 1e-10)
//   VolSDF   volsdf.py:16-35,452-503              sdf -> sigma (Laplace CDF) -> p = exp(-relu(sigma delta)) -> tau = (1 - p + 1e-10) * excl-cumprod(p)
//   UNISURF  unisurf.py:54-62,216-240             logit -> alpha = e^-x / (1 + e^-x) -> w as NeuS
// each followed by rgb = sum w c, acc = sum w, depth = sum w d / (acc + 1e-10), normals = sum w normalize(nabla).
//
// The forward is the inference kernel (nr_*_composite with its per-sample outputs, which are what the trainers log
// anyway); this file is its hand-written adjoint: ONE launch per ray batch, one warp per ray, three sweeps over the ray's
// samples held in a per-warp shared-memory scratch:
//   A  (forward scan)   transmittance T_i = prod_{j<i} q_j from the saved alpha / p           (multiplicative warp scan)
//   B  (reverse scan)   gw_i = dL/dw_i from the ray gradients;  S_i = sum_{k>i} gw_k w_k       (additive warp scan, reversed)
//                       dL/da_i = gw_i T_i,  dL/dq_i = S_i / q_i;  writes g_radiance and g_nablas
//   D  (elementwise)    chain rule to the framework's inputs: g_sdf (+ per-ray partials of g_s, or g_alpha / g_beta),
//                       g_sigma_out of the NeRF++ samples, g_logit
// Upstream gradients taken: rgb, depth_volume, mask_volume, normals_volume, visibility_weights (any may be NULL).  The
// other per-sample outputs (alpha, cdf, p_i, sigma) are returned without a gradient path; no loss of the reference
// consumes them.
//
// Zeros in the product (VolSDF only: p = exp(-x) underflows for x > 103; NeuS / UNISURF factors are >= 1e-10): with z the
// first zero, d/dp_i for i < z needs no change (all w_k behind z vanish) and i > z get 0.  d/dp_z itself -- the one entry
// torch.cumprod's backward computes through its "first zero" path -- is never needed: the chain rule to the inputs
// multiplies it by dp_z/dx = -p_z = 0.  The sweep only has to keep 0/0 out (S_i / q_i with q_i = 0 -> 0).
#include "common.cuh"

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarps = 4;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// inclusive warp scans over the 32 lanes
__device__ __forceinline__ float scan_mul(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v *= t;
  }
  return v;
}
__device__ __forceinline__ float scan_add_rev(float v, int lane) {   // v_l <- sum_{m >= l} v_m
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float t = __shfl_down_sync(kFull, v, o);
    if (lane + o < 32) v += t;
  }
  return v;
}

enum { MODE_NEUS = 0, MODE_VOLSDF = 1, MODE_UNISURF = 2 };

struct BwdArgs {
  int64_t R;
  int K;                  // number of weights per ray
  // saved by the forward
  const float* a_saved;   // [R,K]  alpha (NeuS, UNISURF) or p (VolSDF)
  const float* w;         // [R,K]
  const float* acc;       // [R]
  const float* depth;     // [R]
  // per-weight colour and depth: segment 0 covers [0, K0), segment 1 the rest (VolSDF's NeRF++ samples); NeuS with the
  // background passes the blended radiance of the forward as segment 0
  const float* c0; const float* c1; int K0c; int ld_c0, ld_c1;
  const float* d0; const float* d1; int K0d; int ld_d0, ld_d1;
  const float* nablas; int Mn;       // [R,Mn,3] or NULL; normals use i < min(K, Mn)
  int white_bkgd;
  // upstream
  const float* g_rgb; const float* g_depth; const float* g_acc; const float* g_normals; const float* g_w;
  // outputs common
  float* g_c0; float* g_c1;          // same segmentation as c0 / c1 (g_c1 may be NULL)
  float* g_nablas;                   // [R,Mn,3] or NULL
  // ---- NeuS
  const float* sdf; const float* cdf; const float* s_dev; int M;     // sdf, cdf [R,M]
  const float* rays_o; const float* dirs; const float* d_vals; const float* sigma_out; float radius; int n_out;
  const float* rad_in;               // bg: inside radiance [R,M-1,3] is c-source only where inside; gradient routed by mask
  float* g_sdf; float* g_s_part;     // [R,M], [R]
  float* g_sigma_out;                // bg: [R,T]
  float* g_rad_in;                   // bg: [R,M-1,3] (g_c0 then is g_radiance_out [R,T,3])
  // ---- VolSDF
  const float* alpha_dev; const float* beta_dev; const float* sigma_all; int M_in, M_out;
  float* g_alpha_part; float* g_beta_part; float* g_sigma_o;   // [R], [R], [R,M_out]
  // ---- UNISURF
  const float* logits; float* g_logits;
};

template <int MODE>
__global__ void __launch_bounds__(kWarps * 32) composite_bwd_kernel(const BwdArgs a) {
  extern __shared__ float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t ray = blockIdx.x * (int64_t)kWarps + warp;
  if (ray >= a.R) return;
  const int K = a.K;
  float* sT = smem_f + (size_t)warp * 2 * K;    // transmittance
  float* sG = sT + K;                           // dL/d alpha_i (NeuS, UNISURF) or dL/d p_i (VolSDF)
  const float* as = a.a_saved + ray * (int64_t)K;

  // ---- A: transmittance ------------------------------------------------------------------------------------------
  float carry = 1.0f;
  int z = K;                                    // first zero factor (VolSDF)
  for (int base = 0; base < K; base += 32) {
    const int i = base + lane;
    const bool ok = i < K;
    float q = 1.0f;
    if (ok) q = MODE == MODE_VOLSDF ? as[i] : (1.0f - as[i]) + 1e-10f;
    const float incl = scan_mul(q, lane);
    float excl = __shfl_up_sync(kFull, incl, 1);
    if (lane == 0) excl = 1.0f;
    if (ok) sT[i] = carry * excl;
    carry *= __shfl_sync(kFull, incl, 31);
    if (MODE == MODE_VOLSDF) {
      const unsigned zero = __ballot_sync(kFull, ok && q == 0.0f);
      if (zero && z == K) z = base + __ffs(zero) - 1;
    }
  }
  __syncwarp();

  // ---- B: dL/dw, suffix sums, dL/da and dL/dq ------------------------------------------------------------------------
  const float A = a.acc[ray], Dp = a.depth ? a.depth[ray] : 0.0f;
  float gr[3] = {0.f, 0.f, 0.f}, gn[3] = {0.f, 0.f, 0.f};
  if (a.g_rgb) { gr[0] = a.g_rgb[3 * ray]; gr[1] = a.g_rgb[3 * ray + 1]; gr[2] = a.g_rgb[3 * ray + 2]; }
  if (a.g_normals) { gn[0] = a.g_normals[3 * ray]; gn[1] = a.g_normals[3 * ray + 1]; gn[2] = a.g_normals[3 * ray + 2]; }
  float g_acc = a.g_acc ? a.g_acc[ray] : 0.0f;
  if (a.white_bkgd) g_acc -= gr[0] + gr[1] + gr[2];             // rgb += 1 - acc
  const float g_dep = (a.g_depth ? a.g_depth[ray] : 0.0f) / (A + 1e-10f);
  const int Nn = a.nablas ? min(K, a.Mn) : 0;
  float suffix = 0.0f;
  for (int base = ((K - 1) / 32) * 32; base >= 0; base -= 32) {
    const int i = base + lane;
    const bool ok = i < K;
    float gw = 0.0f, wi = 0.0f, T = 0.0f, q = 1.0f;
    if (ok) {
      T = sT[i];
      const float s = as[i];
      q = MODE == MODE_VOLSDF ? s : (1.0f - s) + 1e-10f;
      // the weight is rebuilt from THIS sweep's transmittance instead of read back from the forward's output: dL/d alpha_i
      // = gw_i T_i - S_i / q_i is a difference of nearly equal terms wherever neighbouring samples have similar colours
      // (it telescopes to gw T_end / q_i), and a forward whose scan rounded T a few ulps differently leaves a residual
      // 1e3 times the rounding (measured: 2e-3 on sum(g_logit) of a UNISURF step)
      wi = (MODE == MODE_VOLSDF ? (1.0f - s) + 1e-10f : s) * T;
      const float* c = i < a.K0c ? a.c0 + (ray * (int64_t)a.ld_c0 + i) * 3 : a.c1 + (ray * (int64_t)a.ld_c1 + (i - a.K0c)) * 3;
      const float d = i < a.K0d ? a.d0[ray * (int64_t)a.ld_d0 + i] : a.d1[ray * (int64_t)a.ld_d1 + (i - a.K0d)];
      gw = gr[0] * c[0] + gr[1] * c[1] + gr[2] * c[2] + g_acc + g_dep * (d - Dp);
      if (a.g_w) gw += a.g_w[ray * (int64_t)K + i];
      // colour gradient (NeuS with background: routed by the inside mask in sweep D, which knows the mask)
      if (!(MODE == MODE_NEUS && a.n_out > 0)) {
        float* gc = i < a.K0c ? a.g_c0 + (ray * (int64_t)a.ld_c0 + i) * 3 : (a.g_c1 ? a.g_c1 + (ray * (int64_t)a.ld_c1 + (i - a.K0c)) * 3 : nullptr);
        if (gc) { gc[0] = wi * gr[0]; gc[1] = wi * gr[1]; gc[2] = wi * gr[2]; }
      }
      if (i < Nn) {
        const float* nb = a.nablas + (ray * (int64_t)a.Mn + i) * 3;
        const float x = nb[0], y = nb[1], zc = nb[2];
        const float nrm = sqrtf(x * x + y * y + zc * zc), den = fmaxf(nrm, 1e-12f), inv = 1.0f / den;
        const float ux = x * inv, uy = y * inv, uz = zc * inv;
        gw += gn[0] * ux + gn[1] * uy + gn[2] * uz;
        if (a.g_nablas) {
          // F.normalize backward: g / den - (nrm > eps) * x (x . g) / (den^2 nrm)
          const float hx = wi * gn[0], hy = wi * gn[1], hz = wi * gn[2];
          const float dot = nrm > 1e-12f ? (ux * hx + uy * hy + uz * hz) : 0.0f;
          float* go = a.g_nablas + (ray * (int64_t)a.Mn + i) * 3;
          go[0] = (hx - ux * dot) * inv; go[1] = (hy - uy * dot) * inv; go[2] = (hz - uz * dot) * inv;
        }
      }
    }
    const float h = gw * wi;
    const float incl = scan_add_rev(h, lane);                     // sum over lanes >= this one
    const float S = (incl - h) + suffix;                          // sum over k > i
    suffix += __shfl_sync(kFull, incl, 0);
    if (ok) {
      const float ga = gw * T;                                    // dL/d a_i
      float gq = q != 0.0f ? S / q : 0.0f;                        // dL/d q_i
      if (MODE == MODE_VOLSDF && i >= z) gq = 0.0f;
      sG[i] = MODE == MODE_VOLSDF ? gq - ga : ga - gq;            // a = 1 - p + eps, q = p  |  a = alpha, q = 1 - alpha + eps
    }
  }
  if (a.g_nablas) {   // samples past the weights carry no normal gradient
    for (int i = Nn + lane; i < a.Mn; i += 32) {
      float* go = a.g_nablas + (ray * (int64_t)a.Mn + i) * 3;
      go[0] = 0.f; go[1] = 0.f; go[2] = 0.f;
    }
  }
  __syncwarp();
  // ---- D: chain to the inputs -------------------------------------------------------------------------------------------
  if (MODE == MODE_NEUS) {
    const int M = a.M, M1 = M - 1;
    const float s = *a.s_dev;
    const float* sd = a.sdf + ray * (int64_t)M;
    const float* cd = a.cdf + ray * (int64_t)M;
    const bool bg = a.n_out > 0;
    float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0;
    if (bg) {
      ox = a.rays_o[3 * ray]; oy = a.rays_o[3 * ray + 1]; oz = a.rays_o[3 * ray + 2];
      dx = a.dirs[3 * ray]; dy = a.dirs[3 * ray + 1]; dz = a.dirs[3 * ray + 2];
    }
    auto inside = [&](int i) {   // the forward kernel's test (neus.py:328: |p_mid| <= radius)
      if (!bg) return true;
      const float d = a.d_vals[ray * (int64_t)K + i];
      const float px = ox + d * dx, py = oy + d * dy, pz = oz + d * dz;
      return sqrtf(px * px + py * py + pz * pz) <= a.radius;
    };
    // dL/dr_i with r = (Phi_i - Phi_{i+1}) / (Phi_i + eps), alpha = max(r, 0) (gradient passes at r >= 0 like clamp_min)
    auto g_r = [&](int i) {
      if (i < 0 || i >= M1 || !inside(i)) return 0.0f;
      const float r = __fdiv_rn(__fsub_rn(cd[i], cd[i + 1]), __fadd_rn(cd[i], 1e-10f));
      return r >= 0.0f ? sG[i] : 0.0f;
    };
    float gs = 0.0f;
    for (int i = lane; i < M; i += 32) {
      const float ph = cd[i];
      float gphi = 0.0f;
      if (i < M1) { const float den = ph + 1e-10f; gphi += g_r(i) * (cd[i + 1] + 1e-10f) / (den * den); }
      if (i > 0) gphi -= g_r(i - 1) / (cd[i - 1] + 1e-10f);
      const float gt = gphi * ph * (1.0f - ph);
      a.g_sdf[ray * (int64_t)M + i] = s * gt;
      gs += sd[i] * gt;
    }
    gs = warp_sum(gs);
    if (lane == 0) a.g_s_part[ray] = gs;
    if (bg) {
      for (int i = lane; i < K; i += 32) {
        const bool in = i < M1 && inside(i);
        const float sa = as[i];
        const float wi = sa * sT[i];
        const float gcx = wi * gr[0], gcy = wi * gr[1], gcz = wi * gr[2];
        float* go = a.g_c0 + (ray * (int64_t)K + i) * 3;          // g_radiance_out [R,T,3]
        go[0] = in ? 0.f : gcx; go[1] = in ? 0.f : gcy; go[2] = in ? 0.f : gcz;
        if (i < M1) {
          float* gi = a.g_rad_in + (ray * (int64_t)M1 + i) * 3;
          gi[0] = in ? gcx : 0.f; gi[1] = in ? gcy : 0.f; gi[2] = in ? gcz : 0.f;
        }
        float gso = 0.0f;
        if (!in) {
          const float d = a.d_vals[ray * (int64_t)K + i];
          const float dist = i + 1 < K ? a.d_vals[ray * (int64_t)K + i + 1] - d : 1e10f;
          const float so = a.sigma_out[ray * (int64_t)K + i];
          const float sp = so > 20.0f ? so : log1pf(expf(so));
          const float dsp = so > 20.0f ? 1.0f : 1.0f / (1.0f + expf(-so));
          gso = sG[i] * expf(-sp * dist) * dist * dsp;            // alpha_out = 1 - exp(-softplus(sigma) dist)
        }
        a.g_sigma_out[ray * (int64_t)K + i] = gso;
      }
    }
  } else if (MODE == MODE_VOLSDF) {
    const int M_in = a.M_in, Mt = a.M_in + a.M_out;
    const float al = *a.alpha_dev, be = *a.beta_dev;
    float ga_sum = 0.0f, gb_sum = 0.0f;
    for (int j = lane; j < Mt; j += 32) {
      float gsig = 0.0f;
      if (j < K) {
        const float dj = j < a.K0d ? a.d0[ray * (int64_t)a.ld_d0 + j] : a.d1[ray * (int64_t)a.ld_d1 + (j - a.K0d)];
        const int j1 = j + 1;
        const float dn = j1 < M_in ? a.d0[ray * (int64_t)a.ld_d0 + j1] : a.d1[ray * (int64_t)a.ld_d1 + (j1 - M_in)];
        const float delta = dn - dj;
        const float x = a.sigma_all[ray * (int64_t)Mt + j] * delta;
        gsig = x > 0.0f ? -sG[j] * as[j] * delta : 0.0f;          // p = exp(-relu(x))
      }
      if (j < M_in) {
        // sigma = alpha * psi, psi = s >= 0 ? e/2 : 1 - e/2, e = exp(-|s| / beta)
        const float sv = a.sdf[ray * (int64_t)M_in + j];
        const float e = 0.5f * expf(-fabsf(sv) / be);
        const float psi = sv >= 0.0f ? e : 1.0f - e;
        const float sgn = sv > 0.0f ? 1.0f : (sv < 0.0f ? -1.0f : 0.0f);   // d|s|/ds as autograd has it (0 at 0)
        const float de_ds = -e * sgn / be;                                 // d e / d s
        const float de_db = e * fabsf(sv) / (be * be);
        const float dpsi_ds = sv >= 0.0f ? de_ds : -de_ds;
        const float dpsi_db = sv >= 0.0f ? de_db : -de_db;
        a.g_sdf[ray * (int64_t)M_in + j] = gsig * al * dpsi_ds;
        ga_sum += gsig * psi;
        gb_sum += gsig * al * dpsi_db;
      } else if (a.g_sigma_o) {
        a.g_sigma_o[ray * (int64_t)a.M_out + (j - M_in)] = gsig;
      }
    }
    ga_sum = warp_sum(ga_sum); gb_sum = warp_sum(gb_sum);
    if (lane == 0) { a.g_alpha_part[ray] = ga_sum; a.g_beta_part[ray] = gb_sum; }
    // the last sample's colour has no weight
    {
      const int j = Mt - 1;
      float* gc = j < a.K0c ? a.g_c0 + (ray * (int64_t)a.ld_c0 + j) * 3 : (a.g_c1 ? a.g_c1 + (ray * (int64_t)a.ld_c1 + (j - a.K0c)) * 3 : nullptr);
      if (gc && lane < 3) gc[lane] = 0.0f;
    }
  } else {
    for (int i = lane; i < K; i += 32) {
      const float odds = expf(-a.logits[ray * (int64_t)K + i]);
      const float den = 1.0f + odds;
      a.g_logits[ray * (int64_t)K + i] = -sG[i] * (odds / (den * den));   // alpha = odds / (1 + odds)
    }
  }
}

template <int MODE>
int launch_bwd(const BwdArgs& a, const char* name, void* stream) {
  if (a.R == 0) return NR_OK;
  const size_t smem = (size_t)kWarps * 2 * a.K * sizeof(float);
  NR_CHECK_ARG(smem <= 200 * 1024, "%s: %d weights per ray exceed the shared-memory scratch", name, a.K);
  if (smem > 48 * 1024)
    NR_CHECK_CUDA(cudaFuncSetAttribute(composite_bwd_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  composite_bwd_kernel<MODE><<<(unsigned)nr_cdiv(a.R, kWarps), kWarps * 32, smem, (cudaStream_t)stream>>>(a);
  NR_CHECK_LAUNCH(name);
  return NR_OK;
}

}  // namespace

extern "C" int nr_neus_composite_bwd(const float* sdf, const float* cdf, const float* alpha, const float* weights,
                                     const float* radiance_used, const float* d_vals, const float* nablas,
                                     const float* s_dev, const float* acc, const float* depth, int64_t R, int32_t M,
                                     int32_t n_out, const float* rays_o, const float* dirs, const float* sigma_out,
                                     float radius, int32_t white_bkgd, const float* g_rgb, const float* g_depth,
                                     const float* g_acc, const float* g_normals, const float* g_weights, float* g_sdf,
                                     float* g_s_part, float* g_radiance, float* g_nablas, float* g_sigma_out,
                                     float* g_radiance_out, void* stream) {
  NR_CHECK_ARG(R >= 0 && M >= 2 && n_out >= 0, "nr_neus_composite_bwd: bad sizes");
  NR_CHECK_ARG(sdf && cdf && alpha && weights && radiance_used && d_vals && s_dev && acc && g_sdf && g_s_part && g_radiance,
               "nr_neus_composite_bwd: null pointer");
  NR_CHECK_ARG(!g_depth || depth, "nr_neus_composite_bwd: g_depth needs the forward's depth");
  NR_CHECK_ARG(!g_normals || nablas, "nr_neus_composite_bwd: g_normals needs nablas");
  NR_CHECK_ARG(n_out == 0 || (rays_o && dirs && sigma_out && g_sigma_out && g_radiance_out),
               "nr_neus_composite_bwd: background tensors missing");
  BwdArgs a = {};
  a.R = R; a.K = M - 1 + n_out; a.a_saved = alpha; a.w = weights; a.acc = acc; a.depth = depth;
  a.c0 = radiance_used; a.c1 = nullptr; a.K0c = a.K; a.ld_c0 = a.K; a.ld_c1 = 0;
  a.d0 = d_vals; a.d1 = nullptr; a.K0d = a.K; a.ld_d0 = a.K; a.ld_d1 = 0;
  a.nablas = (g_normals && nablas) ? nablas : nullptr; a.Mn = M; a.white_bkgd = white_bkgd;
  a.g_rgb = g_rgb; a.g_depth = g_depth; a.g_acc = g_acc; a.g_normals = g_normals; a.g_w = g_weights;
  a.g_c0 = n_out > 0 ? g_radiance_out : g_radiance; a.g_c1 = nullptr; a.g_nablas = a.nablas ? g_nablas : nullptr;
  a.sdf = sdf; a.cdf = cdf; a.s_dev = s_dev; a.M = M;
  a.rays_o = rays_o; a.dirs = dirs; a.d_vals = d_vals; a.sigma_out = sigma_out; a.radius = radius; a.n_out = n_out;
  a.g_sdf = g_sdf; a.g_s_part = g_s_part; a.g_sigma_out = g_sigma_out; a.g_rad_in = g_radiance;
  return launch_bwd<MODE_NEUS>(a, "neus_composite_bwd_kernel", stream);
}

extern "C" int nr_volsdf_composite_bwd(const float* sdf, const float* sigma_all, const float* p, const float* tau,
                                       const float* radiance, const float* d_in, const float* nablas,
                                       const float* alpha_dev, const float* beta_dev, const float* acc,
                                       const float* depth, int64_t R, int32_t M_in, const float* radiance_out,
                                       const float* d_out, int32_t M_out, int32_t white_bkgd, const float* g_rgb,
                                       const float* g_depth, const float* g_acc, const float* g_normals,
                                       const float* g_weights, float* g_sdf, float* g_alpha_part, float* g_beta_part,
                                       float* g_radiance, float* g_nablas, float* g_sigma_out, float* g_radiance_out,
                                       void* stream) {
  NR_CHECK_ARG(R >= 0 && M_in >= 2 && M_out >= 0, "nr_volsdf_composite_bwd: bad sizes");
  NR_CHECK_ARG(sdf && sigma_all && p && tau && radiance && d_in && alpha_dev && beta_dev && acc && g_sdf && g_alpha_part &&
                   g_beta_part && g_radiance, "nr_volsdf_composite_bwd: null pointer");
  NR_CHECK_ARG(!g_depth || depth, "nr_volsdf_composite_bwd: g_depth needs the forward's depth");
  NR_CHECK_ARG(!g_normals || nablas, "nr_volsdf_composite_bwd: g_normals needs nablas");
  NR_CHECK_ARG(M_out == 0 || (radiance_out && d_out && g_sigma_out && g_radiance_out),
               "nr_volsdf_composite_bwd: outside samples missing");
  BwdArgs a = {};
  a.R = R; a.K = M_in + M_out - 1; a.a_saved = p; a.w = tau; a.acc = acc; a.depth = depth;
  a.c0 = radiance; a.c1 = radiance_out; a.K0c = M_in; a.ld_c0 = M_in; a.ld_c1 = M_out;
  a.d0 = d_in; a.d1 = d_out; a.K0d = M_in; a.ld_d0 = M_in; a.ld_d1 = M_out;
  a.nablas = (g_normals && nablas) ? nablas : nullptr; a.Mn = M_in; a.white_bkgd = white_bkgd;
  a.g_rgb = g_rgb; a.g_depth = g_depth; a.g_acc = g_acc; a.g_normals = g_normals; a.g_w = g_weights;
  a.g_c0 = g_radiance; a.g_c1 = g_radiance_out; a.g_nablas = a.nablas ? g_nablas : nullptr;
  a.sdf = sdf; a.alpha_dev = alpha_dev; a.beta_dev = beta_dev; a.sigma_all = sigma_all; a.M_in = M_in; a.M_out = M_out;
  a.g_sdf = g_sdf; a.g_alpha_part = g_alpha_part; a.g_beta_part = g_beta_part; a.g_sigma_o = g_sigma_out;
  return launch_bwd<MODE_VOLSDF>(a, "volsdf_composite_bwd_kernel", stream);
}

extern "C" int nr_unisurf_composite_bwd(const float* logits, const float* alpha, const float* weights,
                                        const float* radiance, const float* d_all, const float* nablas,
                                        const float* acc, const float* depth, int64_t R, int32_t M, int32_t white_bkgd,
                                        const float* g_rgb, const float* g_depth, const float* g_acc,
                                        const float* g_normals, const float* g_weights, float* g_logits,
                                        float* g_radiance, float* g_nablas, void* stream) {
  NR_CHECK_ARG(R >= 0 && M >= 1, "nr_unisurf_composite_bwd: bad sizes");
  NR_CHECK_ARG(logits && alpha && weights && radiance && d_all && acc && g_logits && g_radiance,
               "nr_unisurf_composite_bwd: null pointer");
  NR_CHECK_ARG(!g_depth || depth, "nr_unisurf_composite_bwd: g_depth needs the forward's depth");
  NR_CHECK_ARG(!g_normals || nablas, "nr_unisurf_composite_bwd: g_normals needs nablas");
  BwdArgs a = {};
  a.R = R; a.K = M; a.a_saved = alpha; a.w = weights; a.acc = acc; a.depth = depth;
  a.c0 = radiance; a.K0c = M; a.ld_c0 = M; a.d0 = d_all; a.K0d = M; a.ld_d0 = M;
  a.nablas = (g_normals && nablas) ? nablas : nullptr; a.Mn = M; a.white_bkgd = white_bkgd;
  a.g_rgb = g_rgb; a.g_depth = g_depth; a.g_acc = g_acc; a.g_normals = g_normals; a.g_w = g_weights;
  a.g_c0 = g_radiance; a.g_nablas = a.nablas ? g_nablas : nullptr;
  a.logits = logits; a.g_logits = g_logits;
  return launch_bwd<MODE_UNISURF>(a, "unisurf_composite_bwd_kernel", stream);
}
