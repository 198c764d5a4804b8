// fp32 tier of the SDF / radiance MLPs (<= 1e-4 relative to the reference).
//
// One register-blocked SGEMM kernel (128x128x8 tiles, 8x8 per thread, double-buffered smem)
// with fused epilogues: bias + Softplus(beta=100) [+ its derivative for the tangents], ReLU,
// Sigmoid, the tangent product t' = softplus'(z) * (W t) of forward-mode differentiation, and
// the router of the last SDF layer (column 0 -> sdf, columns 1.. -> geometry feature).
//
// The analytic normal is computed in FORWARD mode: the three tangents d h / d x_c ride through
// the same weights as 3n extra GEMM rows, so forward_with_nablas is four row blocks through one
// kernel and nothing but the current layer has to be kept (models/base.py:243-282 does the same
// job with autograd.grad).  The tcgen05 path (mlp_umma.cu) uses the same formulation.
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8;
constexpr int kThreads = 256;
constexpr int kPadM = BM + 4;

enum EpiMode : int {
  EPI_NONE = 0,       // y = acc + b
  EPI_SOFTPLUS = 1,   // y = softplus100(acc + b); optional S = softplus100'(acc + b)
  EPI_RELU = 2,
  EPI_SIGMOID = 3,
  EPI_TANGENT = 4,    // y = acc * Sin[row % m_val, col]   (no bias)
  EPI_TANGENT_LIN = 5,// y = acc                           (no bias, last layer)
  EPI_SDF_LAST = 6,   // col 0 -> out0[row]; col >= 1 -> Y[row, col-1] (if Y)
};

struct GemmArgs {
  const float* A; int lda;       // [M, K] activations, row-major
  const float* W; int ldw;       // [N, K] weights, row-major (so C = A W^T)
  const float* bias;             // [N] or null
  int M, N, K;
  float* Y; int ldy;             // [M, N]
  float* S; int lds;             // softplus derivative out (EPI_SOFTPLUS) or null
  const float* Sin; int ldsin;   // EPI_TANGENT multiplier [m_val, N]
  int m_val;
  float* out0;                   // EPI_SDF_LAST column 0
  int mode;
};

__device__ __forceinline__ void load_tile4(const float* __restrict__ base, int ld, int row, int nrows, int k, int K,
                                           float (&v)[4]) {
  v[0] = v[1] = v[2] = v[3] = 0.0f;
  if (row >= nrows) return;
  const float* p = base + (size_t)row * ld + k;
  if (k + 3 < K) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (k + j < K) v[j] = p[j];
  }
}

__global__ void __launch_bounds__(kThreads) gemm_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[2][BK][kPadM];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float ra[4], rb[4];
  load_tile4(g.A, g.lda, m0 + lrow, g.M, lk, g.K, ra);
  load_tile4(g.W, g.ldw, n0 + lrow, g.N, lk, g.K, rb);
#pragma unroll
  for (int j = 0; j < 4; ++j) { As[0][lk + j][lrow] = ra[j]; Bs[0][lk + j][lrow] = rb[j]; }
  __syncthreads();

  const int nk = (g.K + BK - 1) / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load_tile4(g.A, g.lda, m0 + lrow, g.M, (kt + 1) * BK + lk, g.K, ra);
      load_tile4(g.W, g.ldw, n0 + lrow, g.N, (kt + 1) * BK + lk, g.K, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { As[cur ^ 1][lk + j][lrow] = ra[j]; Bs[cur ^ 1][lk + j][lrow] = rb[j]; }
    }
    __syncthreads();
  }

  // ---- epilogue ----
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= g.M) continue;
    const int srow = (g.mode == EPI_TANGENT) ? row % g.m_val : 0;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col0 = n0 + jh * 64 + tx * 4;
      float y[4], s[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = col0 + j;
        float z = acc[i][jh * 4 + j];
        s[j] = 0.0f;
        if (col < g.N) {
          switch (g.mode) {
            case EPI_NONE: case EPI_SDF_LAST: z += g.bias[col]; break;
            case EPI_SOFTPLUS: z += g.bias[col]; s[j] = nr_softplus100_grad(z); z = nr_softplus100(z); break;
            case EPI_RELU: z = fmaxf(z + g.bias[col], 0.0f); break;
            case EPI_SIGMOID: z = nr_sigmoid(z + g.bias[col]); break;
            case EPI_TANGENT: z *= g.Sin[(size_t)srow * g.ldsin + col]; break;
            default: break;
          }
        }
        y[j] = z;
      }
      if (g.mode == EPI_SDF_LAST) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = col0 + j;
          if (col >= g.N) continue;
          if (col == 0) g.out0[row] = y[j];
          else if (g.Y) g.Y[(size_t)row * g.ldy + col - 1] = y[j];
        }
      } else if (col0 + 3 < g.N && (g.ldy & 3) == 0) {
        *reinterpret_cast<float4*>(&g.Y[(size_t)row * g.ldy + col0]) = make_float4(y[0], y[1], y[2], y[3]);
        if (g.S && g.mode == EPI_SOFTPLUS)
          *reinterpret_cast<float4*>(&g.S[(size_t)row * g.lds + col0]) = make_float4(s[0], s[1], s[2], s[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = col0 + j;
          if (col >= g.N) continue;
          g.Y[(size_t)row * g.ldy + col] = y[j];
          if (g.S && g.mode == EPI_SOFTPLUS) g.S[(size_t)row * g.lds + col] = s[j];
        }
      }
    }
  }
}

// out[row*out_stride + out_off] = dot(A[row, 0:K], w[0:K]) (+ bias); one warp per row.
__global__ void rowdot_kernel(const float* __restrict__ A, int lda, const float* __restrict__ w,
                              const float* __restrict__ bias_dev, int64_t rows, int K, float* __restrict__ out,
                              int64_t m_val, int out_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t row = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* a = A + row * (int64_t)lda;
  float acc = 0.0f;
  for (int k = lane; k < K; k += 32) acc = fmaf(a[k], w[k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    // row = c * m_val + m  ->  out[m * out_stride + c]
    const int64_t c = row / m_val, m = row % m_val;
    out[m * out_stride + c] = acc + (bias_dev ? bias_dev[0] : 0.0f);
  }
}

// Positional encoding (Embedder.forward, models/base.py:46-64) and, optionally, its three
// tangents d pe / d x_c stacked as rows [c*n + m].  One thread per (point, column).
__global__ void embed_kernel(const float* __restrict__ x, int64_t n, int in_dim, int multires, float* __restrict__ pe,
                             int ld, int col_off, float* __restrict__ tpe, int ldt, int tcol_off, float scale) {
  const int pe_dim = multires < 0 ? in_dim : in_dim * (1 + 2 * multires);
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * pe_dim) return;
  const int64_t m = idx / pe_dim;
  const int j = (int)(idx % pe_dim);
  float val, dval;
  int comp;
  if (j < in_dim) {
    comp = j; val = x[m * in_dim + j]; dval = 1.0f;
  } else {
    const int q = (j - in_dim) / (2 * in_dim), r = (j - in_dim) % (2 * in_dim);
    comp = r % in_dim;
    const float f = (float)(1 << q);
    const float a = x[m * in_dim + comp] * f;
    if (r < in_dim) { val = sinf(a); dval = f * cosf(a); }
    else { val = cosf(a); dval = -f * sinf(a); }
  }
  if (pe) pe[m * ld + col_off + j] = val * scale;
  if (tpe) {
    for (int c = 0; c < 3; ++c) tpe[(c * n + m) * ldt + tcol_off + j] = (c == comp) ? dval * scale : 0.0f;
  }
}

// dst[row, dst_off + j] = src[row, src_off + j] * scale, j < ncols
__global__ void copy_cols_kernel(const float* __restrict__ src, int lds, int src_off, float* __restrict__ dst, int ldd,
                                 int dst_off, int64_t rows, int ncols, float scale) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= rows * ncols) return;
  const int64_t r = idx / ncols;
  const int j = (int)(idx % ncols);
  dst[r * ldd + dst_off + j] = src[r * lds + src_off + j] * scale;
}

int launch_gemm(const GemmArgs& g, cudaStream_t st) {
  if (g.M == 0) return NR_OK;
  dim3 grid((unsigned)nr_cdiv(g.M, BM), (unsigned)nr_cdiv(g.N, BN));
  gemm_kernel<<<grid, kThreads, 0, st>>>(g);
  NR_CHECK_LAUNCH("gemm_kernel");
  return NR_OK;
}

int launch_embed(const float* x, int64_t n, int in_dim, int multires, float* pe, int ld, int col_off, float* tpe,
                 int ldt, int tcol_off, float scale, cudaStream_t st) {
  const int pe_dim = multires < 0 ? in_dim : in_dim * (1 + 2 * multires);
  const int64_t tot = n * pe_dim;
  if (tot == 0) return NR_OK;
  embed_kernel<<<(unsigned)nr_cdiv(tot, 256), 256, 0, st>>>(x, n, in_dim, multires, pe, ld, col_off, tpe, ldt,
                                                            tcol_off, scale);
  NR_CHECK_LAUNCH("embed_kernel");
  return NR_OK;
}

int launch_copy_cols(const float* src, int lds, int src_off, float* dst, int ldd, int dst_off, int64_t rows, int ncols,
                     float scale, cudaStream_t st) {
  const int64_t tot = rows * ncols;
  if (tot == 0) return NR_OK;
  copy_cols_kernel<<<(unsigned)nr_cdiv(tot, 256), 256, 0, st>>>(src, lds, src_off, dst, ldd, dst_off, rows, ncols,
                                                                scale);
  NR_CHECK_LAUNCH("copy_cols_kernel");
  return NR_OK;
}

int launch_rowdot(const float* A, int lda, const float* w, const float* bias_dev, int64_t rows, int K, float* out,
                  int64_t m_val, int out_stride, cudaStream_t st) {
  if (rows == 0) return NR_OK;
  rowdot_kernel<<<(unsigned)nr_cdiv(rows, 8), 256, 0, st>>>(A, lda, w, bias_dev, rows, K, out, m_val, out_stride);
  NR_CHECK_LAUNCH("rowdot_kernel");
  return NR_OK;
}

int sdf_pe_dim(const nr_sdf_net_t* net) { return net->multires < 0 ? 3 : 3 * (1 + 2 * net->multires); }

int check_sdf_net(const nr_sdf_net_t* net) {
  NR_CHECK_ARG(net, "sdf net: null descriptor");
  NR_CHECK_ARG(net->n_layers >= 2 && net->n_layers <= NR_MAX_LAYERS, "sdf net: n_layers=%d unsupported", net->n_layers);
  const int pe = sdf_pe_dim(net);
  NR_CHECK_ARG(net->in_dim[0] == pe, "sdf net: layer 0 in_dim %d != embedding dim %d", net->in_dim[0], pe);
  for (int i = 0; i < net->n_layers; ++i) {
    NR_CHECK_ARG(net->W[i] && net->b[i], "sdf net: layer %d has null weights", i);
    if (i > 0) {
      const int expect = net->out_dim[i - 1] + (i == net->skip_layer ? pe : 0);
      NR_CHECK_ARG(net->in_dim[i] == expect, "sdf net: layer %d in_dim %d != %d", i, net->in_dim[i], expect);
    }
  }
  NR_CHECK_ARG(net->skip_layer != 0 && net->skip_layer < net->n_layers - 1, "sdf net: skip_layer %d unsupported",
               net->skip_layer);
  return NR_OK;
}

int max_width(const int32_t* in_dim, const int32_t* out_dim, int L) {
  int w = 0;
  for (int i = 0; i < L; ++i) { w = max(w, nr_pad4(in_dim[i])); w = max(w, nr_pad4(out_dim[i])); }
  return w;
}

struct SdfWs {
  float *pe, *tpe, *h[2], *t[2], *s;
  int ldpe, ldh;
  size_t bytes;
};

SdfWs carve_sdf_ws(const nr_sdf_net_t* net, int64_t n, bool nablas, void* ws) {
  SdfWs w{};
  w.ldpe = nr_pad4(sdf_pe_dim(net));
  w.ldh = max_width(net->in_dim, net->out_dim, net->n_layers - 1);  // hidden buffers (last layer goes to outputs)
  w.ldh = max(w.ldh, nr_pad4(net->in_dim[net->n_layers - 1]));
  char* p = (char*)ws;
  size_t off = 0;
  auto take = [&](size_t nfloats) { float* r = (float*)(p + off); off += nr_align(nfloats * sizeof(float)); return r; };
  w.pe = take((size_t)n * w.ldpe);
  w.h[0] = take((size_t)n * w.ldh);
  w.h[1] = take((size_t)n * w.ldh);
  if (nablas) {
    w.tpe = take((size_t)3 * n * w.ldpe);
    w.t[0] = take((size_t)3 * n * w.ldh);
    w.t[1] = take((size_t)3 * n * w.ldh);
    w.s = take((size_t)n * w.ldh);
  }
  w.bytes = off;
  return w;
}

int run_sdf(const nr_sdf_net_t* net, const float* x, int64_t n, float* sdf, float* nabla, float* feat, int64_t feat_ld,
            void* ws, size_t ws_bytes, cudaStream_t st) {
  int rc = check_sdf_net(net);
  if (rc) return rc;
  NR_CHECK_ARG(x && sdf && n >= 0, "sdf forward: bad arguments");
  NR_CHECK_ARG(n < (int64_t)1 << 28, "sdf forward: n too large for one call, chunk it");
  if (n == 0) return NR_OK;
  const bool nab = nabla != nullptr;
  SdfWs w = carve_sdf_ws(net, n, nab, ws);
  if (w.bytes > ws_bytes || !ws) {
    nr_set_error("sdf forward: workspace %zu bytes < required %zu", ws_bytes, w.bytes);
    return NR_ERR_WORKSPACE;
  }
  const int L = net->n_layers, pe = sdf_pe_dim(net);
  if ((rc = launch_embed(x, n, 3, net->multires, w.pe, w.ldpe, 0, nab ? w.tpe : nullptr, w.ldpe, 0, 1.0f, st))) return rc;
  const float* hin = w.pe; int ldin = w.ldpe;
  const float* tin = w.tpe;
  for (int i = 0; i < L - 1; ++i) {
    float* hout = w.h[i & 1];
    float* tout = nab ? w.t[i & 1] : nullptr;
    if (i == net->skip_layer) {
      // input = cat([h, pe]) / sqrt2 ; the 1/sqrt2 is folded into W[i] by the host packer
      const int off = net->out_dim[i - 1];
      if ((rc = launch_copy_cols(w.pe, w.ldpe, 0, (float*)hin, ldin, off, n, pe, 1.0f, st))) return rc;
      if (nab && (rc = launch_copy_cols(w.tpe, w.ldpe, 0, (float*)tin, ldin, off, 3 * n, pe, 1.0f, st))) return rc;
    }
    GemmArgs g{};
    g.A = hin; g.lda = ldin; g.W = net->W[i]; g.ldw = nr_pad4(net->in_dim[i]); g.bias = net->b[i];
    g.M = (int)n; g.N = net->out_dim[i]; g.K = net->in_dim[i];
    g.Y = hout; g.ldy = w.ldh; g.S = nab ? w.s : nullptr; g.lds = w.ldh; g.mode = EPI_SOFTPLUS;
    if ((rc = launch_gemm(g, st))) return rc;
    if (nab) {
      GemmArgs t = g;
      t.A = tin; t.M = (int)(3 * n); t.Y = tout; t.S = nullptr; t.Sin = w.s; t.ldsin = w.ldh; t.m_val = (int)n;
      t.bias = nullptr; t.mode = EPI_TANGENT;
      if ((rc = launch_gemm(t, st))) return rc;
    }
    hin = hout; ldin = w.ldh; tin = tout;
  }
  // last (linear) layer: column 0 = sdf, columns 1.. = geometry feature (base.py:253-257)
  const int i = L - 1;
  if (feat) {
    NR_CHECK_ARG(net->out_dim[i] >= 2, "sdf forward: feature requested but last layer has %d outputs", net->out_dim[i]);
    GemmArgs g{};
    g.A = hin; g.lda = ldin; g.W = net->W[i]; g.ldw = nr_pad4(net->in_dim[i]); g.bias = net->b[i];
    g.M = (int)n; g.N = net->out_dim[i]; g.K = net->in_dim[i];
    g.Y = feat; g.ldy = (int)feat_ld; g.out0 = sdf; g.mode = EPI_SDF_LAST;
    if ((rc = launch_gemm(g, st))) return rc;
  } else {
    // sdf only: row 0 of the last layer is a dot product per point
    if ((rc = launch_rowdot(hin, ldin, net->W[i], net->b[i], n, net->in_dim[i], sdf, n, 1, st))) return rc;
  }
  if (nab) {
    // nabla[m, c] = W_last[0, :] . t[c*n + m, :]
    if ((rc = launch_rowdot(tin, ldin, net->W[i], nullptr, 3 * n, net->in_dim[i], nabla, n, 3, st))) return rc;
  }
  return NR_OK;
}

}  // namespace

extern "C" size_t nr_sdf_forward_f32_workspace(const nr_sdf_net_t* net, int64_t n) {
  if (!net || n < 0) return 0;
  return carve_sdf_ws(net, n, false, nullptr).bytes;
}
extern "C" size_t nr_sdf_forward_nablas_f32_workspace(const nr_sdf_net_t* net, int64_t n) {
  if (!net || n < 0) return 0;
  return carve_sdf_ws(net, n, true, nullptr).bytes;
}

extern "C" int nr_sdf_forward_f32(const nr_sdf_net_t* net, const float* x, int64_t n, float* sdf, float* feat,
                                  int64_t feat_ld, void* ws, size_t ws_bytes, void* stream) {
  return run_sdf(net, x, n, sdf, nullptr, feat, feat_ld, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int nr_sdf_forward_nablas_f32(const nr_sdf_net_t* net, const float* x, int64_t n, float* sdf, float* nabla,
                                         float* feat, int64_t feat_ld, void* ws, size_t ws_bytes, void* stream) {
  NR_CHECK_ARG(nabla, "nr_sdf_forward_nablas_f32: nabla output required");
  return run_sdf(net, x, n, sdf, nabla, feat, feat_ld, ws, ws_bytes, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// RadianceNet.forward (models/base.py:372-391): cat([PE(x), PE_view(v), normals, feat]) ->
// D ReLU layers -> Sigmoid(3).
// ---------------------------------------------------------------------------------------------
namespace {
struct RadWs { float *in0, *h[2]; int ld0, ldh; size_t bytes; };

int rad_dims(const nr_radiance_net_t* net, int* px, int* pv) {
  *px = net->multires < 0 ? 3 : 3 * (1 + 2 * net->multires);
  *pv = net->multires_view < 0 ? 3 : 3 * (1 + 2 * net->multires_view);
  return *px + *pv + 3 + net->feat_dim;
}

RadWs carve_rad_ws(const nr_radiance_net_t* net, int64_t n, void* ws) {
  RadWs w{};
  w.ld0 = nr_pad4(net->in_dim[0]);
  w.ldh = max_width(net->in_dim + 1, net->out_dim, net->n_layers - 1);
  char* p = (char*)ws;
  size_t off = 0;
  auto take = [&](size_t nfloats) { float* r = (float*)(p + off); off += nr_align(nfloats * sizeof(float)); return r; };
  w.in0 = take((size_t)n * w.ld0);
  w.h[0] = take((size_t)n * w.ldh);
  w.h[1] = take((size_t)n * w.ldh);
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" size_t nr_radiance_forward_f32_workspace(const nr_radiance_net_t* net, int64_t n) {
  if (!net || n < 0) return 0;
  return carve_rad_ws(net, n, nullptr).bytes;
}

extern "C" int nr_radiance_forward_f32(const nr_radiance_net_t* net, const float* x, const float* view,
                                       const float* normals, const float* feat, int64_t feat_ld, int64_t n, float* rgb,
                                       void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NR_CHECK_ARG(net && x && view && normals && feat && rgb && n >= 0, "nr_radiance_forward_f32: bad arguments");
  NR_CHECK_ARG(net->n_layers >= 2 && net->n_layers <= NR_MAX_LAYERS, "radiance net: n_layers=%d", net->n_layers);
  int px, pv;
  const int in0 = rad_dims(net, &px, &pv);
  NR_CHECK_ARG(net->in_dim[0] == in0, "radiance net: layer 0 in_dim %d != %d", net->in_dim[0], in0);
  for (int i = 1; i < net->n_layers; ++i)
    NR_CHECK_ARG(net->in_dim[i] == net->out_dim[i - 1], "radiance net: skips are not supported (layer %d)", i);
  if (n == 0) return NR_OK;
  RadWs w = carve_rad_ws(net, n, ws);
  if (w.bytes > ws_bytes || !ws) {
    nr_set_error("radiance forward: workspace %zu bytes < required %zu", ws_bytes, w.bytes);
    return NR_ERR_WORKSPACE;
  }
  int rc;
  if ((rc = launch_embed(x, n, 3, net->multires, w.in0, w.ld0, 0, nullptr, 0, 0, 1.0f, st))) return rc;
  if ((rc = launch_embed(view, n, 3, net->multires_view, w.in0, w.ld0, px, nullptr, 0, 0, 1.0f, st))) return rc;
  if ((rc = launch_copy_cols(normals, 3, 0, w.in0, w.ld0, px + pv, n, 3, 1.0f, st))) return rc;
  if ((rc = launch_copy_cols(feat, (int)feat_ld, 0, w.in0, w.ld0, px + pv + 3, n, net->feat_dim, 1.0f, st))) return rc;
  const float* hin = w.in0; int ldin = w.ld0;
  const int L = net->n_layers;
  for (int i = 0; i < L; ++i) {
    GemmArgs g{};
    g.A = hin; g.lda = ldin; g.W = net->W[i]; g.ldw = nr_pad4(net->in_dim[i]); g.bias = net->b[i];
    g.M = (int)n; g.N = net->out_dim[i]; g.K = net->in_dim[i];
    if (i < L - 1) { g.Y = w.h[i & 1]; g.ldy = w.ldh; g.mode = EPI_RELU; }
    else { g.Y = rgb; g.ldy = net->out_dim[i]; g.mode = EPI_SIGMOID; }
    if ((rc = launch_gemm(g, st))) return rc;
    hin = w.h[i & 1]; ldin = w.ldh;
  }
  return NR_OK;
}

// ---------------------------------------------------------------------------------------------
// NeRF++ background MLP (NeRF.forward, models/base.py:426-453, use_view_dirs=True):
// PE(x4) -> D ReLU layers, the embedding re-concatenated IN FRONT of h after layer `skip`
// (base.py:434-435) -> sigma = alpha_linear(h) (raw) ; feature_linear(h) | PE(view) -> W/2 ReLU ->
// rgb_linear -> sigmoid.
// ---------------------------------------------------------------------------------------------
namespace {
struct NerfWs { float *pe, *h[2], *fv, *hv; int ldpe, ldh, ldfv, ldhv; size_t bytes; };

int pe_dim_of(int in_dim, int multires) { return multires < 0 ? in_dim : in_dim * (1 + 2 * multires); }

NerfWs carve_nerf_ws(const nr_nerf_net_t* net, int64_t n, void* ws) {
  NerfWs w{};
  const int pe = pe_dim_of(net->input_dim, net->multires), pv = pe_dim_of(3, net->multires_view);
  w.ldpe = nr_pad4(pe);
  w.ldh = nr_pad4(net->width + pe);
  w.ldfv = nr_pad4(net->width + pv);
  w.ldhv = nr_pad4(net->width / 2);
  char* p = (char*)ws;
  size_t off = 0;
  auto take = [&](size_t nfloats) { float* r = (float*)(p + off); off += nr_align(nfloats * sizeof(float)); return r; };
  w.pe = take((size_t)n * w.ldpe);
  w.h[0] = take((size_t)n * w.ldh);
  w.h[1] = take((size_t)n * w.ldh);
  w.fv = take((size_t)n * w.ldfv);
  w.hv = take((size_t)n * w.ldhv);
  w.bytes = off;
  return w;
}
}  // namespace

extern "C" size_t nr_nerf_forward_f32_workspace(const nr_nerf_net_t* net, int64_t n) {
  if (!net || n < 0) return 0;
  return carve_nerf_ws(net, n, nullptr).bytes;
}

extern "C" int nr_nerf_forward_f32(const nr_nerf_net_t* net, const float* x, const float* view, int64_t n, float* sigma,
                                   float* rgb, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  NR_CHECK_ARG(net && n >= 0, "nr_nerf_forward_f32: bad arguments");
  NR_CHECK_ARG(net->depth >= 2 && net->depth <= NR_MAX_LAYERS && net->input_dim >= 1 && net->input_dim <= 4,
               "nr_nerf_forward_f32: depth=%d input_dim=%d", net->depth, net->input_dim);
  if (n == 0) return NR_OK;
  NR_CHECK_ARG(x && view && sigma && rgb, "nr_nerf_forward_f32: null pointer");
  NerfWs w = carve_nerf_ws(net, n, ws);
  if (w.bytes > ws_bytes || !ws) {
    nr_set_error("nerf forward: workspace %zu bytes < required %zu", ws_bytes, w.bytes);
    return NR_ERR_WORKSPACE;
  }
  const int pe = pe_dim_of(net->input_dim, net->multires), pv = pe_dim_of(3, net->multires_view), W = net->width;
  int rc;
  if ((rc = launch_embed(x, n, net->input_dim, net->multires, w.pe, w.ldpe, 0, nullptr, 0, 0, 1.0f, st))) return rc;
  const float* hin = w.pe; int ldin = w.ldpe; int kin = pe;
  for (int i = 0; i < net->depth; ++i) {
    // layer i's output goes behind the embedding when the NEXT layer consumes cat([pe, h])
    const bool cat_next = (i == net->skip);
    float* hout = w.h[i & 1];
    GemmArgs g{};
    g.A = hin; g.lda = ldin; g.W = net->pts_W[i]; g.ldw = nr_pad4(kin); g.bias = net->pts_b[i];
    g.M = (int)n; g.N = W; g.K = kin; g.Y = hout + (cat_next ? pe : 0); g.ldy = w.ldh; g.mode = EPI_RELU;
    NR_CHECK_ARG(!cat_next || (pe & 3) == 0, "nr_nerf_forward_f32: embedding width must be a multiple of 4 for the skip");
    if ((rc = launch_gemm(g, st))) return rc;
    if (cat_next && (rc = launch_copy_cols(w.pe, w.ldpe, 0, hout, w.ldh, 0, n, pe, 1.0f, st))) return rc;
    hin = hout; ldin = w.ldh; kin = cat_next ? W + pe : W;
  }
  // sigma = alpha_linear(h)  (raw, base.py:438)
  if ((rc = launch_rowdot(hin, ldin, net->alpha_W, net->alpha_b, n, kin, sigma, n, 1, st))) return rc;
  // feature_linear(h) | PE(view)
  {
    GemmArgs g{};
    g.A = hin; g.lda = ldin; g.W = net->feature_W; g.ldw = nr_pad4(kin); g.bias = net->feature_b;
    g.M = (int)n; g.N = W; g.K = kin; g.Y = w.fv; g.ldy = w.ldfv; g.mode = EPI_NONE;
    if ((rc = launch_gemm(g, st))) return rc;
    if ((rc = launch_embed(view, n, 3, net->multires_view, w.fv, w.ldfv, W, nullptr, 0, 0, 1.0f, st))) return rc;
  }
  {
    GemmArgs g{};
    g.A = w.fv; g.lda = w.ldfv; g.W = net->views_W; g.ldw = nr_pad4(W + pv); g.bias = net->views_b;
    g.M = (int)n; g.N = W / 2; g.K = W + pv; g.Y = w.hv; g.ldy = w.ldhv; g.mode = EPI_RELU;
    if ((rc = launch_gemm(g, st))) return rc;
  }
  {
    GemmArgs g{};
    g.A = w.hv; g.lda = w.ldhv; g.W = net->rgb_W; g.ldw = nr_pad4(W / 2); g.bias = net->rgb_b;
    g.M = (int)n; g.N = 3; g.K = W / 2; g.Y = rgb; g.ldy = 3; g.mode = EPI_SIGMOID;
    if ((rc = launch_gemm(g, st))) return rc;
  }
  return NR_OK;
}

// ---------------------------------------------------------------------------------------------
// Building blocks of the TRAINING path (neurecon_b200/models/autograd.py).  The backward of the
// forward-mode network (value rows h, tangent rows t_c; z = W h + b, u_c = W t_c, h' = sp(z),
// t'_c = sp'(z) u_c) for upstream gradients (g_h', g_t'_c) is
//     g_u_c = g_t'_c * sp'(z)
//     g_z   = g_h' * sp'(z) + sum_c g_t'_c * u_c * sp''(z)          <- second-order (eikonal) path
//     g_W   = g_z^T h + sum_c g_u_c^T t_c ,  g_b = colsum(g_z)
//     g_h   = g_z W ,  g_t_c = g_u_c W
// ---------------------------------------------------------------------------------------------
namespace {

// g_z / g_u in place of g_h' / g_t' (see above).  S = sp'(z) [n, N]; u [3n, N]; sp'' = 100 S (1 - S).
// u_scaled: the array holds the layer's OUTPUT tangent t' = S u (what the forward pass keeps anyway as the next layer's
// input), so u sp'' = 100 (1 - S) t'.
// 4 row groups x 64 column quads per block, kBwdRows rows per block; 16-byte accesses when every leading dimension is a
// multiple of 4 floats and the bases are 16-byte aligned.  `gz_colsum` (optional) += column sums of g_z: the layer's
// bias gradient, which would otherwise re-read g_z in a launch of its own.
constexpr int kBwdRows = 64;

__device__ __forceinline__ float bwd_act_elem(float gh, float s, float g0, float g1, float g2, float u0, float u1, float u2,
                                              int u_scaled, float* o0, float* o1, float* o2) {
  const float s2 = u_scaled ? 100.0f * (1.0f - s) : 100.0f * s * (1.0f - s);
  float gz = gh * s;
  gz += g0 * u0 * s2; *o0 = g0 * s;
  gz += g1 * u1 * s2; *o1 = g1 * s;
  gz += g2 * u2 * s2; *o2 = g2 * s;
  return gz;
}

template <bool kVec>
__global__ void __launch_bounds__(256) sdf_bwd_act_kernel(float* __restrict__ gh, int ldgh, float* __restrict__ gt, int ldgt,
                                                          const float* __restrict__ S, int lds, const float* __restrict__ u,
                                                          int ldu, int64_t n, int N, int u_scaled, float* __restrict__ gz_colsum) {
  __shared__ float red[4][256];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int64_t r0 = blockIdx.x * (int64_t)kBwdRows, r1 = min(n, r0 + kBwdRows);
  for (int j0 = 0; j0 < N; j0 += 256) {
    const int j = j0 + 4 * tx;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < N) {
      const bool quad = kVec && j + 3 < N;
      for (int64_t m = r0 + ty; m < r1; m += 4) {
        if (quad) {
          const float4 s4 = *reinterpret_cast<const float4*>(S + m * lds + j);
          const float4 h4 = *reinterpret_cast<const float4*>(gh + m * ldgh + j);
          float4 g4[3], u4[3], o4[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            g4[c] = *reinterpret_cast<const float4*>(gt + (c * n + m) * ldgt + j);
            u4[c] = *reinterpret_cast<const float4*>(u + (c * n + m) * ldu + j);
          }
          float4 z4;
          z4.x = bwd_act_elem(h4.x, s4.x, g4[0].x, g4[1].x, g4[2].x, u4[0].x, u4[1].x, u4[2].x, u_scaled, &o4[0].x, &o4[1].x, &o4[2].x);
          z4.y = bwd_act_elem(h4.y, s4.y, g4[0].y, g4[1].y, g4[2].y, u4[0].y, u4[1].y, u4[2].y, u_scaled, &o4[0].y, &o4[1].y, &o4[2].y);
          z4.z = bwd_act_elem(h4.z, s4.z, g4[0].z, g4[1].z, g4[2].z, u4[0].z, u4[1].z, u4[2].z, u_scaled, &o4[0].z, &o4[1].z, &o4[2].z);
          z4.w = bwd_act_elem(h4.w, s4.w, g4[0].w, g4[1].w, g4[2].w, u4[0].w, u4[1].w, u4[2].w, u_scaled, &o4[0].w, &o4[1].w, &o4[2].w);
          *reinterpret_cast<float4*>(gh + m * ldgh + j) = z4;
#pragma unroll
          for (int c = 0; c < 3; ++c) *reinterpret_cast<float4*>(gt + (c * n + m) * ldgt + j) = o4[c];
          acc[0] += z4.x; acc[1] += z4.y; acc[2] += z4.z; acc[3] += z4.w;
        } else {
          for (int k = 0; k < 4 && j + k < N; ++k) {
            const int jj = j + k;
            float o[3];
            const float gz = bwd_act_elem(gh[m * ldgh + jj], S[m * lds + jj], gt[m * ldgt + jj], gt[(n + m) * ldgt + jj],
                                          gt[(2 * n + m) * ldgt + jj], u[m * ldu + jj], u[(n + m) * ldu + jj],
                                          u[(2 * n + m) * ldu + jj], u_scaled, &o[0], &o[1], &o[2]);
            gh[m * ldgh + jj] = gz;
#pragma unroll
            for (int c = 0; c < 3; ++c) gt[(c * n + m) * ldgt + jj] = o[c];
            acc[k] += gz;
          }
        }
      }
    }
    if (gz_colsum) {                                       // block-uniform
#pragma unroll
      for (int k = 0; k < 4; ++k) red[ty][4 * tx + k] = acc[k];
      __syncthreads();
      const int jc = j0 + (int)threadIdx.x;
      if (jc < N) atomicAdd(gz_colsum + jc, (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]));
      __syncthreads();
    }
  }
}

// dW[i, j] += sum_r G[r, i] * X[r, j]   (i < N, j < K), split over row ranges, fp32 atomics.
__global__ void __launch_bounds__(kThreads) gemm_tn_kernel(const float* __restrict__ G, int ldg,
                                                           const float* __restrict__ X, int ldx, int64_t rows, int N,
                                                           int K, float* __restrict__ dW, int lddw, int rows_per_block) {
  __shared__ __align__(16) float As[2][BK][kPadM];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int i0 = blockIdx.x * BM, j0 = blockIdx.y * BN;
  const int64_t r_begin = (int64_t)blockIdx.z * rows_per_block;
  const int64_t r_end = min(rows, r_begin + rows_per_block);
  const int lr = tid >> 5, lc = (tid & 31) * 4;  // 8 rows x 128 columns per tile: one float4 per thread
  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
  auto load = [&](const float* base, int ld, int64_t r, int c0, int cmax, float (&v)[4]) {
    v[0] = v[1] = v[2] = v[3] = 0.0f;
    if (r >= r_end) return;
    const float* p = base + r * ld + c0;
    if (c0 + 3 < cmax && (ld & 3) == 0) {
      const float4 q = *reinterpret_cast<const float4*>(p);
      v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (c0 + j < cmax) v[j] = p[j];
    }
  };
  float ra[4], rb[4];
  const int nk = (int)((r_end - r_begin + BK - 1) / BK);
  if (nk <= 0) return;
  load(G, ldg, r_begin + lr, i0 + lc, N, ra);
  load(X, ldx, r_begin + lr, j0 + lc, K, rb);
  *reinterpret_cast<float4*>(&As[0][lr][lc]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
  *reinterpret_cast<float4*>(&Bs[0][lr][lc]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      load(G, ldg, r_begin + (int64_t)(kt + 1) * BK + lr, i0 + lc, N, ra);
      load(X, ldx, r_begin + (int64_t)(kt + 1) * BK + lr, j0 + lc, K, rb);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      *reinterpret_cast<float4*>(&As[cur ^ 1][lr][lc]) = make_float4(ra[0], ra[1], ra[2], ra[3]);
      *reinterpret_cast<float4*>(&Bs[cur ^ 1][lr][lc]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = i0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = j0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (col < K) atomicAdd(&dW[(size_t)row * lddw + col], acc[i][j]);
    }
  }
}

// out[j] += sum_r G[r, j]
__global__ void colsum_kernel(const float* __restrict__ G, int ldg, int64_t rows, int N, float* __restrict__ out,
                              int rows_per_block) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s = 0.0f;
  for (int64_t r = r0; r < r1; ++r) s += G[r * ldg + j];
  atomicAdd(&out[j], s);
}

// y = x * (ref > 0)   (ReLU backward) / y = x * ref * (1 - ref)   (sigmoid backward), in place on x
__global__ void act_bwd_kernel(float* __restrict__ x, int ldx, const float* __restrict__ ref, int ldr, int64_t rows,
                               int N, int mode) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= rows * N) return;
  const int64_t r = idx / N;
  const int j = (int)(idx % N);
  const float y = ref[r * ldr + j];
  x[r * ldx + j] *= (mode == 0) ? (y > 0.0f ? 1.0f : 0.0f) : y * (1.0f - y);
}
}  // namespace

extern "C" int nr_gemm_f32(const float* A, int32_t lda, const float* W, int32_t ldw, const float* bias, int64_t M,
                           int32_t N, int32_t K, float* Y, int32_t ldy, int32_t mode, float* S, int32_t lds,
                           const float* aux, int32_t ldaux, int64_t m_val, void* stream) {
  NR_CHECK_ARG(M >= 0 && M < ((int64_t)1 << 31) && N >= 1 && K >= 1, "nr_gemm_f32: bad sizes");
  if (M == 0) return NR_OK;
  NR_CHECK_ARG(A && W && Y, "nr_gemm_f32: null pointer");
  NR_CHECK_ARG(mode >= EPI_NONE && mode <= EPI_TANGENT_LIN, "nr_gemm_f32: mode %d", mode);
  NR_CHECK_ARG((lda & 3) == 0 && (ldw & 3) == 0, "nr_gemm_f32: lda / ldw must be multiples of 4");
  NR_CHECK_ARG(mode == EPI_TANGENT || mode == EPI_TANGENT_LIN || bias, "nr_gemm_f32: bias required");
  NR_CHECK_ARG(mode != EPI_TANGENT || (aux && m_val > 0), "nr_gemm_f32: tangent mode needs the derivative matrix");
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.bias = bias; g.M = (int)M; g.N = N; g.K = K; g.Y = Y; g.ldy = ldy;
  g.S = S; g.lds = lds; g.Sin = aux; g.ldsin = ldaux; g.m_val = (int)m_val; g.mode = mode;
  return launch_gemm(g, (cudaStream_t)stream);
}

extern "C" int nr_gemm_tn_f32(const float* G, int32_t ldg, const float* X, int32_t ldx, int64_t rows, int32_t N,
                              int32_t K, float* dW, int32_t lddw, void* stream) {
  NR_CHECK_ARG(rows >= 0 && N >= 1 && K >= 1, "nr_gemm_tn_f32: bad sizes");
  if (rows == 0) return NR_OK;
  NR_CHECK_ARG(G && X && dW, "nr_gemm_tn_f32: null pointer");
  const int tiles = (int)(nr_cdiv(N, BM) * nr_cdiv(K, BN));
  int splits = (int)std::min<int64_t>(std::max<int64_t>(1, (148 * 4) / tiles), nr_cdiv(rows, 256));
  const int rpb = (int)(nr_cdiv(nr_cdiv(rows, splits), BK) * BK);
  splits = (int)nr_cdiv(rows, rpb);
  dim3 grid((unsigned)nr_cdiv(N, BM), (unsigned)nr_cdiv(K, BN), (unsigned)splits);
  gemm_tn_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(G, ldg, X, ldx, rows, N, K, dW, lddw, rpb);
  NR_CHECK_LAUNCH("gemm_tn_kernel");
  return NR_OK;
}

extern "C" int nr_colsum_f32(const float* G, int32_t ldg, int64_t rows, int32_t N, float* out, void* stream) {
  NR_CHECK_ARG(rows >= 0 && N >= 1, "nr_colsum_f32: bad sizes");
  if (rows == 0) return NR_OK;
  NR_CHECK_ARG(G && out, "nr_colsum_f32: null pointer");
  const int rpb = 512;
  dim3 grid((unsigned)nr_cdiv(N, 128), (unsigned)nr_cdiv(rows, rpb));
  colsum_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(G, ldg, rows, N, out, rpb);
  NR_CHECK_LAUNCH("colsum_kernel");
  return NR_OK;
}

extern "C" int nr_sdf_bwd_act_f32(float* gh, int32_t ldgh, float* gt, int32_t ldgt, const float* S, int32_t lds,
                                  const float* u, int32_t ldu, int64_t n, int32_t N, int32_t u_scaled, float* gz_colsum,
                                  void* stream) {
  NR_CHECK_ARG(n >= 0 && N >= 1, "nr_sdf_bwd_act_f32: bad sizes");
  if (n == 0) return NR_OK;
  NR_CHECK_ARG(gh && gt && S && u, "nr_sdf_bwd_act_f32: null pointer");
  const bool vec = ((ldgh | ldgt | lds | ldu) & 3) == 0 && ((((uintptr_t)gh) | ((uintptr_t)gt) | ((uintptr_t)S) | ((uintptr_t)u)) & 15) == 0;
  const unsigned grid = (unsigned)nr_cdiv(n, kBwdRows);
  if (vec)
    sdf_bwd_act_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(gh, ldgh, gt, ldgt, S, lds, u, ldu, n, N, u_scaled, gz_colsum);
  else
    sdf_bwd_act_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(gh, ldgh, gt, ldgt, S, lds, u, ldu, n, N, u_scaled, gz_colsum);
  NR_CHECK_LAUNCH("sdf_bwd_act_kernel");
  return NR_OK;
}

extern "C" int nr_act_bwd_f32(float* x, int32_t ldx, const float* ref, int32_t ldr, int64_t rows, int32_t N,
                              int32_t mode, void* stream) {
  NR_CHECK_ARG(rows >= 0 && N >= 1 && (mode == 0 || mode == 1), "nr_act_bwd_f32: bad arguments");
  if (rows == 0) return NR_OK;
  NR_CHECK_ARG(x && ref, "nr_act_bwd_f32: null pointer");
  act_bwd_kernel<<<(unsigned)nr_cdiv(rows * N, 256), 256, 0, (cudaStream_t)stream>>>(x, ldx, ref, ldr, rows, N, mode);
  NR_CHECK_LAUNCH("act_bwd_kernel");
  return NR_OK;
}

extern "C" int nr_embed_f32(const float* x, int64_t n, int32_t in_dim, int32_t multires, float* pe, int32_t ld,
                            int32_t col_off, float* tpe, int32_t ldt, int32_t tcol_off, void* stream) {
  NR_CHECK_ARG(n >= 0 && in_dim >= 1 && in_dim <= 4, "nr_embed_f32: bad sizes");
  if (n == 0) return NR_OK;
  NR_CHECK_ARG(x && (pe || tpe), "nr_embed_f32: null pointer");
  return launch_embed(x, n, in_dim, multires, pe, ld, col_off, tpe, ldt, tcol_off, 1.0f, (cudaStream_t)stream);
}
