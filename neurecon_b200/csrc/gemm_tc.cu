// Tensor-core building blocks of the training path (models/autograd.py): the same contracts as nr_gemm_f32 /
// nr_gemm_tn_f32 (fp32 in global memory, fp32 accumulation, fp32 out, fused epilogues), with the operands converted to
// 16 bits on their way into shared memory and the products on tcgen05 / TMEM.
//
// nr_gemm_tc:  Y[M, N] = epilogue(A[M, K] W[N, K]^T + bias)      M = points (x4 with tangents): large; N, K <= 320
//   * W (<= 160 KB as 16-bit) is converted once per CTA into K-major SWIZZLE_128B chunks [Npad x 64 k] and stays
//     resident; CTAs are persistent over 128-row tiles of A.
//   * warps 0-7 load and convert the A tile ([128 x 64 k] per stage, 3 stages; two groups of four warps take alternate
//     k-chunks: two chunks = 64 KB of loads in flight per SM), warp 8 issues the MMAs (M = 128, N = Npad, one
//     accumulator per tile, double buffered in TMEM), warps 9-16 run the epilogue from TMEM: two warps per TMEM lane
//     quarter, each on half of the columns.  ncu: the kernel is latency bound (issue slots 18 % busy, tensor pipe 6 %,
//     long-scoreboard stalls); four -> eight epilogue warps bought 23 %, four -> eight loader warps 3 %.
// nr_gemm_tn_tc: dW[N, K] += G[rows, N]^T X[rows, K]             reduction over the rows, split across CTAs
//   * both operands are MN-major tiles [64 rows x (128 | <=256) columns] converted on the fly; each CTA owns one
//     (128-row slice of dW) x (<=256-column slice) accumulator and a range of row chunks, and adds its partial sum
//     with fp32 atomics (as the SIMT version does).
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int kBM = 128, kKC = 64, kAStages = 3;
constexpr int kGemmThreads = 17 * 32;   // 8 loader warps, 1 MMA warp, 8 epilogue warps
constexpr uint32_t kAStageBytes = kBM * kKC * 2;   // 16 KB

enum : int { M_NONE = 0, M_SOFTPLUS = 1, M_RELU = 2, M_SIGMOID = 3, M_TANGENT = 4, M_LINEAR = 5 };

struct GemmArgs {
  const float* A; int lda;
  const float* W; int ldw;
  const float* bias;
  int64_t M; int N, K;
  float* Y; int ldy;
  int mode;
  float* S; int lds;
  const float* aux; int ldaux; int64_t m_val;
  int npad, n_kc;
};

template <bool kF16>
__device__ __forceinline__ uint4 pack8(const float4& a, const float4& b) {
  uint4 r;
  r.x = umma::pack2<kF16>(a.x, a.y); r.y = umma::pack2<kF16>(a.z, a.w);
  r.z = umma::pack2<kF16>(b.x, b.y); r.w = umma::pack2<kF16>(b.z, b.w);
  return r;
}
// 8 consecutive floats of a row with `valid` readable elements from p (16-byte aligned), zero beyond
__device__ __forceinline__ void load8(const float* p, int valid, float4& a, float4& b) {
  a = make_float4(0.f, 0.f, 0.f, 0.f); b = a;
  if (valid >= 8) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); return; }
  float t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = j < valid ? p[j] : 0.f;
  a = make_float4(t[0], t[1], t[2], t[3]); b = make_float4(t[4], t[5], t[6], t[7]);
}

// 64 consecutive floats of a row -> eight 16-byte chunks of 8 converted values each.  Fast path (everything readable):
// all 16 vector loads are issued before the first conversion, so a thread keeps 16 requests in flight instead of 2 --
// the loaders are latency-bound otherwise (measured: 15x).  `valid` = readable elements from p.
template <bool kF16>
__device__ __forceinline__ void load64(const float* p, int valid, uint4 (&out)[8]) {
  if (valid >= 64) {
    float4 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = *reinterpret_cast<const float4*>(p + 4 * j);
#pragma unroll
    for (int j = 0; j < 8; ++j) out[j] = pack8<kF16>(v[2 * j], v[2 * j + 1]);
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 a, b;
      load8(p + 8 * j, valid - 8 * j, a, b);
      out[j] = pack8<kF16>(a, b);
    }
  }
}

template <bool kF16>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tc_kernel(const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t w_chunk_bytes = (uint32_t)g.npad * 128u;
  uint8_t* sW = smem;                                        // n_kc chunks of [npad x 64 k]
  uint8_t* sA = smem + (size_t)g.n_kc * w_chunk_bytes;        // kAStages x 16 KB (w_chunk_bytes is a multiple of 2 KB)
  uint64_t* bars = (uint64_t*)(sA + kAStages * kAStageBytes);
  uint64_t* a_full = bars;                  // [kAStages] 128 loader threads
  uint64_t* a_empty = bars + kAStages;      // [kAStages] MMA commit
  uint64_t* acc_ready = bars + 2 * kAStages;   // [2] MMA commit
  uint64_t* acc_free = acc_ready + 2;          // [2] 4 epilogue warps
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int64_t n_tiles = (g.M + kBM - 1) / kBM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) { umma::mbar_init(&a_full[s], 128); umma::mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { umma::mbar_init(&acc_ready[b], 1); umma::mbar_init(&acc_free[b], 8); }
    umma::fence_barrier_init();
  }
  if (warp == 8) { umma::tmem_alloc(&tmem_base_s, 512); umma::tmem_relinquish(); }
  // W -> shared memory, K-major 128-byte swizzle, rows >= N and columns >= K zero
  // (four iterations' loads in flight per thread: every CTA converts W itself, ~15 dependent L2 round trips per thread
  // when the loop is not unrolled -- a fifth of a 512-tile launch)
  {
    const int total = g.n_kc * g.npad * 8;
    for (int base = threadIdx.x; base < total; base += 4 * blockDim.x) {
      float4 a[4], b[4];
      int off[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = base + u * blockDim.x;
        off[u] = -1;
        a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < total) {
          const int c8 = idx & 7, n = (idx >> 3) % g.npad, kc = (idx >> 3) / g.npad;
          const int k0 = kc * kKC + c8 * 8;
          load8(g.W + (size_t)n * g.ldw + k0, n < g.N ? g.K - k0 : 0, a[u], b[u]);
          off[u] = kc * (int)w_chunk_bytes + (n >> 3) * 1024 + (n & 7) * 128 + ((c8 ^ (n & 7)) << 4);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (off[u] >= 0) *reinterpret_cast<uint4*>(sW + off[u]) = pack8<kF16>(a[u], b[u]);
    }
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 8) {
    // ===================== A loaders: thread = row of the tile; group 0 / 1 = even / odd chunks of the stream ==========
    const int r = threadIdx.x & 127;
    const uint32_t grp = (uint32_t)warp >> 2;
    uint32_t cnt = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row = tile * kBM + r;
      const float* arow = g.A + (size_t)row * g.lda;
      for (int kc = 0; kc < g.n_kc; ++kc, ++cnt) {
        if ((cnt & 1u) != grp) continue;
        const uint32_t st = cnt % kAStages, ph = (cnt / kAStages) & 1u;
        uint8_t* dst = sA + st * kAStageBytes + r * 128;
        uint4 ch[8];
        load64<kF16>(arow + kc * kKC, row < g.M ? g.K - kc * kKC : 0, ch);    // global loads first, then the slot
        umma::mbar_wait(&a_empty[st], ph ^ 1u);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) *reinterpret_cast<uint4*>(dst + ((c8 ^ (r & 7)) << 4)) = ch[c8];
        umma::fence_proxy_async_smem();
        umma::mbar_arrive(&a_full[st]);
      }
    }
  } else if (warp == 8) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = kF16 ? umma::make_idesc_f16(128, g.npad, 0, 0) : umma::make_idesc_bf16(128, g.npad, 0, 0);
    const uint32_t hi = umma::smem_desc_hi(1024);
    const uint32_t w_lo0 = umma::smem_desc_lo(umma::smem_u32(sW), 16), a_lo0 = umma::smem_desc_lo(umma::smem_u32(sA), 16);
    uint32_t cnt = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      umma::mbar_wait(&acc_free[buf], ((it >> 1) & 1u) ^ 1u);
      umma::tc_fence_after();
      const uint32_t d_addr = tmem_base + buf * 256u;
      for (int kc = 0; kc < g.n_kc; ++kc, ++cnt) {
        const uint32_t st = cnt % kAStages, ph = (cnt / kAStages) & 1u;
        umma::mbar_wait(&a_full[st], ph);
        umma::tc_fence_after();
        const uint32_t a_lo = a_lo0 + st * (kAStageBytes >> 4), w_lo = w_lo0 + (uint32_t)kc * (w_chunk_bytes >> 4);
        if (umma::elect_one()) {
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 2 * ks, hi), umma::desc64(w_lo + 2 * ks, hi), idesc, (kc | (int)ks) ? 1u : 0u);
          umma::mma_commit(&a_empty[st]);
        }
        __syncwarp();
      }
      if (umma::elect_one()) umma::mma_commit(&acc_ready[buf]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue: warps 9..16, TMEM lane quarter = warp % 4, column half = (warp - 9) / 4 ==========
    const int q = warp & 3;
    const int half = (warp - 9) >> 2;
    const int c_lo = half ? ((g.npad >> 1) + 15) & ~15 : 0, c_hi = half ? g.npad : ((g.npad >> 1) + 15) & ~15;
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      umma::mbar_wait(&acc_ready[buf], (it >> 1) & 1u);
      umma::tc_fence_after();
      const int64_t row = tile * kBM + 32 * q + lane;
      const bool rok = row < g.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + buf * 256u;
      float* yrow = g.Y + (size_t)row * g.ldy;
      float* srow = g.S ? g.S + (size_t)row * g.lds : nullptr;
      const float* xrow = g.mode == M_TANGENT ? g.aux + (size_t)(row % g.m_val) * g.ldaux : nullptr;
      const int n_out = (g.N + 3) & ~3;
      // tangent scale: 16-byte loads (ldaux >= pad4(N), rows 16-byte aligned), the NEXT 16 columns' requested while the
      // current ones are processed -- a load issued where it is used stalls the thread for an L2 round trip per 4 outputs
      float4 xn[4];
      auto fetch_scale = [&](int c0) {
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int col = c0 + 4 * j4;
          xn[j4] = (xrow && rok && col < n_out) ? __ldg(reinterpret_cast<const float4*>(xrow + col)) : make_float4(1.f, 1.f, 1.f, 1.f);
        }
      };
      fetch_scale(c_lo);
      for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
        uint32_t raw[16];
        umma::tmem_ld16(taddr + c0, raw);
        float4 xc[4];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) xc[j4] = xn[j4];
        if (c0 + 16 < c_hi) fetch_scale(c0 + 16);
        umma::tmem_ld_wait();
        if (!rok) continue;
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const int col = c0 + 4 * j4;
          if (col >= n_out) break;     // columns [N, pad4(N)) are written as zeros, nothing beyond (Y may be a column slice)
          float v[4], sp[4];
          const float xv[4] = {xc[j4].x, xc[j4].y, xc[j4].z, xc[j4].w};
          float bv[4] = {0.f, 0.f, 0.f, 0.f};
          if (g.bias) {
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = col + j < g.N ? __ldg(g.bias + col + j) : 0.0f;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float z = __uint_as_float(raw[4 * j4 + j]);
            const int cj = col + j;
            sp[j] = 0.f;
            if (cj < g.N) {
              switch (g.mode) {
                case M_NONE: z += bv[j]; break;
                case M_SOFTPLUS: {
                  // 16-bit tiers only (the fp32 tier runs nr_gemm_f32): hardware exponential / logarithm, 2 ulp
                  z += bv[j];
                  const float t = 100.0f * z, e = __expf(-fabsf(t));
                  const float r = __fdividef(1.0f, 1.0f + e);
                  sp[j] = t >= 0.0f ? r : e * r;
                  z = fmaxf(z, 0.0f) + 0.01f * __logf(1.0f + e);
                  break;
                }
                case M_RELU: z = fmaxf(z + bv[j], 0.0f); break;
                case M_SIGMOID: z = __fdividef(1.0f, 1.0f + __expf(-(z + bv[j]))); break;
                case M_TANGENT: z *= xv[j]; break;
                default: break;
              }
            } else {
              z = 0.f;
            }
            v[j] = z;
          }
          *reinterpret_cast<float4*>(yrow + col) = make_float4(v[0], v[1], v[2], v[3]);
          if (srow && g.mode == M_SOFTPLUS) {
            *reinterpret_cast<float4*>(srow + col) = make_float4(sp[0], sp[1], sp[2], sp[3]);   // lds >= pad4(N)
          }
        }
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&acc_free[buf]);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 8) umma::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// dW[N, K] += G[rows, N]^T X[rows, K]
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTnStages = 3;
constexpr uint32_t kTnABytes = 64 * 256 * 2;    // [64 rows x 256 cols]  32 KB: both 128-row M-tiles of dW, four blocks 8 KB apart
constexpr uint32_t kTnBBytes = 64 * 256 * 2;    // [64 rows x 256 cols]  32 KB, four blocks 8 KB apart
constexpr uint32_t kTnLbo = 8192;

struct TnArgs {
  const float* G; int ldg;
  const float* X; int ldx;
  int64_t rows; int N, K;
  float* dW; int lddw;
  int64_t chunks_per_slice;
  int vec4;
};

template <bool kF16>
__global__ void __launch_bounds__(288, 1) gemm_tn_tc_kernel(const TnArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // kTnStages x 32 KB
  uint8_t* sB = smem + kTnStages * kTnABytes;           // kTnStages x 32 KB
  uint64_t* bars = (uint64_t*)(sB + kTnStages * kTnBBytes);
  uint64_t* full = bars;                 // [kTnStages] 128 loader threads
  uint64_t* empty = bars + kTnStages;    // [kTnStages] MMA commit
  uint64_t* acc_ready = bars + 2 * kTnStages;
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * 256, n0 = blockIdx.z * 256;
  const bool two_mt = m0 + 128 < g.N;                               // second 128-row M-tile of dW present
  const int ncols = min(256, (g.K - n0 + 15) / 16 * 16);            // MMA N of this CTA
  const int64_t n_chunks = (g.rows + 63) / 64;
  const int64_t c_begin = blockIdx.x * g.chunks_per_slice;
  const int64_t c_end = min(n_chunks, c_begin + g.chunks_per_slice);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTnStages; ++s) { umma::mbar_init(&full[s], 128); umma::mbar_init(&empty[s], 1); }
    umma::mbar_init(acc_ready, 1);
    umma::fence_barrier_init();
  }
  if (warp == 4) { umma::tmem_alloc(&tmem_base_s, 512); umma::tmem_relinquish(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (c_begin >= c_end) {                 // empty slice (uniform for the CTA)
    __syncthreads();
    if (warp == 4) umma::tmem_dealloc(tmem_base, 512);
    return;
  }

  if (warp < 4) {
    // ===================== loaders: thread = (row of the chunk, column half) =====================
    const int r = threadIdx.x & 63, h = threadIdx.x >> 6;
    uint32_t cnt = 0;
    for (int64_t c = c_begin; c < c_end; ++c, ++cnt) {
      const uint32_t st = cnt % kTnStages, ph = (cnt / kTnStages) & 1u;
      umma::mbar_wait(&empty[st], ph ^ 1u);
      const int64_t row = c * 64 + r;
      const bool rok = row < g.rows;
      const float* grow = g.G + (size_t)row * g.ldg;
      const float* xrow = g.X + (size_t)row * g.ldx;
      uint8_t* a_dst = sA + st * kTnABytes;
      uint8_t* b_dst = sB + st * kTnBBytes;
      // G: 128 of the <= 256 M columns, X: 128 of the <= 256 N columns, 64 floats at a time
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        if (h == 1 && !two_mt) break;
        const int col = m0 + 128 * h + 64 * half;
        uint4 ch[8];
        load64<kF16>(grow + col, rok ? g.N - col : 0, ch);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) *reinterpret_cast<uint4*>(a_dst + umma::b_chunk_offset(r, 16 * h + 8 * half + c8, kTnLbo)) = ch[c8];
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int cb = 128 * h + 64 * half;
        if (cb >= ncols) break;
        const int col = n0 + cb;
        uint4 ch[8];
        load64<kF16>(xrow + col, rok ? g.K - col : 0, ch);
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) *reinterpret_cast<uint4*>(b_dst + umma::b_chunk_offset(r, 16 * h + 8 * half + c8, kTnLbo)) = ch[c8];
      }
      umma::fence_proxy_async_smem();
      umma::mbar_arrive(&full[st]);
    }
  } else if (warp == 4) {
    // ===================== MMA issuer: D[128, ncols] += G^T X over the slice's rows =====================
    const uint32_t idesc = kF16 ? umma::make_idesc_f16(128, ncols, 1, 1) : umma::make_idesc_bf16(128, ncols, 1, 1);
    const uint32_t hi = umma::smem_desc_hi(1024);
    const uint32_t a_lo0 = umma::smem_desc_lo(umma::smem_u32(sA), kTnLbo), b_lo0 = umma::smem_desc_lo(umma::smem_u32(sB), kTnLbo);
    uint32_t cnt = 0;
    for (int64_t c = c_begin; c < c_end; ++c, ++cnt) {
      const uint32_t st = cnt % kTnStages, ph = (cnt / kTnStages) & 1u;
      umma::mbar_wait(&full[st], ph);
      umma::tc_fence_after();
      const uint32_t a_lo = a_lo0 + st * (kTnABytes >> 4), b_lo = b_lo0 + st * (kTnBBytes >> 4);
      if (umma::elect_one()) {
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {    // 16 rows = two 8-row groups = 2048 bytes per k-step
          umma::mma_bf16_ss(tmem_base, umma::desc64(a_lo + 128 * ks, hi), umma::desc64(b_lo + 128 * ks, hi), idesc, (cnt | ks) ? 1u : 0u);
          if (two_mt)                            // M-tile 1 = column blocks 2, 3 of the A stage (+16 KB), accumulator +256
            umma::mma_bf16_ss(tmem_base + 256u, umma::desc64(a_lo + 1024 + 128 * ks, hi), umma::desc64(b_lo + 128 * ks, hi), idesc,
                              (cnt | ks) ? 1u : 0u);
        }
        umma::mma_commit(&empty[st]);
      }
      __syncwarp();
    }
    if (umma::elect_one()) umma::mma_commit(acc_ready);
    __syncwarp();
  } else {
    // ===================== epilogue: fp32 atomics of the partial sum =====================
    const int q = warp & 3;
    umma::mbar_wait(acc_ready, 0);
    umma::tc_fence_after();
    for (int mt = 0; mt < (two_mt ? 2 : 1); ++mt) {
      const int orow = m0 + 128 * mt + 32 * q + lane;
      float* drow = g.dW + (size_t)orow * g.lddw + n0;
      for (int c0 = 0; c0 < ncols; c0 += 16) {
        uint32_t raw[16];
        umma::tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + 256u * mt + c0, raw);
        umma::tmem_ld_wait();
        if (orow < g.N) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int col = n0 + c0 + 4 * j4;
            if (col + 4 <= g.lddw && g.vec4) {        // 16-byte vector reduction (pad columns receive + 0)
              atomicAdd(reinterpret_cast<float4*>(drow + c0 + 4 * j4),
                        make_float4(__uint_as_float(raw[4 * j4]), __uint_as_float(raw[4 * j4 + 1]), __uint_as_float(raw[4 * j4 + 2]),
                                    __uint_as_float(raw[4 * j4 + 3])));
            } else {
              for (int j = 0; j < 4; ++j)
                if (col + j < g.K) atomicAdd(drow + c0 + 4 * j4 + j, __uint_as_float(raw[4 * j4 + j]));
            }
          }
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 4) umma::tmem_dealloc(tmem_base, 512);
}

}  // namespace

extern "C" int nr_gemm_tc(const float* A, int32_t lda, const float* W, int32_t ldw, const float* bias, int64_t M, int32_t N,
                          int32_t K, float* Y, int32_t ldy, int32_t mode, float* S, int32_t lds, const float* aux,
                          int32_t ldaux, int64_t m_val, int32_t operand_f16, void* stream) {
  NR_CHECK_ARG(A && W && Y && M >= 0 && N >= 1 && K >= 1, "nr_gemm_tc: bad arguments");
  NR_CHECK_ARG(mode >= M_NONE && mode <= M_LINEAR, "nr_gemm_tc: mode=%d", mode);
  NR_CHECK_ARG((lda & 3) == 0 && (ldw & 3) == 0 && (ldy & 3) == 0 && lda >= K && ldw >= K && ldy >= N,
               "nr_gemm_tc: leading dimensions must be multiples of 4 and cover the matrix");
  NR_CHECK_ARG((((uintptr_t)A | (uintptr_t)W | (uintptr_t)Y | (uintptr_t)S) & 15) == 0, "nr_gemm_tc: 16-byte alignment");
  NR_CHECK_ARG(mode == M_LINEAR || mode == M_TANGENT || bias, "nr_gemm_tc: bias required");
  NR_CHECK_ARG(mode != M_TANGENT || (aux && m_val > 0 && ldaux >= ((N + 3) & ~3) && ldaux % 4 == 0 && ((uintptr_t)aux & 15) == 0),
               "nr_gemm_tc: tangent mode needs aux with 16-byte aligned rows of at least pad4(N) floats");
  NR_CHECK_ARG(!S || ((lds & 3) == 0 && lds >= ((N + 3) & ~3)), "nr_gemm_tc: lds must be a multiple of 4 covering pad4(N)");
  NR_CHECK_ARG(ldy >= ((N + 3) & ~3), "nr_gemm_tc: ldy must cover pad4(N)");
  if (M == 0) return NR_OK;
  GemmArgs g{A, lda, W, ldw, bias, M, N, K, Y, ldy, mode, S, lds, aux, ldaux, m_val, 0, 0};
  g.npad = (N + 15) / 16 * 16;
  g.n_kc = (K + kKC - 1) / kKC;
  NR_CHECK_ARG(g.npad <= 256, "nr_gemm_tc: N=%d > 256", N);
  const size_t smem = 1024 + (size_t)g.n_kc * g.npad * 128 + kAStages * kAStageBytes + 256;
  NR_CHECK_ARG(smem <= 227 * 1024, "nr_gemm_tc: W (%d x %d) does not fit in shared memory", N, K);
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = nr_cdiv(M, kBM);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  if (operand_f16) {
    NR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tc_kernel<true><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(g);
  } else {
    NR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tc_kernel<false><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(g);
  }
  NR_CHECK_LAUNCH("gemm_tc_kernel");
  return NR_OK;
}

extern "C" int nr_gemm_tn_tc(const float* G, int32_t ldg, const float* X, int32_t ldx, int64_t rows, int32_t N, int32_t K,
                             float* dW, int32_t lddw, int32_t operand_f16, void* stream) {
  NR_CHECK_ARG(G && X && dW && rows >= 0 && N >= 1 && K >= 1, "nr_gemm_tn_tc: bad arguments");
  NR_CHECK_ARG((ldg & 3) == 0 && (ldx & 3) == 0 && ldg >= N && ldx >= K && lddw >= K, "nr_gemm_tn_tc: leading dimensions");
  NR_CHECK_ARG((((uintptr_t)G | (uintptr_t)X) & 15) == 0, "nr_gemm_tn_tc: 16-byte alignment");
  if (rows == 0) return NR_OK;
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n_mt = (N + 255) / 256, n_nt = (K + 255) / 256;
  const int64_t n_chunks = nr_cdiv(rows, 64);
  int64_t slices = sms / (n_mt * n_nt);       // one CTA per SM: the fp32 reduction of the partial sums is the cost to amortise
  if (slices < 1) slices = 1;
  if (slices > n_chunks) slices = n_chunks;
  TnArgs g{G, ldg, X, ldx, rows, N, K, dW, lddw, nr_cdiv(n_chunks, slices), ((((uintptr_t)dW) & 15) == 0 && (lddw & 3) == 0) ? 1 : 0};
  slices = nr_cdiv(n_chunks, g.chunks_per_slice);
  const size_t smem = 1024 + kTnStages * (kTnABytes + kTnBBytes) + 128;
  dim3 grid((unsigned)slices, n_mt, n_nt);
  if (operand_f16) {
    NR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tn_tc_kernel<true><<<grid, 288, smem, (cudaStream_t)stream>>>(g);
  } else {
    NR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tn_tc_kernel<false><<<grid, 288, smem, (cudaStream_t)stream>>>(g);
  }
  NR_CHECK_LAUNCH("gemm_tn_tc_kernel");
  return NR_OK;
}
