// Training GEMMs with 16-BIT TENSORS IN HBM (round 2): the per-layer building blocks of the reverse-mode training path
// of the SDF network (models/autograd_rev.py).  Same machine as gemm_tc.cu (W resident in shared memory, A tiles through a
// 3-stage ring, tcgen05 MMAs with the accumulator double-buffered in TMEM, 8 epilogue warps), but activations, softplus
// derivatives and gradients live in HBM as fp16 rows: the loaders copy 128-byte row pieces straight into the swizzled
// operand (no conversion), the epilogues read their auxiliary rows and write their results as fp16.  Round 1's path kept
// fp32 rows and pushed three tangent rows per point through every layer: 12 n rows of fp32 per layer against 6 n rows of
// fp16 here -- a quarter of the bytes on kernels that run at the HBM roofline.
//
// nr_gemm16:     Y[M, N] = epilogue(A[M, K] W[N, K]^T)     A fp16 rows (lda halves, a multiple of 64 covering K rounded
//                up to 64, pad columns finite: they meet zero weights), W fp32 [N, ldw] (converted per CTA), Y fp16 or fp32
//   G_LINEAR     y = acc + bias
//   G_SOFTPLUS   y = softplus100(acc + bias),  out2 = softplus100'                                   (forward sweep)
//   G_SCALE      y = aux_a * acc (+ aux_b)                                      (reverse sweep; backprop with the 2nd-order addend)
//   G_ADJ        y = aux_a * acc,  out2 = 100 (1 - aux_a) * aux_b * acc         (adjoint of the reverse sweep: aux_a = softplus',
//                aux_b = p = softplus' * g of the reverse sweep;  out2 = softplus'' * g * acc without a division)
//   G_RELU / G_SIGMOID / G_MASK (y = aux_a > 0 ? acc : 0)                                            (radiance net)
// nr_gemm16_tn:  dW[N, K] += scale * G[rows, N]^T X[rows, K]    G, X fp16 rows; fp32 atomics of the CTA's partial sum
// nr_colsum16:   out[N] += scale * column sums of a fp16 matrix
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cuda_fp16.h>

namespace {

// cuTensorMapEncodeTiled through the runtime's driver entry point query: the library must not link against libcuda.so.1
// (it is loaded, and its symbols are checked, on machines without a driver -- the CPU half of the test-suite)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}
// 2-D tensor map of a row-major fp16 matrix [rows, ld], box [box_rows x 64 columns], 128-byte swizzle, zero fill
int make_map(CUtensorMap* map, const void* base, int64_t rows, int ld, int box_rows, const char* who) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { nr_set_error("%s: cuTensorMapEncodeTiled is not available from this driver", who); return NR_ERR_CUDA; }
  const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows}, estr[2] = {1, 1};
  const CUresult rc = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    nr_set_error("%s: cuTensorMapEncodeTiled failed (%d) for %p [%lld x %d]", who, (int)rc, base, (long long)rows, ld);
    return NR_ERR_CUDA;
  }
  return NR_OK;
}

constexpr int kBM = 128, kKC = 64;
#ifndef NR_G16_STAGES
#define NR_G16_STAGES 6
#endif
#ifndef NR_G16_PRE
#define NR_G16_PRE 2
#endif
constexpr int kMaxAStages = NR_G16_STAGES;          // A ring: as many 16 KB stages as fit next to W (3 .. 6): bytes in flight are what a
                                        // latency-bound loader has
constexpr int kGemmThreads = 18 * 32;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-17: epilogue
constexpr int kEpiWarp0 = 2, kEpiWarps = 16;
constexpr uint32_t kAStageBytes = kBM * kKC * 2;   // 16 KB

enum : int { G_LINEAR = 0, G_SOFTPLUS = 1, G_SCALE = 2, G_ADJ = 3, G_RELU = 4, G_SIGMOID = 5, G_MASK = 6 };

struct G16Args {
  const __half* A; int lda;
  const float* W; int ldw;
  const uint8_t* Wimg;     // or: W already packed by nr_gemm16_pack_w (n_kc chunks of [npad x 64 k] fp16, swizzled)
  int a_stages;
  const float* bias;
  int64_t M; int N, K;
  void* Y; int ldy; int y_half;
  int mode;
  const __half* aux_a; int ld_a;
  const __half* aux_b; int ld_b;
  __half* out2; int ld_o2;
  int npad, n_kc;
  // split-precision forward GEMMs (nr_gemm16_split): A = [hi | lo], split_p = P 64-column chunks each; the W image of a
  // column block holds [W_hi (P chunks) | W_lo (P chunks)]; A chunk j < P (hi) multiplies W_hi chunk j AND W_lo chunk j,
  // A chunk P + j (lo) multiplies W_hi chunk j -- (hi + lo)(W_hi + W_lo)^T without the lo x lo term, every A chunk loaded
  // once.  blockIdx.y = block of `npad` output columns with its own W image, lo_off = column offset of the lo part of a
  // fp16 result (0: none)
  int split_p, lo_off, w_block_bytes;
};

__device__ __forceinline__ uint4 pack8h(const float4& a, const float4& b) {
  uint4 r;
  r.x = umma::pack_f16(a.x, a.y); r.y = umma::pack_f16(a.z, a.w);
  r.z = umma::pack_f16(b.x, b.y); r.w = umma::pack_f16(b.z, b.w);
  return r;
}
__device__ __forceinline__ void load8(const float* p, int valid, float4& a, float4& b) {
  a = make_float4(0.f, 0.f, 0.f, 0.f); b = a;
  if (valid >= 8) { a = *reinterpret_cast<const float4*>(p); b = *reinterpret_cast<const float4*>(p + 4); return; }
  float t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = j < valid ? p[j] : 0.f;
  a = make_float4(t[0], t[1], t[2], t[3]); b = make_float4(t[4], t[5], t[6], t[7]);
}
// 256-bit global accesses (sm_100: LDG.256 / STG.256): a thread's 16 fp16 columns are exactly one 32-byte sector, so a
// row-per-thread epilogue writes FULL sectors -- with two 16-byte stores per sector the write path ran at 1.15 TB/s
struct U8 { uint32_t r[8]; };
__device__ __forceinline__ U8 ldg256(const void* p) {
  U8 v;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v.r[0]), "=r"(v.r[1]), "=r"(v.r[2]), "=r"(v.r[3]), "=r"(v.r[4]), "=r"(v.r[5]), "=r"(v.r[6]), "=r"(v.r[7]) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256(void* p, const U8& v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.r[0]), "r"(v.r[1]), "r"(v.r[2]), "r"(v.r[3]),
               "r"(v.r[4]), "r"(v.r[5]), "r"(v.r[6]), "r"(v.r[7]) : "memory");
}
__device__ __forceinline__ void unpack16h(const U8& v, float (&f)[16]) {
  const __half2* h = reinterpret_cast<const __half2*>(v.r);
#pragma unroll
  for (int j = 0; j < 8; ++j) { const float2 t = __half22float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ U8 pack16f(const float (&f)[16]) {
  U8 v;
#pragma unroll
  for (int j = 0; j < 8; ++j) v.r[j] = umma::pack_f16(f[2 * j], f[2 * j + 1]);
  return v;
}
__device__ __forceinline__ void unpack8h(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int j = 0; j < 4; ++j) { const float2 t = __half22float2(h[j]); f[2 * j] = t.x; f[2 * j + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8f(const float (&f)[8]) {
  uint4 r;
  r.x = umma::pack_f16(f[0], f[1]); r.y = umma::pack_f16(f[2], f[3]); r.z = umma::pack_f16(f[4], f[5]); r.w = umma::pack_f16(f[6], f[7]);
  return r;
}

// one 2-D tile of a row-major fp16 matrix -> shared memory through the TMA unit (tensor map: box [128 rows x 64 columns],
// SWIZZLE_128B: exactly the K-major operand tile; rows past the matrix arrive as zeros)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int col, int row, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   umma::smem_u32(smem_dst)),
               "l"(map), "r"(col), "r"(row), "r"(umma::smem_u32(bar))
               : "memory");
}

template <int MODE, bool Y_HALF>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm16_kernel(const G16Args g, const __grid_constant__ CUtensorMap map_a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const uint32_t w_chunk_bytes = (uint32_t)g.npad * 128u;
  uint8_t* sW = smem;                                        // n_kc chunks of [npad x 64 k]
  const int kAStages = g.a_stages;
  uint8_t* sA = smem + (size_t)g.n_kc * w_chunk_bytes;        // kAStages x 16 KB
  uint64_t* bars = (uint64_t*)(sA + kAStages * kAStageBytes);
  uint64_t* a_full = bars;                  // [kAStages] TMA transaction bytes
  uint64_t* a_empty = bars + kMaxAStages;   // [kAStages] MMA commit
  uint64_t* acc_ready = bars + 2 * kMaxAStages;   // [2] MMA commit
  uint64_t* acc_free = acc_ready + 2;          // [2] 16 epilogue warps
  uint64_t* w_bar = acc_free + 2;              // packed W image landed
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int64_t n_tiles = (g.M + kBM - 1) / kBM;
  const int col0 = (int)blockIdx.y * g.npad;        // first output column of this CTA (column blocks: nr_gemm16_split only)

  if (threadIdx.x == 0) {
    for (int s = 0; s < kAStages; ++s) { umma::mbar_init(&a_full[s], 1); umma::mbar_init(&a_empty[s], 1); }
    for (int b = 0; b < 2; ++b) { umma::mbar_init(&acc_ready[b], 1); umma::mbar_init(&acc_free[b], kEpiWarps); }
    umma::mbar_init(w_bar, 1);
    umma::fence_barrier_init();
    if (g.Wimg) {                          // the packed image: one bulk copy per 64-k chunk
      const uint32_t bytes = (uint32_t)g.n_kc * w_chunk_bytes;
      const uint8_t* img = g.Wimg + (size_t)blockIdx.y * g.w_block_bytes;
      umma::mbar_arrive_expect_tx(w_bar, bytes);
      for (int kc = 0; kc < g.n_kc; ++kc)
        umma::bulk_g2s(sW + (size_t)kc * w_chunk_bytes, img + (size_t)kc * w_chunk_bytes, w_chunk_bytes, w_bar);
    }
  }
  if (warp == 1) { umma::tmem_alloc(&tmem_base_s, 512); umma::tmem_relinquish(); }
  // W -> shared memory as fp16, K-major 128-byte swizzle, rows >= N and columns >= K zero
  if (!g.Wimg) {
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
        if (off[u] >= 0) *reinterpret_cast<uint4*>(sW + off[u]) = pack8h(a[u], b[u]);
    }
  }
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== A producer: one thread, one TMA tile load per 64-k chunk =====================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n_a = g.split_p ? 2 * g.split_p : g.n_kc;       // A chunks per tile
        for (int kc = 0; kc < n_a; ++kc, ++cnt) {
          const uint32_t st = cnt % kAStages, ph = (cnt / kAStages) & 1u;
          umma::mbar_wait(&a_empty[st], ph ^ 1u);
          umma::mbar_arrive_expect_tx(&a_full[st], kAStageBytes);
          tma_load_2d(sA + st * kAStageBytes, &map_a, kc * kKC, (int)(tile * kBM), &a_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma::make_idesc_f16(128, g.npad, 0, 0);
    const uint32_t hi = umma::smem_desc_hi(1024);
    const uint32_t w_lo0 = umma::smem_desc_lo(umma::smem_u32(sW), 16), a_lo0 = umma::smem_desc_lo(umma::smem_u32(sA), 16);
    uint32_t cnt = 0, it = 0;
    if (g.Wimg) { umma::mbar_wait(w_bar, 0); umma::tc_fence_after(); }
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      umma::mbar_wait(&acc_free[buf], ((it >> 1) & 1u) ^ 1u);
      umma::tc_fence_after();
      const uint32_t d_addr = tmem_base + buf * 256u;
      const int P = g.split_p, n_a = P ? 2 * P : g.n_kc;
      for (int kc = 0; kc < n_a; ++kc, ++cnt) {
        const uint32_t st = cnt % kAStages, ph = (cnt / kAStages) & 1u;
        umma::mbar_wait(&a_full[st], ph);
        umma::tc_fence_after();
        const uint32_t a_lo = a_lo0 + st * (kAStageBytes >> 4);
        const int wc = (P && kc >= P) ? kc - P : kc;          // split: a lo chunk meets W_hi
        const uint32_t w_lo = w_lo0 + (uint32_t)wc * (w_chunk_bytes >> 4);
        if (umma::elect_one()) {
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 2 * ks, hi), umma::desc64(w_lo + 2 * ks, hi), idesc, (kc | (int)ks) ? 1u : 0u);
          if (P && kc < P) {                                  // split: a hi chunk also meets W_lo
            const uint32_t w2 = w_lo0 + (uint32_t)(P + kc) * (w_chunk_bytes >> 4);
#pragma unroll
            for (uint32_t ks = 0; ks < 4; ++ks)
              umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 2 * ks, hi), umma::desc64(w2 + 2 * ks, hi), idesc, 1u);
          }
          umma::mma_commit(&a_empty[st]);
        }
        __syncwarp();
      }
      if (umma::elect_one()) umma::mma_commit(&acc_ready[buf]);
      __syncwarp();
    }
  } else {
    // ===================== epilogue: warps 2..17, TMEM lane quarter = warp % 4, column quarter = (warp - 2) / 4 ==========
    // thread = one row of the tile, 16 columns (one 32-byte sector of every fp16 row) at a time.  Sixteen warps, not
    // eight: the epilogue is a chain of long-latency steps (TMEM load, auxiliary rows from HBM, MUFU), and it is the number
    // of warps in flight that hides them -- the loaders' eight warps are gone since the TMA unit fetches A
    const int q = warp & 3;
    const int part = (warp - kEpiWarp0) >> 2;
    const int per = ((g.npad + 3) / 4 + 15) & ~15;                 // columns per quarter, a multiple of 16
    const int c_lo = min(part * per, g.npad), c_hi = min(c_lo + per, g.npad);
    const int n16 = min(g.npad, ((g.N + 15) & ~15) - col0);        // columns (of this block) written: [N, n16) as zeros
    uint32_t it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t buf = it & 1u;
      umma::mbar_wait(&acc_ready[buf], (it >> 1) & 1u);
      umma::tc_fence_after();
      const int64_t row = tile * kBM + 32 * q + lane;
      const bool rok = row < g.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + buf * 256u;
      const __half* arow = g.aux_a ? g.aux_a + (size_t)row * g.ld_a + col0 : nullptr;
      const __half* brow = g.aux_b ? g.aux_b + (size_t)row * g.ld_b + col0 : nullptr;
      __half* orow = g.out2 ? g.out2 + (size_t)row * g.ld_o2 + col0 : nullptr;
      // auxiliary rows: kPre iterations (16 columns each) requested ahead of their use -- a thread that asks for 32 bytes and
      // waits an HBM round trip for them moves 5 GB/s per SM; this is what bounded the first version of these kernels
      U8 an[2], bn[2];                           // two prefetch slots (16 columns = 32 bytes each), compile-time indices only
      auto fetch = [&](int slot, int c0) {
        const bool ok = rok && c0 < n16 && c0 < c_hi;
        U8 zero = {};
        an[slot] = (arow && ok) ? ldg256(arow + c0) : zero;
        bn[slot] = (brow && ok) ? ldg256(brow + c0) : zero;
      };
      auto step = [&](int slot, int c0) {        // 16 columns: accumulators from TMEM, auxiliaries from slot, refill the slot
        if (c0 >= c_hi) return;
        uint32_t raw[16];
        umma::tmem_ld16(taddr + c0, raw);
        const U8 ac = an[slot], bc = bn[slot];
        fetch(slot, c0 + 32);
        umma::tmem_ld_wait();
        if (!rok || c0 >= n16) return;
        float av[16], bv[16], y[16], o2[16];
        unpack16h(ac, av);
        unpack16h(bc, bv);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int cj = col0 + c0 + j;
          float z = __uint_as_float(raw[j]);
          o2[j] = 0.f;
          if (cj < g.N) {
            const float b = (MODE == G_LINEAR || MODE == G_SOFTPLUS || MODE == G_RELU || MODE == G_SIGMOID) && g.bias ? __ldg(g.bias + cj) : 0.0f;
            if (MODE == G_LINEAR) {
              z += b;
            } else if (MODE == G_SOFTPLUS) {
              z += b;
              const float t = 100.0f * z, e = __expf(-fabsf(t));
              const float r = __fdividef(1.0f, 1.0f + e);
              o2[j] = t >= 0.0f ? r : e * r;
              z = fmaxf(z, 0.0f) + 0.01f * __logf(1.0f + e);
            } else if (MODE == G_SCALE) {
              z = av[j] * z + (brow ? bv[j] : 0.0f);
            } else if (MODE == G_ADJ) {
              o2[j] = 100.0f * (1.0f - av[j]) * bv[j] * z;
              z = av[j] * z;
            } else if (MODE == G_RELU) {
              z = fmaxf(z + b, 0.0f);
            } else if (MODE == G_SIGMOID) {
              z = __fdividef(1.0f, 1.0f + __expf(-(z + b)));
            } else {
              z = av[j] > 0.0f ? z : 0.0f;
            }
          } else {
            z = 0.f;
          }
          y[j] = z;
        }
        if (Y_HALF) {
          __half* yrow = reinterpret_cast<__half*>(g.Y) + (size_t)row * g.ldy + col0 + c0;
          const U8 hi = pack16f(y);
          stg256(yrow, hi);
          if (g.lo_off) {                       // split result: lo = fp16(y - hi) next to it
            float hf[16], lo[16];
            unpack16h(hi, hf);
#pragma unroll
            for (int j = 0; j < 16; ++j) lo[j] = y[j] - hf[j];
            stg256(yrow + g.lo_off, pack16f(lo));
          }
        } else {
          float* yrow = reinterpret_cast<float*>(g.Y) + (size_t)row * g.ldy + col0 + c0;
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) *reinterpret_cast<float4*>(yrow + 4 * j4) = make_float4(y[4 * j4], y[4 * j4 + 1], y[4 * j4 + 2], y[4 * j4 + 3]);
        }
        if ((MODE == G_SOFTPLUS || MODE == G_ADJ) && orow) stg256(orow + c0, pack16f(o2));
      };
      // auxiliary rows are requested 32 columns ahead of their use: a thread that asks for 32 bytes and waits an HBM round
      // trip for them moves ~5 GB/s per SM
      fetch(0, c_lo);
      fetch(1, c_lo + 16);
#pragma unroll 1
      for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
        step(0, c0);
        step(1, c0 + 16);
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&acc_free[buf]);
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// dW[N, K] += scale * G[rows, N]^T X[rows, K]        G, X fp16 rows
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTnStages = 3;
constexpr uint32_t kTnABytes = 64 * 256 * 2;    // [64 rows x 256 cols]  32 KB: both 128-row M-tiles of dW, four blocks 8 KB apart
constexpr uint32_t kTnBBytes = 64 * 256 * 2;
constexpr uint32_t kTnLbo = 8192;

struct Tn16Args {
  const __half* G; int ldg;
  const __half* X; int ldx;
  int64_t rows; int N, K;
  float* dW; int lddw;
  int64_t chunks_per_slice;
  float scale;
  int vec4;
  int n_pairs;
};

// n_pairs = 2: dW += scale * (G^T X + G2^T X2) -- both products accumulate in the same TMEM tile and leave through ONE pass of
// atomics (the atomics of 146 partial sums are a third of this kernel's time: the SDF layers' dW = zb^T h + p^T gb as two
// launches paid them twice)
__global__ void __launch_bounds__(192, 1) gemm16_tn_kernel(const Tn16Args g, const __grid_constant__ CUtensorMap map_g,
                                                            const __grid_constant__ CUtensorMap map_x,
                                                            const __grid_constant__ CUtensorMap map_g2,
                                                            const __grid_constant__ CUtensorMap map_x2) {
  // warp 0: TMA producer (one thread), warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue (TMEM lane quarter = warp % 4)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kTnStages * kTnABytes;
  uint64_t* bars = (uint64_t*)(sB + kTnStages * kTnBBytes);
  uint64_t* full = bars;                 // [kTnStages] TMA transaction bytes
  uint64_t* empty = bars + kTnStages;    // [kTnStages] MMA commit
  uint64_t* acc_ready = bars + 2 * kTnStages;
  __shared__ uint32_t tmem_base_s;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * 256, n0 = blockIdx.z * 256;
  const bool two_mt = m0 + 128 < g.N;
  const int ncols = min(256, (g.K - n0 + 15) / 16 * 16);
  const int64_t n_chunks = (g.rows + 63) / 64;
  const int64_t c_begin = blockIdx.x * g.chunks_per_slice;
  const int64_t c_end = min(n_chunks, c_begin + g.chunks_per_slice);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTnStages; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1); }
    umma::mbar_init(acc_ready, 1);
    umma::fence_barrier_init();
  }
  if (warp == 1) { umma::tmem_alloc(&tmem_base_s, 512); umma::tmem_relinquish(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (c_begin >= c_end) {
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem_base, 512);
    return;
  }

  if (warp == 0) {
    // a 64-row chunk of G (the 128 / 256 columns of this CTA's M-tiles) and of X (ncols columns), as 64-column blocks:
    // a [64 rows x 64 columns] box with the 128-byte swizzle IS one block of the MN-major operand
    if (lane == 0) {
      const int g_blocks = two_mt ? 4 : 2, x_blocks = (ncols + 63) / 64;
      uint32_t cnt = 0;
      for (int pair = 0; pair < g.n_pairs; ++pair) {
        const CUtensorMap* mg = pair ? &map_g2 : &map_g;
        const CUtensorMap* mx = pair ? &map_x2 : &map_x;
        for (int64_t c = c_begin; c < c_end; ++c, ++cnt) {
          const uint32_t st = cnt % kTnStages, ph = (cnt / kTnStages) & 1u;
          umma::mbar_wait(&empty[st], ph ^ 1u);
          umma::mbar_arrive_expect_tx(&full[st], (uint32_t)(g_blocks + x_blocks) * kTnLbo);
          for (int b = 0; b < g_blocks; ++b) tma_load_2d(sA + st * kTnABytes + b * kTnLbo, mg, m0 + 64 * b, (int)(c * 64), &full[st]);
          for (int b = 0; b < x_blocks; ++b) tma_load_2d(sB + st * kTnBBytes + b * kTnLbo, mx, n0 + 64 * b, (int)(c * 64), &full[st]);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma::make_idesc_f16(128, ncols, 1, 1);
    const uint32_t hi = umma::smem_desc_hi(1024);
    const uint32_t a_lo0 = umma::smem_desc_lo(umma::smem_u32(sA), kTnLbo), b_lo0 = umma::smem_desc_lo(umma::smem_u32(sB), kTnLbo);
    uint32_t cnt = 0;
    const int64_t n_it = (c_end - c_begin) * g.n_pairs;
    for (int64_t c = 0; c < n_it; ++c, ++cnt) {
      const uint32_t st = cnt % kTnStages, ph = (cnt / kTnStages) & 1u;
      umma::mbar_wait(&full[st], ph);
      umma::tc_fence_after();
      const uint32_t a_lo = a_lo0 + st * (kTnABytes >> 4), b_lo = b_lo0 + st * (kTnBBytes >> 4);
      if (umma::elect_one()) {
#pragma unroll
        for (uint32_t ks = 0; ks < 4; ++ks) {
          umma::mma_bf16_ss(tmem_base, umma::desc64(a_lo + 128 * ks, hi), umma::desc64(b_lo + 128 * ks, hi), idesc, (cnt | ks) ? 1u : 0u);
          if (two_mt)
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
            if (g.vec4 && col + 4 <= g.lddw) {        // 16-byte vector reduction (pad columns of dW receive scale * 0)
              atomicAdd(reinterpret_cast<float4*>(drow + c0 + 4 * j4),
                        make_float4(g.scale * __uint_as_float(raw[4 * j4]), g.scale * __uint_as_float(raw[4 * j4 + 1]),
                                    g.scale * __uint_as_float(raw[4 * j4 + 2]), g.scale * __uint_as_float(raw[4 * j4 + 3])));
            } else {
              for (int j = 0; j < 4; ++j)
                if (col + j < g.K) atomicAdd(drow + c0 + 4 * j4 + j, g.scale * __uint_as_float(raw[4 * j4 + j]));
            }
          }
        }
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc(tmem_base, 512);
}

// W fp32 [N, ldw] -> the fp16 image nr_gemm16 keeps in shared memory: n_kc chunks of [npad rows x 64 k], K-major, 128-byte
// swizzle, rows >= N and columns >= K zero.  One launch per weight matrix and training step instead of one conversion
// per CTA and GEMM launch.
__global__ void gemm16_pack_w_kernel(const float* __restrict__ W, int ldw, int N, int K, int npad, int n_kc, uint8_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_kc * npad * 8) return;
  const int c8 = idx & 7, n = (idx >> 3) % npad, kc = (idx >> 3) / npad;
  const int k0 = kc * kKC + c8 * 8;
  float4 a, b;
  load8(W + (size_t)n * ldw + k0, n < N ? K - k0 : 0, a, b);
  *reinterpret_cast<uint4*>(img + (size_t)kc * npad * 128 + (n >> 3) * 1024 + (n & 7) * 128 + ((c8 ^ (n & 7)) << 4)) = pack8h(a, b);
}

// [W_hi | W_lo] of nr_gemm16_split, all column blocks of a weight matrix in one launch: image = n_blocks x
// (2 Kp / 64 chunks) x [128 rows x 64 k] fp16, swizzled; hi = fp16(w), lo = fp16(w - hi)
constexpr int kSplitBlock = 128;      // output columns per CTA of nr_gemm16_split
__global__ void gemm16_pack_w_split_kernel(const float* __restrict__ W, int ldw, int N, int K, int kp, int n_blocks,
                                           uint8_t* __restrict__ img) {
  const int n_kc = 2 * kp / kKC, per_part = kp / kKC;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_blocks * n_kc * kSplitBlock * 8) return;
  const int c8 = idx & 7, r = (idx >> 3) & (kSplitBlock - 1), kc = (idx >> 10) % n_kc, b = (idx >> 10) / n_kc;
  const int part = kc / per_part, k0 = (kc % per_part) * kKC + c8 * 8, n = kSplitBlock * b + r;
  float4 a4, b4;
  load8(W + (size_t)n * ldw + k0, n < N ? K - k0 : 0, a4, b4);
  float v[8] = {a4.x, a4.y, a4.z, a4.w, b4.x, b4.y, b4.z, b4.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    if (part == 0) {
      o[j] = *reinterpret_cast<const uint32_t*>(&h);
    } else {
      const float2 hf = __half22float2(h);
      o[j] = umma::pack_f16(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
    }
  }
  *reinterpret_cast<uint4*>(img + ((size_t)b * n_kc + kc) * (kSplitBlock * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4)) =
      make_uint4(o[0], o[1], o[2], o[3]);
}

// dst[r, c] = fp16(scale * src[r, c]) for a block of columns (fp32 rows -> the fp16 rows of the training GEMMs), optionally the
// lo part fp16(scale * src - hi) `lo_off` columns to the right: one pass instead of torch's mul + cast + strided copy
__global__ void cast_cols16_kernel(const float* __restrict__ src, int64_t ld_src, int64_t n, int ncols, float scale,
                                   __half* __restrict__ dst, int64_t ld_dst, int lo_off) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (idx >= n * ncols) return;
  const int64_t r = idx / ncols;
  const int c = (int)(idx - r * ncols);
  const float v = scale * src[r * ld_src + c];
  const __half h = __float2half_rn(v);
  dst[r * ld_dst + c] = h;
  if (lo_off) dst[r * ld_dst + lo_off + c] = __float2half_rn(v - __half2float(h));
}

// out[c] += scale * sum_r A[r, c]  (fp16 rows): block = 256 columns x 4 row lanes, grid over row slabs
__global__ void colsum16_kernel(const __half* __restrict__ A, int lda, int64_t rows, int N, float scale, float* __restrict__ out) {
  const int c = threadIdx.x & 255, rl = threadIdx.x >> 8;
  const int64_t r0 = (int64_t)blockIdx.x * 256, r1 = min(rows, r0 + 256);
  float acc = 0.0f;
  if (c < N)
    for (int64_t r = r0 + rl; r < r1; r += 4) acc += __half2float(A[(size_t)r * lda + c]);
  __shared__ float part[4][256];
  part[rl][c] = acc;
  __syncthreads();
  if (rl == 0 && c < N) atomicAdd(out + c, scale * (part[0][c] + part[1][c] + part[2][c] + part[3][c]));
}

// ---- embedding helpers of the reverse-mode path (Embedder.forward, base.py:46-64, and its Jacobian) ----------------
__device__ __forceinline__ void pe_and_jac(int j, int multires, const float* x3, float& val, int& comp, float& jac) {
  if (j < 3) { val = x3[j]; comp = j; jac = 1.0f; return; }
  const int qf = (j - 3) / 6, r = (j - 3) % 6;
  comp = r % 3;
  const float f = (float)(1 << qf);
  float s, c;
  sincosf(x3[comp] * f, &s, &c);
  val = r < 3 ? s : c;
  jac = r < 3 ? f * c : -f * s;
}
// e [n, ld] <- PE(x) as fp16 (columns [pe_dim, width) zero); optionally the same values into e2 at column offset off2
__global__ void pe16_kernel(const float* __restrict__ x, int64_t n, int multires, int pe_dim, __half* __restrict__ e, int ld,
                            int width, __half* __restrict__ e2, int ld2, int off2, int lo_off, int lo_off2) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int64_t p = i / width;
  const int j = (int)(i - p * width);
  float v = 0.0f;
  if (j < pe_dim) {
    const float x3[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    int comp; float jac;
    pe_and_jac(j, multires, x3, v, comp, jac);
  }
  const __half h = __float2half_rn(v);
  e[(size_t)p * ld + j] = h;
  if (e2 && j < pe_dim) e2[(size_t)p * ld2 + off2 + j] = h;
  if (lo_off) {                            // split-precision rows: lo = fp16(v - hi), lo_off columns to the right
    const __half l = __float2half_rn(v - __half2float(h));
    e[(size_t)p * ld + lo_off + j] = l;
    if (e2 && j < pe_dim) e2[(size_t)p * ld2 + lo_off2 + off2 + j] = l;
  }
}
// nabla[p, c] = sum_j dPE_j/dx_c (g0[p, j] + ge[p, j])          (g0 fp32 [n, ldg0]; ge fp16 view or NULL)
__global__ void pe_jac_t_kernel(const float* __restrict__ x, int64_t n, int multires, int pe_dim, const float* __restrict__ g0,
                                int ldg0, const __half* __restrict__ ge, int ldge, int ge_lo_off, float* __restrict__ nabla) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float x3[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
  float acc[3] = {0.f, 0.f, 0.f};
  for (int j = 0; j < pe_dim; ++j) {
    float v, jac; int comp;
    pe_and_jac(j, multires, x3, v, comp, jac);
    float gv = g0[(size_t)p * ldg0 + j];
    if (ge) gv += __half2float(ge[(size_t)p * ldge + j]) + (ge_lo_off ? __half2float(ge[(size_t)p * ldge + ge_lo_off + j]) : 0.0f);
    acc[comp] += jac * gv;
  }
  nabla[3 * p] = acc[0]; nabla[3 * p + 1] = acc[1]; nabla[3 * p + 2] = acc[2];
}
// gbar[p, j] = scale * dPE_j/dx_c nbar[p, c(j)]  as fp16 (columns [pe_dim, width) zero); optionally into g2 at off2
__global__ void pe_jac_kernel(const float* __restrict__ x, int64_t n, int multires, int pe_dim, const float* __restrict__ nbar,
                              float scale, __half* __restrict__ gbar, int ld, int width, __half* __restrict__ g2, int ld2, int off2) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n * width) return;
  const int64_t p = i / width;
  const int j = (int)(i - p * width);
  float v = 0.0f;
  if (j < pe_dim) {
    const float x3[3] = {x[3 * p], x[3 * p + 1], x[3 * p + 2]};
    float val, jac; int comp;
    pe_and_jac(j, multires, x3, val, comp, jac);
    v = scale * jac * nbar[3 * p + comp];
  }
  gbar[(size_t)p * ld + j] = __float2half_rn(v);
  if (g2 && j < pe_dim) g2[(size_t)p * ld2 + off2 + j] = __float2half_rn(v);
}

}  // namespace

extern "C" size_t nr_gemm16_pack_w_bytes(int32_t N, int32_t K) {
  return (size_t)((K + kKC - 1) / kKC) * ((N + 15) / 16 * 16) * 128;
}

extern "C" int nr_gemm16_pack_w(const float* W, int32_t ldw, int32_t N, int32_t K, void* img, void* stream) {
  NR_CHECK_ARG(W && img && N >= 1 && N <= 256 && K >= 1 && (ldw & 3) == 0 && ldw >= K, "nr_gemm16_pack_w: bad arguments");
  NR_CHECK_ARG((((uintptr_t)W | (uintptr_t)img) & 15) == 0, "nr_gemm16_pack_w: 16-byte alignment");
  const int npad = (N + 15) / 16 * 16, n_kc = (K + kKC - 1) / kKC;
  gemm16_pack_w_kernel<<<(unsigned)nr_cdiv((int64_t)n_kc * npad * 8, 256), 256, 0, (cudaStream_t)stream>>>(W, ldw, N, K, npad, n_kc,
                                                                                                    (uint8_t*)img);
  NR_CHECK_LAUNCH("gemm16_pack_w_kernel");
  return NR_OK;
}

extern "C" size_t nr_gemm16_pack_w_split_bytes(int32_t N, int32_t K) {
  const int kp = (K + kKC - 1) / kKC * kKC;
  return (size_t)((N + kSplitBlock - 1) / kSplitBlock) * (2 * kp / kKC) * kSplitBlock * 128;
}

extern "C" int nr_gemm16_pack_w_split(const float* W, int32_t ldw, int32_t N, int32_t K, void* img, void* stream) {
  NR_CHECK_ARG(W && img && N >= 1 && K >= 1 && (ldw & 3) == 0 && ldw >= K, "nr_gemm16_pack_w_split: bad arguments");
  NR_CHECK_ARG((((uintptr_t)W | (uintptr_t)img) & 15) == 0, "nr_gemm16_pack_w_split: 16-byte alignment");
  const int kp = (K + kKC - 1) / kKC * kKC, n_blocks = (N + kSplitBlock - 1) / kSplitBlock;
  const int64_t total = (int64_t)n_blocks * (2 * kp / kKC) * kSplitBlock * 8;
  gemm16_pack_w_split_kernel<<<(unsigned)nr_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(W, ldw, N, K, kp, n_blocks, (uint8_t*)img);
  NR_CHECK_LAUNCH("gemm16_pack_w_split_kernel");
  return NR_OK;
}

extern "C" int nr_gemm16(const void* A, int32_t lda, const float* W, int32_t ldw, const float* bias, int64_t M, int32_t N,
                         int32_t K, void* Y, int32_t ldy, int32_t y_half, int32_t mode, const void* aux_a, int32_t ld_a,
                         const void* aux_b, int32_t ld_b, void* out2, int32_t ld_o2, int32_t w_packed, void* stream) {
  NR_CHECK_ARG(A && W && Y && M >= 0 && N >= 1 && K >= 1, "nr_gemm16: bad arguments");
  NR_CHECK_ARG(mode >= G_LINEAR && mode <= G_MASK, "nr_gemm16: mode=%d", mode);
  const int n8 = (N + 15) & ~15, kpad = (K + kKC - 1) / kKC * kKC;     // columns are handled 16 at a time (32-byte accesses)
  NR_CHECK_ARG(lda % 16 == 0 && lda >= kpad, "nr_gemm16: lda=%d must be a multiple of 16 covering K rounded up to 64 (%d)", lda, kpad);
  NR_CHECK_ARG(w_packed || ((ldw & 3) == 0 && ldw >= K), "nr_gemm16: ldw");
  NR_CHECK_ARG(ldy >= n8 && ldy % (y_half ? 16 : 4) == 0, "nr_gemm16: ldy=%d must cover N rounded up to 16", ldy);
  NR_CHECK_ARG((((uintptr_t)A | (uintptr_t)aux_a | (uintptr_t)aux_b | (uintptr_t)out2) & 31) == 0 &&
                   ((uintptr_t)Y & (y_half ? 31 : 15)) == 0 && ((uintptr_t)W & 15) == 0,
               "nr_gemm16: fp16 operands must be 32-byte aligned, fp32 ones 16-byte");
  NR_CHECK_ARG(!aux_a || (ld_a % 16 == 0 && ld_a >= n8), "nr_gemm16: ld_a");
  NR_CHECK_ARG(!aux_b || (ld_b % 16 == 0 && ld_b >= n8), "nr_gemm16: ld_b");
  NR_CHECK_ARG(!out2 || (ld_o2 % 16 == 0 && ld_o2 >= n8), "nr_gemm16: ld_o2");
  NR_CHECK_ARG((mode != G_SCALE && mode != G_ADJ && mode != G_MASK) || aux_a, "nr_gemm16: mode %d needs aux_a", mode);
  NR_CHECK_ARG(mode != G_ADJ || (aux_b && out2), "nr_gemm16: G_ADJ needs aux_b and out2");
  NR_CHECK_ARG(mode != G_SOFTPLUS || out2, "nr_gemm16: G_SOFTPLUS needs out2");
  if (M == 0) return NR_OK;
  G16Args g{(const __half*)A, lda, w_packed ? nullptr : W, ldw, w_packed ? (const uint8_t*)W : nullptr, 3, bias, M, N, K, Y, ldy,
            y_half, mode, (const __half*)aux_a, ld_a, (const __half*)aux_b, ld_b, (__half*)out2, ld_o2, 0, 0, 0, 0, 0};
  g.npad = (N + 15) / 16 * 16;
  g.n_kc = (K + kKC - 1) / kKC;
  NR_CHECK_ARG(g.npad <= 256, "nr_gemm16: N=%d > 256", N);
  const size_t fixed = 1024 + (size_t)g.n_kc * g.npad * 128 + 256;
  NR_CHECK_ARG(fixed + 3 * kAStageBytes <= 227 * 1024, "nr_gemm16: W (%d x %d) does not fit in shared memory", N, K);
  g.a_stages = (int)((227 * 1024 - fixed) / kAStageBytes);
  if (g.a_stages > kMaxAStages) g.a_stages = kMaxAStages;
  const size_t smem = fixed + (size_t)g.a_stages * kAStageBytes;
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = nr_cdiv(M, kBM);
  const int grid = (int)(n_tiles < sms ? n_tiles : sms);
  CUtensorMap map_a;      // A: [M rows, lda columns] fp16 row-major, box = [128 rows x 64 columns]
  if (int rc = make_map(&map_a, A, M, lda, kBM, "nr_gemm16")) return rc;
#define NR_G16_LAUNCH(MODE_, YH_)                                                                                         \
  do {                                                                                                                    \
    NR_CHECK_CUDA(cudaFuncSetAttribute(gemm16_kernel<MODE_, YH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gemm16_kernel<MODE_, YH_><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(g, map_a);                                 \
  } while (0)
#define NR_G16_CASE(MODE_) case MODE_: if (y_half) NR_G16_LAUNCH(MODE_, true); else NR_G16_LAUNCH(MODE_, false); break
  switch (mode) {
    NR_G16_CASE(G_LINEAR); NR_G16_CASE(G_SOFTPLUS); NR_G16_CASE(G_SCALE); NR_G16_CASE(G_ADJ); NR_G16_CASE(G_RELU);
    NR_G16_CASE(G_SIGMOID); NR_G16_CASE(G_MASK);
    default: break;
  }
  NR_CHECK_LAUNCH("gemm16_kernel");
  return NR_OK;
}

// Split-precision forward GEMM:  Y = epilogue((A_hi + A_lo) (W_hi + W_lo)^T) without the lo x lo term, accumulated in ONE
// TMEM accumulator: every A chunk is loaded once, a hi chunk multiplies the W_hi and the W_lo chunk of its k range, a lo chunk
// the W_hi chunk.  A: fp16 rows [M, lda] holding [hi (Kp columns) | lo (Kp columns)], Kp = K rounded up to 64; Wimg:
// nr_gemm16_pack_w_split (per block of 128 output columns [W_hi | W_lo]); Y fp16 with the lo part of the result lo_off columns
// to the right (lo_off = 0: hi only), or fp32.  grid = (tiles, column blocks): the blocks of a row tile run on different SMs at
// the same time and share its A chunks through L2.
extern "C" int nr_gemm16_split(const void* A, int32_t lda, const void* Wimg, const float* bias, int64_t M, int32_t N, int32_t K,
                               void* Y, int32_t ldy, int32_t y_half, int32_t lo_off, int32_t mode, void* out2, int32_t ld_o2,
                               const void* aux_a, int32_t ld_a, void* stream) {
  NR_CHECK_ARG(A && Wimg && Y && M >= 0 && N >= 1 && K >= 1, "nr_gemm16_split: bad arguments");
  NR_CHECK_ARG(mode == G_LINEAR || mode == G_SOFTPLUS || mode == G_RELU || mode == G_SIGMOID || mode == G_SCALE,
               "nr_gemm16_split: mode=%d", mode);
  NR_CHECK_ARG(mode != G_SCALE || aux_a, "nr_gemm16_split: G_SCALE needs aux_a");
  NR_CHECK_ARG(!aux_a || (ld_a % 16 == 0 && ld_a >= ((N + 15) & ~15) && ((uintptr_t)aux_a & 31) == 0), "nr_gemm16_split: aux_a");
  const int kp = (K + kKC - 1) / kKC * kKC, n_blocks = (N + kSplitBlock - 1) / kSplitBlock, n16 = (N + 15) & ~15;
  NR_CHECK_ARG(lda % 16 == 0 && lda >= 2 * kp, "nr_gemm16_split: lda=%d must cover [hi | lo] = 2 x %d columns", lda, kp);
  NR_CHECK_ARG(ldy % (y_half ? 16 : 4) == 0 && ldy >= (lo_off ? lo_off + n16 : n16) && (lo_off == 0 || (y_half && lo_off % 16 == 0 && lo_off >= n16)),
               "nr_gemm16_split: ldy / lo_off");
  NR_CHECK_ARG((((uintptr_t)A | (uintptr_t)out2) & 31) == 0 && ((uintptr_t)Y & (y_half ? 31 : 15)) == 0 && ((uintptr_t)Wimg & 15) == 0,
               "nr_gemm16_split: alignment");
  NR_CHECK_ARG(mode != G_SOFTPLUS || out2, "nr_gemm16_split: G_SOFTPLUS needs out2");
  NR_CHECK_ARG(!out2 || (ld_o2 % 16 == 0 && ld_o2 >= n16), "nr_gemm16_split: ld_o2");
  if (M == 0) return NR_OK;
  G16Args g{(const __half*)A, lda, nullptr, 0, (const uint8_t*)Wimg, 3, bias, M, N, 2 * kp, Y, ldy, y_half, mode, (const __half*)aux_a,
            ld_a, nullptr, 0, (__half*)out2, ld_o2, 0, 0, 0, 0, 0};
  g.npad = kSplitBlock;
  g.n_kc = 2 * kp / kKC;          // W chunks of a block: [W_hi | W_lo]
  g.split_p = kp / kKC;
  g.lo_off = lo_off;
  g.w_block_bytes = g.n_kc * g.npad * 128;
  const size_t fixed = 1024 + (size_t)g.n_kc * g.npad * 128 + 256;
  NR_CHECK_ARG(fixed + 3 * kAStageBytes <= 227 * 1024, "nr_gemm16_split: W (128 x %d) does not fit in shared memory", 2 * kp);
  g.a_stages = (int)((227 * 1024 - fixed) / kAStageBytes);
  if (g.a_stages > kMaxAStages) g.a_stages = kMaxAStages;
  const size_t smem = fixed + (size_t)g.a_stages * kAStageBytes;
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_tiles = nr_cdiv(M, kBM);
  int64_t gx = sms / n_blocks;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  CUtensorMap map_a;
  if (int rc = make_map(&map_a, A, M, lda, kBM, "nr_gemm16_split")) return rc;
  const dim3 grid((unsigned)gx, (unsigned)n_blocks);
#define NR_G16S_LAUNCH(MODE_, YH_)                                                                                        \
  do {                                                                                                                    \
    NR_CHECK_CUDA(cudaFuncSetAttribute(gemm16_kernel<MODE_, YH_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gemm16_kernel<MODE_, YH_><<<grid, kGemmThreads, smem, (cudaStream_t)stream>>>(g, map_a);                                 \
  } while (0)
#define NR_G16S_CASE(MODE_) case MODE_: if (y_half) NR_G16S_LAUNCH(MODE_, true); else NR_G16S_LAUNCH(MODE_, false); break
  switch (mode) {
    NR_G16S_CASE(G_LINEAR); NR_G16S_CASE(G_SOFTPLUS); NR_G16S_CASE(G_RELU); NR_G16S_CASE(G_SIGMOID); NR_G16S_CASE(G_SCALE);
    default: break;
  }
  NR_CHECK_LAUNCH("gemm16_kernel (split)");
  return NR_OK;
}

static int gemm16_tn_launch(const void* G, int32_t ldg, const void* X, int32_t ldx, const void* G2, int32_t ldg2, const void* X2,
                            int32_t ldx2, int64_t rows, int32_t N, int32_t K, float* dW, int32_t lddw, float scale, void* stream);

extern "C" int nr_gemm16_tn(const void* G, int32_t ldg, const void* X, int32_t ldx, int64_t rows, int32_t N, int32_t K,
                            float* dW, int32_t lddw, float scale, void* stream) {
  return gemm16_tn_launch(G, ldg, X, ldx, nullptr, 0, nullptr, 0, rows, N, K, dW, lddw, scale, stream);
}

extern "C" int nr_gemm16_tn2(const void* G, int32_t ldg, const void* X, int32_t ldx, const void* G2, int32_t ldg2, const void* X2,
                             int32_t ldx2, int64_t rows, int32_t N, int32_t K, float* dW, int32_t lddw, float scale, void* stream) {
  NR_CHECK_ARG(G2 && X2, "nr_gemm16_tn2: null pointer");
  return gemm16_tn_launch(G, ldg, X, ldx, G2, ldg2, X2, ldx2, rows, N, K, dW, lddw, scale, stream);
}

static int gemm16_tn_launch(const void* G, int32_t ldg, const void* X, int32_t ldx, const void* G2, int32_t ldg2, const void* X2,
                            int32_t ldx2, int64_t rows, int32_t N, int32_t K, float* dW, int32_t lddw, float scale, void* stream) {
  NR_CHECK_ARG(G && X && dW && rows >= 0 && N >= 1 && K >= 1, "nr_gemm16_tn: bad arguments");
  NR_CHECK_ARG(ldg % 64 == 0 && ldx % 64 == 0 && ldg >= N && ldx >= K && lddw >= K,
               "nr_gemm16_tn: ldg / ldx must be multiples of 64 covering N / K (pad columns finite), lddw >= K");
  NR_CHECK_ARG((((uintptr_t)G | (uintptr_t)X) & 15) == 0, "nr_gemm16_tn: 16-byte alignment");
  NR_CHECK_ARG(!G2 || (ldg2 % 64 == 0 && ldx2 % 64 == 0 && ldg2 >= N && ldx2 >= K && (((uintptr_t)G2 | (uintptr_t)X2) & 15) == 0),
               "nr_gemm16_tn2: second operand pair");
  if (rows == 0) return NR_OK;
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int n_mt = (N + 255) / 256, n_nt = (K + 255) / 256;
  const int64_t n_chunks = nr_cdiv(rows, 64);
  int64_t slices = sms / (n_mt * n_nt);
  if (slices < 1) slices = 1;
  if (slices > n_chunks) slices = n_chunks;
  Tn16Args g{(const __half*)G, ldg, (const __half*)X, ldx, rows, N, K, dW, lddw, nr_cdiv(n_chunks, slices), scale,
             ((((uintptr_t)dW) & 15) == 0 && (lddw & 3) == 0) ? 1 : 0, G2 ? 2 : 1};
  slices = nr_cdiv(n_chunks, g.chunks_per_slice);
  const size_t smem = 1024 + kTnStages * (kTnABytes + kTnBBytes) + 128;
  dim3 grid((unsigned)slices, n_mt, n_nt);
  CUtensorMap map_g, map_x, map_g2, map_x2;
  if (int rc = make_map(&map_g, G, rows, ldg, 64, "nr_gemm16_tn")) return rc;
  if (int rc = make_map(&map_x, X, rows, ldx, 64, "nr_gemm16_tn")) return rc;
  if (int rc = make_map(&map_g2, G2 ? G2 : G, rows, G2 ? ldg2 : ldg, 64, "nr_gemm16_tn")) return rc;
  if (int rc = make_map(&map_x2, X2 ? X2 : X, rows, X2 ? ldx2 : ldx, 64, "nr_gemm16_tn")) return rc;
  NR_CHECK_CUDA(cudaFuncSetAttribute(gemm16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gemm16_tn_kernel<<<grid, 192, smem, (cudaStream_t)stream>>>(g, map_g, map_x, map_g2, map_x2);
  NR_CHECK_LAUNCH("gemm16_tn_kernel");
  return NR_OK;
}

extern "C" int nr_colsum16(const void* A, int32_t lda, int64_t rows, int32_t N, float scale, float* out, void* stream) {
  NR_CHECK_ARG(A && out && rows >= 0 && N >= 1 && N <= 256 && lda >= N, "nr_colsum16: bad arguments");
  if (rows == 0) return NR_OK;
  colsum16_kernel<<<(unsigned)nr_cdiv(rows, 256), 1024, 0, (cudaStream_t)stream>>>((const __half*)A, lda, rows, N, scale, out);
  NR_CHECK_LAUNCH("colsum16_kernel");
  return NR_OK;
}

extern "C" int nr_cast_cols16(const float* src, int64_t ld_src, int64_t n, int32_t ncols, float scale, void* dst, int64_t ld_dst,
                              int32_t lo_off, void* stream) {
  NR_CHECK_ARG(src && dst && n >= 0 && ncols >= 1 && ld_src >= 0 && ld_dst >= ncols && lo_off >= 0, "nr_cast_cols16: bad arguments");
  if (n == 0) return NR_OK;
  cast_cols16_kernel<<<(unsigned)nr_cdiv(n * ncols, 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, n, ncols, scale, (__half*)dst,
                                                                                          ld_dst, lo_off);
  NR_CHECK_LAUNCH("cast_cols16_kernel");
  return NR_OK;
}

extern "C" int nr_pe16(const float* x, int64_t n, int32_t multires, void* e, int32_t ld, int32_t width, void* e2, int32_t ld2,
                       int32_t off2, void* stream) {
  const int pe_dim = multires < 0 ? 3 : 3 + 6 * multires;
  NR_CHECK_ARG(x && e && n >= 0 && width >= pe_dim && ld >= width, "nr_pe16: bad arguments");
  if (n == 0) return NR_OK;
  pe16_kernel<<<(unsigned)nr_cdiv(n * width, 256), 256, 0, (cudaStream_t)stream>>>(x, n, multires < 0 ? 0 : multires, pe_dim, (__half*)e, ld,
                                                                                  width, (__half*)e2, ld2, off2, 0, 0);
  NR_CHECK_LAUNCH("pe16_kernel");
  return NR_OK;
}

// nr_pe16 with split-precision output: the lo parts lo_off (lo_off2 for e2) columns to the right of the hi parts
extern "C" int nr_pe16_split(const float* x, int64_t n, int32_t multires, void* e, int32_t ld, int32_t width, int32_t lo_off, void* e2,
                             int32_t ld2, int32_t off2, int32_t lo_off2, void* stream) {
  const int pe_dim = multires < 0 ? 3 : 3 + 6 * multires;
  NR_CHECK_ARG(x && e && n >= 0 && width >= pe_dim && lo_off >= width && ld >= lo_off + width, "nr_pe16_split: bad arguments");
  NR_CHECK_ARG(!e2 || (lo_off2 > 0 && ld2 >= lo_off2 + off2 + pe_dim), "nr_pe16_split: second output");
  if (n == 0) return NR_OK;
  pe16_kernel<<<(unsigned)nr_cdiv(n * width, 256), 256, 0, (cudaStream_t)stream>>>(x, n, multires < 0 ? 0 : multires, pe_dim, (__half*)e, ld,
                                                                                  width, (__half*)e2, ld2, off2, lo_off, lo_off2);
  NR_CHECK_LAUNCH("pe16_kernel (split)");
  return NR_OK;
}

extern "C" int nr_pe_jac_t(const float* x, int64_t n, int32_t multires, const float* g0, int32_t ldg0, const void* ge,
                           int32_t ldge, int32_t ge_lo_off, float* nabla, void* stream) {
  const int pe_dim = multires < 0 ? 3 : 3 + 6 * multires;
  NR_CHECK_ARG(x && g0 && nabla && n >= 0 && ldg0 >= pe_dim, "nr_pe_jac_t: bad arguments");
  if (n == 0) return NR_OK;
  pe_jac_t_kernel<<<(unsigned)nr_cdiv(n, 128), 128, 0, (cudaStream_t)stream>>>(x, n, multires < 0 ? 0 : multires, pe_dim, g0, ldg0,
                                                                            (const __half*)ge, ldge, ge_lo_off, nabla);
  NR_CHECK_LAUNCH("pe_jac_t_kernel");
  return NR_OK;
}

extern "C" int nr_pe_jac(const float* x, int64_t n, int32_t multires, const float* nbar, float scale, void* gbar, int32_t ld,
                         int32_t width, void* g2, int32_t ld2, int32_t off2, void* stream) {
  const int pe_dim = multires < 0 ? 3 : 3 + 6 * multires;
  NR_CHECK_ARG(x && nbar && gbar && n >= 0 && width >= pe_dim && ld >= width, "nr_pe_jac: bad arguments");
  if (n == 0) return NR_OK;
  pe_jac_kernel<<<(unsigned)nr_cdiv(n * width, 256), 256, 0, (cudaStream_t)stream>>>(x, n, multires < 0 ? 0 : multires, pe_dim, nbar, scale,
                                                                                    (__half*)gbar, ld, width, (__half*)g2, ld2, off2);
  NR_CHECK_LAUNCH("pe_jac_kernel");
  return NR_OK;
}
