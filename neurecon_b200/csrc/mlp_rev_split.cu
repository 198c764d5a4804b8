// SPLIT-PRECISION fused SDF MLP with reverse-mode normals: the <= 1e-4 tier on the tensor pipe (precision 'fp16x2').
//
// Same machine and program format as mlp_rev.cu (weights = A operand through a 4-stage bulk-copy ring, activations /
// gradients = B operand resident in shared memory, accumulators in TMEM, two tile slots ping-ponging, 640 threads, the
// softplus' values parked in a per-CTA global scratch between the forward and the backward sweep) -- but every operand is
// a PAIR of fp16 numbers, x = hi + lo with hi = fp16(x), lo = fp16(x - hi): 22 mantissa bits instead of 11.  A product
// keeps three of the four terms (the dropped lo * lo is 2^-22 relative):
//
//     W h  ~=  W_hi h_hi + W_hi h_lo + W_lo h_hi
//
// Layout that makes this two MMAs per k-step instead of three: a tile slot holds 64 points; the operand buffer's two
// 64-column blocks (the N = 128 MN-major operand of mlp_rev.cu) are [h_hi | h_lo] of the SAME 64 points.  Then
//     MMA 1:  A = W_hi chunk, B = both blocks (N = 128)  ->  D[:, 0:64) = W_hi h_hi,  D[:, 64:128) = W_hi h_lo
//     MMA 2:  A = W_lo chunk, B = block 0     (N = 64)   ->  D[:, 64:128) += W_lo h_hi
// and the epilogue adds the two column halves of its TMEM lane.  The lo parts are stored scaled by 2^12 (a lo part
// of an O(0.1) number would be an fp16 subnormal with 2^-25 absolute resolution; scaled it is a normal number), so the
// second column half carries that factor and the epilogue computes  z = D[:, c] + 2^-12 D[:, 64 + c].
// Per 16-wide k-step and 64 points that is 64 + 48 cycles of the tensor pipe (the N = 64 product is bound by the 4 KB
// fetch of A) against 32 for the plain fp16 kernel: 3.5 x the tensor time, which leaves room for an exact epilogue:
// softplus through ex2 + a degree-6
// polynomial of log1p (1.5e-8 absolute), softplus' through a reciprocal, kept as 16-bit fixed-point codes (exact at 0 and
// 1, |err| <= 7.7e-6), sincosf for the embedding and its Jacobian.  The weight image interleaves (hi, lo) chunks
// (umma_pack.pack_a_tiles_split); the ring protocol is mlp_rev.cu's with an even number of chunks per
// M-tile and k-chunk.  The radiance pass consumes the fp16 `hi` image exactly as from mlp_rev.cu.
//
// A program may end at EPI_SDF_OUT (sdf only: no codes are written, no backward sweep).
#include "mlp_epilogue.cuh"

namespace {

constexpr int kStages = 4;
constexpr int kStagesLog2 = 2;
constexpr int kThreads = 640;            // 4 control warps + 2 groups x 8 epilogue warps
constexpr int kEpiPerTile = 256;
constexpr int kEpiWarpsPerTile = 8;
constexpr int kEpiWarp0 = 4;
constexpr int kPts = 64;                 // points per tile slot (the other 64 operand columns are their lo parts)
constexpr uint32_t kSigBytes = 32768;    // softplus' of one layer of one slot: [4 column chunks][256 features][16 x u16]
constexpr uint32_t kSigChunk = 8192;
constexpr int kRedLd = 41;
constexpr int kCh = 2;                   // 16-column chunks of a slot a warp drains per wide step (4 chunks, two groups)
constexpr float kLoScale = 4096.0f;      // lo parts are kept as fp16((x - hi) * 2^12): normal numbers instead of subnormals

struct SmemSplit {
  static constexpr uint32_t act = 0;                                 // 2 x 64 KB: [256 k][64 hi | 64 lo]
  static constexpr uint32_t ring = 2 * kActBytes;                    // kStages x 16 KB
  static constexpr uint32_t xs = ring + kStages * kChunkBytes;       // 2 x 64 x 3 floats
  static constexpr uint32_t pes = xs + 2 * 256 * 4;                  // 2 x 40 rows x 64 floats: embedding / skip-gradient stash
  static constexpr uint32_t bars = pes + 2 * kPeStashRows * 256;
  static constexpr uint32_t total = bars + 256;
};
static_assert(kPts * kRedLd * 4 <= kActBytes, "embedding-gradient scratch must fit in the slot's operand buffer");

struct SplitArgs {
  const uint8_t* image;
  const float* bias;
  const float* x;      // [n,3]
  int64_t n;
  float* sdf;          // [n] or null
  float* nabla;        // [n,3] (null for a forward-only program)
  float* feat;         // [n, feat_ld] or null
  int64_t feat_ld;
  uint8_t* feat_img;   // null, or [ceil(n/128)][64 KB]: last hidden activations (hi part) as the radiance pass's operand image
  uint8_t* sig;        // [grid][2][n_sig][32 KB]
  int n_sig;
  int store_sig;       // 0: forward-only program
};

__device__ __forceinline__ void publish(uint64_t* bar) {
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bar);
}
__device__ __forceinline__ void wait_tag(uint64_t* bar, uint32_t parity, int tag) { umma::mbar_wait(bar, parity, tag); }

// 16 fp32 values of operand row `ra`, columns [16 c, 16 c + 16) -> hi into block 0, lo into block 1
__device__ __forceinline__ void split_store16(const RowAddr& ra, int c, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __half2 h = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((v[2 * j] - hf.x) * kLoScale, (v[2 * j + 1] - hf.y) * kLoScale);
    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
    lo[j] = *reinterpret_cast<const uint32_t*>(&l);
  }
  st_shared_v4(ra.chunk(2 * c), hi[0], hi[1], hi[2], hi[3]);
  st_shared_v4(ra.chunk(2 * c + 1), hi[4], hi[5], hi[6], hi[7]);
  st_shared_v4(ra.chunk(8 + 2 * c), lo[0], lo[1], lo[2], lo[3]);
  st_shared_v4(ra.chunk(8 + 2 * c + 1), lo[4], lo[5], lo[6], lo[7]);
}
__device__ __forceinline__ void zero_store16(const RowAddr& ra, int c) {
  st_shared_v4(ra.chunk(2 * c), 0, 0, 0, 0);
  st_shared_v4(ra.chunk(2 * c + 1), 0, 0, 0, 0);
  st_shared_v4(ra.chunk(8 + 2 * c), 0, 0, 0, 0);
  st_shared_v4(ra.chunk(8 + 2 * c + 1), 0, 0, 0, 0);
}
// one element (k, n) of a slot's operand, split
__device__ __forceinline__ void split_elem(uint8_t* act, int k, int n, float v) {
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn((v - __half2float(h)) * kLoScale);
  *reinterpret_cast<__half*>(act + umma::b_chunk_offset(k, n >> 3, kLbo) + (n & 7) * 2) = h;
  *reinterpret_cast<__half*>(act + umma::b_chunk_offset(k, (kPts + n) >> 3, kLbo) + (n & 7) * 2) = l;
}

// softplus(beta = 100)(z) and its derivative for a pair, z = acc + bias, t = 100 z log2(e):
//   u = 2^-|t|;  softplus = max(t, 0) ln2 / 100 + u P5(u)  (|err| 1.5e-8 absolute);  sigmoid = 1/2 + copysign(1/(1+u) - 1/2, t)
__device__ __forceinline__ void softplus_sig_exact2(f32x2 acc2, f32x2 b144, float& sp0, float& sp1, float& s0, float& s1) {
  const f32x2 t2 = fma2(acc2, splat2(144.26950408889634f), b144);
  float t0, t1;
  upk2(t2, t0, t1);
  const f32x2 u2 = pk2(ex2_approx(-fabsf(t0)), ex2_approx(-fabsf(t1)));
  f32x2 p = fma2(u2, splat2(-0.00018338880909141153f), splat2(0.0008556997636333108f));
  p = fma2(p, u2, splat2(-0.001937609864398837f));
  p = fma2(p, u2, splat2(0.0031764907762408257f));
  p = fma2(p, u2, splat2(-0.004978750832378864f));
  p = fma2(p, u2, splat2(0.009999016299843788f));
  const f32x2 q = mul2(p, u2);
  upk2(fma2(pk2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), splat2(0.006931471805599453f), q), sp0, sp1);
  float d0, d1;
  upk2(add2(u2, splat2(1.0f)), d0, d1);
  float r0, r1;
  upk2(add2(pk2(rcp_approx(d0), rcp_approx(d1)), splat2(-0.5f)), r0, r1);
  r0 = __uint_as_float(__float_as_uint(r0) | (__float_as_uint(t0) & 0x80000000u));
  r1 = __uint_as_float(__float_as_uint(r1) | (__float_as_uint(t1) & 0x80000000u));
  upk2(add2(pk2(r0, r1), splat2(0.5f)), s0, s1);
}
// 16-bit fixed-point codes of two derivatives in [0, 1]: fma(s, 65535, 2^23) leaves round(65535 s) in the low mantissa half
__device__ __forceinline__ uint32_t code2(float s0, float s1) {
  float a, b;
  upk2(fma2(pk2(s0, s1), splat2(65535.0f), splat2(8388608.0f)), a, b);
  return __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x5410);
}
// the two derivatives of a code word: 2^23 + code is exact, the subtraction too, one rounding in the scale
__device__ __forceinline__ f32x2 decode2(uint32_t w) {
  const float lo = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7410));
  const float hi = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7432));
  return mul2(add2(pk2(lo, hi), splat2(-8388608.0f)), splat2(1.0f / 65535.0f));
}
__device__ __forceinline__ uint4 ldcg16(const uint8_t* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void stcg16(uint8_t* p, uint4 v) { __stcg(reinterpret_cast<uint4*>(p), v); }
__device__ __forceinline__ void discard_line(const uint8_t* p) { asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory"); }

// z[j] = a[j] + b[j] for the 16 columns of a chunk: the two halves of the split product
__device__ __forceinline__ void ld_sum16(uint32_t taddr, float (&z)[16]) {
  uint32_t ra[16], rb[16];
  umma::tmem_ld16(taddr, ra);
  umma::tmem_ld16(taddr + kPts, rb);
  umma::tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < 16; ++j) z[j] = fmaf(__uint_as_float(rb[j]), 1.0f / kLoScale, __uint_as_float(ra[j]));
}

__global__ void __launch_bounds__(kThreads, 1) mlp_rev_split_kernel(const __grid_constant__ DevProgram prog, const SplitArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + SmemSplit::bars);
  uint64_t* w_full = bars;                   // [kStages]
  uint64_t* w_empty = bars + kStages;        // [kStages]
  uint64_t* in_ready = bars + 2 * kStages;   // [2]
  uint64_t* acc_ready = in_ready + 2;        // [2]
  __shared__ uint32_t tmem_base_s;

  const nr_umma_program_t& P = prog.p;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.n + kPts - 1) / kPts;
  const int64_t n_pairs = ((n_tiles + 1) / 2 + gridDim.x - 1) / gridDim.x * gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { umma::mbar_init(&w_full[s], 1); umma::mbar_init(&w_empty[s], 1); }
    for (int t = 0; t < 2; ++t) {
      umma::mbar_init(&in_ready[t], 2 * kEpiWarpsPerTile);
      umma::mbar_init(&acc_ready[t], 2);
    }
    umma::fence_barrier_init();
  }
  if (warp == 2) {
    umma::tmem_alloc(&tmem_base_s, 512);
    umma::tmem_relinquish();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== weight producer: (hi, lo) chunk pairs in (k-chunk, M-tile) order =====================
    uint32_t cnt = 0;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      for (int s = 0; s < P.n_steps; ++s) {
        const int nch = 2 * P.steps[s].n_mt * (P.steps[s].k_steps >> 2);
        const uint8_t* src = a.image + (size_t)P.steps[s].chunk_begin * kChunkBytes;
        for (int t = 0; t < 2; ++t) {
          for (int c = 0; c < nch; ++c, ++cnt) {
            const uint32_t stage = cnt & (kStages - 1);
            wait_tag(&w_empty[stage], ((cnt >> kStagesLog2) & 1u) ^ 1u, 1000 + s);
            NR_INJECT_DELAY(P.debug_flags, P.steps[s].n_mt, t);
            if (umma::elect_one()) {
              umma::mbar_arrive_expect_tx(&w_full[stage], kChunkBytes);
              umma::bulk_g2s(smem + SmemSplit::ring + stage * kChunkBytes, src + (size_t)c * kChunkBytes, kChunkBytes,
                             &w_full[stage]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers: warp 1 owns M-tile 0, warp 3 owns M-tile 1 =====================
    // Ring lockstep exactly as in mlp_rev.cu (rules 1 and 2); an issuer takes the (hi, lo) pair of its
    // M-tile, i.e. two consecutive stages per k-chunk.
    const uint32_t my_mt = warp == 1 ? 0u : 1u;
    uint32_t cnt = 0;
    uint32_t in_par = 0;
    int prev_t = -1;
    uint32_t prev_par = 0;
    const uint32_t a_hi = umma::smem_desc_hi(1024), b_hi = umma::smem_desc_hi(1024);
    const uint32_t ring_lo = umma::smem_desc_lo(umma::smem_u32(smem + SmemSplit::ring), 16);
    const uint32_t act_lo0 = umma::smem_desc_lo(umma::smem_u32(smem + SmemSplit::act), kLbo);
    const uint32_t idesc128 = umma::make_idesc_f16(128, 128, 0, 1), idesc64 = umma::make_idesc_f16(128, 64, 0, 1);
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      for (int s = 0; s < P.n_steps; ++s) {
        const uint32_t n_mt = P.steps[s].n_mt, nkc = P.steps[s].k_steps >> 2;
        for (int t = 0; t < 2; ++t) {
          const uint32_t par = (in_par >> t) & 1u;
          wait_tag(&in_ready[t], par, 2000 + s);
          in_par ^= 1u << t;
          umma::tc_fence_after();
          if (n_mt == 1 && my_mt == 0 && prev_t >= 0) wait_tag(&acc_ready[prev_t], prev_par, 6000 + s);   // rule 2
          if (my_mt < n_mt) {
            const uint32_t d_addr = tmem_base + (uint32_t)t * 256u + my_mt * 128u;
            uint32_t b_lo = act_lo0 + (uint32_t)t * (kActBytes >> 4);
            uint32_t c = cnt + 2u * my_mt;
#pragma unroll 1
            for (uint32_t kc = 0; kc < nkc; ++kc, c += 2u * n_mt, b_lo += 512) {
              {   // W_hi x [h_hi | h_lo]
                const uint32_t st = c & (kStages - 1);
                wait_tag(&w_full[st], (c >> kStagesLog2) & 1u, 3000 + s);
                umma::tc_fence_after();
                const uint32_t a_lo = ring_lo + st * (kChunkBytes >> 4);
                if (umma::elect_one()) {
                  umma::mma_bf16_ss(d_addr, umma::desc64(a_lo, a_hi), umma::desc64(b_lo, b_hi), idesc128, kc ? 1u : 0u);
                  umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 2, a_hi), umma::desc64(b_lo + 128, b_hi), idesc128, 1u);
                  umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 4, a_hi), umma::desc64(b_lo + 256, b_hi), idesc128, 1u);
                  umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 6, a_hi), umma::desc64(b_lo + 384, b_hi), idesc128, 1u);
                  umma::mma_commit(&w_empty[st]);
                }
                __syncwarp();
              }
              {   // W_lo x h_hi onto the lo columns (both lo products carry the factor 2^12)
                const uint32_t c1 = c + 1u;
                const uint32_t st = c1 & (kStages - 1);
                wait_tag(&w_full[st], (c1 >> kStagesLog2) & 1u, 3500 + s);
                umma::tc_fence_after();
                const uint32_t a_lo = ring_lo + st * (kChunkBytes >> 4);
                if (umma::elect_one()) {
                  umma::mma_bf16_ss(d_addr + kPts, umma::desc64(a_lo, a_hi), umma::desc64(b_lo, b_hi), idesc64, 1u);
                  umma::mma_bf16_ss(d_addr + kPts, umma::desc64(a_lo + 2, a_hi), umma::desc64(b_lo + 128, b_hi), idesc64, 1u);
                  umma::mma_bf16_ss(d_addr + kPts, umma::desc64(a_lo + 4, a_hi), umma::desc64(b_lo + 256, b_hi), idesc64, 1u);
                  umma::mma_bf16_ss(d_addr + kPts, umma::desc64(a_lo + 6, a_hi), umma::desc64(b_lo + 384, b_hi), idesc64, 1u);
                  umma::mma_commit(&w_empty[st]);
                }
                __syncwarp();
              }
            }
          }
          if (umma::elect_one()) umma::mma_commit(&acc_ready[t]);
          __syncwarp();
          if (my_mt >= n_mt) wait_tag(&acc_ready[t], par, 5000 + s);   // rule 1
          prev_t = t;
          prev_par = par;
          cnt += 2u * n_mt * nkc;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: 16 warps; group g = warps 4-11 / 12-19 owns tile slot g =====================
    // wide steps of BOTH slots are drained by BOTH groups (owner: column chunks 0-1, the other group: 2-3), as in mlp_rev.cu
    const int e = warp - kEpiWarp0;
    const int g = e >> 3;
    const int mo = (e >> 2) & 1;
    const int q = warp & 3;
    const int etid = (e & 7) * 32 + lane;
    const int F = mo * 128 + 32 * q + lane;
    uint32_t acc_par = 0;
    const int pe_dim = P.multires < 0 ? 3 : 3 + 6 * P.multires;

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      {
        // ---- prologue (owner): stage the points, embedding (hi, lo) into operand rows [0, k0), fp32 copy to the stash ----
        const int64_t p0 = (2 * pair + g) * kPts;
        uint8_t* act = smem + SmemSplit::act + g * kActBytes;
        float* xs = (float*)(smem + SmemSplit::xs) + g * 256;
        float* pes = (float*)(smem + SmemSplit::pes) + g * (kPeStashRows * kPts);
        for (int i = etid; i < 3 * kPts; i += kEpiPerTile) {
          const int64_t gi = p0 * 3 + i;
          xs[i] = gi < a.n * 3 ? a.x[gi] : 0.0f;
        }
        named_bar_sync(1 + g, kEpiPerTile);
        const int n = etid & (kPts - 1);
        const int part = etid >> 6;                    // 0..3: splits the rows
        const float x3[3] = {xs[3 * n], xs[3 * n + 1], xs[3 * n + 2]};
        const int k0 = P.steps[0].k_steps * 16;
        auto put = [&](int j, float val) {
          split_elem(act, j, n, val);
          if (j < kPeStashRows) pes[j * kPts + n] = val;
        };
        if (part == 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) put(j, x3[j]);
        }
        for (int qf = part; qf < P.multires; qf += 4) {
          const float f = (float)(1 << qf);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincosf(x3[c] * f, &sn, &cs);
            put(3 + 6 * qf + c, sn);
            put(3 + 6 * qf + 3 + c, cs);
          }
        }
        for (int j = pe_dim + part; j < k0; j += 4) split_elem(act, j, n, 0.f);
        publish(&in_ready[g]);
        if (lane == 0) umma::mbar_arrive(&in_ready[g ^ 1]);
      }

      for (int s = 0; s < P.n_steps; ++s) {
        const nr_umma_step_t& S = P.steps[s];
        const bool mine = mo < S.n_mt;
        const bool is_h = F < S.out_rows;
        const bool is_pe = S.pe_fill && F >= S.out_rows && F < S.out_rows + pe_dim;
        const bool last_step = s + 1 == P.n_steps;
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
          const bool own = t == g;
          const int c0 = own ? 0 : kCh;
          const int64_t tile = 2 * pair + t;
          const int64_t p0 = tile * kPts;
          uint8_t* act = smem + SmemSplit::act + t * kActBytes;
          float* xs = (float*)(smem + SmemSplit::xs) + t * 256;
          float* pes = (float*)(smem + SmemSplit::pes) + t * (kPeStashRows * kPts);
          const uint32_t tslot = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * 256);
          const uint32_t taddr = tslot + (uint32_t)(mo * 128) + 16u * c0;
          const RowAddr ra(umma::smem_u32(act), F);
          uint8_t* sig_slot = a.sig + (((size_t)blockIdx.x * 2 + t) * a.n_sig + S.sig_slot) * kSigBytes + (size_t)F * 32 +
                              (size_t)c0 * kSigChunk;

          uint4 sg[kCh][2];
          const bool use_sig = (S.epi == EPI_BWD && mine && is_h) || (S.epi == EPI_SDF_OUT && !last_step);
          if (use_sig) {
#pragma unroll
            for (int k = 0; k < kCh; ++k) {
              sg[k][0] = ldcg16(sig_slot + k * kSigChunk);
              sg[k][1] = ldcg16(sig_slot + k * kSigChunk + 16);
            }
          }

          wait_tag(&acc_ready[t], (acc_par >> t) & 1u, 4000 + s);
          acc_par ^= 1u << t;
          umma::tc_fence_after();

          if (S.epi == EPI_HIDDEN) {
            if (mine) {
              const f32x2 b144 = splat2(a.bias[S.bias_off + F] * 144.26950408889634f);
              const int jpe = F - S.out_rows;
#pragma unroll 1
              for (int k = 0; k < kCh; ++k) {
                const int c = c0 + k;
                float z[16];
                ld_sum16(taddr + 16 * k, z);
                if (!is_pe) {
                  float vv[16], sv[16];
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    softplus_sig_exact2(pk2(z[2 * j], z[2 * j + 1]), b144, vv[2 * j], vv[2 * j + 1], sv[2 * j], sv[2 * j + 1]);
                  split_store16(ra, c, vv);
                  if (S.to_rad && a.feat_img && tile < n_tiles)
                    img_store16<true>(a.feat_img, tile >> 1, F, 16 * c + kPts * (int)(tile & 1), vv);
                  if (a.store_sig) {
                    stcg16(sig_slot + k * kSigChunk, make_uint4(code2(sv[0], sv[1]), code2(sv[2], sv[3]), code2(sv[4], sv[5]),
                                                                code2(sv[6], sv[7])));
                    stcg16(sig_slot + k * kSigChunk + 16, make_uint4(code2(sv[8], sv[9]), code2(sv[10], sv[11]),
                                                                     code2(sv[12], sv[13]), code2(sv[14], sv[15])));
                  }
                } else {   // skip layer: its input rows [out_rows, out_rows + pe_dim) are the embedding
                  float vv[16];
                  const float4* src = reinterpret_cast<const float4*>(pes + jpe * kPts + 16 * c);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const float4 f4 = src[j];
                    vv[4 * j] = f4.x; vv[4 * j + 1] = f4.y; vv[4 * j + 2] = f4.z; vv[4 * j + 3] = f4.w;
                  }
                  split_store16(ra, c, vv);
                }
              }
            }
          } else if (S.epi == EPI_FEAT) {
            if (mine) {
              const float b = a.bias[S.bias_off + F];
#pragma unroll 1
              for (int k = 0; k < kCh; ++k) {
                float z[16];
                ld_sum16(taddr + 16 * k, z);
                if (a.feat && F < S.out_rows) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const int64_t gp = p0 + 16 * (c0 + k) + j;
                    if (gp < a.n) a.feat[gp * a.feat_ld + F] = z[j] + b;
                  }
                }
              }
            }
          } else if (S.epi == EPI_SDF_OUT) {
            // rows 0..31 of M-tile 0 all hold the sdf row: lane l keeps column l of each 32-column half (owner's warp 0)
            if (own && mo == 0 && q == 0 && a.sdf) {
              const float b = a.bias[S.bias_off];
#pragma unroll 1
              for (int c = 0; c < 2; ++c) {
                uint32_t ra32[32], rb32[32];
                umma::tmem_ld32(tslot + 32 * c, ra32);
                umma::tmem_ld32(tslot + kPts + 32 * c, rb32);
                umma::tmem_ld_wait();
                float m = 0.0f;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  m = (lane == j) ? fmaf(__uint_as_float(rb32[j]), 1.0f / kLoScale, __uint_as_float(ra32[j])) : m;
                const int64_t gp = p0 + 32 * c + lane;
                if (gp < a.n) a.sdf[gp] = m + b;
              }
            }
            if (!last_step) {
              // start of the backward sweep: operand row F <- softplus'(z_last)[F, :] * w_sdf[F]
              const f32x2 wF = splat2(a.bias[S.aux_off + F]);
#pragma unroll
              for (int k = 0; k < kCh; ++k) {
                const uint32_t w8[8] = {sg[k][0].x, sg[k][0].y, sg[k][0].z, sg[k][0].w, sg[k][1].x, sg[k][1].y, sg[k][1].z, sg[k][1].w};
                float vv[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) upk2(mul2(decode2(w8[j]), wF), vv[2 * j], vv[2 * j + 1]);
                split_store16(ra, c0 + k, vv);
              }
            }
          } else if (S.epi == EPI_BWD) {
            if (mine) {
              const int jpe = F - S.out_rows;
#pragma unroll
              for (int k = 0; k < kCh; ++k) {
                const int c = c0 + k;
                float z[16];
                ld_sum16(taddr + 16 * k, z);
                if (is_h) {
                  const uint32_t w8[8] = {sg[k][0].x, sg[k][0].y, sg[k][0].z, sg[k][0].w, sg[k][1].x, sg[k][1].y, sg[k][1].z, sg[k][1].w};
                  float vv[16];
#pragma unroll
                  for (int j = 0; j < 8; ++j) upk2(mul2(decode2(w8[j]), pk2(z[2 * j], z[2 * j + 1])), vv[2 * j], vv[2 * j + 1]);
                  split_store16(ra, c, vv);
                } else {
                  if (is_pe) {   // gradient w.r.t. the skip connection's copy of the embedding: kept (fp32) for EPI_NABLA
                    float4* dst = reinterpret_cast<float4*>(pes + jpe * kPts + 16 * c);
#pragma unroll
                    for (int j = 0; j < 4; ++j) dst[j] = make_float4(z[4 * j], z[4 * j + 1], z[4 * j + 2], z[4 * j + 3]);
                  }
                  zero_store16(ra, c);       // these rows meet zero weights; keep them finite
                }
              }
            }
          } else if (S.epi == EPI_NABLA && own) {
            // acc rows [0, pe_dim) = d sdf / d PE(x) through layer 0 (+ the skip layer's share from the stash);
            // nabla_c = sum_j dPE_j/dx_c * g_j: products to a [point][row] scratch in the (now free) operand buffer
            // Phase A, all 256 threads: the Jacobian dPE_j/dx of every row and point, one sincosf per (point, frequency, component)
            // (it was evaluated per (row, point) element, with the stash read through a 32-way bank conflict: see mlp_rev.cu)
            float* red = reinterpret_cast<float*>(act);
            {
              const int n = etid & (kPts - 1), part = etid >> 6;
              const float x3[3] = {xs[3 * n], xs[3 * n + 1], xs[3 * n + 2]};
              float* rn = red + n * kRedLd;
              if (part == 0) rn[0] = rn[1] = rn[2] = 1.0f;
              for (int qf = part; qf < P.multires; qf += 4) {
                const float f = (float)(1 << qf);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  float sn, cs;
                  sincosf(x3[c] * f, &sn, &cs);
                  rn[3 + 6 * qf + c] = f * cs;
                  rn[3 + 6 * qf + 3 + c] = -f * sn;
                }
              }
            }
            named_bar_sync(1 + g, kEpiPerTile);
            // Phase B: scale by the gradient rows (TMEM lane = row; the two warps of a lane quarter take 32 columns each)
            if (32 * q < pe_dim) {   // warp-uniform
              const int Rr = 32 * q + lane;
              const float* srow = pes + Rr * kPts;
#pragma unroll 1
              for (int c = 2 * mo; c < 2 * mo + 2; ++c) {
                float z[16];
                ld_sum16(tslot + 16 * c, z);
                if (Rr < pe_dim) {
                  float sv[16] = {};
                  if (S.pe_fill) {   // 16 stash values as four 16-byte loads
                    const float4* sp = reinterpret_cast<const float4*>(srow + 16 * c);
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                      const float4 w = sp[j4];
                      sv[4 * j4] = w.x; sv[4 * j4 + 1] = w.y; sv[4 * j4 + 2] = w.z; sv[4 * j4 + 3] = w.w;
                    }
                  }
#pragma unroll
                  for (int j = 0; j < 16; ++j) red[(16 * c + j) * kRedLd + Rr] *= z[j] + sv[j];
                }
              }
            }
            named_bar_sync(1 + g, kEpiPerTile);
            if (etid < kPts) {
              const float* r = red + etid * kRedLd;
              float g0 = r[0], g1 = r[1], g2 = r[2];
              for (int j = 3; j < pe_dim; j += 3) { g0 += r[j]; g1 += r[j + 1]; g2 += r[j + 2]; }
              const int64_t gp = p0 + etid;
              if (gp < a.n) { a.nabla[gp * 3] = g0; a.nabla[gp * 3 + 1] = g1; a.nabla[gp * 3 + 2] = g2; }
            }
          }
          __syncwarp();
          if (use_sig && (lane & 3) == 0) {   // the codes are dead once read: drop the lines instead of writing them back
#pragma unroll
            for (int k = 0; k < kCh; ++k) discard_line(sig_slot + k * kSigChunk);
          }
          if (!last_step) publish(&in_ready[t]);
        }
      }
      umma::tc_fence_before();
      named_bar_sync(1 + g, kEpiPerTile);
    }
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 2) umma::tmem_dealloc(tmem_base, 512);
}

int count_sig_slots(const nr_umma_program_t* p) {
  int n = 0;
  for (int s = 0; s < p->n_steps; ++s)
    if (p->steps[s].epi == EPI_HIDDEN) ++n;
  return n;
}

}  // namespace

extern "C" size_t nr_mlp_split_reverse_workspace(const nr_umma_program_t* prog, int64_t n) {
  if (!prog || n <= 0) return 0;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return (size_t)sms * 2 * count_sig_slots(prog) * kSigBytes;
}

extern "C" int nr_mlp_split_reverse(const nr_umma_program_t* prog, const void* image, size_t image_bytes, const float* bias,
                                    size_t bias_floats, const float* x, int64_t n, float* sdf, float* nabla, float* feat,
                                    int64_t feat_ld, void* feat_img, void* workspace, size_t workspace_bytes, void* stream) {
  NR_CHECK_ARG(prog && image && bias && x, "nr_mlp_split_reverse: null pointer");
  NR_CHECK_ARG(n >= 0, "nr_mlp_split_reverse: n < 0");
  NR_CHECK_ARG(prog->reverse == 1 && prog->tangents == 0 && prog->input_mode == 0 && prog->operand_f16 == 1,
               "nr_mlp_split_reverse: needs a reverse-mode fp16 program on value tiles");
  NR_CHECK_ARG(prog->n_steps >= 2 && prog->n_steps <= NR_UMMA_MAX_STEPS, "nr_mlp_split_reverse: n_steps=%d", prog->n_steps);
  NR_CHECK_ARG(((uintptr_t)image & 15) == 0 && ((uintptr_t)workspace & 127) == 0 && ((uintptr_t)feat_img & 15) == 0,
               "nr_mlp_split_reverse: image and feat_img must be 16-byte, the workspace 128-byte aligned");
  const int pe_dim = prog->multires < 0 ? 3 : 3 + 6 * prog->multires;
  NR_CHECK_ARG(pe_dim <= kPeStashRows && pe_dim % 3 == 0, "nr_mlp_split_reverse: embedding of %d rows", pe_dim);
  NR_CHECK_ARG(prog->steps[0].k_steps * 16 >= pe_dim, "nr_mlp_split_reverse: step 0 K does not cover the embedding");
  const int n_sig = count_sig_slots(prog);
  // order: hidden layers, optional feature, sdf row [, backward layers, embedding Jacobian]
  int phase = 0, n_sdf = 0, n_nabla = 0;
  for (int s = 0; s < prog->n_steps; ++s) {
    const nr_umma_step_t& S = prog->steps[s];
    const int nch = 2 * S.n_mt * (S.k_steps / 4);
    NR_CHECK_ARG(S.k_steps % 4 == 0 && S.k_steps >= 4 && S.k_steps <= 16, "step %d: k_steps=%d", s, S.k_steps);
    NR_CHECK_ARG(S.n_mt >= 1 && S.n_mt <= 2, "step %d: n_mt=%d", s, S.n_mt);
    NR_CHECK_ARG(S.n_cols == 128 && !S.accumulate, "step %d: split steps use the whole operand buffer, no split-K", s);
    NR_CHECK_ARG(S.chunk_begin >= 0 && (size_t)(S.chunk_begin + nch) * kChunkBytes <= image_bytes,
                 "step %d: weight chunks [%d,%d) exceed the image", s, S.chunk_begin, S.chunk_begin + nch);
    const bool last = s == prog->n_steps - 1;
    switch (S.epi) {
      case EPI_HIDDEN:
        NR_CHECK_ARG(phase == 0 && S.sig_slot == s, "step %d: hidden layers come first, sig_slot = layer", s);
        NR_CHECK_ARG(S.bias_off >= 0 && (size_t)S.bias_off + S.n_mt * 128 <= bias_floats, "step %d: bias range", s);
        NR_CHECK_ARG(!S.pe_fill || S.out_rows + pe_dim <= S.n_mt * 128, "step %d: skip operand too wide", s);
        break;
      case EPI_FEAT:
        NR_CHECK_ARG(phase == 0 && s > 0 && !S.to_rad, "step %d: EPI_FEAT follows the hidden layers", s);
        NR_CHECK_ARG(S.bias_off >= 0 && (size_t)S.bias_off + S.n_mt * 128 <= bias_floats, "step %d: bias range", s);
        break;
      case EPI_SDF_OUT:
        NR_CHECK_ARG(phase == 0 && s > 0 && S.n_mt == 1 && S.sig_slot == n_sig - 1, "step %d: EPI_SDF_OUT placement", s);
        NR_CHECK_ARG(S.bias_off >= 0 && (size_t)S.bias_off < bias_floats && S.aux_off >= 0 &&
                         (size_t)S.aux_off + 256 <= bias_floats, "step %d: bias / w_sdf range", s);
        phase = 1;
        ++n_sdf;
        break;
      case EPI_BWD:
        NR_CHECK_ARG(phase == 1 && !last && S.sig_slot >= 0 && S.sig_slot < n_sig, "step %d: EPI_BWD placement / slot", s);
        NR_CHECK_ARG(!S.pe_fill || S.out_rows + pe_dim <= S.n_mt * 128, "step %d: skip operand too wide", s);
        break;
      case EPI_NABLA:
        NR_CHECK_ARG(phase == 1 && last && S.n_mt == 1, "step %d: EPI_NABLA is the last step", s);
        ++n_nabla;
        break;
      default:
        NR_CHECK_ARG(false, "step %d: epi=%d is not a reverse-mode step", s, S.epi);
    }
  }
  const bool fwd_only = prog->steps[prog->n_steps - 1].epi == EPI_SDF_OUT;
  NR_CHECK_ARG(n_sdf == 1 && n_sig >= 1 && (fwd_only ? n_nabla == 0 : n_nabla == 1),
               "nr_mlp_split_reverse: program needs one EPI_SDF_OUT, last or followed by the backward sweep up to EPI_NABLA");
  NR_CHECK_ARG(fwd_only || nabla, "nr_mlp_split_reverse: nabla is null");
  if (n == 0) return NR_OK;
  if (!fwd_only) {
    const size_t need = nr_mlp_split_reverse_workspace(prog, n);
    NR_CHECK_ARG(workspace && workspace_bytes >= need, "nr_mlp_split_reverse: workspace of %zu bytes needed, %zu given", need,
                 workspace_bytes);
  }
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_pairs = (nr_cdiv(n, kPts) + 1) / 2;
  const int grid = (int)(n_pairs < sms ? n_pairs : sms);
  const size_t smem = SmemSplit::total + 1024;
  static unsigned long long attr_set = 0;  // per-device bit: the attribute is per (function, device)
  if (!(attr_set >> (dev & 63) & 1ull)) {
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_rev_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set |= 1ull << (dev & 63);
  }
  DevProgram dp;
  dp.p = *prog;
  SplitArgs ka{(const uint8_t*)image, bias, x, n, sdf, nabla, feat, feat_ld, (uint8_t*)feat_img, (uint8_t*)workspace, n_sig,
               fwd_only ? 0 : 1};
  mlp_rev_split_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(dp, ka);
  NR_CHECK_LAUNCH("mlp_rev_split_kernel");
  return NR_OK;
}
