// Fused positional-encoding + SDF MLP + REVERSE-MODE analytic normals on tcgen05 / TMEM.
//
// Same machine as mlp_umma.cu (weights = A operand streamed through a 4-stage ring, activations = B operand resident
// in shared memory, accumulators in TMEM, two 128-point tiles ping-ponging through mbarriers, 640 threads), but the
// normal  d sdf / d x  is computed the way autograd does it (models/base.py:265-282): one forward pass that remembers
// softplus'(z) of every hidden unit, then one backward pass  g_{l-1} = softplus'(z_{l-1}) * (W_l^T g_l)  from the sdf
// row down to the embedding, whose Jacobian is applied in the last epilogue.  The forward-mode kernel pushes three
// tangent columns per point through every layer (4.03 MFLOP per query on the tensor pipe, 32 points per tile); this
// one executes 2.0 MFLOP per query on full 128-point tiles and halves the activation traffic through shared memory.
//
// The 8 x 256 derivatives per point do not fit on chip next to the operands, so they go through a per-CTA scratch in
// global memory as 8-bit codes 128 + round(254 (s - 1/2)): 32 KB per (tile slot, layer), written by the forward epilogue and read
// back a few microseconds later by THE SAME THREAD in the backward epilogue (no fences needed), 76 MB for the whole
// grid, which stays in the 126 MB L2 (with 16-bit codes the 148 MB scratch cycled through HBM: ncu, profiles/).
// softplus(beta = 100)' is a logistic that sits at 0 or 1 for most units, which the code represents exactly; the
// quantisation (|err| <= 1/508) costs 1.2e-3 rms on the normal (fp64 model of the kernel: tests/test_rev_code_model.py).
// All loads of a step are issued before the thread waits for its accumulator: their latency hides under the MMAs.
//
// Steps (nr_umma_program_t, reverse = 1): EPI_HIDDEN x L (sig_slot = layer), [EPI_FEAT], EPI_SDF_OUT (sdf to global,
// then operand <- softplus'(z_{L-1}) * w_sdf), EPI_BWD x (L-1) (A = W_l^T; the skip layer's embedding rows go to the
// stash), EPI_NABLA (A = W_0^T; (acc + stash) * dPE/dx summed over the embedding rows -> nabla).
#include "mlp_epilogue.cuh"

namespace {

constexpr int kStages = 4;
constexpr int kStagesLog2 = 2;
constexpr int kThreads = 640;            // 4 control warps + 2 tiles x 8 epilogue warps
#ifndef NR_REV_DEFAULT_MODE
#define NR_REV_DEFAULT_MODE 1
#endif
constexpr int kThreadsFat = 384;         // 4 control warps + 8 epilogue warps of 168 registers (kMode 2)
constexpr int kEpiPerTile = 256;
constexpr int kEpiWarpsPerTile = 8;
constexpr int kEpiWarp0 = 4;
constexpr uint32_t kSigBytes = 32768;    // softplus' of one layer of one tile: [8 column chunks][256 features][16 x u8]
constexpr uint32_t kSigChunk = 4096;     // one column chunk: a warp's 16-byte loads / stores cover 512 contiguous bytes
constexpr int kRedLd = 41;               // row stride (floats) of the embedding-gradient scratch: conflict-free both ways

struct SmemRev {
  static constexpr uint32_t act = 0;                                 // 2 x 64 KB
  static constexpr uint32_t ring = 2 * kActBytes;                    // kStages x 16 KB
  static constexpr uint32_t xs = ring + kStages * kChunkBytes;       // 2 x 128 x 3 floats
  static constexpr uint32_t pes = xs + 2 * 512 * 4;                  // 2 x 40 rows x 256 B: embedding / skip-gradient stash
  static constexpr uint32_t bars = pes + 2 * kPeStashRows * 256;
  static constexpr uint32_t total = bars + 256;
};
static_assert(128 * kRedLd * 4 <= kActBytes, "embedding-gradient scratch must fit in the tile's operand buffer");

struct RevArgs {
  const uint8_t* image;
  const float* bias;
  const float* x;      // [n,3]
  int64_t n;
  float* sdf;          // [n] or null
  float* nabla;        // [n,3]
  float* feat;         // [n, feat_ld] or null
  int64_t feat_ld;
  uint8_t* feat_img;   // null, or [ceil(n/128)][64 KB]: last hidden activations as the radiance pass's operand image
  uint8_t* sig;        // [grid][2][n_sig][32 KB]
  int n_sig;
  int store_sig;       // 0: forward-only program (ends with the sdf row): no softplus' codes
  long long* trace;    // test twin only: [8][kTraceCap][4] (event, step * 2 + tile, clock, pair) from CTA 0 (tools/trace_rev.py)
};

constexpr int kTraceCap = 1024;
__device__ __forceinline__ void trace_ev(long long* tr, int region, int& cnt, int ev, int st, long long pair) {
  if (!tr || blockIdx.x != 0 || cnt >= kTraceCap) return;
  long long* p = tr + ((size_t)region * kTraceCap + cnt) * 4;
  p[0] = ev; p[1] = st; p[2] = clock64(); p[3] = pair;
  ++cnt;
}

__device__ __forceinline__ void publish(uint64_t* bar) {
  umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive(bar);
}

__device__ __forceinline__ void wait_tag(uint64_t* bar, uint32_t parity, int tag) { umma::mbar_wait(bar, parity, tag); }

// softplus' codes: byte = 128 + round(254 (s - 1/2)) in [1, 255], exact at s = 0, 1/2 and 1.  The producer hands over
// s - 1/2 (softplus_sigq2); fma(., 254, 1.5 * 2^23 + 128) leaves the code in the low mantissa byte, three byte permutes
// gather four of them.
// Forward activation: 0 = one ex2, log1p cubic + 1/(1+u) quartic on the FMA pipe (softplus_sigq2); 1 = ex2 for softplus and
// tanh for its derivative (softplus_sigt2: two transcendentals, XU bound, measured equal to 0); 2 / 3 = ONE tanh for both
// (softplus_th2, relu on the ALU / on the FMA pipe): profiles/r2_mlp_rev_epilogue.md
#ifndef NR_SIG_TANH
#define NR_SIG_TANH 3
#endif
#ifdef NR_FAULT_INJECT
constexpr bool kProbe = true;     // the test twin (libneurecon_b200_inject.so) also carries the epilogue probes: debug_flags 16 / 32 / 256
#else
constexpr bool kProbe = false;
#endif
constexpr float kSigPackScale = NR_SIG_TANH ? 127.0f : 254.0f;   // the producer hands over tanh(50 z) = 2 (s - 1/2), or s - 1/2
__device__ __forceinline__ uint32_t sig_pack4(f32x2 a, f32x2 b) {
  float a0, a1, b0, b1;
  upk2(fma2(a, splat2(kSigPackScale), splat2(12583040.0f)), a0, a1);
  upk2(fma2(b, splat2(kSigPackScale), splat2(12583040.0f)), b0, b1);
  const uint32_t p0 = __byte_perm(__float_as_uint(a0), __float_as_uint(a1), 0x0040);
  const uint32_t p1 = __byte_perm(__float_as_uint(b0), __float_as_uint(b1), 0x0040);
  return __byte_perm(p0, p1, 0x5410);
}
// codes of byte pair (2 i, 2 i + 1) of w -> the two derivatives times a scale (fp32): 2^23 + code is exact, then one
// FFMA2 with k = scale / 254 and offs = scale / 2 - (2^23 + 128) k
template <int kPair>
__device__ __forceinline__ f32x2 sig_unpack2(uint32_t w, f32x2 k, f32x2 offs) {
  const uint32_t lo = __byte_perm(w, 0x4B000000u, kPair ? 0x7442 : 0x7440);
  const uint32_t hi = __byte_perm(w, 0x4B000000u, kPair ? 0x7443 : 0x7441);
  return fma2(pk2(__uint_as_float(lo), __uint_as_float(hi)), k, offs);
}
// 16 accumulators (or 1.0) times the 16 derivatives coded in q, times scale -> 16-bit operand row chunk
template <bool kF16, bool kAcc>
__device__ __forceinline__ void apply_sig16(const RowAddr& ra, int c, const uint32_t (&r)[16], uint4 q, float scale) {
  const float kf = scale * (1.0f / 254.0f);
  const f32x2 k = splat2(kf), o = splat2(fmaf(-8388736.0f, kf, 0.5f * scale));
  const uint32_t qw[4] = {q.x, q.y, q.z, q.w};
  uint32_t h[8];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f32x2 s0 = sig_unpack2<0>(qw[j], k, o), s1 = sig_unpack2<1>(qw[j], k, o);
    if (kAcc) {
      s0 = mul2(s0, pk2(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1])));
      s1 = mul2(s1, pk2(__uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
    }
    float a0, a1, b0, b1;
    upk2(s0, a0, a1);
    upk2(s1, b0, b1);
    h[2 * j] = umma::pack2<kF16>(a0, a1);
    h[2 * j + 1] = umma::pack2<kF16>(b0, b1);
  }
  st_shared_v4(ra.chunk(2 * c), h[0], h[1], h[2], h[3]);
  st_shared_v4(ra.chunk(2 * c + 1), h[4], h[5], h[6], h[7]);
}
__device__ __forceinline__ uint4 ldcg16(const uint8_t* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }
// The softplus' codes are dead once the backward sweep has read them (the next tile pair overwrites the slot before it
// reads it again).  Left alone, every line is eventually evicted from L2 DIRTY: 0.79 GB of write-backs per 524 288-point
// launch (round 1: dram bytes 57 x the algorithmic traffic).  discard.global.L2 drops the line without the write-back.
// One lane per 128-byte line (8 features x 16 B), after the whole warp has consumed its loads.
#ifndef NR_SIG_DISCARD
#define NR_SIG_DISCARD 1
#endif
__device__ __forceinline__ void discard_line(const uint8_t* p) {
#if NR_SIG_DISCARD
  asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
#endif
}
__device__ __forceinline__ void stcg16(uint8_t* p, uint4 v) { __stcg(reinterpret_cast<uint4*>(p), v); }

// kMode 0: every tile slot's epilogue on its own 8 warps (measurements); 1: both 8-warp groups drain every tile, 64 columns
// each; 2: EIGHT epilogue warps in all, every one draining its 32 features x 128 columns of both tiles -- 384 threads, so 168
// registers per thread instead of 96: the sweep's per-thread invariants stay in registers next to the accumulator chunks, and a
// (layer, slot) visit's bookkeeping is paid by 8 warps per tile instead of 16 (two warps per scheduler keep the activation's
// instruction rate: tools/probe_epi.py variants 20 / 21).
template <bool kF16, int kMode>
__global__ void __launch_bounds__(kMode == 2 ? kThreadsFat : kThreads, 1)
mlp_rev_kernel(const __grid_constant__ DevProgram prog, const RevArgs a) {
  constexpr bool kShare = kMode == 1;
  constexpr bool kFat = kMode == 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + SmemRev::bars);
  uint64_t* w_full = bars;                   // [kStages]
  uint64_t* w_empty = bars + kStages;        // [kStages]
  uint64_t* in_ready = bars + 2 * kStages;   // [2]
  uint64_t* acc_ready = in_ready + 2;        // [2]
  __shared__ uint32_t tmem_base_s;

  const nr_umma_program_t& P = prog.p;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int64_t n_tiles = (a.n + 127) / 128;
  // every CTA runs the same number of tile pairs and always both slots (tiles past the end compute on zeros and store
  // nothing): one static schedule for the producer and the MMA issuers
  const int64_t n_pairs = ((n_tiles + 1) / 2 + gridDim.x - 1) / gridDim.x * gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { umma::mbar_init(&w_full[s], 1); umma::mbar_init(&w_empty[s], 1); }
    for (int t = 0; t < 2; ++t) {
      umma::mbar_init(&in_ready[t], kShare ? 2 * kEpiWarpsPerTile : kEpiWarpsPerTile);   // kFat: its 8 warps
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
    // ===================== weight producer =====================
    uint32_t cnt = 0;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      for (int s = 0; s < P.n_steps; ++s) {
        const int nch = P.steps[s].n_mt * (P.steps[s].k_steps >> 2);
        const uint8_t* src = a.image + (size_t)P.steps[s].chunk_begin * kChunkBytes;
        for (int t = 0; t < 2; ++t) {
          for (int c = 0; c < nch; ++c, ++cnt) {
            const uint32_t stage = cnt & (kStages - 1);
            wait_tag(&w_empty[stage], ((cnt >> kStagesLog2) & 1u) ^ 1u, 1000 + s);
            NR_INJECT_DELAY(P.debug_flags, P.steps[s].n_mt, t);
            if (umma::elect_one()) {
              if (kProbe && (P.debug_flags & 1024)) {      // probe: no weight stream (the MMAs run on whatever the ring holds)
                umma::mbar_arrive(&w_full[stage]);
              } else {
                umma::mbar_arrive_expect_tx(&w_full[stage], kChunkBytes);
                umma::bulk_g2s(smem + SmemRev::ring + stage * kChunkBytes, src + (size_t)c * kChunkBytes, kChunkBytes,
                               &w_full[stage]);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers: warp 1 owns M-tile 0, warp 3 owns M-tile 1 (see mlp_umma.cu) ==============
    const uint32_t my_mt = warp == 1 ? 0u : 1u;
    uint32_t cnt = 0;
    uint32_t in_par = 0;
    int tcnt = 0;
    int prev_t = -1;                   // tile slot and acc_ready parity of the previous (step, tile) visit
    uint32_t prev_par = 0;
    const uint32_t a_hi = umma::smem_desc_hi(1024), b_hi = umma::smem_desc_hi(1024);
    const uint32_t ring_lo = umma::smem_desc_lo(umma::smem_u32(smem + SmemRev::ring), 16);
    const uint32_t act_lo0 = umma::smem_desc_lo(umma::smem_u32(smem + SmemRev::act), kLbo);
    const uint32_t idesc = kF16 ? umma::make_idesc_f16(128, 128, 0, 1) : umma::make_idesc_bf16(128, 128, 0, 1);
    // Both tile slots run the same step back to back.  Measured alternatives, all slower or equal (profiles/): slot 1
    // lagging half a program behind (forward epilogue of one slot against the backward MMAs of the other: +15 %), one
    // weight chunk feeding both slots' MMAs (halves the weight stream: +0 % MMA-only, slower with epilogues), and
    // multicasting the weight stream across 2- or 4-CTA clusters (+0 % / +15 %): the weight stream is not the limiter.
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      for (int s = 0; s < P.n_steps; ++s) {
        const uint32_t n_mt = P.steps[s].n_mt, nkc = P.steps[s].k_steps >> 2;
        for (int t = 0; t < 2; ++t) {
          const uint32_t par = (in_par >> t) & 1u;   // in_ready[t] and acc_ready[t] complete one phase per visit of tile t
          wait_tag(&in_ready[t], par, 2000 + s);
          in_par ^= 1u << t;
          umma::tc_fence_after();
          if (kProbe && lane == 0) trace_ev(a.trace, (int)my_mt, tcnt, 11, s * 2 + t, pair);
          // RING LOCKSTEP, rule 2 (rule 1 and the story are below): a step with one M-tile is issued by warp 1 alone, also
          // from the stages whose previous chunks were warp 3's.  Before it touches them it waits for the previous
          // visit's accumulator: both issuers have committed there, so every chunk issued before this visit has landed.
          if (n_mt == 1 && my_mt == 0 && prev_t >= 0 && !(P.debug_flags & 128)) wait_tag(&acc_ready[prev_t], prev_par, 6000 + s);
          if (my_mt < n_mt) {
            const uint32_t d_addr = tmem_base + (uint32_t)t * 256u + my_mt * 128u;
            uint32_t b_lo = act_lo0 + (uint32_t)t * (kActBytes >> 4);
            uint32_t c = cnt + my_mt;
#pragma unroll 1
            for (uint32_t kc = 0; kc < nkc; ++kc, c += n_mt, b_lo += 512) {
              const uint32_t st = c & (kStages - 1);
              wait_tag(&w_full[st], (c >> kStagesLog2) & 1u, 3000 + s);
              umma::tc_fence_after();
              if (kProbe && lane == 0 && (kc == 0 || kc + 1 == nkc)) trace_ev(a.trace, (int)my_mt, tcnt, kc == 0 ? 13 : 14, s * 2 + t, pair);
              const uint32_t a_lo = ring_lo + st * (kChunkBytes >> 4);
              if (umma::elect_one()) {
                if (!(kProbe && (P.debug_flags & 512))) {  // probe 512: no MMAs (the commits still arrive)
                umma::mma_bf16_ss(d_addr, umma::desc64(a_lo, a_hi), umma::desc64(b_lo, b_hi), idesc, kc ? 1u : 0u);
                umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 2, a_hi), umma::desc64(b_lo + 128, b_hi), idesc, 1u);
                umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 4, a_hi), umma::desc64(b_lo + 256, b_hi), idesc, 1u);
                umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 6, a_hi), umma::desc64(b_lo + 384, b_hi), idesc, 1u);
                }
                umma::mma_commit(&w_empty[st]);
              }
              __syncwarp();
            }
          }
          if (umma::elect_one()) umma::mma_commit(&acc_ready[t]);
          __syncwarp();
          if (kProbe && lane == 0) trace_ev(a.trace, (int)my_mt, tcnt, 12, s * 2 + t, pair);
          // RING LOCKSTEP, rule 1.  A parity wait is sound only if the waiter knows that the barrier's PREVIOUS phase has
          // completed: met one phase too early, the barrier shows the opposite parity and the wait falls through.  A
          // consumer that takes every chunk of a stage knows it from its own last wait; with two issuers sharing the
          // ring that breaks where a stage changes hands -- at the single-M-tile steps (sdf row, Jacobian), which warp 1
          // issues alone.  Round 1 let warp 3 skip such a step; when tile 1's chunks of it were late (weights evicted
          // from L2) warp 3 reached the w_full barrier of its next chunk while the chunk four positions earlier had not
          // landed, fell through, multiplied stale weights, and its extra w_empty arrival put the ring out of step for
          // good -- the driver's "wait 3016 / 4016 / 2000 timed out" (tools/repro_r1_trap.sh reproduces it with the
          // producer's fault injection).  Rule 1: the issuer without an M-tile in a visit waits for that visit's
          // accumulator, i.e. for the other issuer's MMAs on all of its chunks, before it moves on.  Rules 1 + 2 give
          // every w_full wait a happens-before edge from the landing of the chunk four positions earlier.
          if (my_mt >= n_mt && !(P.debug_flags & 128)) wait_tag(&acc_ready[t], par, 5000 + s);
          prev_t = t;
          prev_par = par;
          cnt += n_mt * nkc;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: 16 warps; group g = warps 4-11 / 12-19 owns tile slot g =====================
    // The owner runs the tile's prologue, its sdf read-out and the Jacobian step.  The wide steps (activation, feature,
    // backward scaling) of BOTH tiles are drained by BOTH groups, tile 0 first, the owner taking columns [0, 64) and the
    // other group [64, 128): the activation epilogue is what the tensor pipe waits for, and with it bound to the tile's
    // own 8 warps only half of the epilogue warps have work while the other tile's MMAs run.
    const int e = warp - kEpiWarp0;                 // 0..15
    const int g = kFat ? 0 : e >> 3;                // group = the tile slot it owns (kFat: one group owns both)
    const int mo = (e >> 2) & 1;                    // M-tile this warpgroup owns
    const int q = warp & 3;                         // TMEM lane quarter (hardware: warp id % 4)
    const int etid = (e & 7) * 32 + lane;           // 0..255 inside the group
    const int F = mo * 128 + 32 * q + lane;         // feature = TMEM lane of the owned M-tile
    uint32_t acc_par = 0;                           // bit t: parity of acc_ready[t]
    const int pe_dim = P.multires < 0 ? 3 : 3 + 6 * P.multires;
    constexpr int kCh = kShare ? 4 : 8;             // 16-column chunks of a tile this warp drains per wide step
    constexpr int kVisits = (kShare || kFat) ? 2 : 1;   // tile slots a warp visits per step
    int tcnt = 0;
    const int treg = e == 0 ? 2 : e == 3 ? 3 : e == 8 ? 4 : e == 15 ? 5 : e == 5 ? 6 : -1;   // traced warps (test twin)
    const bool tracer = kProbe && lane == 0 && treg >= 0;

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
#pragma unroll 1
      for (int pg = g; pg < (kFat ? 2 : g + 1); ++pg) {
        // ---- prologue (owner): stage the points, evaluate the embedding into operand rows [0, k0) and the stash ----
        const int64_t p0 = (2 * pair + pg) * 128;
        uint8_t* act = smem + SmemRev::act + pg * kActBytes;
        float* xs = (float*)(smem + SmemRev::xs) + pg * 512;
        uint8_t* pes = smem + SmemRev::pes + pg * (kPeStashRows * 256);
        for (int i = etid; i < 384; i += kEpiPerTile) {
          const int64_t gi = p0 * 3 + i;
          xs[i] = gi < a.n * 3 ? a.x[gi] : 0.0f;
        }
        named_bar_sync(1 + g, kEpiPerTile);
        const int n = etid & 127;                      // operand column = point
        const int part = etid >> 7;                    // 0..1: splits the rows
        const float x3[3] = {xs[3 * n], xs[3 * n + 1], xs[3 * n + 2]};
        const int k0 = P.steps[0].k_steps * 16;
        uint16_t* stash = reinterpret_cast<uint16_t*>(pes) + n;   // row j at stash[j * 128]
        auto put = [&](int j, float val) {
          store_elem<kF16>(act, j, n, val);
          if (j < kPeStashRows) stash[j * 128] = umma::pack1<kF16>(val);
        };
        if (part == 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) put(j, x3[j]);
        }
        for (int qf = part; qf < P.multires; qf += 2) {
          const float f = (float)(1 << qf);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            __sincosf(x3[c] * f, &sn, &cs);
            put(3 + 6 * qf + c, sn);
            put(3 + 6 * qf + 3 + c, cs);
          }
        }
        for (int j = pe_dim + part; j < k0; j += 2) store_elem<kF16>(act, j, n, 0.f);
        publish(&in_ready[pg]);
        if (kShare && lane == 0) umma::mbar_arrive(&in_ready[pg ^ 1]);   // nothing of the other tile's operand is ours yet
      }

      float b_ahead = a.bias[P.steps[0].bias_off + F];
      for (int s = 0; s < P.n_steps; ++s) {
        const nr_umma_step_t& S = P.steps[s];
        const bool mine = mo < S.n_mt;
        const bool is_h = F < S.out_rows;
        const bool is_pe = S.pe_fill && F >= S.out_rows && F < S.out_rows + pe_dim;
        // this layer's bias (both slots use it), requested a layer ahead: loaded inside the visit it sat exposed behind the
        // accumulator wait (ncu: 170 cycles per forward visit); 1.042 -> 1.029 ms ('rev'), 1.13 -> 1.09 ms ('rev_img')
        const float b_step = b_ahead;
        if (s + 1 < P.n_steps && P.steps[s + 1].epi == EPI_HIDDEN) b_ahead = a.bias[P.steps[s + 1].bias_off + F];
#pragma unroll 1
        for (int v = 0; v < kVisits; ++v) {
          const int t = kVisits == 2 ? v : g;         // tile slot worked on
          const bool own = kFat || t == g;
          const int c0 = kShare && !own ? 4 : 0;      // first chunk of this warp's share
          const int64_t tile = 2 * pair + t;
          const int64_t p0 = tile * 128;
          uint8_t* act = smem + SmemRev::act + t * kActBytes;
          float* xs = (float*)(smem + SmemRev::xs) + t * 512;
          uint8_t* pes = smem + SmemRev::pes + t * (kPeStashRows * 256);
          const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * 256 + mo * 128) + 16u * c0;
          const RowAddr ra(umma::smem_u32(act), F);
          uint8_t* sig_slot = a.sig + (((size_t)blockIdx.x * 2 + t) * a.n_sig + S.sig_slot) * kSigBytes + (size_t)F * 16 +
                              (size_t)c0 * kSigChunk;   // written (EPI_HIDDEN) and read (EPI_BWD, EPI_SDF_OUT) by this thread

          // softplus' codes of this warp's columns: requested before the accumulator wait (latency under the MMAs)
          uint4 sg[kCh];
          const bool use_sig = ((S.epi == EPI_BWD && mine && is_h) || (S.epi == EPI_SDF_OUT && s + 1 < P.n_steps)) && !(P.debug_flags & 2);
          if (use_sig) {
#pragma unroll
            for (int k = 0; k < kCh; ++k) sg[k] = ldcg16(sig_slot + k * kSigChunk);
          }

          wait_tag(&acc_ready[t], (acc_par >> t) & 1u, 4000 + s);
          acc_par ^= 1u << t;
          umma::tc_fence_after();
          if (tracer) trace_ev(a.trace, treg, tcnt, 21, s * 2 + t, pair);

          if (S.epi == EPI_HIDDEN) {
            if (mine && !(P.debug_flags & 8)) {
              const float b = b_step;
              const f32x2 b144 = splat2(b * (NR_SIG_TANH >= 2 ? 50.0f : 144.26950408889634f));
              const int jpe = F - S.out_rows;
              // loop-invariant decisions and addresses, taken out of the per-chunk code (it costs issue slots there: the
              // forward epilogue is bound by what its FMA / ALU pipes and issue port can take, profiles/r2_mlp_rev_epilogue.md)
              const bool do_img = S.to_rad && a.feat_img && tile < n_tiles;
              uint8_t* const img_row = a.feat_img + (size_t)tile * kActBytes + (F >> 3) * 1024 + (F & 7) * 128;
              const bool no_codes = !a.store_sig || (kProbe && (P.debug_flags & 1));
              uint32_t raw[16], rawB[16];
              auto values = [&](const uint32_t (&r)[16], int k) {
                const int c = c0 + k;
                if (!is_pe) {
                  float vv[16];
                  f32x2 d2[8];
                  if (kProbe && (P.debug_flags & 32)) {      // probe: no activation math
#pragma unroll
                    for (int j = 0; j < 8; ++j) { vv[2 * j] = __uint_as_float(r[2 * j]); vv[2 * j + 1] = __uint_as_float(r[2 * j + 1]); d2[j] = 0ull; }
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                      if (NR_SIG_TANH >= 2) softplus_th2<NR_SIG_TANH == 3>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), b144, vv[2 * j], vv[2 * j + 1], d2[j]);
                      else if (NR_SIG_TANH) softplus_sigt2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), b144, vv[2 * j], vv[2 * j + 1], d2[j]);
                      else softplus_sigq2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), b144, vv[2 * j], vv[2 * j + 1], d2[j]);
                  }
                  store_row16<kF16>(ra, 16 * c, vv, kProbe && (P.debug_flags & 16));   // probe: no operand stores
                  if (do_img) img_row_store16<kF16>(img_row, F & 7, 16 * c, vv);
                  if (!no_codes) {
                    const uint4 w = make_uint4(sig_pack4(d2[0], d2[1]), sig_pack4(d2[2], d2[3]), sig_pack4(d2[4], d2[5]),
                                               sig_pack4(d2[6], d2[7]));
                    stcg16(sig_slot + k * kSigChunk, w);
                  }
                } else {
                  copy_row16(ra, pes, jpe, 16 * c);
                }
              };
              const bool no_ld = kProbe && (P.debug_flags & 256);              // probe: no TMEM loads
              if (no_ld) {
#pragma unroll
                for (int j = 0; j < 16; ++j) raw[j] = rawB[j] = 0x3c23d70au + j;
              }
              if (!no_ld) umma::tmem_ld16(taddr, raw);
#pragma unroll
              for (int k = 0; k < kCh; k += 2) {
                umma::tmem_ld_wait();
                if (!no_ld) umma::tmem_ld16(taddr + 16 * (k + 1), rawB);
                if (tracer && k == 0) trace_ev(a.trace, treg, tcnt, 24, s * 2 + t, pair);
                values(raw, k);
                umma::tmem_ld_wait();
                if (k + 2 < kCh && !no_ld) umma::tmem_ld16(taddr + 16 * (k + 2), raw);
                values(rawB, k + 1);
                if (tracer && k == 0) trace_ev(a.trace, treg, tcnt, 25, s * 2 + t, pair);
              }
            }
          } else if (S.epi == EPI_FEAT) {
            if (mine) {
              const float b = a.bias[S.bias_off + F];
#pragma unroll 1
              for (int k = 0; k < kCh; ++k) {
                uint32_t raw[16];
                umma::tmem_ld16(taddr + 16 * k, raw);
                umma::tmem_ld_wait();
                if (a.feat && F < S.out_rows) {
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const int64_t gp = p0 + 16 * (c0 + k) + j;
                    if (gp < a.n) a.feat[gp * a.feat_ld + F] = __uint_as_float(raw[j]) + b;
                  }
                }
              }
            }
          } else if (S.epi == EPI_SDF_OUT) {
            // rows 0..31 of M-tile 0 all hold the sdf row: lane l keeps column l of a 32-column chunk.  The four chunks are split
            // over the warps of lane quarter 0 that visit this tile (one warp did all four, with a 32-deep select chain each: 2.7 k
            // cycles on the critical path of the tile's first backward step)
            if (q == 0 && a.sdf) {
              constexpr int kQ0 = kShare ? 4 : 2;                      // warps of quarter 0 visiting the tile
              const int iq = kShare ? (e >> 2) : ((e & 7) >> 2);
              const float b = a.bias[S.bias_off];
#pragma unroll 1
              for (int c = iq * (4 / kQ0); c < (iq + 1) * (4 / kQ0); ++c) {
                uint32_t raw[32];
                umma::tmem_ld32(tmem_base + (uint32_t)(t * 256 + 32 * c), raw);
                umma::tmem_ld_wait();
                uint32_t s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                for (int j = 0; j < 16; ++j) s16[j] = (lane & 1) ? raw[2 * j + 1] : raw[2 * j];
#pragma unroll
                for (int j = 0; j < 8; ++j) s8[j] = (lane & 2) ? s16[2 * j + 1] : s16[2 * j];
#pragma unroll
                for (int j = 0; j < 4; ++j) s4[j] = (lane & 4) ? s8[2 * j + 1] : s8[2 * j];
                s2[0] = (lane & 8) ? s4[1] : s4[0];
                s2[1] = (lane & 8) ? s4[3] : s4[2];
                const float m = __uint_as_float((lane & 16) ? s2[1] : s2[0]);
                const int64_t gp = p0 + 32 * c + lane;
                if (gp < a.n) a.sdf[gp] = m + b;
              }
            }
            // start of the backward pass: operand row F <- softplus'(z_last)[F, :] * w_sdf[F]  (d sdf / d z_last)
            if (s + 1 < P.n_steps) {
              const float wF = a.bias[S.aux_off + F];
              const uint32_t none[16] = {};
#pragma unroll
              for (int k = 0; k < kCh; ++k) apply_sig16<kF16, false>(ra, c0 + k, none, sg[k], wF);
            }
          } else if (S.epi == EPI_BWD) {
            if (mine && !(P.debug_flags & 4)) {
              uint32_t raw[16], rawB[16];
              const int jpe = F - S.out_rows;
              auto apply = [&](const uint32_t (&r)[16], int k) {
                const int c = c0 + k;
                if (is_h) {
                  apply_sig16<kF16, true>(ra, c, r, sg[k], 1.0f);
                } else {
                  if (is_pe) {   // gradient w.r.t. the skip connection's copy of the embedding: kept for EPI_NABLA
                    uint32_t h[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) h[j] = umma::pack2<kF16>(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                    uint4* dst = reinterpret_cast<uint4*>(pes + jpe * 256 + c * 32);
                    dst[0] = make_uint4(h[0], h[1], h[2], h[3]);
                    dst[1] = make_uint4(h[4], h[5], h[6], h[7]);
                  }
                  st_shared_v4(ra.chunk(2 * c), 0, 0, 0, 0);       // these rows meet zero weights; keep them finite
                  st_shared_v4(ra.chunk(2 * c + 1), 0, 0, 0, 0);
                }
              };
              umma::tmem_ld16(taddr, raw);
#pragma unroll
              for (int k = 0; k < kCh; k += 2) {
                umma::tmem_ld_wait();
                umma::tmem_ld16(taddr + 16 * (k + 1), rawB);
                apply(raw, k);
                umma::tmem_ld_wait();
                if (k + 2 < kCh) umma::tmem_ld16(taddr + 16 * (k + 2), raw);
                apply(rawB, k + 1);
              }
            }
          } else if (S.epi == EPI_NABLA && own) {
            // acc rows [0, pe_dim) = d sdf / d PE(x) through layer 0 (+ the skip layer's share from the stash);
            // nabla_c = sum_j dPE_j/dx_c * g_j over a [point][row] scratch in the (now free) operand buffer.
            // Phase A, all 256 threads: the Jacobian dPE_j/dx of every row and point, one sincos per (point, frequency,
            // component).  (Until late in round 2 every (row, point) element evaluated its own sincos and read its stash value
            // through a 32-way bank conflict, on four warps: 18 - 23 k cycles per tile pair, a seventh of the kernel.)
            float* red = reinterpret_cast<float*>(act);
            {
              const int n = etid & 127, part = etid >> 7;
              const float x3[3] = {xs[3 * n], xs[3 * n + 1], xs[3 * n + 2]};
              float* rn = red + n * kRedLd;
              if (part == 0) rn[0] = rn[1] = rn[2] = 1.0f;
              for (int qf = part; qf < P.multires; qf += 2) {
                const float f = (float)(1 << qf);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  float sn, cs;
                  __sincosf(x3[c] * f, &sn, &cs);
                  rn[3 + 6 * qf + c] = f * cs;
                  rn[3 + 6 * qf + 3 + c] = -f * sn;
                }
              }
            }
            named_bar_sync(1 + g, kEpiPerTile);
            // Phase B: scale by the gradient rows (TMEM lane = row; both warps of a lane quarter work, each on half of the columns)
            if (32 * q < pe_dim) {   // warp-uniform (the TMEM loads are .sync.aligned)
              const int Rr = 32 * q + lane;
              const uint32_t taddr0 = taddr - (uint32_t)(mo * 128);
#pragma unroll 1
              for (int c = 4 * mo; c < 4 * mo + 4; ++c) {
                uint32_t raw[16];
                umma::tmem_ld16(taddr0 + 16 * c, raw);
                umma::tmem_ld_wait();
                if (Rr < pe_dim) {
                  uint32_t hv[8] = {};
                  if (S.pe_fill) {   // 16 stash values as two 16-byte loads
                    const uint4* sp = reinterpret_cast<const uint4*>(pes + Rr * 256 + c * 32);
                    const uint4 s0 = sp[0], s1 = sp[1];
                    hv[0] = s0.x; hv[1] = s0.y; hv[2] = s0.z; hv[3] = s0.w; hv[4] = s1.x; hv[5] = s1.y; hv[6] = s1.z; hv[7] = s1.w;
                  }
#pragma unroll
                  for (int j = 0; j < 16; ++j) {
                    const int col = 16 * c + j;
                    float gv = __uint_as_float(raw[j]);
                    if (S.pe_fill) {
                      const uint16_t h16 = (uint16_t)(hv[j >> 1] >> (16 * (j & 1)));
                      gv += kF16 ? __half2float(__ushort_as_half(h16)) : __uint_as_float((uint32_t)h16 << 16);
                    }
                    red[col * kRedLd + Rr] *= gv;
                  }
                }
              }
            }
            named_bar_sync(1 + g, kEpiPerTile);
            if (etid < 128) {
              const float* r = red + etid * kRedLd;
              float g0 = r[0], g1 = r[1], g2 = r[2];
              for (int j = 3; j < pe_dim; j += 3) { g0 += r[j]; g1 += r[j + 1]; g2 += r[j + 2]; }
              const int64_t gp = p0 + etid;
              if (gp < a.n) { a.nabla[gp * 3] = g0; a.nabla[gp * 3 + 1] = g1; a.nabla[gp * 3 + 2] = g2; }
            }
          }
          if (NR_SIG_DISCARD) {
            __syncwarp();                                   // every lane of the warp has consumed its codes (use_sig differs
            if (use_sig && (lane & 7) == 0) {               // between lanes around out_rows: the barrier stays outside)
#pragma unroll
              for (int k = 0; k < kCh; ++k) discard_line(sig_slot + k * kSigChunk);
            }
          }
          if (tracer) trace_ev(a.trace, treg, tcnt, 22, s * 2 + t, pair);
          if (s + 1 < P.n_steps) publish(&in_ready[t]);
          if (tracer) trace_ev(a.trace, treg, tcnt, 23, s * 2 + t, pair);
        }
      }
      umma::tc_fence_before();
      // The owner's staging buffers and TMEM slot are free before its next prologue.  Its own group suffices also when
      // the wide steps are shared: the other group last touched this tile in the last backward step, which completed
      // (all 16 arrivals on in_ready) before the Jacobian step's MMAs were even issued.
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

#ifdef NR_FAULT_INJECT
static long long* g_rev_trace = nullptr;
// test twin only: timestamps of the MMA <-> epilogue hand-offs of CTA 0 ([8][1024][4] int64, device memory)
extern "C" int nr_mlp_rev_set_trace(void* buf) { g_rev_trace = (long long*)buf; return NR_OK; }
#else
static long long* const g_rev_trace = nullptr;
#endif

extern "C" size_t nr_mlp_umma_reverse_workspace(const nr_umma_program_t* prog, int64_t n) {
  if (!prog || n <= 0) return 0;
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  (void)n;   // sized for the largest grid, so one buffer serves every n
  return (size_t)sms * 2 * count_sig_slots(prog) * kSigBytes;
}

extern "C" int nr_mlp_umma_reverse(const nr_umma_program_t* prog, const void* image, size_t image_bytes,
                                   const float* bias, size_t bias_floats, const float* x, int64_t n, float* sdf,
                                   float* nabla, float* feat, int64_t feat_ld, void* feat_img, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  NR_CHECK_ARG(prog && image && bias && x, "nr_mlp_umma_reverse: null pointer");
  NR_CHECK_ARG(n >= 0, "nr_mlp_umma_reverse: n < 0");
  NR_CHECK_ARG(prog->reverse == 1 && prog->tangents == 0 && prog->input_mode == 0,
               "nr_mlp_umma_reverse: needs a reverse-mode program on value tiles");
  NR_CHECK_ARG(prog->n_steps >= 2 && prog->n_steps <= NR_UMMA_MAX_STEPS, "nr_mlp_umma_reverse: n_steps=%d", prog->n_steps);
  NR_CHECK_ARG(((uintptr_t)image & 15) == 0 && ((uintptr_t)workspace & 15) == 0 && ((uintptr_t)feat_img & 15) == 0,
               "nr_mlp_umma_reverse: image, workspace and feat_img must be 16-byte aligned");
  const int pe_dim = prog->multires < 0 ? 3 : 3 + 6 * prog->multires;
  NR_CHECK_ARG(pe_dim <= kPeStashRows && pe_dim % 3 == 0, "nr_mlp_umma_reverse: embedding of %d rows", pe_dim);
  NR_CHECK_ARG(prog->steps[0].k_steps * 16 >= pe_dim, "nr_mlp_umma_reverse: step 0 K does not cover the embedding");
  const int n_sig = count_sig_slots(prog);
  // order: hidden layers, optional feature, sdf row, backward layers, embedding Jacobian
  int phase = 0, n_sdf = 0, n_nabla = 0;
  for (int s = 0; s < prog->n_steps; ++s) {
    const nr_umma_step_t& S = prog->steps[s];
    const int nch = S.n_mt * (S.k_steps / 4);
    NR_CHECK_ARG(S.k_steps % 4 == 0 && S.k_steps >= 4 && S.k_steps <= 16, "step %d: k_steps=%d", s, S.k_steps);
    NR_CHECK_ARG(S.n_mt >= 1 && S.n_mt <= 2, "step %d: n_mt=%d", s, S.n_mt);
    NR_CHECK_ARG(S.n_cols == 128 && !S.accumulate, "step %d: reverse-mode steps are 128 columns wide, no split-K", s);
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
  // a program that ends with the sdf row is the forward sweep alone (sdf [+ feature] of every point: no codes, no normals)
  const bool fwd_only = prog->steps[prog->n_steps - 1].epi == EPI_SDF_OUT;
  NR_CHECK_ARG(n_sdf == 1 && n_sig >= 1 && (fwd_only ? n_nabla == 0 : n_nabla == 1),
               "nr_mlp_umma_reverse: program needs one EPI_SDF_OUT, last or followed by the backward sweep up to a final EPI_NABLA");
  NR_CHECK_ARG(fwd_only || nabla, "nr_mlp_umma_reverse: nabla is null");
  if (n == 0) return NR_OK;
  if (!fwd_only) {
    const size_t need = nr_mlp_umma_reverse_workspace(prog, n);
    NR_CHECK_ARG(workspace && workspace_bytes >= need, "nr_mlp_umma_reverse: workspace of %zu bytes needed, %zu given", need,
                 workspace_bytes);
  }
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_pairs = (nr_cdiv(n, 128) + 1) / 2;
  const int grid = (int)(n_pairs < sms ? n_pairs : sms);
  const size_t smem = SmemRev::total + 1024;
  static unsigned long long attr_set = 0;  // per-device bit: the attribute is per (function, device)
  if (!(attr_set >> (dev & 63) & 1ull)) {
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_rev_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_rev_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_rev_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_rev_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_rev_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_rev_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set |= 1ull << (dev & 63);
  }
  DevProgram dp;
  dp.p = *prog;
  RevArgs ka{(const uint8_t*)image, bias, x, n, sdf, nabla, feat, feat_ld, (uint8_t*)feat_img, (uint8_t*)workspace, n_sig, fwd_only ? 0 : 1, g_rev_trace};
  // NEURECON_B200_REV_SHARE (measurements): 0 = every tile's epilogue on its own 8 warps, 1 = both 8-warp groups on every tile,
  // 2 = eight 168-register epilogue warps in all
  static int mode = -1;
  if (mode < 0) { const char* ev = getenv("NEURECON_B200_REV_SHARE"); mode = ev ? atoi(ev) : NR_REV_DEFAULT_MODE; if (mode < 0 || mode > 2) mode = NR_REV_DEFAULT_MODE; }
  const cudaStream_t st = (cudaStream_t)stream;
  if (prog->operand_f16) {
    if (mode == 2) mlp_rev_kernel<true, 2><<<grid, kThreadsFat, smem, st>>>(dp, ka);
    else if (mode == 1) mlp_rev_kernel<true, 1><<<grid, kThreads, smem, st>>>(dp, ka);
    else mlp_rev_kernel<true, 0><<<grid, kThreads, smem, st>>>(dp, ka);
  } else {
    if (mode == 2) mlp_rev_kernel<false, 2><<<grid, kThreadsFat, smem, st>>>(dp, ka);
    else if (mode == 1) mlp_rev_kernel<false, 1><<<grid, kThreads, smem, st>>>(dp, ka);
    else mlp_rev_kernel<false, 0><<<grid, kThreads, smem, st>>>(dp, ka);
  }
  NR_CHECK_LAUNCH("mlp_rev_kernel");
  return NR_OK;
}
