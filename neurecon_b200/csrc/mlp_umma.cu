// Fused positional-encoding + SDF MLP (+ analytic normals) + radiance MLP on tcgen05 / TMEM.
//
// Orientation: D[feature, column] = W[feature, k] * H[k, column].
//   * A operand = weights, streamed per layer from a pre-swizzled bf16 image in global memory
//     (L2 resident, ~1.6 MB) by 16 KB bulk async copies (TMA unit) into a 4-stage smem ring.
//   * B operand = activations, RESIDENT in shared memory for the whole network: every layer's
//     epilogue writes the next layer's operand in place (MN-major, 128-byte swizzle), so hidden
//     activations never touch HBM.  PE is evaluated on-chip from the 12-byte point.
//   * Accumulators live in TMEM: 2 point-tiles x 2 M-tiles x 128 columns = all 512 columns.
//   * A column is one (point, component): with normals a tile holds 32 points x {value, d/dx,
//     d/dy, d/dz} (forward-mode tangents ride through the same weights); without, 128 points.
//     TMEM lane = feature, so one thread owns a feature for all columns: the tangent epilogue
//     t' = softplus'(z) * (W t) is register-local, no shuffles.
//   * Warp roles (640 threads): warp 0 bulk-copy producer, warps 1 and 3 MMA issuers (one elected
//     lane each, one M-tile each), warp 2 TMEM allocator, warps 4-11 / 12-19 epilogue warps of tile A / tile B (one warpgroup
//     per M-tile, so every SM sub-partition holds 4 epilogue warps to hide TMEM-load and MUFU
//     latency).  The two tiles ping-pong: while the tensor core runs layer l of tile B, tile A's
//     warps apply layer l's activation and publish its next operand.
//
// Reference semantics: models/base.py:46-64 (Embedder), :243-282 (ImplicitSurface.forward /
// forward_with_nablas), :372-391 (RadianceNet.forward).
#include "mlp_epilogue.cuh"

namespace {

constexpr int kStages = 4;             // weight-ring depth (3 / 4 / 5 measured: profiles/mlp_umma_r1_history.md; 5 does not fit with the stash)
constexpr int kStagesLog2 = 2;
static_assert(kStages == (1 << kStagesLog2), "the ring position arithmetic assumes a power-of-two ring");
constexpr int kThreads = 640;            // 4 control warps + 2 tiles x 8 epilogue warps
constexpr int kEpiPerTile = 256;
constexpr int kEpiWarpsPerTile = 8;
constexpr int kEpiWarp0 = 4;
// kShare (template parameter of the kernel): the wide activation steps (EPI_HIDDEN, EPI_RELU) of a tile are drained by all 16
// epilogue warps, each group taking half of the points; otherwise by the tile's own 8 warps.  Chosen per program by the host:
// the softplus programs on value tiles (sdf only) gain 9 % from sharing, the ReLU programs (radiance pass) lose 14 %
// (tools/bench_mlp.py, round 2); NR_SHARE_EPILOGUE = 0 / 1 forces it for measurements.
#ifndef NR_SHARE_EPILOGUE
#define NR_SHARE_EPILOGUE -1
#endif

struct SmemLayout {
  // offsets from the 1024-aligned base
  static constexpr uint32_t act = 0;                                 // 2 x 64 KB
  static constexpr uint32_t ring = 2 * kActBytes;                    // kStages x 16 KB
  static constexpr uint32_t xs = ring + kStages * kChunkBytes;       // 2 x 128 x (3 | 4) floats
  static constexpr uint32_t vs = xs + 2 * 512 * 4;                   // 2 x 128 x 3 floats (view dirs)
  static constexpr uint32_t nabs = vs + 2 * 384 * 4;                 // 2 x 128 x 3 floats (normal stash)
  static constexpr uint32_t pes = nabs + 2 * 384 * 4;                // 2 x 40 rows x 256 B: embedding stash (skip)
  static constexpr uint32_t bars = pes + 2 * kPeStashRows * 256;     // mbarriers
  static constexpr uint32_t total = bars + 256;
};

// Epilogue -> MMA hand-off: every lane makes its generic-proxy smem writes visible to the async
// proxy and orders its TMEM reads, then one lane per warp arrives (8 arrivals per tile).
__device__ __forceinline__ void publish(uint64_t* bar, int debug_flags, uint32_t count = 1) {
  if (!(debug_flags & 32)) umma::fence_proxy_async_smem();
  umma::tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive_n(bar, count);
}

constexpr int kTraceCap = 2048;
__device__ __forceinline__ void trace_ev(long long* tr, int region, int& cnt, int ev, int st, long long pair) {
  if (!tr || blockIdx.x != 0 || cnt >= kTraceCap) return;
  long long* p = tr + ((size_t)region * kTraceCap + cnt) * 4;
  p[0] = ev; p[1] = st; p[2] = clock64(); p[3] = pair;
  ++cnt;
}

template <bool kF16, bool kShare>
__global__ void __launch_bounds__(kThreads, 1) mlp_umma_kernel(const __grid_constant__ DevProgram prog, const KArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + SmemLayout::bars);
  uint64_t* w_full = bars;                   // [kStages]
  uint64_t* w_empty = bars + kStages;        // [kStages]
  uint64_t* in_ready = bars + 2 * kStages;   // [2]
  uint64_t* acc_ready = in_ready + 2;        // [2]
  uint64_t* feat_full = acc_ready + 2;       // [2] radiance-only mode: feature image landed in the operand buffer
  __shared__ uint32_t tmem_base_s;

  const nr_umma_program_t& P = prog.p;
  // warp index through a shuffle: tells the compiler it is warp-uniform, so the role loops below keep their ring
  // positions, barrier addresses and MMA descriptors in uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const int tang = P.tangents;
  const bool rad_only = P.input_mode == 1;
  const bool nerf = P.input_mode == 2;
  const int xdim = nerf ? P.input_dim : 3;
  const int ppt = tang ? 32 : 128;                       // points per tile
  const int64_t n_tiles = (a.n + ppt - 1) / ppt;
  const int64_t n_pairs = (n_tiles + 1) / 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { umma::mbar_init(&w_full[s], 1); umma::mbar_init(&w_empty[s], 1); }
    for (int t = 0; t < 2; ++t) {
      umma::mbar_init(&in_ready[t], kShare ? 2 * kEpiWarpsPerTile : kEpiWarpsPerTile);
      umma::mbar_init(&acc_ready[t], 2);
      umma::mbar_init(&feat_full[t], 1);
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
    // ===================== weight producer (warp-uniform loop, one elected lane copies) ==========
    // Ring position = running chunk count: stage = cnt % kStages, phase parity = (cnt / kStages) & 1.
    uint32_t cnt = 0;
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int ntl = (2 * pair + 1 < n_tiles) ? 2 : 1;
      for (int s = 0; s < P.n_steps; ++s) {
        const int nch = P.steps[s].n_mt * (P.steps[s].k_steps >> 2);
        const uint8_t* src = a.image + (size_t)P.steps[s].chunk_begin * kChunkBytes;
        for (int t = 0; t < ntl; ++t) {
          for (int c = 0; c < nch; ++c, ++cnt) {
            const uint32_t stage = cnt & (kStages - 1);
            umma::mbar_wait(&w_empty[stage], ((cnt >> kStagesLog2) & 1u) ^ 1u, 1000 + s);
            NR_INJECT_DELAY(P.debug_flags, P.steps[s].n_mt, t);
            if (umma::elect_one()) {
              if (P.debug_flags & 1) {
                umma::mbar_arrive(&w_full[stage]);
              } else {
                umma::mbar_arrive_expect_tx(&w_full[stage], kChunkBytes);
                umma::bulk_g2s(smem + SmemLayout::ring + stage * kChunkBytes, src + (size_t)c * kChunkBytes,
                               kChunkBytes, &w_full[stage]);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 || warp == 3) {
    // ===================== MMA issuers: warp 1 owns M-tile 0, warp 3 owns M-tile 1 =====================
    // Two issuing warps: the tensor pipe accepts an MMA only every 64 cycles (N = 128) with a queue about two deep
    // (tools/umma_queue_depth.py), so whatever else the issuing thread does between MMAs (barrier probes, descriptor
    // arithmetic, at a fifth of the issue slots of a sub-partition shared with four epilogue warps) must fit
    // under the other warp's MMAs.  Different M-tiles are different accumulators, so the relative order of the two
    // warps' MMAs is irrelevant.  Chunks are interleaved (k-chunk major, M-tile minor) in the ring: warp w consumes
    // chunk kc*n_mt + w of every step.
    const uint32_t my_mt = warp == 1 ? 0u : 1u;
    int tcnt = 0;
    uint32_t cnt = 0;                  // chunks consumed by the CTA before the current (step, tile)
    uint32_t in_par = 0;               // bit t: parity of in_ready[t]
    int prev_t = -1;                   // tile slot and acc_ready parity of the previous (step, tile) visit
    uint32_t prev_par = 0;
    const uint32_t a_hi = umma::smem_desc_hi(1024), b_hi = umma::smem_desc_hi(1024);
    const uint32_t ring_lo = umma::smem_desc_lo(umma::smem_u32(smem + SmemLayout::ring), 16);
    const uint32_t act_lo0 = umma::smem_desc_lo(umma::smem_u32(smem + SmemLayout::act), kLbo);
    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int ntl = (2 * pair + 1 < n_tiles) ? 2 : 1;
      for (int s = 0; s < P.n_steps; ++s) {
        const uint32_t n_mt = P.steps[s].n_mt, nkc = P.steps[s].k_steps >> 2;
        const uint32_t acc0 = P.steps[s].accumulate != 0 ? 1u : 0u;
        const uint32_t idesc = kF16 ? umma::make_idesc_f16(128, P.steps[s].n_cols, 0, 1)
                                    : umma::make_idesc_bf16(128, P.steps[s].n_cols, 0, 1);
        for (int t = 0; t < ntl; ++t) {
          const uint32_t par = (in_par >> t) & 1u;   // in_ready[t] and acc_ready[t] complete one phase per visit of tile t
          umma::mbar_wait(&in_ready[t], par, 2000 + s);
          in_par ^= 1u << t;
          umma::tc_fence_after();
          // ring lockstep, rule 2 (mlp_rev.cu has the full story): before warp 1 issues a single-M-tile step from stages
          // whose previous chunks were warp 3's, the previous visit's accumulator must be complete
          if (n_mt == 1 && my_mt == 0 && prev_t >= 0 && !(P.debug_flags & 128)) umma::mbar_wait(&acc_ready[prev_t], prev_par, 6000 + s);
          if (lane == 0) trace_ev(a.trace, my_mt, tcnt, 11, s * 2 + t, pair);
          if (my_mt < n_mt) {
            const uint32_t d_addr = tmem_base + (uint32_t)t * 256u + my_mt * 128u;
            uint32_t b_lo = act_lo0 + (uint32_t)t * (kActBytes >> 4);
            uint32_t c = cnt + my_mt;
#pragma unroll 1
            for (uint32_t kc = 0; kc < nkc; ++kc, c += n_mt, b_lo += 512) {
              const uint32_t st = c & (kStages - 1);
              umma::mbar_wait(&w_full[st], (c >> kStagesLog2) & 1u, 3000 + s);
              umma::tc_fence_after();
              const uint32_t a_lo = ring_lo + st * (kChunkBytes >> 4);
              if (umma::elect_one()) {
                umma::mma_bf16_ss(d_addr, umma::desc64(a_lo, a_hi), umma::desc64(b_lo, b_hi), idesc, (kc | acc0) ? 1u : 0u);
                umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 2, a_hi), umma::desc64(b_lo + 128, b_hi), idesc, 1u);
                umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 4, a_hi), umma::desc64(b_lo + 256, b_hi), idesc, 1u);
                umma::mma_bf16_ss(d_addr, umma::desc64(a_lo + 6, a_hi), umma::desc64(b_lo + 384, b_hi), idesc, 1u);
                umma::mma_commit(&w_empty[st]);
              }
              __syncwarp();
            }
          }
          if (umma::elect_one()) umma::mma_commit(&acc_ready[t]);
          __syncwarp();
          if (lane == 0) trace_ev(a.trace, my_mt, tcnt, 12, s * 2 + t, pair);
          // ring lockstep, rule 1: the issuer without an M-tile in this visit waits for the visit's accumulator (the other
          // issuer's MMAs on all of its chunks), so that it never meets a w_full barrier a phase early
          if (my_mt >= n_mt && !(P.debug_flags & 128)) umma::mbar_wait(&acc_ready[t], par, 5000 + s);
          prev_t = t;
          prev_par = par;
          cnt += n_mt * nkc;
        }
      }
    }
  } else if (warp >= kEpiWarp0) {
    // ===================== epilogue: 16 warps; group g = e >> 3 owns tile slot g =====================
    // The owner group runs the tile's prologue and its narrow output steps.  With kShare the wide activation steps
    // are drained by BOTH groups (owner: first half of the points, other group: second half): a step's activation is
    // a latency chain (TMEM load -> MUFU -> pack -> st.shared) that two warps per sub-partition cannot hide, and
    // halving it is what lets the other tile's MMA burst cover it.
    const int e = warp - kEpiWarp0;                 // 0..15
    const int g = e >> 3;                           // group
    const int mo = (e >> 2) & 1;                    // M-tile this warpgroup owns
    const int q = warp & 3;                         // TMEM lane quarter (hardware: warp id % 4)
    const int etid = (e & 7) * 32 + lane;           // 0..255 inside the group
    uint32_t feat_par = 0;
    bool img_ahead = false;                         // radiance pass: this tile's feature image was requested during the previous pair
    uint32_t acc_par = 0;                           // bit t: parity of acc_ready[t]
    int tcnt = 0;
    const bool tracer = (mo == 0 && q == 0 && lane == 0 && g == 0);
    const int pe_dim = P.multires < 0 ? 3 : 3 + 6 * P.multires;
    const bool no_st = P.debug_flags & 8;

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      const int ntl = (2 * pair + 1 < n_tiles) ? 2 : 1;
      if (!kShare && g >= ntl) continue;
      if (g < ntl) {
      const int t = g;
      const int64_t tile = 2 * pair + t;
      const int64_t p0 = tile * ppt;
      uint8_t* act = smem + SmemLayout::act + t * kActBytes;
      float* xs = (float*)(smem + SmemLayout::xs) + t * 512;
      float* vs = (float*)(smem + SmemLayout::vs) + t * 384;
      float* nabs = (float*)(smem + SmemLayout::nabs) + t * 384;
      uint8_t* pes = smem + SmemLayout::pes + t * (kPeStashRows * 256);

      // ---- prologue: stage the points, evaluate the embedding into operand rows [0, k0) ----
      for (int i = etid; i < ppt * 3; i += kEpiPerTile) {
        const int64_t gi = p0 * 3 + i;
        if (!nerf) xs[i] = gi < a.n * 3 ? a.x[gi] : 0.0f;
        if (a.view && (rad_only || nerf || i < 96)) vs[i] = gi < a.n * 3 ? a.view[gi] : 0.0f;
        if (rad_only) nabs[i] = gi < a.n * 3 ? a.nabla[gi] * (a.normal_scale ? a.normal_scale[i % 3] : 1.0f) : 0.0f;
      }
      if (nerf) {
        for (int i = etid; i < ppt * xdim; i += kEpiPerTile) {
          const int64_t gi = p0 * xdim + i;
          xs[i] = gi < a.n * xdim ? a.x[gi] : 0.0f;
        }
      }
      if (rad_only) {
        // operand rows [0,256) = this tile's 64 KB block of the feature image, four 16 KB bulk copies (already on their way if
        // the previous pair's colour step requested them: the 64 KB come from HBM, 4 - 5 k cycles that the prologue sat out)
        if (etid == 0 && !img_ahead) {
          umma::mbar_arrive_expect_tx(&feat_full[t], kActBytes);
          const uint8_t* src = a.feat_img + (size_t)tile * kActBytes;
#pragma unroll
          for (int c = 0; c < 4; ++c) umma::bulk_g2s(act + c * kChunkBytes, src + c * kChunkBytes, kChunkBytes, &feat_full[t]);
        }
        umma::mbar_wait(&feat_full[t], feat_par, 7000);
        feat_par ^= 1;
        named_bar_sync(1 + t, kEpiPerTile);
      } else if (nerf) {
        // NeRF++: operand rows [0, K0) = PE(x) of the tile's 128 points (x has xdim components), zero padded
        named_bar_sync(1 + t, kEpiPerTile);
        const int p = etid & 127, k0 = P.steps[0].k_steps * 16;
        const int npe = P.multires < 0 ? xdim : xdim * (1 + 2 * P.multires);
        for (int r = etid >> 7; r < k0; r += 2) store_elem<kF16>(act, r, p, r < npe ? pe_row_nd(r, P.multires, xs + xdim * p, xdim) : 0.0f);
      } else {
      named_bar_sync(1 + t, kEpiPerTile);
      {
        const int n = etid & 127;                      // operand column
        const int part = etid >> 7;                    // 0..1: splits the rows
        const int p = tang ? (n & 31) : n;
        const int ct = tang ? (n >> 5) - 1 : -1;       // -1: value column, 0..2: tangent component
        const float x3[3] = {xs[3 * p], xs[3 * p + 1], xs[3 * p + 2]};
        const int k0 = P.steps[0].k_steps * 16;
        uint16_t* stash = reinterpret_cast<uint16_t*>(pes) + n;   // row j at stash[j * 128]
        auto put = [&](int j, float val) {             // operand row j of layer 0 + copy kept for the skip layer
          store_elem<kF16>(act, j, n, val);
          if (j < kPeStashRows) stash[j * 128] = umma::pack1<kF16>(val);
        };
        if (part == 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) put(j, ct < 0 ? x3[j] : (ct == j ? 1.f : 0.f));
        }
        for (int qf = part; qf < P.multires; qf += 2) {   // one sincos per (frequency, component)
          const float f = (float)(1 << qf);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            float sv = 0.f, cv = 0.f;
            if (ct < 0 || ct == c) {
              float sn, cs;
              __sincosf(x3[c] * f, &sn, &cs);
              sv = ct < 0 ? sn : f * cs;
              cv = ct < 0 ? cs : -f * sn;
            }
            put(3 + 6 * qf + c, sv);
            put(3 + 6 * qf + 3 + c, cv);
          }
        }
        for (int j = pe_dim + part; j < k0; j += 2) store_elem<kF16>(act, j, n, 0.f);
      }
      }
      publish(&in_ready[t], P.debug_flags, kShare ? 2 : 1);
      }

      for (int s = 0; s < P.n_steps; ++s) {
       for (int t = kShare ? 0 : g; t < (kShare ? ntl : g + 1); ++t) {
        const nr_umma_step_t& S = P.steps[s];
        const bool own = (t == g);
        const int64_t tile = 2 * pair + t;
        const int64_t p0 = tile * ppt;
        uint8_t* act = smem + SmemLayout::act + t * kActBytes;
        float* xs = (float*)(smem + SmemLayout::xs) + t * 512;
        float* vs = (float*)(smem + SmemLayout::vs) + t * 384;
        float* nabs = (float*)(smem + SmemLayout::nabs) + t * 384;
        uint8_t* pes = smem + SmemLayout::pes + t * (kPeStashRows * 256);
        const uint32_t tmem_tile = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * 256);
        umma::mbar_wait(&acc_ready[t], (acc_par >> t) & 1u, 4000 + s);
        acc_par ^= 1u << t;
        umma::tc_fence_after();
        if (tracer) trace_ev(a.trace, 2, tcnt, 21, s * 2 + t, pair);
        const int F = mo * 128 + 32 * q + lane;         // feature = TMEM lane of the owned M-tile
        const uint32_t taddr = tmem_tile + (uint32_t)(mo * 128);
        const RowAddr ra(umma::smem_u32(act), F);

        if (P.debug_flags & 2) {
          // profiling: MMA + weight pipeline only
        } else if (S.epi == EPI_HIDDEN) {
          if (mo < S.n_mt) {
            const float b = a.bias[S.bias_off + F];
            const bool is_pe = S.pe_fill && F >= S.out_rows && F < S.out_rows + pe_dim;
            uint32_t raw[16];
            float v[16];
            if (P.debug_flags & 4) {
              for (int c = (kShare && !own) ? 4 : 0; c < ((kShare && own) ? 4 : 8); ++c) {
                umma::tmem_ld16(taddr + 16 * c, raw);
                umma::tmem_ld_wait();
                if (__uint_as_float(raw[0]) == 123.456f) act[0] = 1;
              }
            } else if (tang) {
              // columns: [0,32) values of 32 points, [32c, 32c+32) d/dx_c; two 16-point halves.  TMEM loads are
              // double-buffered: the load of chunk k+1 is in flight while chunk k is processed.
              uint32_t rawB[16];
              f32x2 sg[8];
              const int jpe = F - S.out_rows;
              const f32x2 b144 = splat2(b * 144.26950408889634f);
              auto tangent = [&](const uint32_t (&r)[16], int col0) {   // t' = sigmoid * (W t), 16 points of one component
                if (!is_pe) {
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    upk2(mul2(pk2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), sg[j]), v[2 * j], v[2 * j + 1]);
                  store_row16<kF16>(ra, col0, v, no_st);
                } else {
                  copy_row16(ra, pes, jpe, col0);
                }
              };
              const int h0 = kShare ? (own ? 0 : 1) : 0, h1 = kShare ? h0 + 1 : 2;   // 16-point halves this warp drains
              umma::tmem_ld16(taddr + 16 * h0, raw);
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h < h0 || h >= h1) continue;
                umma::tmem_ld_wait();
                umma::tmem_ld16(taddr + 32 + 16 * h, rawB);
                if (!is_pe) {
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    if (P.debug_flags & 16) {   // profiling: no activation math, loads + packs + stores only
                      v[2 * j] = __uint_as_float(raw[2 * j]); v[2 * j + 1] = __uint_as_float(raw[2 * j + 1]); sg[j] = b144;
                    } else {
                      softplus_sig2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]), b144, v[2 * j], v[2 * j + 1], sg[j]);
                    }
                  }
                  store_row16<kF16>(ra, 16 * h, v, no_st);
                  // last hidden layer of 'nablas_img': its value activations are what the radiance pass starts from
                  if (S.to_rad && a.feat_img) img_store16<kF16>(a.feat_img, tile >> 2, F, 32 * (int)(tile & 3) + 16 * h, v);
                } else {
                  copy_row16(ra, pes, jpe, 16 * h);
                }
                umma::tmem_ld_wait();
                umma::tmem_ld16(taddr + 64 + 16 * h, raw);
                tangent(rawB, 32 + 16 * h);
                umma::tmem_ld_wait();
                umma::tmem_ld16(taddr + 96 + 16 * h, rawB);
                tangent(raw, 64 + 16 * h);
                umma::tmem_ld_wait();
                if (h + 1 < h1) umma::tmem_ld16(taddr + 16 * (h + 1), raw);
                tangent(rawB, 96 + 16 * h);
              }
            } else {
              uint32_t rawB[16];
              const int jpe = F - S.out_rows;
              const f32x2 b144 = splat2(b * 144.26950408889634f);
              auto values = [&](const uint32_t (&r)[16], int col0) {
                if (!is_pe) {
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    softplus2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), b144, v[2 * j], v[2 * j + 1]);
                  store_row16<kF16>(ra, col0, v, no_st);
                } else {
                  copy_row16(ra, pes, jpe, col0);
                }
              };
              const int c0 = kShare ? (own ? 0 : 4) : 0, c1 = kShare ? c0 + 4 : 8;   // 16-point chunks this warp drains
              umma::tmem_ld16(taddr + 16 * c0, raw);
#pragma unroll
              for (int c = 0; c < 8; c += 2) {
                if (c < c0 || c >= c1) continue;
                umma::tmem_ld_wait();
                umma::tmem_ld16(taddr + 16 * (c + 1), rawB);
                values(raw, 16 * c);
                umma::tmem_ld_wait();
                if (c + 2 < c1) umma::tmem_ld16(taddr + 16 * (c + 2), raw);
                values(rawB, 16 * (c + 1));
              }
            }
          }
        } else if (!own && S.epi != EPI_RELU && S.epi != EPI_LINEAR) {
          // narrow output steps: the owner group alone
        } else if (S.epi == EPI_SDF_OUT) {
          // rows 0..31 of this M-tile all hold the sdf row: lane l keeps column l of each 32-column chunk
          if (mo == 0 && q == 0) {
            const float b = a.bias[S.bias_off];
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t raw[32];
              umma::tmem_ld32(tmem_tile + 32 * c, raw);
              umma::tmem_ld_wait();
              float mine = 0.0f;
#pragma unroll
              for (int j = 0; j < 32; ++j) mine = (lane == j) ? __uint_as_float(raw[j]) : mine;
              if (tang) {
                const int64_t gp = p0 + lane;
                if (c == 0) {
                  if (a.sdf && gp < a.n) a.sdf[gp] = mine + b;
                } else {
                  nabs[3 * lane + (c - 1)] = a.normal_scale ? mine * a.normal_scale[c - 1] : mine;
                  if (a.nabla && gp < a.n) a.nabla[gp * 3 + (c - 1)] = mine;
                }
              } else {
                const int64_t gp = p0 + 32 * c + lane;
                if (a.sdf && gp < a.n) a.sdf[gp] = mine + b;
              }
            }
          }
        } else if (S.epi == EPI_FEAT) {
          if (mo < S.n_mt) {
            const float b = a.bias[S.bias_off + F];
            const int nchunk = S.n_cols >> 4;
#pragma unroll 1
            for (int c = 0; c < nchunk; ++c) {
              uint32_t raw[16];
              float v[16];
              umma::tmem_ld16(taddr + 16 * c, raw);
              umma::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(raw[j]) + b;
              if (a.feat && F < S.out_rows) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int64_t gp = p0 + 16 * c + j;
                  if (gp < a.n) a.feat[gp * a.feat_ld + F] = v[j];
                }
              }
              if (S.to_rad) store_row16<kF16>(ra, 16 * c, v);
              if (a.feat_img && F < S.out_rows)   // tangent tiles hold 32 points (block column 32 * (tile & 3)), value tiles 128
                img_store16<kF16>(a.feat_img, tang ? (tile >> 2) : tile, F, (tang ? 32 * (int)(tile & 3) : 0) + 16 * c, v);
            }
          }
          if (S.to_rad) {
            // operand rows [256, 256 + extras): [PE(x) | PE(view) | normals | 0-pad]  (tangent tiles)
            named_bar_sync(1 + t, kEpiPerTile);  // normal stash of EPI_SDF_OUT visible
            const int p = etid & 31, g = etid >> 5;
            const int px = P.rad_multires < 0 ? 3 : 3 + 6 * P.rad_multires;
            const int pv = P.rad_multires_view < 0 ? 3 : 3 + 6 * P.rad_multires_view;
            const int extra = P.rad_extra_rows;
            for (int r = g; r < extra; r += 8) {
              float val = 0.0f;
              if (r < px) val = pe_row(r, P.rad_multires, xs + 3 * p, -1);
              else if (r < px + pv) val = pe_row(r - px, P.rad_multires_view, vs + 3 * p, -1);
              else if (r < px + pv + 3) val = nabs[3 * p + (r - px - pv)];
              store_elem<kF16>(act, 256 + r, p, val);
            }
          }
        } else if (S.epi == EPI_EXTRAS) {
          // split-K layer, after its first part: the rows that part read are dead; rows [0, K of the next step) become the
          // second operand, whose product the next step accumulates onto the same TMEM columns.
          const int p = etid & 127, pad = P.steps[s + 1].k_steps * 16;
          if (S.to_rad == 0) {       // radiance net: [PE(x) | PE(view) | normals | 0-pad]
            // two threads per point (rows split by frequency): ONE sincos per (frequency, component) writes its sin and its cos
            // row (until late in round 2: pe_row() per row -- a sincos, a division and a modulo for every single row; the trace
            // showed this epilogue at 8.3 k cycles per tile, a sixth of the radiance pass)
            const int px = P.rad_multires < 0 ? 3 : 3 + 6 * P.rad_multires;
            const int pv = P.rad_multires_view < 0 ? 3 : 3 + 6 * P.rad_multires_view;
            const int part = etid >> 7;
            const float x3[3] = {xs[3 * p], xs[3 * p + 1], xs[3 * p + 2]};
            const float v3[3] = {vs[3 * p], vs[3 * p + 1], vs[3 * p + 2]};
            if (part == 0) {
#pragma unroll
              for (int c = 0; c < 3; ++c) store_elem<kF16>(act, c, p, x3[c]);
#pragma unroll
              for (int c = 0; c < 3; ++c) store_elem<kF16>(act, px + pv + c, p, nabs[3 * p + c]);
            } else {
#pragma unroll
              for (int c = 0; c < 3; ++c) store_elem<kF16>(act, px + c, p, v3[c]);
            }
            for (int qf = part; qf < P.rad_multires; qf += 2) {
              const float f = (float)(1 << qf);
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float sn, cs;
                __sincosf(x3[c] * f, &sn, &cs);
                store_elem<kF16>(act, 3 + 6 * qf + c, p, sn);
                store_elem<kF16>(act, 3 + 6 * qf + 3 + c, p, cs);
              }
            }
            for (int qf = part; qf < P.rad_multires_view; qf += 2) {
              const float f = (float)(1 << qf);
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                float sn, cs;
                __sincosf(v3[c] * f, &sn, &cs);
                store_elem<kF16>(act, px + 3 + 6 * qf + c, p, sn);
                store_elem<kF16>(act, px + 3 + 6 * qf + 3 + c, p, cs);
              }
            }
            for (int r = px + pv + 3 + part; r < pad; r += 2) store_elem<kF16>(act, r, p, 0.0f);
          } else if (S.to_rad == 1) {   // PE(x) again (skip connection cat([PE(x), h]))
            const int npe = P.multires < 0 ? xdim : xdim * (1 + 2 * P.multires);
            for (int r = etid >> 7; r < pad; r += 2)
              store_elem<kF16>(act, r, p, r < npe ? pe_row_nd(r, P.multires, xs + xdim * p, xdim) : 0.0f);
          } else {                      // PE(view) (cat([feature, PE(view)]))
            const int npe = P.rad_multires_view < 0 ? 3 : 3 + 6 * P.rad_multires_view;
            for (int r = etid >> 7; r < pad; r += 2)
              store_elem<kF16>(act, r, p, r < npe ? pe_row_nd(r, P.rad_multires_view, vs + 3 * p, 3) : 0.0f);
          }
        } else if (S.epi == EPI_RELU || S.epi == EPI_LINEAR) {
          const float lo = S.epi == EPI_RELU ? 0.0f : -INFINITY;
          if (mo < S.n_mt) {
            const float b = a.bias[S.bias_off + F];
            const int nall = S.n_cols >> 4;              // 16-column chunks; split between the groups when >= 4
            const bool split = kShare && nall >= 4;
            const int cb = split ? (own ? 0 : nall >> 1) : 0, nchunk = split ? cb + (nall >> 1) : (own ? nall : 0);
            uint32_t raw[16], rawB[16];
            float v[16];
            if (cb < nchunk) umma::tmem_ld16(taddr + 16 * cb, raw);
            for (int c = cb; c < nchunk; c += 2) {
              umma::tmem_ld_wait();
              umma::tmem_ld16(taddr + 16 * (c + 1), rawB);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(__uint_as_float(raw[j]) + b, lo);
              store_row16<kF16>(ra, 16 * c, v);
              umma::tmem_ld_wait();
              if (c + 2 < nchunk) umma::tmem_ld16(taddr + 16 * (c + 2), raw);
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(__uint_as_float(rawB[j]) + b, lo);
              store_row16<kF16>(ra, 16 * (c + 1), v);
            }
          }
        } else if (S.epi == EPI_RGB) {
          // rows 0..2 of the accumulator = the three colour logits of the tile's points.  Lanes 0..2 of one warp park them (+ bias)
          // in the (by now dead) view-direction staging buffer, then all 256 threads of the group apply the sigmoid and write the interleaved [point][3] output
          // coalesced (until late in round 2 three lanes did all of it: 8.6 k cycles per tile, a sixth of the radiance pass)
          if (rad_only && own && etid == 0) {
            // the operand buffer is dead once this step's MMAs have read it: request the next pair's feature image now
            const int64_t next_pair = pair + gridDim.x, next_tile = 2 * next_pair + t;
            img_ahead = next_pair < n_pairs && next_tile < n_tiles;
            if (img_ahead) {
              umma::mbar_arrive_expect_tx(&feat_full[t], kActBytes);
              const uint8_t* src = a.feat_img + (size_t)next_tile * kActBytes;
#pragma unroll
              for (int c = 0; c < 4; ++c) umma::bulk_g2s(act + c * kChunkBytes, src + c * kChunkBytes, kChunkBytes, &feat_full[t]);
            }
          }
          if (own) {
            float* stage = vs;                                 // [3][n_cols] floats (n_cols <= 128)
            const int ncols = S.n_cols, nch = ncols >> 5;
            if (q == 0) {                                      // the two warps of lane quarter 0 share the 32-column chunks
              const float b = lane < 3 ? a.bias[S.bias_off + lane] : 0.0f;
#pragma unroll 1
              for (int c = mo; c < nch; c += 2) {
                uint32_t raw[32];
                umma::tmem_ld32(tmem_tile + 32 * c, raw);
                umma::tmem_ld_wait();
                if (lane < 3) {
                  float4* dst = reinterpret_cast<float4*>(stage + lane * ncols + 32 * c);
#pragma unroll
                  for (int j = 0; j < 8; ++j)
                    dst[j] = make_float4(__uint_as_float(raw[4 * j]) + b, __uint_as_float(raw[4 * j + 1]) + b,
                                         __uint_as_float(raw[4 * j + 2]) + b, __uint_as_float(raw[4 * j + 3]) + b);
                }
              }
            }
            named_bar_sync(1 + t, kEpiPerTile);
            if (a.rgb) {
              const int64_t n_left = a.n - p0;                  // points of this tile that exist
              const int n_out = 3 * (int)(n_left < ncols ? (n_left > 0 ? n_left : 0) : ncols);
              for (int i = etid; i < n_out; i += kEpiPerTile) {
                const int pt = i / 3, ch = i - 3 * pt;
                a.rgb[p0 * 3 + i] = sigmoid_fast(stage[ch * ncols + pt]);
              }
            }
          }
        }
        if (tracer) trace_ev(a.trace, 2, tcnt, 22, s * 2 + t, pair);
        if (s + 1 < P.n_steps) publish(&in_ready[t], P.debug_flags);
       }
      }
      umma::tc_fence_before();
      if (kShare) named_bar_sync(3, 2 * kEpiPerTile);   // staging buffers and TMEM slots free before the next pair
      else named_bar_sync(1 + g, kEpiPerTile);
    }
  }

  umma::tc_fence_before();
  __syncthreads();
  if (warp == 2) umma::tmem_dealloc(tmem_base, 512);
}

}  // namespace

static long long* g_trace = nullptr;
// profiling hook: timestamps of the MMA <-> epilogue hand-offs of CTA 0 ([3][2048][4] int64, device memory)
extern "C" int nr_mlp_umma_set_trace(void* buf) { g_trace = (long long*)buf; return NR_OK; }

extern "C" int nr_mlp_umma_forward(const nr_umma_program_t* prog, const void* image, size_t image_bytes,
                                   const float* bias, size_t bias_floats, const float* x, const float* view,
                                   int64_t n, float* sdf, float* nabla, float* feat, int64_t feat_ld, float* rgb,
                                   const float* normal_scale, void* feat_img, void* stream) {
  NR_CHECK_ARG(prog && image && bias && x, "nr_mlp_umma_forward: null pointer");
  NR_CHECK_ARG(n >= 0, "nr_mlp_umma_forward: n < 0");
  NR_CHECK_ARG(!prog->reverse, "nr_mlp_umma_forward: reverse-mode programs run through nr_mlp_umma_reverse");
  NR_CHECK_ARG(prog->n_steps >= 1 && prog->n_steps <= NR_UMMA_MAX_STEPS, "nr_mlp_umma_forward: n_steps=%d", prog->n_steps);
  NR_CHECK_ARG(((uintptr_t)image & 15) == 0, "nr_mlp_umma_forward: image must be 16-byte aligned");
  bool has_rad = false;
  for (int s = 0; s < prog->n_steps; ++s) {
    const nr_umma_step_t& S = prog->steps[s];
    const int nch = S.n_mt * (S.k_steps / 4);
    NR_CHECK_ARG(S.k_steps % 4 == 0, "step %d: k_steps=%d must be a multiple of 4 (K padded to 64)", s, S.k_steps);
    NR_CHECK_ARG(S.n_mt >= 1 && S.n_mt <= 2, "step %d: n_mt=%d", s, S.n_mt);
    NR_CHECK_ARG(S.k_steps >= 1 && S.k_steps <= 32, "step %d: k_steps=%d", s, S.k_steps);
    NR_CHECK_ARG(S.n_cols == 32 || S.n_cols == 64 || S.n_cols == 128, "step %d: n_cols=%d", s, S.n_cols);
    NR_CHECK_ARG(S.k_steps <= 16 || S.n_cols <= 64, "step %d: K > 256 needs n_cols <= 64", s);
    NR_CHECK_ARG(S.chunk_begin >= 0 && (size_t)(S.chunk_begin + nch) * kChunkBytes <= image_bytes,
                 "step %d: weight chunks [%d,%d) exceed the image", s, S.chunk_begin, S.chunk_begin + nch);
    NR_CHECK_ARG(S.bias_off >= 0 && (size_t)S.bias_off + S.n_mt * 128 <= bias_floats, "step %d: bias range", s);
    NR_CHECK_ARG(S.epi >= EPI_HIDDEN && S.epi <= EPI_LINEAR, "step %d: epi=%d", s, S.epi);
    NR_CHECK_ARG(!S.accumulate || (s > 0 && prog->steps[s - 1].n_mt == S.n_mt && prog->steps[s - 1].n_cols == S.n_cols &&
                                   prog->steps[s - 1].epi == EPI_EXTRAS),
                 "step %d: accumulate needs a preceding EPI_EXTRAS step of the same shape", s);
    NR_CHECK_ARG(S.epi != EPI_EXTRAS || (prog->input_mode >= 1 && s + 1 < prog->n_steps && prog->steps[s + 1].accumulate &&
                                         S.to_rad >= 0 && S.to_rad <= 2 && (S.to_rad == 0) == (prog->input_mode == 1)),
                 "step %d: EPI_EXTRAS needs a value-tile program and a following accumulate step", s);
    if (prog->input_mode != 2 && (S.epi == EPI_RELU || S.epi == EPI_RGB || (S.epi == EPI_FEAT && S.to_rad))) has_rad = true;
  }
  if (prog->input_mode == 1) {
    NR_CHECK_ARG(!prog->tangents && view && nabla && feat_img && prog->steps[0].k_steps == 16 &&
                     (prog->steps[0].epi == EPI_EXTRAS || prog->steps[0].epi == EPI_LINEAR),
                 "nr_mlp_umma_forward: radiance-only programs need value tiles, view dirs, normals, the feature image and a "
                 "K = 256 first step");
    NR_CHECK_ARG(((uintptr_t)feat_img & 15) == 0, "nr_mlp_umma_forward: feat_img must be 16-byte aligned");
  } else if (prog->input_mode == 2) {
    NR_CHECK_ARG(!prog->tangents && view && prog->input_dim >= 1 && prog->input_dim <= 4,
                 "nr_mlp_umma_forward: NeRF++ programs need value tiles, view dirs and 1..4 input components");
    NR_CHECK_ARG(prog->steps[0].k_steps * 16 >= prog->input_dim * (prog->multires < 0 ? 1 : 1 + 2 * prog->multires),
                 "nr_mlp_umma_forward: step 0 K does not cover the embedding");
  } else {
    NR_CHECK_ARG(prog->input_mode == 0, "nr_mlp_umma_forward: input_mode=%d", prog->input_mode);
    NR_CHECK_ARG(prog->steps[0].k_steps * 16 >= (prog->multires < 0 ? 3 : 3 + 6 * prog->multires),
                 "nr_mlp_umma_forward: step 0 K does not cover the embedding");
  }
  for (int s = 0; s < prog->n_steps; ++s)
    if (prog->steps[s].pe_fill)
      NR_CHECK_ARG((prog->multires < 0 ? 3 : 3 + 6 * prog->multires) <= kPeStashRows,
                   "nr_mlp_umma_forward: skip connections need an embedding of at most %d rows", kPeStashRows);
  if (has_rad && prog->input_mode == 0)
    NR_CHECK_ARG(prog->tangents && view, "nr_mlp_umma_forward: the radiance steps need tangent tiles and view dirs");
  if (n == 0) return NR_OK;
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ppt = prog->tangents ? 32 : 128;
  const int64_t n_pairs = (nr_cdiv(n, ppt) + 1) / 2;
  const int grid = (int)(n_pairs < sms ? n_pairs : sms);
  const size_t smem = SmemLayout::total + 1024;
  static unsigned long long attr_set = 0;  // per-device bit: the attribute is per (function, device)
  if (!(attr_set >> (dev & 63) & 1ull)) {
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_umma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_umma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_umma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set |= 1ull << (dev & 63);
  }
  DevProgram dp;
  dp.p = *prog;
  KArgs ka{(const uint8_t*)image, bias, x, view, n, sdf, nabla, feat, feat_ld, rgb, normal_scale, (uint8_t*)feat_img, g_trace};
  const bool share = NR_SHARE_EPILOGUE >= 0 ? NR_SHARE_EPILOGUE != 0 : (!has_rad && prog->tangents == 0 && prog->input_mode == 0);
  if (prog->operand_f16) {
    if (share) mlp_umma_kernel<true, true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(dp, ka);
    else mlp_umma_kernel<true, false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(dp, ka);
  } else {
    if (share) mlp_umma_kernel<false, true><<<grid, kThreads, smem, (cudaStream_t)stream>>>(dp, ka);
    else mlp_umma_kernel<false, false><<<grid, kThreads, smem, (cudaStream_t)stream>>>(dp, ka);
  }
  NR_CHECK_LAUNCH("mlp_umma_kernel");
  return NR_OK;
}
