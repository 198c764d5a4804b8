// Fused positional-encoding + SDF MLP (+ analytic normals, + feature image) on CTA PAIRS:
// tcgen05.mma.cta_group::2, M = 256 = all output features of a layer in one instruction.
//
// Why pairs: the one-CTA kernel (mlp_umma.cu) is bound by shared-memory bandwidth, not by the tensor pipe -- per
// 128-column tile and layer it moves 256 KB of MMA operand reads (every M-tile re-reads the whole activation tile),
// 128 KB of weight-ring writes and 64 KB of activation stores through one SM's shared memory in the 2048 cycles the
// MMAs need (tools/smem_contention.py: the stores get 47 B/clk next to a 128-column MMA stream).  In a pair each CTA
// holds HALF of the layer's weights (its 128 output features) and HALF of the columns (its own points):
//   per CTA and layer: 128 KB operand reads + 64 KB weight writes + 64 KB activation stores = 57 % of the above,
//   half the MMA instructions, each 128 cycles long, issued by one thread of the leader CTA.
// The price: a CTA's epilogue owns 128 features x 256 columns, so half of the activations it produces belong to the
// peer's operand and are written through distributed shared memory (st.shared::cluster).
//
// Layout per CTA (rank r of the pair): operand act[t] = 256 k-rows x its own 128 columns (MN-major, 128-byte
// swizzle, as in mlp_umma.cu); weight ring = 16 KB chunks of ITS M-tile (chunk index kc * 2 + r of a step);
// TMEM slot t = columns [256 t, 256 t + 256): lanes = its 128 features, columns = [CTA 0's 128 | CTA 1's 128].
// Barriers: w_full[stage] (leader: local bulk copy + the peer's relay arrival; peer: local copy only),
// w_empty[stage] / acc_ready[t] (tcgen05.commit multicast to both CTAs), in_ready[t] in the leader (16 warps of
// both CTAs arrive with release.cluster after fence.proxy.async; the MMA thread acquires at cluster scope).
//
// Reference semantics: models/base.py:46-64 (Embedder), :243-282 (ImplicitSurface.forward / forward_with_nablas).
#include "mlp_epilogue.cuh"

namespace {

constexpr int kStages2 = 4;
constexpr int kStages2Log2 = 2;
constexpr int kThreads2 = 640;        // producer, MMA issuer / relay, TMEM allocator, spare + 2 slots x 8 epilogue warps

struct Smem2 {
  static constexpr uint32_t act = 0;                                  // 2 x 64 KB
  static constexpr uint32_t ring = 2 * kActBytes;                     // kStages2 x 16 KB
  static constexpr uint32_t xs = ring + kStages2 * kChunkBytes;       // 2 x 128 x 3 floats
  static constexpr uint32_t nabs = xs + 2 * 384 * 4;                  // 2 x 32 x 3 floats (normal stash, unused so far)
  static constexpr uint32_t pes = nabs + 2 * 96 * 4;                  // 2 x 40 rows x 256 B: embedding stash (skip)
  static constexpr uint32_t bars = pes + 2 * kPeStashRows * 256;
  static constexpr uint32_t total = bars + 256;
};

__device__ __forceinline__ void st_row16_cluster(uint32_t xbase, int col0, const float (&v)[16], bool f16) {
#pragma unroll
  for (int j4 = 0; j4 < 2; ++j4) {
    const int n8 = (col0 >> 3) + j4;
    const uint32_t addr = (xbase ^ (((uint32_t)n8 & 7u) << 4)) + (uint32_t)(n8 >> 3) * kLbo;
    if (f16)
      umma::st_cluster_v4(addr, umma::pack_f16(v[8 * j4 + 0], v[8 * j4 + 1]), umma::pack_f16(v[8 * j4 + 2], v[8 * j4 + 3]),
                          umma::pack_f16(v[8 * j4 + 4], v[8 * j4 + 5]), umma::pack_f16(v[8 * j4 + 6], v[8 * j4 + 7]));
    else
      umma::st_cluster_v4(addr, umma::pack_bf16(v[8 * j4 + 0], v[8 * j4 + 1]), umma::pack_bf16(v[8 * j4 + 2], v[8 * j4 + 3]),
                          umma::pack_bf16(v[8 * j4 + 4], v[8 * j4 + 5]), umma::pack_bf16(v[8 * j4 + 6], v[8 * j4 + 7]));
  }
}

// Epilogue -> MMA hand-off across the pair: writes (local and DSMEM) made visible to the async proxy, TMEM reads
// ordered, then one arrival per warp on the LEADER's barrier with release at cluster scope.
__device__ __forceinline__ void publish2(uint32_t leader_bar_caddr) {
  asm volatile("fence.proxy.async;" ::: "memory");
  umma::tc_fence_before();
  __syncwarp();
  if ((threadIdx.x & 31) == 0) umma::mbar_arrive_cluster(leader_bar_caddr);
}

template <bool kF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2, 1)
mlp_umma2_kernel(const __grid_constant__ DevProgram prog, const KArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + Smem2::bars);
  uint64_t* w_full = bars;                    // [kStages2]
  uint64_t* w_empty = bars + kStages2;        // [kStages2]
  uint64_t* in_ready = bars + 2 * kStages2;   // [2]  (the leader's copy is the one in use)
  uint64_t* acc_ready = in_ready + 2;         // [2]
  __shared__ uint32_t tmem_base_s;

  const nr_umma_program_t& P = prog.p;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t rank = umma::cluster_ctarank();
  const int tang = P.tangents;
  const int ppt = tang ? 32 : 128;                         // points per CTA and slot
  const int64_t n_sub = (a.n + ppt - 1) / ppt;             // one-CTA tiles
  const int64_t n_pt = (n_sub + 1) / 2;                    // pair tiles (one slot of the pair)
  const int64_t n_pp = (n_pt + 1) / 2;                     // pair iterations (both slots)
  const int64_t cluster = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages2; ++s) { umma::mbar_init(&w_full[s], rank == 0 ? 2 : 1); umma::mbar_init(&w_empty[s], 1); }
    for (int t = 0; t < 2; ++t) { umma::mbar_init(&in_ready[t], 16); umma::mbar_init(&acc_ready[t], 1); }
    umma::fence_barrier_init();
  }
  __syncthreads();
  umma::cluster_sync_all();          // nobody arrives on a peer barrier before it exists
  if (warp == 2) {
    umma::tmem_alloc2(&tmem_base_s, 512);
    umma::tmem_relinquish2();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== weight producer: this CTA's M-tile of every chunk pair =====================
    uint32_t cnt = 0;
    for (int64_t pp = cluster; pp < n_pp; pp += n_clusters) {
      const int ntl = (2 * pp + 1 < n_pt) ? 2 : 1;
      for (int s = 0; s < P.n_steps; ++s) {
        const int nkc = P.steps[s].k_steps >> 2;
        const uint8_t* src = a.image + ((size_t)P.steps[s].chunk_begin + rank) * kChunkBytes;
        for (int t = 0; t < ntl; ++t) {
          for (int kc = 0; kc < nkc; ++kc, ++cnt) {
            const uint32_t stage = cnt & (kStages2 - 1);
            umma::mbar_wait_tag(&w_empty[stage], ((cnt >> kStages2Log2) & 1u) ^ 1u, 100 + (int)cnt);
            if (umma::elect_one()) {
              if (P.debug_flags & 1) {   // profiling: no weight traffic
                umma::mbar_arrive(&w_full[stage]);
              } else {
                umma::mbar_arrive_expect_tx(&w_full[stage], kChunkBytes);
                umma::bulk_g2s(smem + Smem2::ring + stage * kChunkBytes, src + (size_t)kc * 2 * kChunkBytes, kChunkBytes,
                               &w_full[stage]);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA): 256 x N x 16 per instruction, 128 cycles at N = 256 ============
    uint32_t cnt = 0, in_par = 0;
    const uint32_t a_hi = umma::smem_desc_hi(1024), b_hi = umma::smem_desc_hi(1024);
    const uint32_t ring_lo = umma::smem_desc_lo(umma::smem_u32(smem + Smem2::ring), 16);
    const uint32_t act_lo0 = umma::smem_desc_lo(umma::smem_u32(smem + Smem2::act), kLbo);
    for (int64_t pp = cluster; pp < n_pp; pp += n_clusters) {
      const int ntl = (2 * pp + 1 < n_pt) ? 2 : 1;
      for (int s = 0; s < P.n_steps; ++s) {
        const uint32_t nkc = P.steps[s].k_steps >> 2;
        const uint32_t idesc = kF16 ? umma::make_idesc_f16(256, 2 * P.steps[s].n_cols, 0, 1)
                                    : umma::make_idesc_bf16(256, 2 * P.steps[s].n_cols, 0, 1);
        for (int t = 0; t < ntl; ++t) {
          umma::mbar_wait_cluster(&in_ready[t], (in_par >> t) & 1u, 200 + s * 2 + t);
          in_par ^= 1u << t;
          umma::tc_fence_after();
          const uint32_t d_addr = tmem_base + (uint32_t)t * 256u;
          uint32_t b_lo = act_lo0 + (uint32_t)t * (kActBytes >> 4);
#pragma unroll 1
          for (uint32_t kc = 0; kc < nkc; ++kc, ++cnt, b_lo += 512) {
            const uint32_t st = cnt & (kStages2 - 1);
            umma::mbar_wait_cluster(&w_full[st], (cnt >> kStages2Log2) & 1u, 300 + (int)cnt);
            umma::tc_fence_after();
            const uint32_t a_lo = ring_lo + st * (kChunkBytes >> 4);
            if (umma::elect_one()) {
              umma::mma2_f16_ss(d_addr, umma::desc64(a_lo, a_hi), umma::desc64(b_lo, b_hi), idesc, kc ? 1u : 0u);
              umma::mma2_f16_ss(d_addr, umma::desc64(a_lo + 2, a_hi), umma::desc64(b_lo + 128, b_hi), idesc, 1u);
              umma::mma2_f16_ss(d_addr, umma::desc64(a_lo + 4, a_hi), umma::desc64(b_lo + 256, b_hi), idesc, 1u);
              umma::mma2_f16_ss(d_addr, umma::desc64(a_lo + 6, a_hi), umma::desc64(b_lo + 384, b_hi), idesc, 1u);
              umma::mma2_commit_mc(&w_empty[st], 3);
            }
            __syncwarp();
          }
          if (umma::elect_one()) umma::mma2_commit_mc(&acc_ready[t], 3);
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ===================== relay (peer CTA): "my half of chunk k has landed" -> the leader's w_full ==============
    uint32_t cnt = 0;
    const uint32_t leader_full0 = umma::map_to_cta(umma::smem_u32(&w_full[0]), 0);
    for (int64_t pp = cluster; pp < n_pp; pp += n_clusters) {
      const int ntl = (2 * pp + 1 < n_pt) ? 2 : 1;
      for (int s = 0; s < P.n_steps; ++s) {
        const int nkc = P.steps[s].k_steps >> 2;
        for (int i = 0; i < ntl * nkc; ++i, ++cnt) {
          const uint32_t st = cnt & (kStages2 - 1);
          umma::mbar_wait_tag(&w_full[st], (cnt >> kStages2Log2) & 1u, 400 + (int)cnt);
          if (lane == 0) umma::mbar_arrive_cluster(leader_full0 + st * 8);
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: slot t, destination half ch, lane quarter q =====================
    const int e = warp - 4;
    const int t = e >> 3;                            // slot
    const uint32_t ch = (uint32_t)(e >> 2) & 1u;     // columns [128 ch, 128 ch + 128) of the pair tile = CTA ch's points
    const int q = warp & 3;
    const int etid = (e & 7) * 32 + lane;            // 0..255 inside the slot group
    const int Fl = 32 * q + lane;                    // TMEM lane = local feature
    const int F = 128 * (int)rank + Fl;              // feature of the layer
    uint8_t* act = smem + Smem2::act + t * kActBytes;                    // this CTA's operand (prologue, PE rows)
    float* xs = (float*)(smem + Smem2::xs) + t * 384;
    uint8_t* pes = smem + Smem2::pes + t * (kPeStashRows * 256);
    const uint32_t act_dst = umma::map_to_cta(umma::smem_u32(act), ch);   // operand of the CTA that owns these columns
    const uint32_t row_dst = act_dst + (uint32_t)((F >> 3) * 1024 + (F & 7) * 128);
    const uint32_t xrow_dst = row_dst ^ ((uint32_t)(F & 7) << 4);
    const uint32_t leader_in = umma::map_to_cta(umma::smem_u32(&in_ready[t]), 0);
    const uint32_t tmem_q = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(t * 256);
    const int pe_dim = P.multires < 0 ? 3 : 3 + 6 * P.multires;
    uint32_t acc_par = 0;

    for (int64_t pp = cluster; pp < n_pp; pp += n_clusters) {
      if (2 * pp + t >= n_pt) continue;
      const int64_t tile = (2 * pp + t) * 2 + rank;          // this CTA's own 32 / 128 points
      const int64_t p0 = tile * ppt;
      const int64_t tile_ch = (2 * pp + t) * 2 + ch;         // the points whose columns this warp drains
      const int64_t p0_ch = tile_ch * ppt;

      // ---- prologue: own points -> embedding rows [0, k0) of the own operand (+ stash for the skip layer) ----
      for (int i = etid; i < ppt * 3; i += 256) {
        const int64_t gi = p0 * 3 + i;
        xs[i] = gi < a.n * 3 ? a.x[gi] : 0.0f;
      }
      named_bar_sync(1 + t, 256);
      {
        const int n = etid & 127, part = etid >> 7;
        const int p = tang ? (n & 31) : n;
        const int ct = tang ? (n >> 5) - 1 : -1;
        const float x3[3] = {xs[3 * p], xs[3 * p + 1], xs[3 * p + 2]};
        const int k0 = P.steps[0].k_steps * 16;
        uint16_t* stash = reinterpret_cast<uint16_t*>(pes) + n;
        auto put = [&](int j, float val) {
          store_elem<kF16>(act, j, n, val);
          if (j < kPeStashRows) stash[j * 128] = umma::pack1<kF16>(val);
        };
        if (part == 0) {
#pragma unroll
          for (int j = 0; j < 3; ++j) put(j, ct < 0 ? x3[j] : (ct == j ? 1.f : 0.f));
        }
        for (int qf = part; qf < P.multires; qf += 2) {
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
      publish2(leader_in);

      for (int s = 0; s < P.n_steps; ++s) {
        const nr_umma_step_t& S = P.steps[s];
        umma::mbar_wait_tag(&acc_ready[t], acc_par, 500 + s * 2 + t);
        acc_par ^= 1;
        umma::tc_fence_after();
        const uint32_t taddr = tmem_q + ch * (uint32_t)S.n_cols;   // this warp's columns: CTA ch's n_cols

        if (P.debug_flags & 2) {
          // profiling: MMA + weight pipeline only
        } else if (S.epi == EPI_HIDDEN) {
          const float b = a.bias[S.bias_off + F];
          const bool is_pe = S.pe_fill && F >= S.out_rows;          // embedding rows of the skip operand: no math
          uint32_t raw[16], rawB[16];
          float v[16];
          const f32x2 b144 = splat2(b * 144.26950408889634f);
          // (tcgen05.ld is warp-collective: the embedding-row lanes of a mixed warp load too and skip only math + stores)
          if (tang) {
            f32x2 sg[8];
            auto tangent = [&](const uint32_t (&r)[16], int col0) {
              if (is_pe) return;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                upk2(mul2(pk2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), sg[j]), v[2 * j], v[2 * j + 1]);
              st_row16_cluster(xrow_dst, col0, v, kF16);
            };
            umma::tmem_ld16(taddr, raw);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              umma::tmem_ld_wait();
              umma::tmem_ld16(taddr + 32 + 16 * h, rawB);
              if (!is_pe) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  softplus_sig2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]), b144, v[2 * j], v[2 * j + 1], sg[j]);
                st_row16_cluster(xrow_dst, 16 * h, v, kF16);
              }
              umma::tmem_ld_wait();
              umma::tmem_ld16(taddr + 64 + 16 * h, raw);
              tangent(rawB, 32 + 16 * h);
              umma::tmem_ld_wait();
              umma::tmem_ld16(taddr + 96 + 16 * h, rawB);
              tangent(raw, 64 + 16 * h);
              umma::tmem_ld_wait();
              if (h == 0) umma::tmem_ld16(taddr + 16, raw);
              tangent(rawB, 96 + 16 * h);
            }
          } else {
            auto values = [&](const uint32_t (&r)[16], int col0) {
              if (is_pe) return;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                softplus2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]), b144, v[2 * j], v[2 * j + 1]);
              st_row16_cluster(xrow_dst, col0, v, kF16);
            };
            umma::tmem_ld16(taddr, raw);
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
              umma::tmem_ld_wait();
              umma::tmem_ld16(taddr + 16 * (c + 1), rawB);
              values(raw, 16 * c);
              umma::tmem_ld_wait();
              if (c + 2 < 8) umma::tmem_ld16(taddr + 16 * (c + 2), raw);
              values(rawB, 16 * (c + 1));
            }
          }
          // skip layer: embedding rows [out_rows, out_rows + pe_dim) of the OWN operand from the own stash; every CTA
          // does it for its own columns with the threads whose lane maps to row 128 + Fl
          if (S.pe_fill && ch == rank) {
            const int prow = 128 + Fl, jpe = prow - S.out_rows;
            if (jpe >= 0 && jpe < pe_dim) {
              const RowAddr rl(umma::smem_u32(act), prow);
#pragma unroll
              for (int c = 0; c < 8; ++c) copy_row16(rl, pes, jpe, 16 * c);
            }
          }
        } else if (S.epi == EPI_SDF_OUT) {
          // both M-tiles carry the sdf row replicated over lanes 0..31: each CTA writes its own points
          if (ch == rank && q == 0) {
            const float b = a.bias[S.bias_off];
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t raw[32];
              umma::tmem_ld32(taddr + 32 * c, raw);
              umma::tmem_ld_wait();
              float mine = 0.0f;
#pragma unroll
              for (int j = 0; j < 32; ++j) mine = (lane == j) ? __uint_as_float(raw[j]) : mine;
              if (tang) {
                const int64_t gp = p0 + lane;
                if (c == 0) {
                  if (a.sdf && gp < a.n) a.sdf[gp] = mine + b;
                } else if (a.nabla && gp < a.n) {
                  a.nabla[gp * 3 + (c - 1)] = mine;
                }
              } else {
                const int64_t gp = p0 + 32 * c + lane;
                if (a.sdf && gp < a.n) a.sdf[gp] = mine + b;
              }
            }
          }
        } else if (S.epi == EPI_FEAT) {
          // this CTA's 128 features of the points of CTA ch: fp32 rows and / or the radiance operand image
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
                const int64_t gp = p0_ch + 16 * c + j;
                if (gp < a.n) a.feat[gp * a.feat_ld + F] = v[j];
              }
            }
            if (a.feat_img && F < S.out_rows) {
              const int col0 = (tang ? 32 * (int)(tile_ch & 3) : 0) + 16 * c;
              uint8_t* blk = a.feat_img + (size_t)(tang ? (tile_ch >> 2) : tile_ch) * kActBytes + (F >> 3) * 1024 + (F & 7) * 128 +
                             (size_t)(col0 >> 6) * kLbo;
#pragma unroll
              for (int j4 = 0; j4 < 2; ++j4) {
                uint4 w;
                w.x = umma::pack2<kF16>(v[8 * j4 + 0], v[8 * j4 + 1]);
                w.y = umma::pack2<kF16>(v[8 * j4 + 2], v[8 * j4 + 3]);
                w.z = umma::pack2<kF16>(v[8 * j4 + 4], v[8 * j4 + 5]);
                w.w = umma::pack2<kF16>(v[8 * j4 + 6], v[8 * j4 + 7]);
                *reinterpret_cast<uint4*>(blk + ((((((col0 & 63) >> 3) + j4) & 7) ^ (F & 7)) << 4)) = w;
              }
            }
          }
        }
        if (s + 1 < P.n_steps) publish2(leader_in);
      }
      umma::tc_fence_before();
      named_bar_sync(1 + t, 256);   // staging buffers and the TMEM slot are free before the next pair tile
    }
  }

  umma::tc_fence_before();
  __syncthreads();
  umma::cluster_sync_all();          // the peer may still be writing into this CTA's shared memory / barriers
  if (warp == 2) umma::tmem_dealloc2(tmem_base, 512);
}

}  // namespace

// Same contract as nr_mlp_umma_forward for programs made of EPI_HIDDEN / EPI_SDF_OUT / EPI_FEAT steps whose weight
// chunks all come in M-tile pairs (n_mt = 2; the sdf row replicated into both M-tiles).
extern "C" int nr_mlp_umma2_forward(const nr_umma_program_t* prog, const void* image, size_t image_bytes,
                                    const float* bias, size_t bias_floats, const float* x, int64_t n, float* sdf,
                                    float* nabla, float* feat, int64_t feat_ld, void* feat_img, void* stream) {
  NR_CHECK_ARG(prog && image && bias && x, "nr_mlp_umma2_forward: null pointer");
  NR_CHECK_ARG(n >= 0, "nr_mlp_umma2_forward: n < 0");
  NR_CHECK_ARG(prog->n_steps >= 1 && prog->n_steps <= NR_UMMA_MAX_STEPS, "nr_mlp_umma2_forward: n_steps=%d", prog->n_steps);
  NR_CHECK_ARG(((uintptr_t)image & 15) == 0 && ((uintptr_t)feat_img & 15) == 0, "nr_mlp_umma2_forward: 16-byte alignment");
  NR_CHECK_ARG(prog->input_mode == 0, "nr_mlp_umma2_forward: input_mode=%d", prog->input_mode);
  const int pe_dim = prog->multires < 0 ? 3 : 3 + 6 * prog->multires;
  for (int s = 0; s < prog->n_steps; ++s) {
    const nr_umma_step_t& S = prog->steps[s];
    NR_CHECK_ARG(S.n_mt == 2, "step %d: the pair kernel needs both M-tiles (n_mt=%d)", s, S.n_mt);
    NR_CHECK_ARG(S.k_steps % 4 == 0 && S.k_steps >= 4 && S.k_steps <= 16, "step %d: k_steps=%d", s, S.k_steps);
    NR_CHECK_ARG(S.n_cols == 32 || S.n_cols == 64 || S.n_cols == 128, "step %d: n_cols=%d", s, S.n_cols);
    NR_CHECK_ARG(S.chunk_begin >= 0 && (size_t)(S.chunk_begin + 2 * (S.k_steps / 4)) * kChunkBytes <= image_bytes,
                 "step %d: weight chunks exceed the image", s);
    NR_CHECK_ARG(S.bias_off >= 0 && (size_t)S.bias_off + 256 <= bias_floats, "step %d: bias range", s);
    NR_CHECK_ARG(S.epi == EPI_HIDDEN || S.epi == EPI_SDF_OUT || (S.epi == EPI_FEAT && !S.to_rad), "step %d: epi=%d", s, S.epi);
    NR_CHECK_ARG(!S.accumulate, "step %d: accumulate is not supported by the pair kernel", s);
    if (S.pe_fill)
      NR_CHECK_ARG(pe_dim <= kPeStashRows && S.out_rows >= 128 && S.out_rows + pe_dim <= 256, "step %d: skip operand layout", s);
  }
  NR_CHECK_ARG(prog->steps[0].k_steps * 16 >= pe_dim, "nr_mlp_umma2_forward: step 0 K does not cover the embedding");
  if (n == 0) return NR_OK;
  int dev = 0, sms = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  NR_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ppt = prog->tangents ? 32 : 128;
  const int64_t n_pp = ((nr_cdiv(n, ppt) + 1) / 2 + 1) / 2;
  const int clusters = (int)(n_pp < sms / 2 ? n_pp : sms / 2);
  const size_t smem = Smem2::total + 1024;
  static unsigned long long attr_set = 0;
  if (!(attr_set >> (dev & 63) & 1ull)) {
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_umma2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CHECK_CUDA(cudaFuncSetAttribute(mlp_umma2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set |= 1ull << (dev & 63);
  }
  DevProgram dp;
  dp.p = *prog;
  KArgs ka{(const uint8_t*)image, bias, x, nullptr, n, sdf, nabla, feat, feat_ld, nullptr, nullptr, (uint8_t*)feat_img, nullptr};
  if (prog->operand_f16) mlp_umma2_kernel<true><<<2 * clusters, kThreads2, smem, (cudaStream_t)stream>>>(dp, ka);
  else mlp_umma2_kernel<false><<<2 * clusters, kThreads2, smem, (cudaStream_t)stream>>>(dp, ka);
  NR_CHECK_LAUNCH("mlp_umma2_kernel");
  return NR_OK;
}
