// Shared pieces of the fused tcgen05 MLP kernels (mlp_umma.cu: one CTA per SM; mlp_umma2.cu: CTA pairs):
// step vocabulary, operand geometry, the packed-fp32x2 activation math, the operand-row store helpers.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace {

constexpr int kPeStashRows = 40;             // embedding rows kept for the skip layer (multires <= 6)
constexpr uint32_t kChunkBytes = 16384;      // one A tile: 128 features x 64 k, 16-bit
constexpr uint32_t kActBytes = 65536;        // one tile's B operand: 256 k-rows x 128 columns, 16-bit
constexpr uint32_t kLbo = 32768;             // bytes between 64-column blocks of the B operand

enum : int32_t {
  EPI_HIDDEN = 0,   // softplus(beta=100) hidden layer (+ tangents), next operand in smem
  EPI_SDF_OUT = 1,  // row-replicated sdf row: sdf / nabla to global (+ normal stash)
  EPI_FEAT = 2,     // geometry feature: to global and/or radiance operand rows [0,256) + extras
  EPI_RELU = 3,     // radiance hidden layer
  EPI_RGB = 4,      // sigmoid, rows 0..2 to global
  EPI_EXTRAS = 5,   // no accumulator read: operand rows [0, K of the next step) <- the second operand of a split-K layer
                    // (to_rad 0: [PE(x)|PE(view)|normals|0] of the radiance net, 1: PE(x), 2: PE(view)); next step accumulates
  EPI_LINEAR = 6,   // bias only (NeRF++ feature_linear), next operand in smem
  EPI_BWD = 7,      // reverse mode (mlp_rev.cu): next operand = softplus'(z) * accumulator
  EPI_NABLA = 8,    // reverse mode, last step: embedding Jacobian, nabla to global
};

struct DevProgram {
  nr_umma_program_t p;
};

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2): one issue slot for two lanes of math.  The epilogue is
// bound by instruction issue (an epilogue warp issues ~0.3 instr/clk, ncu round 1), not by a math pipe, so the
// activation is written on register pairs wherever the two values take the same path.
#ifndef NR_SIGMOID_NEWTON
#define NR_SIGMOID_NEWTON 0
#endif
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 splat2(float x) { return pk2(x, x); }

// softplus(beta=100)(z) and its derivative sigmoid(100 z), z = acc + bias, for a pair of accumulators.
//   t  = 100 z log2(e) = fma(acc, 144.27, bias * 144.27)          (b144 = splat(bias * 144.27))
//   u  = 2^-|t| in (0, 1]                                          1 MUFU per value
//   softplus = max(t, 0) ln2/100 + log1p(u)/100,  log1p(u)/100 = u (c0 + c1 u + c2 u^2 + c3 u^3): |err| < 7.1e-7
//     absolute, a fifth of a half-ulp of the 16-bit operand it is rounded to; nn.Softplus's threshold-20 linear
//     branch needs no select (beyond it log1p(u) < 2.1e-9)
//   sigmoid = 1/2 + copysign(r - 1/2, t),  r = 1 / (1 + u) in [1/2, 1)   1 MUFU per value
__device__ __forceinline__ void softplus_sig2(float a0, float a1, f32x2 b144, float& sp0, float& sp1, f32x2& sg) {
  const f32x2 t2 = fma2(pk2(a0, a1), splat2(144.26950408889634f), b144);
  float t0, t1;
  upk2(t2, t0, t1);
  const f32x2 u2 = pk2(ex2_approx(-fabsf(t0)), ex2_approx(-fabsf(t1)));
  f32x2 p = fma2(u2, splat2(-0.05875718221068382e-2f), splat2(0.22568579018115997e-2f));
  p = fma2(p, u2, splat2(-0.4713013470172882e-2f));
  p = fma2(p, u2, splat2(0.9974489808082581e-2f));
  const f32x2 q = mul2(p, u2);
  const f32x2 sp2 = fma2(pk2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), splat2(0.006931471805599453f), q);
  upk2(sp2, sp0, sp1);
  float s0, s1;
#if NR_SIGMOID_NEWTON
  // r = 1/(1+u) on the FMA pipe (the XU pipe issues a warp instruction only every 8 cycles per sub-partition and the
  // ex2 above already fills half of it): cubic minimax guess (2.5e-3) + one Newton step -> 6.4e-6 relative
  f32x2 r = fma2(u2, splat2(-0.23549793660640717f), splat2(0.6862913966178894f));
  r = fma2(r, u2, splat2(-0.950793445110321f));
  r = fma2(r, u2, splat2(0.9987373352050781f));
  const f32x2 d = add2(u2, splat2(1.0f));
  const f32x2 e = fma2(d, r ^ 0x8000000080000000ull, splat2(1.0f));   // 1 - d r
  r = fma2(r, e, r);
  upk2(add2(r, splat2(-0.5f)), s0, s1);
#else
  float d0, d1;
  upk2(add2(u2, splat2(1.0f)), d0, d1);
  upk2(add2(pk2(rcp_approx(d0), rcp_approx(d1)), splat2(-0.5f)), s0, s1);
#endif
  s0 = __uint_as_float(__float_as_uint(s0) | (__float_as_uint(t0) & 0x80000000u));
  s1 = __uint_as_float(__float_as_uint(s1) | (__float_as_uint(t1) & 0x80000000u));
  sg = add2(pk2(s0, s1), splat2(0.5f));
}
// Same softplus, with sg = derivative - 1/2 good to 2.2e-4 absolute from ONE transcendental: r = 1/(1+u) as a minimax
// quartic in u on the FMA pipe (the MUFU unit issues 16 lanes/clk/SM: two per value would cost a 128-point tile 4096
// cycles per layer, twice its MMA time).  For consumers that quantise the derivative to 8 bits (mlp_rev.cu).
__device__ __forceinline__ void softplus_sigq2(float a0, float a1, f32x2 b144, float& sp0, float& sp1, f32x2& sg) {
  const f32x2 t2 = fma2(pk2(a0, a1), splat2(144.26950408889634f), b144);
  float t0, t1;
  upk2(t2, t0, t1);
  const f32x2 u2 = pk2(ex2_approx(-fabsf(t0)), ex2_approx(-fabsf(t1)));
  f32x2 p = fma2(u2, splat2(-0.05875718221068382e-2f), splat2(0.22568579018115997e-2f));
  p = fma2(p, u2, splat2(-0.4713013470172882e-2f));
  p = fma2(p, u2, splat2(0.9974489808082581e-2f));
  const f32x2 q = mul2(p, u2);
  upk2(fma2(pk2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), splat2(0.006931471805599453f), q), sp0, sp1);
  f32x2 r = fma2(u2, splat2(0.16162029f), splat2(-0.55180615f));
  r = fma2(r, u2, splat2(0.87791824f));
  r = fma2(r, u2, splat2(-0.9872991f));
  r = fma2(r, u2, splat2(0.99978334f - 0.5f));          // r - 1/2 in [0, 1/2]
  float s0, s1;
  upk2(r, s0, s1);
  s0 = __uint_as_float(__float_as_uint(s0) | (__float_as_uint(t0) & 0x80000000u));
  s1 = __uint_as_float(__float_as_uint(s1) | (__float_as_uint(t1) & 0x80000000u));
  sg = pk2(s0, s1);
}
// Same softplus; the derivative as th = tanh(50 z) = 2 sigmoid(100 z) - 1 from a SECOND transcendental (tanh.approx, 2^-11
// relative) instead of the quartic: 3 operations fewer per pair on the FMA pipe and 2 fewer on the ALU pipe (no sign
// transfer), 1 more on the XU pipe.  The FMA pipe is what the forward epilogue of mlp_rev.cu saturates (88 FFMA2 of 238
// instructions per 16 values at one per two cycles and scheduler; the XU pipe idles at 22 %).
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void softplus_sigt2(float a0, float a1, f32x2 b144, float& sp0, float& sp1, f32x2& th) {
  const f32x2 t2 = fma2(pk2(a0, a1), splat2(144.26950408889634f), b144);
  float t0, t1;
  upk2(t2, t0, t1);
  const f32x2 u2 = pk2(ex2_approx(-fabsf(t0)), ex2_approx(-fabsf(t1)));
  f32x2 p = fma2(u2, splat2(-0.05875718221068382e-2f), splat2(0.22568579018115997e-2f));
  p = fma2(p, u2, splat2(-0.4713013470172882e-2f));
  p = fma2(p, u2, splat2(0.9974489808082581e-2f));
  const f32x2 q = mul2(p, u2);
  upk2(fma2(pk2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), splat2(0.006931471805599453f), q), sp0, sp1);
  float h0, h1;
  upk2(mul2(t2, splat2(0.34657359027997264f)), h0, h1);          // t ln2 / 2 = 50 z
  th = pk2(tanh_approx(h0), tanh_approx(h1));
}
// softplus AND its derivative from ONE transcendental, th = tanh(50 z) = 2 sigmoid(100 z) - 1 (MUFU.TANH, 2^-11 relative):
//   softplus' - 1/2 = th / 2 comes signed out of the XU pipe (no quartic, no sign transfer), and with 1 + u = 2 / (1 + |th|)
//   softplus(z) = relu(z) + log1p(u) / 100 = relu(z) + (ln 2 - log1p(|th|)) / 100: the same degree-3 log1p, evaluated in
//   |th| (a free operand modifier of FFMA2), relu(t) = (t + |t|) / 2 on the FMA pipe as well.  7 FMA-pipe operations per
//   pair of values instead of 11, no ALU-pipe operation instead of 4 (tools/probe_epi.py: 990 -> 690 cycles per 16-value
//   chunk and warp).  |err| of softplus: 7.1e-7 (polynomial) + 2.4e-6 (th), of the derivative 2.4e-4: for consumers that
//   round the activation to 16 bits and the derivative to 8 (mlp_rev.cu).  b50 = splat(50 bias); th = tanh(50 z) in [-1, 1].
template <bool kReluFma>
__device__ __forceinline__ void softplus_th2(float a0, float a1, f32x2 b50, float& sp0, float& sp1, f32x2& th) {
  const f32x2 t2 = fma2(pk2(a0, a1), splat2(50.0f), b50);
  float t0, t1;
  upk2(t2, t0, t1);
  const float h0 = tanh_approx(t0), h1 = tanh_approx(t1);
  th = pk2(h0, h1);
  const f32x2 x = pk2(fabsf(h0), fabsf(h1));
  f32x2 q = fma2(x, splat2(0.05875718221068382e-2f), splat2(-0.22568579018115997e-2f));
  q = fma2(q, x, splat2(0.4713013470172882e-2f));
  q = fma2(q, x, splat2(-0.9974489808082581e-2f));
  q = fma2(q, x, splat2(0.006931471805599453f));
  if (kReluFma) {
    q = fma2(pk2(fabsf(t0), fabsf(t1)), splat2(0.01f), q);
    upk2(fma2(t2, splat2(0.01f), q), sp0, sp1);
  } else {
    upk2(fma2(pk2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), splat2(0.02f), q), sp0, sp1);
  }
}
__device__ __forceinline__ void softplus2(float a0, float a1, f32x2 b144, float& sp0, float& sp1) {
  const f32x2 t2 = fma2(pk2(a0, a1), splat2(144.26950408889634f), b144);
  float t0, t1;
  upk2(t2, t0, t1);
  const f32x2 u2 = pk2(ex2_approx(-fabsf(t0)), ex2_approx(-fabsf(t1)));
  f32x2 p = fma2(u2, splat2(-0.05875718221068382e-2f), splat2(0.22568579018115997e-2f));
  p = fma2(p, u2, splat2(-0.4713013470172882e-2f));
  p = fma2(p, u2, splat2(0.9974489808082581e-2f));
  const f32x2 q = mul2(p, u2);
  upk2(fma2(pk2(fmaxf(t0, 0.0f), fmaxf(t1, 0.0f)), splat2(0.006931471805599453f), q), sp0, sp1);
}
__device__ __forceinline__ float sigmoid_fast(float x) { return rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * x)); }

// Row j of Embedder.forward(x) (models/base.py:53-61): [x, sin(2^0 x), cos(2^0 x), sin(2^1 x), ...]
// value (comp_t < 0) or derivative w.r.t. x[comp_t].
__device__ __forceinline__ float pe_row(int j, int multires, const float* x3, int comp_t) {
  if (multires < 0 || j < 3) {
    if (j >= 3) return 0.0f;
    return comp_t < 0 ? x3[j] : (comp_t == j ? 1.0f : 0.0f);
  }
  const int q = (j - 3) / 6, r = (j - 3) % 6, comp = r % 3;
  if (q >= multires) return 0.0f;
  const float f = (float)(1 << q);
  float s, c;
  __sincosf(x3[comp] * f, &s, &c);
  if (comp_t < 0) return r < 3 ? s : c;
  if (comp_t != comp) return 0.0f;
  return r < 3 ? f * c : -f * s;
}

// Row j of Embedder.forward(x) for a `dim`-component input (models/base.py:53-61): [x, sin(2^0 x), cos(2^0 x), ...]
__device__ __forceinline__ float pe_row_nd(int j, int multires, const float* x, int dim) {
  if (j < dim) return x[j];
  if (multires < 0) return 0.0f;
  const int q = (j - dim) / (2 * dim), r = (j - dim) % (2 * dim);
  if (q >= multires) return 0.0f;
  float s, c;
  __sincosf(x[r % dim] * (float)(1 << q), &s, &c);
  return r < dim ? s : c;
}

template <bool kF16>
__device__ __forceinline__ void store_row32(uint8_t* act, int k, int col0, const float (&v)[32], bool skip = false) {
  if (skip) {  // profiling: keep the math alive without touching shared memory
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) acc += v[j];
    if (acc == 123.456f) *reinterpret_cast<float*>(act) = acc;
    return;
  }
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    uint4 w;
    w.x = umma::pack2<kF16>(v[8 * j4 + 0], v[8 * j4 + 1]);
    w.y = umma::pack2<kF16>(v[8 * j4 + 2], v[8 * j4 + 3]);
    w.z = umma::pack2<kF16>(v[8 * j4 + 4], v[8 * j4 + 5]);
    w.w = umma::pack2<kF16>(v[8 * j4 + 6], v[8 * j4 + 7]);
    *reinterpret_cast<uint4*>(act + umma::b_chunk_offset(k, (col0 >> 3) + j4, kLbo)) = w;
  }
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
// Row k of the operand as a 32-bit shared address + its swizzle key; a 16-byte chunk n8 of that row
// then lives at row_addr + (n8 >> 3) * LBO + (((n8 & 7) ^ key) << 4)  (see umma::b_chunk_offset).
struct RowAddr {
  uint32_t base, xbase;   // xbase = base ^ (key << 4): bits 4..6 of base are zero (1024-byte aligned operand, 128-byte rows)
  __device__ __forceinline__ RowAddr(uint32_t act_s, int k)
      : base(act_s + (uint32_t)((k >> 3) * 1024 + (k & 7) * 128)), xbase(base ^ ((uint32_t)(k & 7) << 4)) {}
  __device__ __forceinline__ uint32_t chunk(int n8) const {   // one LOP3 + an immediate offset for a compile-time n8
    return (xbase ^ (((uint32_t)n8 & 7u) << 4)) + (uint32_t)(n8 >> 3) * kLbo;
  }
};
template <bool kF16>
__device__ __forceinline__ void store_row16(const RowAddr& ra, int col0, const float (&v)[16], bool skip = false) {
  if (skip) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc += v[j];
    if (acc == 123.456f) st_shared_v4(ra.base, 0, 0, 0, 0);
    return;
  }
#pragma unroll
  for (int j4 = 0; j4 < 2; ++j4)
    st_shared_v4(ra.chunk((col0 >> 3) + j4), umma::pack2<kF16>(v[8 * j4 + 0], v[8 * j4 + 1]),
                 umma::pack2<kF16>(v[8 * j4 + 2], v[8 * j4 + 3]), umma::pack2<kF16>(v[8 * j4 + 4], v[8 * j4 + 5]),
                 umma::pack2<kF16>(v[8 * j4 + 6], v[8 * j4 + 7]));
}
// Copy 16 columns [col0, col0+16) of stash row j (linear, 256 B per row) into operand row ra.
__device__ __forceinline__ void copy_row16(const RowAddr& ra, const uint8_t* stash, int j, int col0) {
  const uint4* src = reinterpret_cast<const uint4*>(stash + j * 256 + col0 * 2);
  const uint4 c0 = src[0], c1 = src[1];
  st_shared_v4(ra.chunk(col0 >> 3), c0.x, c0.y, c0.z, c0.w);
  st_shared_v4(ra.chunk((col0 >> 3) + 1), c1.x, c1.y, c1.z, c1.w);
}
// 16 consecutive columns [col0, col0+16) of operand row F of 128-point block `blk` of the radiance operand image in
// global memory (the same 16-bit, 128-byte-swizzled layout the kernel keeps in shared memory).
template <bool kF16>
__device__ __forceinline__ void img_store16(uint8_t* img, int64_t blk, int F, int col0, const float (&v)[16]) {
  uint8_t* row = img + (size_t)blk * kActBytes + (F >> 3) * 1024 + (F & 7) * 128 + (size_t)(col0 >> 6) * kLbo;
#pragma unroll
  for (int j4 = 0; j4 < 2; ++j4) {
    uint4 w;
    w.x = umma::pack2<kF16>(v[8 * j4 + 0], v[8 * j4 + 1]);
    w.y = umma::pack2<kF16>(v[8 * j4 + 2], v[8 * j4 + 3]);
    w.z = umma::pack2<kF16>(v[8 * j4 + 4], v[8 * j4 + 5]);
    w.w = umma::pack2<kF16>(v[8 * j4 + 6], v[8 * j4 + 7]);
    *reinterpret_cast<uint4*>(row + ((((((col0 & 63) >> 3) + j4) & 7) ^ (F & 7)) << 4)) = w;
  }
}

// the same store from a precomputed row pointer (row F of block `blk`) and swizzle key F & 7: with a compile-time col0 the
// chunk positions are immediates XOR the key
template <bool kF16>
__device__ __forceinline__ void img_row_store16(uint8_t* row, int key, int col0, const float (&v)[16]) {
  uint8_t* blk = row + (size_t)(col0 >> 6) * kLbo;
#pragma unroll
  for (int j4 = 0; j4 < 2; ++j4) {
    uint4 w;
    w.x = umma::pack2<kF16>(v[8 * j4 + 0], v[8 * j4 + 1]);
    w.y = umma::pack2<kF16>(v[8 * j4 + 2], v[8 * j4 + 3]);
    w.z = umma::pack2<kF16>(v[8 * j4 + 4], v[8 * j4 + 5]);
    w.w = umma::pack2<kF16>(v[8 * j4 + 6], v[8 * j4 + 7]);
    *reinterpret_cast<uint4*>(blk + ((((((col0 & 63) >> 3) + j4) & 7) ^ key) << 4)) = w;
  }
}

template <bool kF16>
__device__ __forceinline__ void store_elem(uint8_t* act, int k, int n, float v) {
  *reinterpret_cast<uint16_t*>(act + umma::b_chunk_offset(k, n >> 3, kLbo) + (n & 7) * 2) = umma::pack1<kF16>(v);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct KArgs {
  const uint8_t* image;
  const float* bias;
  const float* x;      // [n,3]
  const float* view;   // [n,3] or null
  int64_t n;
  float* sdf;          // [n] or null
  float* nabla;        // [n,3] or null
  float* feat;         // [n, feat_ld] or null
  int64_t feat_ld;
  float* rgb;          // [n,3] or null
  const float* normal_scale;  // [3] or null: normals fed to the radiance net = nabla * scale (UNISURF)
  uint8_t* feat_img;   // [ceil(n/128)][64 KB] geometry feature as the radiance operand image (out: mode 0, in: mode 1)
  long long* trace;    // profiling only: [3][kTraceCap][4] (event, step*2+tile, clock, pair) from CTA 0
};
}  // namespace
