/*
 * neurecon_b200_devtools -- self-tests and micro-architecture probes (tcgen05 operand-layout self-test, tensor-pipe /
 * TMEM / ALU rate probes).  Development tooling, NOT part of the drop-in boundary: built into its own library
 * (neurecon_b200/lib/libneurecon_b200_devtools.so) so that the production library carries only the hot path.
 */
#ifndef NEURECON_B200_DEVTOOLS_H
#define NEURECON_B200_DEVTOOLS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------
 * tcgen05 self-test: D[128,N] = A[128,K] B[K,N] through the operand layouts / descriptors of
 * the fused MLP kernel.  a_image: bf16 K-major SWIZZLE_128B tiles (16 KB per 64 columns of K),
 * B: fp32 [K,N] row-major, D: fp32 [128,N].  variant = 0 is the production encoding.
 * ------------------------------------------------------------------------------------------ */
int nr_selftest_umma(const void* a_image, const float* B, int32_t K, int32_t N, float* D,
                     int32_t variant, void* stream);

/* CTA-pair variant: D[256,N] = A[256,K] B[K,N] with tcgen05.mma.cta_group::2 on a 2-CTA cluster (a_image: fp16 tiles
 * ordered k-chunk major / M-tile minor; N = 64, 128 or 256).  variant bit 0: each CTA writes its peer's half of B
 * through distributed shared memory. */
int nr_selftest_umma2(const void* a_image, const float* B, int32_t K, int32_t N, float* D,
                      int32_t variant, void* stream);

/* Tensor-pipe rate probe (tools/bench_umma_rate.py): cycles for n_mmas back-to-back 128 x N x 16 MMAs on operands
 * resident in shared memory, optionally under concurrent shared-memory store / bulk-copy traffic.
 * gsrc: >= 1 MiB of device memory (copy source); out: [3][grid] int64: [2*b], [2*b+1] = {issue..completion, issue loop}
 * cycles of block b, [2*grid + b] = 512-byte stores retired by one store warp.  n_mmas < 0: no MMAs, window of -n_mmas cycles. */
int nr_bench_umma(int32_t N, int32_t n_mmas, int32_t store_warps, int32_t bulk_copies, const void* gsrc,
                  int32_t grid, long long* out, void* stream);

/* Issue-rate probe (tools/probe_alu.py): cycles for iters x 8 independent ops per thread; op 0 ex2, 1 rcp,
 * 2 cvt.f16x2.f32, 3 fma, 4 cvt.bf16x2.f32, 5 lg2.  cycles: [grid] int64. */
int nr_probe_alu(int32_t op, int32_t threads, int32_t iters, int32_t grid, float* out, long long* cycles, void* stream);

/* TMEM read-rate probe: `warps` warps x iters x 4 tcgen05.ld.32x32b.x16 (2 KB each); out: [grid] int64 cycles. */
int nr_bench_ldtm(int32_t warps, int32_t iters, int32_t grid, long long* out, float* sink, void* stream);

/* Forward epilogue of mlp_rev_kernel in isolation (tools/probe_epi.py): 16 warps x iters chunks of 16 values per thread, the
 * accumulators read from shared memory; variant 0 = shipped math + codes + stores, 1 = no stores, 2 = lower-degree
 * polynomials, 3 = no bias FMA, 4 = scalar FFMA, 5 = math only, 6 = codes / conversions / stores only.
 * bias: [256] floats, scratch: 32 * 8192 * grid bytes, cycles: [grid] int64. */
int nr_probe_epi(int32_t variant, int32_t iters, int32_t grid, const float* bias, void* scratch, long long* cycles, void* stream);

/* Pipe-overlap probe (tools/probe_mix.py): iters x (8 x op_a interleaved with 8 x op_b) per thread on independent registers;
 * ops: 0 none, 1 fma.f32, 2 fma.f32x2, 3 ex2.approx, 4 lop3, 5 max.f32, 6 prmt, 7 cvt.f16x2.f32, 8 fma.f16x2, 9 add.s32,
 * 10 fma.f32x2 with distinct operands.  cycles: [grid] int64. */
int nr_probe_mix(int32_t op_a, int32_t op_b, int32_t threads, int32_t iters, int32_t grid, uint32_t* out, long long* cycles,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NEURECON_B200_DEVTOOLS_H */
