/*
 * neurecon_b200 -- C-ABI of the B200-native ray-marched SDF volume-rendering hot path.
 *
 * This is the drop-in boundary.  The reference (SuwoongHeo/neurecon) is pure
 * Python/PyTorch and has no FFI of its own; each entry point below replaces the
 * reference Python function named in its comment (paths relative to the
 * reference root).  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller (PyTorch); the library
 *    never allocates or frees caller memory.  Tensors are contiguous fp32 unless
 *    stated.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *  - Return value: 0 = ok, negative = error (NR_ERR_*); the message is available
 *    through nr_last_error().  Nothing throws across the ABI.  No host syncs:
 *    all entry points are CUDA-graph capturable.
 *  - There is no CPU fallback anywhere behind this interface.
 */
#ifndef NEURECON_B200_H
#define NEURECON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NR_OK 0
#define NR_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define NR_ERR_CUDA (-2)      /* CUDA runtime error */
#define NR_ERR_WORKSPACE (-3) /* workspace too small */

#define NR_MAX_LAYERS 16

/* activation codes */
#define NR_ACT_NONE 0
#define NR_ACT_SOFTPLUS100 1 /* nn.Softplus(beta=100, threshold=20), models/base.py:202 */
#define NR_ACT_RELU 2
#define NR_ACT_SIGMOID 3

int nr_version(void);
/* copies the calling thread's last error message (NUL terminated) into buf */
int nr_last_error(char* buf, size_t n);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
long long nr_launch_count(void);
/* device properties the host side needs (SM count, max dynamic smem) */
int nr_device_info(int* sm_count, int* smem_optin, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * Network descriptors.  Weights are the EFFECTIVE matrices W = g*v/||v|| (weight_norm is
 * evaluated in PyTorch so autograd owns g and v; models/base.py:226-227), fp32 row-major
 * [out_dim, in_pad] with in_pad = in_dim rounded up to a multiple of 4 (zero filled).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t n_layers;                /* D+1 linear layers (models/base.py:178) */
  int32_t multires;                /* Embedder frequencies, <0 = identity (base.py:67-81) */
  int32_t skip_layer;              /* layer whose input is cat([h, pe])/sqrt2 (base.py:248-250), -1 = none */
  int32_t width;                   /* hidden width W */
  int32_t in_dim[NR_MAX_LAYERS];   /* logical input width of layer i */
  int32_t out_dim[NR_MAX_LAYERS];  /* logical output width of layer i */
  const float* W[NR_MAX_LAYERS];   /* [out_dim, pad4(in_dim)]; 1/sqrt2 of the skip already folded in */
  const float* b[NR_MAX_LAYERS];   /* [out_dim] */
  const void* umma_image;          /* bf16 pre-swizzled operand image for the tcgen05 path (or NULL) */
  const float* umma_bias;          /* fp32 bias table for the tcgen05 path (or NULL) */
} nr_sdf_net_t;                    /* replaces ImplicitSurface parameters, models/base.py:131-282 */

typedef struct {
  int32_t n_layers;                /* D+1 (models/base.py:340) */
  int32_t multires;                /* PE of x (embed_multires) */
  int32_t multires_view;           /* PE of view dirs (embed_multires_view) */
  int32_t feat_dim;                /* W_geo_feat */
  int32_t in_dim[NR_MAX_LAYERS];
  int32_t out_dim[NR_MAX_LAYERS];
  const float* W[NR_MAX_LAYERS];
  const float* b[NR_MAX_LAYERS];
  const void* umma_image;
  const float* umma_bias;
} nr_radiance_net_t;               /* replaces RadianceNet parameters, models/base.py:312-391 */

/* ------------------------------------------------------------------------------------------
 * fp32 tier of the MLPs (<= 1e-4 relative to the reference).  Workspace is caller-allocated;
 * query the size first.  n = number of points.
 * ------------------------------------------------------------------------------------------ */
/* ImplicitSurface.forward(x, return_h) -- models/base.py:243-263.  feat may be NULL. */
size_t nr_sdf_forward_f32_workspace(const nr_sdf_net_t* net, int64_t n);
int nr_sdf_forward_f32(const nr_sdf_net_t* net, const float* x, int64_t n, float* sdf,
                       float* feat, int64_t feat_ld, void* ws, size_t ws_bytes, void* stream);

/* ImplicitSurface.forward_with_nablas(x) -- models/base.py:265-282 (analytic forward-mode
 * d sdf / d x instead of autograd.grad).  feat may be NULL. */
size_t nr_sdf_forward_nablas_f32_workspace(const nr_sdf_net_t* net, int64_t n);
int nr_sdf_forward_nablas_f32(const nr_sdf_net_t* net, const float* x, int64_t n, float* sdf,
                              float* nabla, float* feat, int64_t feat_ld, void* ws,
                              size_t ws_bytes, void* stream);

/* RadianceNet.forward(x, view_dirs, normals, feat) -- models/base.py:372-391. */
size_t nr_radiance_forward_f32_workspace(const nr_radiance_net_t* net, int64_t n);
int nr_radiance_forward_f32(const nr_radiance_net_t* net, const float* x, const float* view,
                            const float* normals, const float* feat, int64_t feat_ld, int64_t n,
                            float* rgb, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Building blocks of the training path (fp32).  neurecon_b200/models/autograd.py composes them into
 * torch.autograd.Functions that replace autograd through ImplicitSurface.forward_with_nablas
 * (create_graph=True, models/base.py:265-282) and RadianceNet.forward: the forward-mode network is
 * differentiated by hand, including the second-order path the eikonal loss needs (neus.py:458).
 * ------------------------------------------------------------------------------------------ */
/* Y[M,N] = epilogue(A[M,K] W[N,K]^T + bias).  mode: 0 none, 1 softplus(beta=100) (S, if given, receives
 * its derivative), 2 ReLU, 3 sigmoid, 4 tangent: Y = (A W^T) * aux[row % m_val, col] (no bias),
 * 5 tangent-linear: Y = A W^T (no bias).  lda/ldw multiples of 4. */
int nr_gemm_f32(const float* A, int32_t lda, const float* W, int32_t ldw, const float* bias, int64_t M,
                int32_t N, int32_t K, float* Y, int32_t ldy, int32_t mode, float* S, int32_t lds,
                const float* aux, int32_t ldaux, int64_t m_val, void* stream);
/* Same contract on the tensor cores (csrc/gemm_tc.cu): operands converted to fp16 (operand_f16 = 1) or bf16 on their
 * way into shared memory, fp32 accumulation in TMEM; N <= 256, W must fit in shared memory as 16-bit. */
int nr_gemm_tc(const float* A, int32_t lda, const float* W, int32_t ldw, const float* bias, int64_t M, int32_t N,
               int32_t K, float* Y, int32_t ldy, int32_t mode, float* S, int32_t lds, const float* aux,
               int32_t ldaux, int64_t m_val, int32_t operand_f16, void* stream);
/* dW[N,K] += G[rows,N]^T X[rows,K]  (weight gradient; split over rows, fp32 atomics) */
int nr_gemm_tn_f32(const float* G, int32_t ldg, const float* X, int32_t ldx, int64_t rows, int32_t N,
                   int32_t K, float* dW, int32_t lddw, void* stream);
/* Same on the tensor cores (csrc/gemm_tc.cu): both operands converted to 16 bits, MN-major tiles, fp32 atomics. */
int nr_gemm_tn_tc(const float* G, int32_t ldg, const float* X, int32_t ldx, int64_t rows, int32_t N, int32_t K,
                  float* dW, int32_t lddw, int32_t operand_f16, void* stream);
/* out[N] += column sums of G[rows,N]  (bias gradient) */
int nr_colsum_f32(const float* G, int32_t ldg, int64_t rows, int32_t N, float* out, void* stream);
/* in place: gh <- g_z = gh*S + sum_c gt_c*u_c*100*S*(1-S),  gt_c <- g_u_c = gt_c*S.
 * gh,S: [n,N]; gt,u: [3n,N] (component-major row blocks).  u_scaled = 1: `u` holds the layer's output tangent S*u
 * (the next layer's input, kept by the forward pass anyway) and the factor becomes 100*(1-S).
 * gz_colsum (optional, [N]) += column sums of g_z, the layer's bias gradient. */
int nr_sdf_bwd_act_f32(float* gh, int32_t ldgh, float* gt, int32_t ldgt, const float* S, int32_t lds,
                       const float* u, int32_t ldu, int64_t n, int32_t N, int32_t u_scaled, float* gz_colsum,
                       void* stream);
/* in place: x *= (ref > 0) (mode 0, ReLU backward) or x *= ref*(1-ref) (mode 1, sigmoid backward) */
int nr_act_bwd_f32(float* x, int32_t ldx, const float* ref, int32_t ldr, int64_t rows, int32_t N,
                   int32_t mode, void* stream);
/* Embedder.forward (models/base.py:46-64) into pe[:, col_off:], and/or its three tangents d pe/d x_c
 * stacked as row blocks [c*n + m] into tpe[:, tcol_off:] (in_dim = 3 only for the tangents). */
int nr_embed_f32(const float* x, int64_t n, int32_t in_dim, int32_t multires, float* pe, int32_t ld,
                 int32_t col_off, float* tpe, int32_t ldt, int32_t tcol_off, void* stream);

/* NeRF++ background network, NeRF.forward(input_pts, input_views) -- models/base.py:395-453
 * (use_view_dirs=True).  Plain (not weight-normed) layers; weights fp32 [out, pad4(in)]. */
typedef struct {
  int32_t depth;          /* D: number of pts_linears */
  int32_t width;          /* W */
  int32_t input_dim;      /* 4 for the inverted-sphere parametrisation [x/r, 1/r] */
  int32_t multires;       /* PE of the point (10) */
  int32_t multires_view;  /* PE of the view direction (4) */
  int32_t skip;           /* the embedding is re-concatenated in front of h after this layer (base.py:434) */
  const float* pts_W[NR_MAX_LAYERS];
  const float* pts_b[NR_MAX_LAYERS];
  const float* alpha_W;   const float* alpha_b;    /* [1, W] */
  const float* feature_W; const float* feature_b;  /* [W, W] */
  const float* views_W;   const float* views_b;    /* [W/2, W + pe_view] */
  const float* rgb_W;     const float* rgb_b;      /* [3, W/2] */
} nr_nerf_net_t;
size_t nr_nerf_forward_f32_workspace(const nr_nerf_net_t* net, int64_t n);
/* x [n,input_dim], view [n,3] -> sigma [n] (raw), rgb [n,3] (sigmoid) */
int nr_nerf_forward_f32(const nr_nerf_net_t* net, const float* x, const float* view, int64_t n,
                        float* sigma, float* rgb, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Ray geometry and inverse-CDF sampling -- utils/rend_util.py
 * ------------------------------------------------------------------------------------------ */
/* near_far_from_sphere(o, d, r) -- rend_util.py:167-185.  near/far: [R]. */
int nr_near_far_from_sphere(const float* rays_o, const float* rays_d, int64_t R, float r,
                            float* near, float* far, void* stream);

/* get_rays(c2w, intrinsics, H, W, N_rays) -- rend_util.py:95-164 (the step in front of volume_render in
 * all three callers).  pose [B,4,4] camera-to-world, intr [B,5] = fx, fy, cx, cy, skew, select_inds
 * int64 [B,N] (pixel = h*W + w) or NULL for all N = H*W pixels; rays_o, rays_d [B,N,3] (not normalised). */
int nr_get_rays(const float* pose, const float* intr, const int64_t* select_inds, int32_t B, int32_t W,
                int64_t N, float* rays_o, float* rays_d, void* stream);

/* sample_pdf(bins, weights, N, det, eps) -- rend_util.py:255-292.
 * bins [R,M], weights [R,M-1], u [R,N] or NULL (det: torch.linspace(0,1,N) bit-exact),
 * samples [R,N]; optional outputs: below/above int32 [R,N] (the gathered indices) and the
 * CDF [R,M] the search ran on.  cdf_is_given != 0 turns this into sample_cdf
 * (rend_util.py:294-327): `weights` then holds the un-normalised CDF [R,M-1].
 * Even M <= 128 with none of the optional outputs requested runs one thread per ray (64 rows staged per block),
 * everything else one warp per ray; both return the same bits. */
int nr_sample_pdf(const float* bins, const float* weights, const float* u, int64_t R, int32_t M,
                  int32_t N, int32_t cdf_is_given, float eps, float* samples, int32_t* below,
                  int32_t* above, float* cdf_out, void* stream);

/* Lattice points i0 .. i0+count-1 of the N^3 grid of extract_mesh (utils/mesh_util.py:82-100), fp32
 * [count,3], computed in float64 like the reference.  faithful != 0 reproduces the reference's
 * true-division indices (a sheared lattice), 0 gives the intended integer lattice. */
int nr_grid_points(int64_t i0, int64_t count, int32_t N, double volume_size, int32_t faithful, float* pts,
                   void* stream);

/* Iso-surface extraction on the device: what the reference does on the host with skimage.measure.marching_cubes
 * (utils/mesh_util.py:33-35) after copying the grid over PCIe.  Indexed triangle mesh of {vol = level} on the lattice
 * vol [Nx,Ny,Nz] (row-major, x slowest); a sample is inside when vol < level; vertices are the linear edge crossings,
 * (index + t) * spacing per axis, every crossing ONE vertex shared by the triangles around it; case table =
 * neurecon_b200/mc_tables.py (face-consistent disambiguation: closed meshes).  Canonical output order (vertices by owner
 * lattice point then axis, triangles by cell then table order), so a CPU restatement reproduces both arrays bit for bit.
 * nr_mc_count: flags / cases [Nx*Ny*Nz] u8, vlocal [Nx*Ny*Nz] u16 (a point's first vertex inside its block of 256 points),
 *   block_v / block_f [nr_mc_blocks] i32 (first vertex / triangle of a block, after the scan), totals: 2 x int64 ON THE DEVICE
 *   (vertices, triangles) -- the one number the host must read to allocate outputs.  n_tris: the table's u8[256] triangle
 *   counts on the device; workspace: nr_mc_count_workspace bytes.
 * nr_mc_generate: verts [V,3] f32, faces [F,3] i32; tri_table: the table's int8 [256][32] on the device; ascent != 0
 *   flips the winding (default: right-hand normals point towards decreasing values, skimage's 'descent'). */
size_t nr_mc_count_workspace(int32_t Nx, int32_t Ny, int32_t Nz);
int64_t nr_mc_blocks(int32_t Nx, int32_t Ny, int32_t Nz);
int nr_mc_count(const float* vol, int32_t Nx, int32_t Ny, int32_t Nz, float level, const uint8_t* n_tris, uint8_t* flags,
                uint8_t* cases, uint16_t* vlocal, int32_t* block_v, int32_t* block_f, int64_t* totals, void* workspace,
                size_t workspace_bytes, void* stream);
int nr_mc_generate(const float* vol, int32_t Nx, int32_t Ny, int32_t Nz, float level, float spacing_x, float spacing_y,
                   float spacing_z, int32_t ascent, const uint8_t* flags, const uint8_t* cases, const uint16_t* vlocal,
                   const int32_t* block_v, const int32_t* block_f, const uint8_t* n_tris, const int8_t* tri_table, float* verts,
                   int32_t* faces, void* stream);

/* ------------------------------------------------------------------------------------------
 * NeuS -- models/frameworks/neus.py
 * ------------------------------------------------------------------------------------------ */
/* Ray prologue of render_rayschunk (neus.py:169-172,184-210): normalises rays_d, computes
 * near/far (or the bypass values when >= 0... use NAN for "no bypass"), the coarse depths
 * d_coarse = near*(1-t)+far*t with t = linspace(0,1,N_samples), and the coarse points.
 * Outputs: dirs [R,3], near [R], far [R], d_new [R,N_samples], pts_new [R,N_samples,3]. */
int nr_neus_ray_setup(const float* rays_o, const float* rays_d, int64_t R, float radius,
                      float near_bypass, float far_bypass, int32_t n_samples, float* dirs,
                      float* near, float* far, float* d_new, float* pts_new, void* stream);

/* One iteration of the 'official_solution' up-sampler (neus.py:249-277), one warp per ray:
 *   1. merge the n_new freshly evaluated samples (d_new, sdf_new) into the sorted per-ray
 *      state (d_buf, sdf_buf: [R, cap], m_cur valid entries)            -- neus.py:272-276
 *   2. if n_next > 0: slopes -> logistic CDF with s = 64*2^iter -> alpha -> weights ->
 *      sample_pdf(n_next) -> d_next [R,n_next], pts_next [R,n_next,3]   -- neus.py:253-271
 *      (u_next: [R,n_next] uniforms, NULL = deterministic linspace)
 *   3. if n_next == 0 (final): emit pts [R,m,3], d_mid [R,m-1], pts_mid [R,m-1,3]
 *                                                                       -- neus.py:284-288
 * nab_buf [R,cap,3] / nab_new [R,n_new,3] (both or neither): the samples' normals ride along with the merge, so that
 * after the final step (sdf_buf, nab_buf) ARE forward_with_nablas at the sorted samples (neus.py:291) and the render
 * need not evaluate the network there again. */
int nr_neus_upsample_step(const float* rays_o, const float* dirs, int64_t R, float* d_buf,
                          float* sdf_buf, int32_t cap, int32_t m_cur, const float* d_new,
                          const float* sdf_new, int32_t n_new, int32_t iter, int32_t n_next,
                          const float* u_next, float* d_next, float* pts_next, float* pts_all,
                          float* d_mid, float* pts_mid, float* nab_buf, const float* nab_new, void* stream);

/* Alpha, exclusive-cumprod transmittance and compositing (neus.py:28-35,57-70,296,346-381)
 * as one warp-scan pass.  sdf [R,M], nablas [R,M,3] (NULL if !calc_normal), radiance
 * [R,M-1,3], d_mid [R,M-1], inv_s: device scalar s = exp(ln_s*speed) (neus.py:108-109).
 * Per-ray outputs rgb [R,3], depth [R], acc [R], normals [R,3] (or NULL); optional
 * per-sample outputs cdf [R,M], alpha [R,M-1], weights [R,M-1] (NULL to skip). */
int nr_neus_composite(const float* sdf, const float* nablas, const float* radiance,
                      const float* d_mid, const float* s_dev, int64_t R, int32_t M,
                      int32_t white_bkgd, float* rgb, float* depth, float* acc, float* normals,
                      float* cdf_out, float* alpha_out, float* weights_out, void* stream);

/* sdf_to_w with a fixed slope s (neus.py:28-70, the weights of the 'direct_use' / 'direct_more' up-samplers, :216-243):
 * sdf [R,M] -> w [R,M-1] = alpha_i * prod_{j<i} (1 - alpha_j + 1e-10), alpha from sigmoid(sdf * s). */
int nr_neus_sdf_to_w(const float* sdf, float s, int64_t R, int32_t M, float* w, void* stream);

/* NeuS with the NeRF++ background (N_outside > 0, neus.py:303-343).
 * nr_neus_outside_points: d_vals [R, M1+n_out] = cat(d_mid, far / flip(linspace(0,1,n_out+2)[1:-1]))
 * (u [R,n_out]: stratified jitter uniforms or NULL) and the inverted-sphere inputs x_out [R, M1+n_out, 4]
 * = [p/|p|, 1/|p|] of NeRF.forward.
 * nr_neus_composite_bg: like nr_neus_composite, with alpha / radiance taken from the background net
 * (sigma_out raw, radiance_out; both [R, M-1+n_out]) wherever the mid point lies outside the bounding
 * sphere and for the appended samples.  Optional per-sample outputs over M-1+n_out entries. */
int nr_neus_outside_points(const float* rays_o, const float* dirs, const float* far, const float* d_mid,
                           int64_t R, int32_t M1, int32_t n_out, const float* u, float* d_vals, float* x_out,
                           void* stream);
int nr_neus_composite_bg(const float* sdf, const float* nablas, const float* radiance, const float* rays_o,
                         const float* dirs, const float* d_vals, const float* sigma_out,
                         const float* radiance_out, const float* s_dev, float radius, int64_t R, int32_t M,
                         int32_t n_out, int32_t white_bkgd, float* rgb, float* depth, float* acc,
                         float* normals, float* cdf_out, float* alpha_out, float* weights_out,
                         float* radiance_blend_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * VolSDF -- models/frameworks/volsdf.py
 * ------------------------------------------------------------------------------------------ */
/* error_bound(d_vals, sdf, alpha, beta) -- volsdf.py:38-74.  d_vals, sdf [R,M]; alpha / beta are
 * device arrays read at [ray * stride] (stride 0 = one scalar for all rays); bounds [R,M-1] and/or
 * the per-ray maximum bound_max [R] (either may be NULL).  NaN -> inf as in the reference. */
int nr_volsdf_error_bound(const float* d_vals, const float* sdf, int64_t R, int32_t M,
                          const float* alpha, int32_t alpha_stride, const float* beta,
                          int32_t beta_stride, float* bounds, float* bound_max, void* stream);

/* Ray prologue of render_rayschunk (volsdf.py:169-172,402-427): normalised dirs [R,3], fars [R]
 * (`far`, or the exact sphere exit of get_sphere_intersection when sphere_radius > 0 -- rays that
 * miss are counted in *miss_count, the reference asserts on them), the dense initial depths
 * linspace(near, far, n_init) written to d_buf[:, :n_init] (row stride cap) and their points. */
int nr_volsdf_ray_setup(const float* rays_o, const float* rays_d, int64_t R, float near, float far,
                        float sphere_radius, int32_t n_init, float* dirs, float* fars,
                        int32_t* miss_count, float* d_buf, int32_t cap, float* pts_new, void* stream);

/* sdf = min(sdf, radius - |x|): the bounding-sphere background of VolSDF.forward_surface /
 * forward_surface_with_nablas (volsdf.py:310-325). */
int nr_sphere_min(const float* pts, float* sdf, int64_t n, float radius, void* stream);

/* One iteration of fine_sample (volsdf.py:129-270), one warp per ray, no host synchronisation.
 * it = 0: bound check of the m0 initial samples (sdf_new [R,n_new=m0]) with the network's beta;
 *         converged rays get their final samples, the others beta+ (:129) and n_up proposals.
 * it >= 1: merge last iteration's proposals (d_new_in, sdf_new [R,n_new=n_up]) into the sorted state
 *         (d_buf, sdf_buf [R,cap], m_cur valid), re-check, bisect beta+ (max_bisection steps), propose
 *         again -- or, at it == max_iter, sample with the last beta+ (iter_usage = -1).
 * Per-ray state: beta [R], status [R] (0 active / 1 done), outputs iter_usage [R], beta_map [R],
 * d_fine [R,n_final]; proposals d_new_out [R,n_up], pts_new [R,n_up,3] (zeros for finished rays).
 * u_final: [R,n_final] uniforms for the final inverse-CDF draw, NULL = deterministic. */
int nr_volsdf_fine_iter(const float* rays_o, const float* dirs, const float* fars, int64_t R,
                        float* d_buf, float* sdf_buf, int32_t cap, int32_t m_cur, const float* sdf_new,
                        int32_t n_new, const float* d_new_in, const float* alpha_net,
                        const float* beta_net, float eps, int32_t it, int32_t max_iter,
                        int32_t max_bisection, int32_t n_up, int32_t n_final, const float* u_final,
                        int32_t m0, float* beta, int32_t* status, float* iter_usage, float* beta_map,
                        float* d_fine, float* d_new_out, float* pts_new, void* stream);

/* d_all = sort(cat(linspace(near, far, n_coarse), d_fine)) and its points (volsdf.py:436-444). */
int nr_volsdf_merge(const float* rays_o, const float* dirs, const float* fars, int64_t R, float near,
                    int32_t n_coarse, const float* d_fine, int32_t n_fine, float* d_all, float* pts,
                    void* stream);

/* Density + compositing (volsdf.py:452-503).  Inside samples: sdf [R,M_in], nablas [R,M_in,3] or
 * NULL, radiance [R,M_in,3], d_in [R,M_in]; alpha/beta device scalars (forward_ab, :306-308).
 * Optional NeRF++ samples appended after them: sigma_out [R,M_out] (raw), radiance_out
 * [R,M_out,3], d_out [R,M_out] (M_out = 0: none).  Outputs rgb [R,3], depth [R], acc [R], normals
 * [R,3] or NULL; optional per-sample sigma_all [R,M], p_out [R,M-1], tau_out [R,M-1]. */
int nr_volsdf_composite(const float* sdf, const float* nablas, const float* radiance,
                        const float* d_in, const float* alpha_dev, const float* beta_dev, int64_t R,
                        int32_t M_in, const float* sigma_out, const float* radiance_out,
                        const float* d_out, int32_t M_out, int32_t white_bkgd, float* rgb, float* depth,
                        float* acc, float* normals, float* sigma_all, float* p_out, float* tau_out,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * UNISURF -- models/ray_casting.py, models/frameworks/unisurf.py
 * ------------------------------------------------------------------------------------------ */
/* unisurf.py:124-131 + ray_casting.py:69-77: normalised dirs [R,3], near/far [R] from the sphere of
 * interest (NAN bypass = none) and the n_steps uniform proposal points pts [R,n_steps,3]. */
int nr_unisurf_ray_setup(const float* rays_o, const float* rays_d, int64_t R, float radius,
                         float near_bypass, float far_bypass, int32_t n_steps, float* dirs, float* near,
                         float* far, float* pts, void* stream);
/* ray_casting.py:84-137: val [R,n_steps] = surface_query_fn(proposals); first sign change of
 * (val - logit_tau), the three masks (uint8 [R]), the secant bracket and first estimate
 * (state [5][R] = d_low, f_low, d_high, f_high, d_pred) and its point pts_pred [R,3]. */
int nr_unisurf_first_crossing(const float* val, const float* rays_o, const float* dirs, const float* near,
                              const float* far, int64_t R, int32_t n_steps, float logit_tau, float* state,
                              uint8_t* mask, uint8_t* mask_sign_change, uint8_t* mask_0_free,
                              float* pts_pred, void* stream);
/* run_secant_method, one step (ray_casting.py:16-29): f_mid [R] = surface_query_fn(pts_pred). */
int nr_unisurf_secant_step(const float* f_mid, float logit_tau, const float* rays_o, const float* dirs,
                           const uint8_t* mask, int64_t R, float* state, float* pts_pred, void* stream);
/* ray_casting.py:139-151 + unisurf.py:147-207: depth_surface [R] (clamped to [near, far]),
 * surface_pts [R,3], the n_query interval + n_free free-space depths merged and sorted d_all
 * [R,n_query+n_free] and their points.  u_int / u_free: stratified jitter uniforms or NULL. */
int nr_unisurf_sample(const float* rays_o, const float* dirs, const float* near, const float* far,
                      const float* state, const uint8_t* mask, const uint8_t* mask_sign_change,
                      const uint8_t* mask_0_free, int64_t R, float interval, float too_close,
                      int32_t n_query, int32_t n_free, const float* u_int, const float* u_free,
                      float* depth_surface, float* surface_pts, float* d_all, float* pts, void* stream);
/* unisurf.py:216-240: occupancy logits [R,M] -> alpha, exclusive-cumprod weights, rgb / depth / acc /
 * normals; optional per-sample alpha_out, weights_out [R,M]. */
int nr_unisurf_composite(const float* logits, const float* nablas, const float* radiance,
                         const float* d_all, int64_t R, int32_t M, int32_t white_bkgd, float* rgb,
                         float* depth, float* acc, float* normals, float* alpha_out, float* weights_out,
                         void* stream);

/* sphere_tracing_surface_points, one iteration (ray_casting.py:178-183): d[mask] += val[mask]
 * (val = NULL: only emit the points of the current d), mask cleared where d leaves [0, far], pts = o + d*dir. */
int nr_sphere_trace_step(const float* val, const float* rays_o, const float* dirs, float far, int64_t R,
                         float* d, uint8_t* mask, float* pts, void* stream);

/* ------------------------------------------------------------------------------------------
 * NeRF++ background helpers -- utils/rend_util.py:188-234, models/frameworks/volsdf.py:456-467
 * ------------------------------------------------------------------------------------------ */
/* get_sphere_intersection (rend_util.py:188-210): exact ray / sphere near and far (clamped at 0) and the hit mask
 * (bytes, may be NULL).  rays_d normalised.  r is a double because the reference squares it as a Python float before it
 * meets the fp32 tensors. */
int nr_sphere_intersection(const float* rays_o, const float* rays_d, int64_t R, double r, float* near, float* far,
                           uint8_t* mask, void* stream);
/* get_dvals_from_radius (rend_util.py:213-234): depth at which |o + t d| = rs, rs [R,N] -> d_vals [R,N]; far_end
 * selects the far (else the near, clamped at 0) intersection.  The reference asserts rs^2 > |o|^2 - (o.d)^2 on the
 * host (:225); here every violating entry adds 1 to *bad_count (device int, may be NULL) and yields NaN. */
int nr_dvals_from_radius(const float* rays_o, const float* rays_d, const float* rs, int64_t R, int32_t N,
                         int32_t far_end, float* d_vals, int32_t* bad_count, void* stream);
/* VolSDF's inverted-sphere samples (volsdf.py:456-467): radii radius / flip(linspace(0,1,n_out+2)[1:-1]), jittered
 * between the mid points of their neighbours with u [R,n_out] (NULL = none), their depths d_out [R,n_out]
 * (get_dvals_from_radius) and the NeRF++ inputs x_out [R,n_out,4] = [p / rs, 1 / rs]. */
int nr_volsdf_outside_points(const float* rays_o, const float* dirs, int64_t R, float radius, int32_t n_out,
                             const float* u, float* d_out, float* x_out, int32_t* bad_count, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backward of the compositing passes (training): the adjoints of nr_neus_composite[_bg], nr_volsdf_composite and
 * nr_unisurf_composite, one launch each, one warp per ray (csrc/composite_bwd.cu).  Inputs: what the forward read, its
 * per-sample outputs, and the upstream gradients of rgb [R,3], depth [R], acc [R], normals [R,3] and the weights
 * [R,K] (each may be NULL = zero).  Replaces what autograd records for neus.py:296-352, volsdf.py:452-503,
 * unisurf.py:216-240.
 * ------------------------------------------------------------------------------------------ */
/* NeuS.  K = M-1+n_out weights.  radiance_used [R,K,3]: the radiance the forward composited (the blend when n_out > 0),
 * d_vals [R,K] (= d_mid when n_out = 0).  Outputs g_sdf [R,M], g_s_part [R] (sum over rays = dL/ds, s = exp(ln_s*speed)),
 * g_radiance [R,M-1,3], g_nablas [R,M,3] (iff g_normals); with the background also g_sigma_out [R,K] and
 * g_radiance_out [R,K,3]. */
int nr_neus_composite_bwd(const float* sdf, const float* cdf, const float* alpha, const float* weights,
                          const float* radiance_used, const float* d_vals, const float* nablas, const float* s_dev,
                          const float* acc, const float* depth, int64_t R, int32_t M, int32_t n_out,
                          const float* rays_o, const float* dirs, const float* sigma_out, float radius,
                          int32_t white_bkgd, const float* g_rgb, const float* g_depth, const float* g_acc,
                          const float* g_normals, const float* g_weights, float* g_sdf, float* g_s_part,
                          float* g_radiance, float* g_nablas, float* g_sigma_out, float* g_radiance_out, void* stream);
/* VolSDF.  K = M_in+M_out-1.  sigma_all [R,M_in+M_out], p, tau [R,K] from the forward.  Outputs g_sdf [R,M_in],
 * g_alpha_part / g_beta_part [R] (sums = dL/dalpha, dL/dbeta of forward_ab), g_radiance [R,M_in,3], g_nablas
 * [R,M_in,3] (iff g_normals), g_sigma_out [R,M_out], g_radiance_out [R,M_out,3].  A ray whose product of p contains
 * zeros gets the exact one-zero gradient (see the source). */
int nr_volsdf_composite_bwd(const float* sdf, const float* sigma_all, const float* p, const float* tau,
                            const float* radiance, const float* d_in, const float* nablas, const float* alpha_dev,
                            const float* beta_dev, const float* acc, const float* depth, int64_t R, int32_t M_in,
                            const float* radiance_out, const float* d_out, int32_t M_out, int32_t white_bkgd,
                            const float* g_rgb, const float* g_depth, const float* g_acc, const float* g_normals,
                            const float* g_weights, float* g_sdf, float* g_alpha_part, float* g_beta_part,
                            float* g_radiance, float* g_nablas, float* g_sigma_out, float* g_radiance_out,
                            void* stream);
/* UNISURF.  Outputs g_logits [R,M], g_radiance [R,M,3], g_nablas [R,M,3] (iff g_normals). */
int nr_unisurf_composite_bwd(const float* logits, const float* alpha, const float* weights, const float* radiance,
                             const float* d_all, const float* nablas, const float* acc, const float* depth, int64_t R,
                             int32_t M, int32_t white_bkgd, const float* g_rgb, const float* g_depth,
                             const float* g_acc, const float* g_normals, const float* g_weights, float* g_logits,
                             float* g_radiance, float* g_nablas, void* stream);

/* ------------------------------------------------------------------------------------------
 * Training GEMMs with 16-bit tensors in HBM (csrc/gemm16.cu): the per-layer building blocks of the reverse-mode
 * training path of the SDF network (models/autograd_rev.py; reference semantics base.py:265-282 under create_graph
 * + neus.py:443-458).  fp16 rows in, fp16 (or fp32) rows out, fp32 weights, fp32 accumulation on tcgen05.
 * ------------------------------------------------------------------------------------------ */
#define NR_G16_LINEAR 0   /* y = acc + bias */
#define NR_G16_SOFTPLUS 1 /* y = softplus100(acc + bias), out2 = its derivative */
#define NR_G16_SCALE 2    /* y = aux_a * acc (+ aux_b) */
#define NR_G16_ADJ 3      /* y = aux_a * acc, out2 = 100 (1 - aux_a) * aux_b * acc */
#define NR_G16_RELU 4
#define NR_G16_SIGMOID 5
#define NR_G16_MASK 6     /* y = aux_a > 0 ? acc : 0 */
/* Y[M, N] = epilogue(A[M, K] W[N, K]^T).  A: fp16 rows of lda halves (a multiple of 8 covering K rounded up to 64; pad
 * columns finite -- they meet zero weights); W fp32 [N, ldw]; bias fp32 [N] or NULL; Y fp16 (y_half) or fp32 rows of ldy
 * elements, columns [N, round_up(N, 16)) written as zeros; aux_a / aux_b / out2: fp16 rows (NULL where the mode has none);
 * fp16 rows are read and written 32 bytes at a time (32-byte aligned, leading dimensions multiples of 16). */
int nr_gemm16(const void* A, int32_t lda, const float* W, int32_t ldw, const float* bias, int64_t M, int32_t N, int32_t K,
              void* Y, int32_t ldy, int32_t y_half, int32_t mode, const void* aux_a, int32_t ld_a, const void* aux_b,
              int32_t ld_b, void* out2, int32_t ld_o2, int32_t w_packed, void* stream);
/* w_packed != 0: W points to the fp16 shared-memory image of the weight matrix written by nr_gemm16_pack_w
 * (nr_gemm16_pack_w_bytes(N, K) bytes), which a CTA then fetches with bulk copies instead of converting W itself. */
size_t nr_gemm16_pack_w_bytes(int32_t N, int32_t K);
int nr_gemm16_pack_w(const float* W, int32_t ldw, int32_t N, int32_t K, void* img, void* stream);
/* dW[N, K] += scale * G[rows, N]^T X[rows, K]; G, X fp16 rows (ldg, ldx multiples of 64), dW fp32 (atomics). */
int nr_gemm16_tn(const void* G, int32_t ldg, const void* X, int32_t ldx, int64_t rows, int32_t N, int32_t K, float* dW,
                 int32_t lddw, float scale, void* stream);
/* dW += scale * (G^T X + G2^T X2): two products of the same shape in one launch, one pass of atomics (the SDF layers'
 * dW = zb^T h + p^T gb). */
int nr_gemm16_tn2(const void* G, int32_t ldg, const void* X, int32_t ldx, const void* G2, int32_t ldg2, const void* X2,
                  int32_t ldx2, int64_t rows, int32_t N, int32_t K, float* dW, int32_t lddw, float scale, void* stream);
/* Split-precision forward GEMM of the training path (precision 'fp16x2'): Y = epilogue((A_hi + A_lo)(W_hi + W_lo)^T) minus
 * the lo x lo term, in one accumulator: every 64-column chunk of A is loaded once; a hi chunk multiplies the W_hi and the
 * W_lo chunk of its k range, a lo chunk the W_hi chunk.  A: fp16 rows [M, lda] = [hi (Kp columns) | lo (Kp columns)], Kp = K
 * rounded up to 64.  Wimg: nr_gemm16_pack_w_split (per block of 128 output columns [W_hi | W_lo], rows past N zero).
 * mode: 0 linear, 1 softplus100 (+ out2 = its derivative), 2 scale (y = aux_a * acc: the reverse sweep), 4 relu, 5 sigmoid.
 * Y fp16 (y_half) with the lo part of the result lo_off columns to the right of the hi part (lo_off = 0: hi only), or fp32. */
int nr_gemm16_split(const void* A, int32_t lda, const void* Wimg, const float* bias, int64_t M, int32_t N, int32_t K, void* Y,
                    int32_t ldy, int32_t y_half, int32_t lo_off, int32_t mode, void* out2, int32_t ld_o2, const void* aux_a,
                    int32_t ld_a, void* stream);
/* The weight side of nr_gemm16_split for a fp32 matrix W [N, ldw]: all its column blocks in one launch (img:
 * nr_gemm16_pack_w_split_bytes). */
size_t nr_gemm16_pack_w_split_bytes(int32_t N, int32_t K);
int nr_gemm16_pack_w_split(const float* W, int32_t ldw, int32_t N, int32_t K, void* img, void* stream);
/* nr_pe16 with split-precision output: lo = fp16(v - hi) lo_off (lo_off2 for e2) columns to the right of the hi parts. */
int nr_pe16_split(const float* x, int64_t n, int32_t multires, void* e, int32_t ld, int32_t width, int32_t lo_off, void* e2,
                  int32_t ld2, int32_t off2, int32_t lo_off2, void* stream);
/* dst[r, c] = fp16(scale * src[r, c]), r < n, c < ncols (fp32 rows with stride ld_src -> fp16 rows with stride ld_dst, both in
 * elements); lo_off != 0: also the lo part fp16(scale * src - hi) lo_off columns to the right. */
int nr_cast_cols16(const float* src, int64_t ld_src, int64_t n, int32_t ncols, float scale, void* dst, int64_t ld_dst,
                   int32_t lo_off, void* stream);
/* out[N] += scale * column sums of a fp16 matrix (N <= 256). */
int nr_colsum16(const void* A, int32_t lda, int64_t rows, int32_t N, float scale, float* out, void* stream);
/* Embedder.forward (base.py:46-64) as fp16 rows e [n, ld] (columns [pe_dim, width) zero; optionally also into e2 at
 * column off2), the transposed Jacobian nabla = J^T (g0 + ge) and the Jacobian product gbar = scale * J nbar. */
int nr_pe16(const float* x, int64_t n, int32_t multires, void* e, int32_t ld, int32_t width, void* e2, int32_t ld2,
            int32_t off2, void* stream);
int nr_pe_jac_t(const float* x, int64_t n, int32_t multires, const float* g0, int32_t ldg0, const void* ge, int32_t ldge,
                int32_t ge_lo_off /* != 0: ge is a (hi, lo) pair, the lo part that many columns to the right */, float* nabla,
                void* stream);
int nr_pe_jac(const float* x, int64_t n, int32_t multires, const float* nbar, float scale, void* gbar, int32_t ld,
              int32_t width, void* g2, int32_t ld2, int32_t off2, void* stream);

/* Weight normalisation W = scale * g v / ||v||_row (base.py:226-227, nn.utils.weight_norm dim 0) of ALL layers of a
 * network in one launch, and its backward (dW -> dg, dv) in one more.  `table`: n_entries <= 16 records IN HOST MEMORY of
 * twelve 8-byte words {v*, g*, W*, dW*, dv*, dg* (device pointers), rows, cols, ldw, lddw, scale (double), first global
 * row}, passed on to the kernel by value (graph-capturable); one warp per row. */
int nr_weight_norm(const void* table, int32_t n_entries, int64_t total_rows, int32_t backward, void* stream);

/* ------------------------------------------------------------------------------------------
 * After the path in a training step: losses, gradient norm, Adam -- no host syncs.
 * ------------------------------------------------------------------------------------------ */
/* NeuS Trainer.forward losses (neus.py:443-478) and the gradients of their sum:
 *   loss_img  = mean |rgb - target|            (no masks), or  sum(|.| * m) / (sum(m) + 1e-10) with m = target_mask
 *               [& mask_ignore] (with_mask) or m = mask_ignore;
 *   loss_eik  = w_eikonal * mean((|nablas| - 1)^2) over [R,P];
 *   loss_mask = w_mask * mean BCE(clamp(mask_volume, 1e-3, 1-1e-3), target_mask)     (iff target_mask != NULL).
 * rgb/target_rgb [R,3], nablas [R,P,3], mask_volume [R], target_mask / mask_ignore [R] bytes or NULL.
 * sums4: 4 floats of scratch; losses4 = {loss_img, loss_eikonal, loss_mask, total} on the device;
 * g_rgb [R,3], g_nablas [R,P,3], g_mask_volume [R] = d total / d input. */
int nr_neus_loss(const float* rgb, const float* target_rgb, const float* nablas, const float* mask_volume,
                 const uint8_t* target_mask, const uint8_t* mask_ignore, int64_t R, int64_t P, float w_eikonal,
                 float w_mask, float* sums4, float* losses4, float* g_rgb, float* g_nablas, float* g_mask_volume,
                 void* stream);
/* table: n_tensors device records {float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64 numel}.
 * nr_grad_sqsum: out[0] = sum of squared gradients (train_util.calc_grad_norm's total, squared).
 * nr_adam_step: torch.optim.Adam's update (no weight decay / amsgrad), `step` = 1-based step count. */
int nr_grad_sqsum(const void* table, int32_t n_tensors, float* out, void* stream);
int nr_adam_step(const void* table, int32_t n_tensors, float lr, float beta1, float beta2, float eps, int64_t step,
                 void* stream);
/* The same update with the step count and the learning rate in device memory (`*step_dev` is incremented first), so
 * that a CUDA graph of a whole training step (train.py:196-210: forward, backward, optimizer.step(), scheduler) replays
 * with the right bias correction and the scheduler's current lr. */
int nr_adam_step_dev(const void* table, int32_t n_tensors, const float* lr_dev, float beta1, float beta2, float eps,
                     int64_t* step_dev, void* stream);

/* ------------------------------------------------------------------------------------------
 * bf16 tier: fused PE + SDF MLP (+ forward-mode normals) + radiance MLP on tcgen05 / TMEM.
 * The host packs the weights once into a pre-swizzled bf16 image (16 KB chunks = A tiles of
 * 128 features x 64 k) and describes the network as a short program of steps; the kernel
 * keeps all hidden activations on-chip.  Replaces, per launch, batchify_query over
 * ImplicitSurface.forward / forward_with_nablas / NeuS.forward_radiance / VolSDF.forward
 * (utils/train_util.py:23-71, models/base.py:243-282,372-391, neus.py:103-106).
 * ------------------------------------------------------------------------------------------ */
#define NR_UMMA_MAX_STEPS 24
#define NR_UMMA_EPI_HIDDEN 0  /* softplus(beta=100) hidden layer, tangents scaled by its derivative */
#define NR_UMMA_EPI_SDF_OUT 1 /* sdf row (replicated over 32 rows): sdf / nabla to global */
#define NR_UMMA_EPI_FEAT 2    /* geometry feature rows: to global and/or the radiance operand */
#define NR_UMMA_EPI_RELU 3    /* radiance hidden layer */
#define NR_UMMA_EPI_RGB 4     /* sigmoid, 3 rows, to global */
#define NR_UMMA_EPI_EXTRAS 5  /* split-K layer: refill the operand rows with the second operand, next step accumulates */
#define NR_UMMA_EPI_LINEAR 6  /* bias only, next operand in shared memory */
#define NR_UMMA_EPI_BWD 7     /* reverse mode: next operand = softplus'(z) of slot sig_slot times the accumulator (W^T g) */
#define NR_UMMA_EPI_NABLA 8   /* reverse mode, last step: accumulator rows = d sdf / d PE(x); embedding Jacobian -> nabla */

typedef struct {
  int32_t chunk_begin; /* first 16 KB weight chunk of this step; chunks ordered (k-chunk major, M-tile minor) */
  int32_t n_mt;        /* M-tiles of 128 output features (1 or 2) */
  int32_t k_steps;     /* K / 16, a multiple of 4 (K zero-padded to a multiple of 64) */
  int32_t n_cols;      /* MMA N: operand columns consumed (32, 64 or 128) */
  int32_t epi;         /* NR_UMMA_EPI_* */
  int32_t bias_off;    /* offset of n_mt*128 fp32 biases in the bias table */
  int32_t out_rows;    /* valid output features */
  int32_t pe_fill;     /* 1: rows [out_rows, out_rows+pe_dim) of the next operand are the embedding (skip) */
  int32_t to_rad;      /* EPI_FEAT: also build the radiance operand [feat | PE(x) | PE(view) | normals];
                          EPI_EXTRAS: which rows to write: 0 = the radiance extras, 1 = PE(x), 2 = PE(view) */
  int32_t accumulate;  /* 1: the step's MMAs add onto the previous step's accumulators (split-K over two operands) */
  int32_t sig_slot;    /* reverse mode: the 64 KB scratch slot of softplus'(z) this step writes (EPI_HIDDEN) or applies
                          (EPI_BWD, EPI_SDF_OUT) */
  int32_t aux_off;     /* reverse-mode EPI_SDF_OUT: offset in the bias table of the 256 fp32 weights of the sdf row */
} nr_umma_step_t;

typedef struct {
  int32_t n_steps;
  int32_t tangents;          /* 1: tiles of 32 points x (value + 3 tangents); 0: 128 points, values only */
  int32_t multires;          /* embedding of x for the SDF net */
  int32_t rad_multires;      /* embedding of x for the radiance net (<0: identity) */
  int32_t rad_multires_view; /* embedding of the view direction */
  int32_t rad_extra_rows;    /* zero-padded rows after the 256 feature rows of the radiance operand */
  int32_t operand_f16;       /* 1: fp16 operands (image packed as fp16), 0: bf16; fp32 accumulation either way */
  int32_t debug_flags;       /* profiling only (results invalid).  nr_mlp_umma_forward: 1 = no weight copies, 2 = no
                                epilogue math/stores.  nr_mlp_umma_reverse: 1 = no softplus' scratch stores, 2 = no scratch
                                loads, 4 = no backward epilogue, 8 = no forward (hidden-layer) epilogue */
  int32_t input_mode;        /* 0: points -> embedding -> SDF net; 1: radiance net alone on 128-point tiles, its operand
                                rows [0,256) bulk-copied from the feature image a previous launch wrote (feat_img);
                                2: NeRF++ background net on 128-point tiles: x = [n, input_dim] points, operand rows
                                [0, K0) = PE(x) (multires), view dirs embedded with rad_multires_view */
  int32_t input_dim;         /* components of a point (3; 4 for the NeRF++ inverted-sphere parametrisation) */
  int32_t reverse;           /* 1: reverse-mode normals program for nr_mlp_umma_reverse (128-point value tiles) */
  nr_umma_step_t steps[NR_UMMA_MAX_STEPS];
} nr_umma_program_t;

/* x [n,3]; view [n,3] or NULL; outputs (each may be NULL): sdf [n], nabla [n,3], feat [n,feat_ld],
 * rgb [n,3].  image: the packed 16-bit weight chunks, bias: fp32 bias table.  normal_scale: NULL, or
 * 3 device floats multiplied into the normals the radiance net sees (UNISURF's chunk-wide
 * F.normalize, unisurf.py:36); the nabla output stays unscaled.
 * feat_img: NULL, or ceil(n/128) x 64 KB of device memory holding the geometry feature as the radiance net's
 * 16-bit shared-memory operand image (128 points per block, MN-major, 128-byte swizzle): written by an EPI_FEAT
 * step when input_mode = 0, read when input_mode = 1 (then nabla [n,3] is an INPUT: the normals). */
int nr_mlp_umma_forward(const nr_umma_program_t* prog, const void* image, size_t image_bytes,
                        const float* bias, size_t bias_floats, const float* x, const float* view,
                        int64_t n, float* sdf, float* nabla, float* feat, int64_t feat_ld, float* rgb,
                        const float* normal_scale, void* feat_img, void* stream);

/* Weight (A-operand) images of the fused MLP kernels from a fp32 matrix W[row, k] = W[row * row_stride + k * col_stride]
 * (strides in elements: a transposed view or a broadcast row needs no copy): chunks of [128 rows x 64 k], K-major with the
 * 128-byte swizzle, ordered (k-chunk, M-tile); rows past `rows` and k past K are zero, k is padded to k_pad (a multiple of 64).
 * mode 0: fp16, 1: bf16, 2: split precision -- every chunk as a (hi, lo) pair, hi = fp16(w), lo = fp16((w - hi) 2^12). */
size_t nr_umma_pack_a_bytes(int32_t n_mt, int32_t k_pad, int32_t mode);
int nr_umma_pack_a(const float* W, int64_t row_stride, int64_t col_stride, int32_t rows, int32_t K, int32_t n_mt, int32_t k_pad,
                   int32_t mode, void* img, void* stream);

/* sdf + d sdf / d x (+ feature) with REVERSE-mode normals, the arithmetic of ImplicitSurface.forward_with_nablas'
 * autograd.grad (models/base.py:265-282): hidden layers forward on 128-point tiles, softplus' of every unit parked in
 * `workspace` (device memory, nr_mlp_umma_reverse_workspace bytes, contents irrelevant before and after), then the
 * backward sweep from the sdf row to the embedding on the same tiles.  Program: reverse = 1, steps EPI_HIDDEN x L
 * (sig_slot = layer), optional EPI_FEAT, EPI_SDF_OUT (aux_off = the sdf row's weights), EPI_BWD x (L-1) whose chunks
 * hold W_l^T (pe_fill marks the skip layer), EPI_NABLA with W_0^T.  Outputs as nr_mlp_umma_forward; feat_img, if not
 * NULL, receives the LAST HIDDEN activations as the radiance pass's operand image (to_rad on the last hidden step).
 * A program that ENDS with EPI_SDF_OUT is the forward sweep alone (ImplicitSurface.forward, models/base.py:247-263: sdf and
 * optionally the feature of every point): nabla and workspace may then be NULL. */
size_t nr_mlp_umma_reverse_workspace(const nr_umma_program_t* prog, int64_t n);
int nr_mlp_umma_reverse(const nr_umma_program_t* prog, const void* image, size_t image_bytes, const float* bias,
                        size_t bias_floats, const float* x, int64_t n, float* sdf, float* nabla, float* feat,
                        int64_t feat_ld, void* feat_img, void* workspace, size_t workspace_bytes, void* stream);

/* The same program on SPLIT-PRECISION operands (precision tier 'fp16x2': <= 1e-4 against the reference's fp32 on the tensor
 * pipe).  Every weight chunk of the image comes as a (hi, lo) fp16 pair (chunk_begin counts 16 KB chunks of that
 * interleaved image, umma_pack.pack_a_tiles_split), tile slots hold 64 points whose activations / gradients live as
 * [hi | lo] column blocks of the operand buffer; W h ~= W_hi h_hi + W_hi h_lo + W_lo h_hi as one N = 128 and one N = 64
 * MMA per k-step; softplus' as 16-bit codes.  A program may also END at EPI_SDF_OUT (sdf [+ feature] only, nabla and
 * workspace may then be NULL).  feat_img as above (the fp16 `hi` parts of the last hidden activations). */
size_t nr_mlp_split_reverse_workspace(const nr_umma_program_t* prog, int64_t n);
int nr_mlp_split_reverse(const nr_umma_program_t* prog, const void* image, size_t image_bytes, const float* bias,
                         size_t bias_floats, const float* x, int64_t n, float* sdf, float* nabla, float* feat,
                         int64_t feat_ld, void* feat_img, void* workspace, size_t workspace_bytes, void* stream);

/* The same network on CTA pairs (tcgen05.mma.cta_group::2, 2-CTA clusters): programs made of EPI_HIDDEN,
 * EPI_SDF_OUT and EPI_FEAT (to_rad = 0) steps whose weight chunks all come as M-tile pairs (n_mt = 2; the sdf row
 * replicated into both M-tiles).  Halves the shared-memory traffic per SM of nr_mlp_umma_forward; same outputs. */
int nr_mlp_umma2_forward(const nr_umma_program_t* prog, const void* image, size_t image_bytes,
                         const float* bias, size_t bias_floats, const float* x, int64_t n, float* sdf,
                         float* nabla, float* feat, int64_t feat_ld, void* feat_img, void* stream);

/* profiling hook (tools/trace_mlp.py): device buffer [3][2048][4] int64 receiving clock64 stamps of the
 * MMA <-> epilogue hand-offs of CTA 0; NULL disables. */
int nr_mlp_umma_set_trace(void* buf);

#ifdef __cplusplus
}
#endif
#endif /* NEURECON_B200_H */
