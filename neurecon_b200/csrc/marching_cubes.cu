// Iso-surface extraction on the device (the reference calls skimage.measure.marching_cubes on the
// host, utils/mesh_util.py:33-35, after copying the N^3 grid over PCIe).  Indexed mesh, every edge crossing one shared vertex.
//
// The case table comes from neurecon_b200/mc_tables.py (generated from its rule: a corner is inside when value < level,
// ambiguous faces cut off each inside corner on its own -- decided per face, so neighbouring cells agree and the mesh is
// closed).  Two passes over the lattice, one thread per lattice point p = (i, j, k), k fastest:
//   count:     flags[p]  = which of the three edges p owns (+x, +y, +z) carry a vertex;  cases[p] = the 8-bit case of the
//              cell whose origin is p (0 on the upper boundary);  two-level scans: inside each block of 256 points a
//              shuffle scan (vlocal[p], 16 bits: a block owns at most 768 vertices), across the blocks cub's ExclusiveSum
//              on the 1/256-size arrays of block sums -- 4 bytes of bookkeeping per lattice point (the first version
//              scanned two int32 arrays of the full size: 10 bytes per point, 2.3 ms per 512^3 grid);
//   generate:  p writes its vertices (linear interpolation, un-fused fp32 so that a numpy restatement gets the same bits)
//              and its cell's triangles (offsets: the block scan again), looking up the vertex index of edge
//              e = (axis a, owner point q) as block_v[q / 256] + vlocal[q] + popcount(flags[q] & ((1 << a) - 1)).
// Output order is therefore canonical: vertices by (owner point, axis), triangles by (cell, table order) -- a CPU restatement
// (oracle/mesh.py) reproduces vertices and faces bit for bit.  HBM-bound: the volume is read twice (the 8-corner stencil
// hits L1/L2), 4 bytes of bookkeeping per lattice point.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace {

constexpr int kMcBlock = 256;      // lattice points per block: the unit of the two-level scans

__device__ __forceinline__ int popc3(uint32_t f) { return __popc(f & 7u); }

// exclusive scan of one int per thread over a 256-thread block; returns the thread's prefix, *total = the block's sum
__device__ __forceinline__ int block_scan_excl(int v, int* total) {
  __shared__ int warp_sum[kMcBlock / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();                       // the previous use of warp_sum is over
  if (lane == 31) warp_sum[warp] = incl;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kMcBlock / 32; ++w) {
    const int sw = warp_sum[w];
    if (w < warp) base += sw;
    tot += sw;
  }
  *total = tot;
  return base + incl - v;
}

// flags / case of lattice point p (see the file header); 0 / 0 past the end
__device__ __forceinline__ void classify(const float* __restrict__ vol, int64_t p, int64_t n, int Nx, int Ny, int Nz, float level,
                                         uint32_t& f, uint32_t& c) {
  f = 0; c = 0;
  if (p >= n) return;
  // n < 2^31 (checked by the host): 32-bit index arithmetic -- 64-bit divisions were half of this kernel's instructions
  const uint32_t pu = (uint32_t)p, row = pu / (uint32_t)Nz;
  const int k = (int)(pu - row * (uint32_t)Nz), i = (int)(row / (uint32_t)Ny), j = (int)(row - (uint32_t)i * (uint32_t)Ny);
  const int64_t sx = (int64_t)Ny * Nz, sy = Nz;
  const bool hx = i + 1 < Nx, hy = j + 1 < Ny, hz = k + 1 < Nz;
  const bool in0 = vol[p] < level;
  bool c1 = false, c2 = false, c4 = false;
  if (hx) { c1 = vol[p + sx] < level; f |= (c1 != in0) ? 1 : 0; }
  if (hy) { c2 = vol[p + sy] < level; f |= (c2 != in0) ? 2 : 0; }
  if (hz) { c4 = vol[p + 1] < level; f |= (c4 != in0) ? 4 : 0; }
  if (hx && hy && hz) {   // corner c at offset (c & 1, (c >> 1) & 1, (c >> 2) & 1)
    c = (in0 ? 1 : 0) | (c1 ? 2 : 0) | (c2 ? 4 : 0) | (vol[p + sx + sy] < level ? 8 : 0) | (c4 ? 16 : 0) |
        (vol[p + sx + 1] < level ? 32 : 0) | (vol[p + sy + 1] < level ? 64 : 0) | (vol[p + sx + sy + 1] < level ? 128 : 0);
  }
}

__global__ void __launch_bounds__(kMcBlock) mc_count_kernel(const float* __restrict__ vol, int Nx, int Ny, int Nz, float level,
                                                            const uint8_t* __restrict__ n_tris, uint8_t* __restrict__ flags,
                                                            uint8_t* __restrict__ cases, uint16_t* __restrict__ vlocal,
                                                            int32_t* __restrict__ block_v, int32_t* __restrict__ block_f) {
  const int64_t p = blockIdx.x * (int64_t)kMcBlock + threadIdx.x;
  const int64_t n = (int64_t)Nx * Ny * Nz;
  uint32_t f, c;
  classify(vol, p, n, Nx, Ny, Nz, level, f, c);
  int tot_v, tot_f;
  const int pre_v = block_scan_excl(popc3(f), &tot_v);
  block_scan_excl((int)n_tris[c], &tot_f);
  if (p < n) {
    flags[p] = (uint8_t)f;
    cases[p] = (uint8_t)c;
    vlocal[p] = (uint16_t)pre_v;
  }
  if (threadIdx.x == 0) { block_v[blockIdx.x] = tot_v; block_f[blockIdx.x] = tot_f; }
}

// the last block's own sums, saved before the in-place scans overwrite them
__global__ void mc_save_last_kernel(const int32_t* __restrict__ block_v, const int32_t* __restrict__ block_f, int64_t nb,
                                    int64_t* __restrict__ totals) {
  totals[0] = block_v[nb - 1];
  totals[1] = block_f[nb - 1];
}
__global__ void mc_finish_totals_kernel(const int32_t* __restrict__ block_v, const int32_t* __restrict__ block_f, int64_t nb,
                                        int64_t* __restrict__ totals) {
  totals[0] += block_v[nb - 1];
  totals[1] += block_f[nb - 1];
}

__global__ void __launch_bounds__(kMcBlock) mc_generate_kernel(const float* __restrict__ vol, int Nx, int Ny, int Nz, float level,
                                                               float spx, float spy, float spz, int flip,
                                                               const uint8_t* __restrict__ flags, const uint8_t* __restrict__ cases,
                                                               const uint16_t* __restrict__ vlocal, const int32_t* __restrict__ block_v,
                                                               const int32_t* __restrict__ block_f, const uint8_t* __restrict__ n_tris,
                                                               const int8_t* __restrict__ tri_table, float* __restrict__ verts,
                                                               int32_t* __restrict__ faces) {
  const int64_t p = blockIdx.x * (int64_t)kMcBlock + threadIdx.x;
  const int64_t n = (int64_t)Nx * Ny * Nz;
  const bool live = p < n;
  const uint32_t f = live ? flags[p] : 0u, cs = live ? cases[p] : 0u;
  int tot_f;
  const int pre_f = block_scan_excl((int)n_tris[cs], &tot_f);      // triangle offsets: the block scan again (cheaper than an array)
  if (!live) return;
  const uint32_t pu = (uint32_t)p, row = pu / (uint32_t)Nz;
  const int k = (int)(pu - row * (uint32_t)Nz), i = (int)(row / (uint32_t)Ny), j = (int)(row - (uint32_t)i * (uint32_t)Ny);
  const int64_t sx = (int64_t)Ny * Nz, sy = Nz;
  const int64_t stride[3] = {sx, sy, 1};
  if (f) {
    const float v0 = vol[p];
    int32_t vi = block_v[blockIdx.x] + (int32_t)vlocal[p];
    const float base[3] = {(float)i, (float)j, (float)k};
    const float sp[3] = {spx, spy, spz};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (f >> a & 1) {
        const float v1 = vol[p + stride[a]];
        const float t = __fdiv_rn(__fsub_rn(level, v0), __fsub_rn(v1, v0));
        float* out = verts + 3 * (int64_t)vi;
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = __fmul_rn(c == a ? __fadd_rn(base[c], t) : base[c], sp[c]);
        ++vi;
      }
    }
  }
  if (cs != 0 && cs != 255) {
    const int8_t* row = tri_table + 32 * (int)cs;
    int32_t* out = faces + 3 * ((int64_t)block_f[blockIdx.x] + pre_f);
    for (int t = 0; t < 30 && row[t] >= 0; t += 3) {
      int32_t id[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int e = row[t + c], a = e >> 2, b0 = e & 1, b1 = (e >> 1) & 1;
        const int o0 = a == 0 ? 1 : 0, o1 = a == 2 ? 1 : 2;   // the other two axes in increasing order
        const int64_t q = p + b0 * stride[o0] + b1 * stride[o1];
        id[c] = block_v[(uint32_t)q / kMcBlock] + (int32_t)vlocal[q] + __popc(flags[q] & ((1u << a) - 1u));
      }
      out[t] = id[0];
      out[t + 1] = flip ? id[2] : id[1];
      out[t + 2] = flip ? id[1] : id[2];
    }
  }
}

size_t scan_bytes(int64_t nb) {
  size_t a = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, a, (int32_t*)nullptr, (int32_t*)nullptr, (int)nb);
  return a;
}

}  // namespace

extern "C" size_t nr_mc_count_workspace(int32_t Nx, int32_t Ny, int32_t Nz) {
  if (Nx < 1 || Ny < 1 || Nz < 1) return 0;
  return nr_align(scan_bytes(nr_cdiv((int64_t)Nx * Ny * Nz, kMcBlock)));
}

extern "C" int64_t nr_mc_blocks(int32_t Nx, int32_t Ny, int32_t Nz) { return nr_cdiv((int64_t)Nx * Ny * Nz, kMcBlock); }

extern "C" int nr_mc_count(const float* vol, int32_t Nx, int32_t Ny, int32_t Nz, float level, const uint8_t* n_tris,
                           uint8_t* flags, uint8_t* cases, uint16_t* vlocal, int32_t* block_v, int32_t* block_f, int64_t* totals,
                           void* workspace, size_t workspace_bytes, void* stream) {
  NR_CHECK_ARG(Nx >= 2 && Ny >= 2 && Nz >= 2, "nr_mc_count: the volume needs at least 2 samples per axis (%d,%d,%d)", Nx, Ny, Nz);
  const int64_t n = (int64_t)Nx * Ny * Nz;
  NR_CHECK_ARG(n < (1ll << 31), "nr_mc_count: more than 2^31 lattice points");
  NR_CHECK_ARG(vol && n_tris && flags && cases && vlocal && block_v && block_f && totals, "nr_mc_count: null pointer");
  const size_t need = nr_mc_count_workspace(Nx, Ny, Nz);
  NR_CHECK_ARG(workspace && workspace_bytes >= need, "nr_mc_count: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
  const cudaStream_t st = (cudaStream_t)stream;
  const int64_t nb = nr_cdiv(n, kMcBlock);
  mc_count_kernel<<<(unsigned)nb, kMcBlock, 0, st>>>(vol, Nx, Ny, Nz, level, n_tris, flags, cases, vlocal, block_v, block_f);
  NR_CHECK_LAUNCH("mc_count_kernel");
  mc_save_last_kernel<<<1, 1, 0, st>>>(block_v, block_f, nb, totals);
  NR_CHECK_LAUNCH("mc_save_last_kernel");
  size_t bytes = workspace_bytes;
  NR_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(workspace, bytes, block_v, block_v, (int)nb, st));
  bytes = workspace_bytes;
  NR_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(workspace, bytes, block_f, block_f, (int)nb, st));
  mc_finish_totals_kernel<<<1, 1, 0, st>>>(block_v, block_f, nb, totals);
  NR_CHECK_LAUNCH("mc_finish_totals_kernel");
  return NR_OK;
}

extern "C" int nr_mc_generate(const float* vol, int32_t Nx, int32_t Ny, int32_t Nz, float level, float spacing_x, float spacing_y,
                              float spacing_z, int32_t ascent, const uint8_t* flags, const uint8_t* cases, const uint16_t* vlocal,
                              const int32_t* block_v, const int32_t* block_f, const uint8_t* n_tris, const int8_t* tri_table,
                              float* verts, int32_t* faces, void* stream) {
  NR_CHECK_ARG(Nx >= 2 && Ny >= 2 && Nz >= 2, "nr_mc_generate: bad sizes");
  NR_CHECK_ARG(vol && flags && cases && vlocal && block_v && block_f && n_tris && tri_table, "nr_mc_generate: null pointer");
  const int64_t n = (int64_t)Nx * Ny * Nz;
  mc_generate_kernel<<<(unsigned)nr_cdiv(n, kMcBlock), kMcBlock, 0, (cudaStream_t)stream>>>(
      vol, Nx, Ny, Nz, level, spacing_x, spacing_y, spacing_z, ascent, flags, cases, vlocal, block_v, block_f, n_tris, tri_table, verts,
      faces);
  NR_CHECK_LAUNCH("mc_generate_kernel");
  return NR_OK;
}
