// Iso-surface extraction on the device (the reference calls skimage.measure.marching_cubes on the
// host, utils/mesh_util.py:33-35, after copying the N^3 grid over PCIe).  Indexed mesh, every edge crossing one shared vertex.
//
// The case table comes from neurecon_b200/mc_tables.py (generated from its rule: a corner is inside when value < level,
// ambiguous faces cut off each inside corner on its own -- decided per face, so neighbouring cells agree and the mesh is
// closed).  Two passes over the lattice, one thread per lattice point p = (i, j, k), k fastest:
//   count:     flags[p]  = which of the three edges p owns (+x, +y, +z) carry a vertex;  cases[p] = the 8-bit case of the
//              cell whose origin is p (0 on the upper boundary);  then two exclusive scans (cub) give every point the index
//              of its first vertex and every cell the index of its first triangle, and the totals;
//   generate:  p writes its vertices (linear interpolation, un-fused fp32 so that a numpy restatement gets the same bits)
//              and its cell's triangles, looking up the vertex index of edge e = (axis a, owner point q) as
//              vbase[q] + popcount(flags[q] & ((1 << a) - 1)).
// Output order is therefore canonical: vertices by (owner point, axis), triangles by (cell, table order) -- a CPU restatement
// (oracle/mesh.py) reproduces vertices and faces bit for bit.  HBM-bound: the volume is read twice (the 8-corner stencil
// hits L1/L2), 2 + 8 bytes of bookkeeping per lattice point.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "common.cuh"

namespace {

struct PopcLow3 {
  __host__ __device__ __forceinline__ int operator()(const uint8_t& f) const { return (f & 1) + ((f >> 1) & 1) + ((f >> 2) & 1); }
};
struct TriCount {
  const uint8_t* n_tris;
  __device__ __forceinline__ int operator()(const uint8_t& c) const { return n_tris[c]; }
};

__global__ void mc_count_kernel(const float* __restrict__ vol, int Nx, int Ny, int Nz, float level, uint8_t* __restrict__ flags,
                                uint8_t* __restrict__ cases) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)Nx * Ny * Nz;
  if (p >= n) return;
  const int k = (int)(p % Nz), j = (int)((p / Nz) % Ny), i = (int)(p / ((int64_t)Nz * Ny));
  const int64_t sx = (int64_t)Ny * Nz, sy = Nz;
  const bool hx = i + 1 < Nx, hy = j + 1 < Ny, hz = k + 1 < Nz;
  const bool in0 = vol[p] < level;
  uint8_t f = 0;
  bool c1 = false, c2 = false, c4 = false;
  if (hx) { c1 = vol[p + sx] < level; f |= (c1 != in0) ? 1 : 0; }
  if (hy) { c2 = vol[p + sy] < level; f |= (c2 != in0) ? 2 : 0; }
  if (hz) { c4 = vol[p + 1] < level; f |= (c4 != in0) ? 4 : 0; }
  flags[p] = f;
  uint8_t c = 0;
  if (hx && hy && hz) {   // corner c at offset (c & 1, (c >> 1) & 1, (c >> 2) & 1)
    c = (in0 ? 1 : 0) | (c1 ? 2 : 0) | (c2 ? 4 : 0) | (vol[p + sx + sy] < level ? 8 : 0) | (c4 ? 16 : 0) |
        (vol[p + sx + 1] < level ? 32 : 0) | (vol[p + sy + 1] < level ? 64 : 0) | (vol[p + sx + sy + 1] < level ? 128 : 0);
  }
  cases[p] = c;
}

__global__ void mc_totals_kernel(const uint8_t* __restrict__ flags, const uint8_t* __restrict__ cases,
                                 const uint8_t* __restrict__ n_tris, const int32_t* __restrict__ vbase,
                                 const int32_t* __restrict__ fbase, int64_t n, int64_t* __restrict__ totals) {
  totals[0] = (int64_t)vbase[n - 1] + PopcLow3()(flags[n - 1]);
  totals[1] = (int64_t)fbase[n - 1] + n_tris[cases[n - 1]];
}

__global__ void mc_generate_kernel(const float* __restrict__ vol, int Nx, int Ny, int Nz, float level, float spx, float spy,
                                   float spz, int flip, const uint8_t* __restrict__ flags, const uint8_t* __restrict__ cases,
                                   const int32_t* __restrict__ vbase, const int32_t* __restrict__ fbase,
                                   const int8_t* __restrict__ tri_table, float* __restrict__ verts, int32_t* __restrict__ faces) {
  const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t n = (int64_t)Nx * Ny * Nz;
  if (p >= n) return;
  const int k = (int)(p % Nz), j = (int)((p / Nz) % Ny), i = (int)(p / ((int64_t)Nz * Ny));
  const int64_t sx = (int64_t)Ny * Nz, sy = Nz;
  const int64_t stride[3] = {sx, sy, 1};
  const uint8_t f = flags[p];
  if (f) {
    const float v0 = vol[p];
    int32_t vi = vbase[p];
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
  const uint8_t cs = cases[p];
  if (cs != 0 && cs != 255) {
    const int8_t* row = tri_table + 32 * (int)cs;
    int32_t* out = faces + 3 * (int64_t)fbase[p];
    for (int t = 0; t < 30 && row[t] >= 0; t += 3) {
      int32_t id[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int e = row[t + c], a = e >> 2, b0 = e & 1, b1 = (e >> 1) & 1;
        // the other two axes in increasing order
        const int o0 = a == 0 ? 1 : 0, o1 = a == 2 ? 1 : 2;
        const int64_t q = p + b0 * stride[o0] + b1 * stride[o1];
        id[c] = vbase[q] + __popc(flags[q] & ((1u << a) - 1u));
      }
      out[t] = id[0];
      out[t + 1] = flip ? id[2] : id[1];
      out[t + 2] = flip ? id[1] : id[2];
    }
  }
}

size_t scan_bytes(int64_t n) {
  size_t a = 0, b = 0;
  cub::TransformInputIterator<int, PopcLow3, const uint8_t*> it0(nullptr, PopcLow3());
  cub::DeviceScan::ExclusiveSum(nullptr, a, it0, (int32_t*)nullptr, (int)n);
  cub::TransformInputIterator<int, TriCount, const uint8_t*> it1(nullptr, TriCount{nullptr});
  cub::DeviceScan::ExclusiveSum(nullptr, b, it1, (int32_t*)nullptr, (int)n);
  return a > b ? a : b;
}

}  // namespace

extern "C" size_t nr_mc_count_workspace(int32_t Nx, int32_t Ny, int32_t Nz) {
  if (Nx < 1 || Ny < 1 || Nz < 1) return 0;
  return nr_align(scan_bytes((int64_t)Nx * Ny * Nz));
}

extern "C" int nr_mc_count(const float* vol, int32_t Nx, int32_t Ny, int32_t Nz, float level, const uint8_t* n_tris,
                           uint8_t* flags, uint8_t* cases, int32_t* vbase, int32_t* fbase, int64_t* totals, void* workspace,
                           size_t workspace_bytes, void* stream) {
  NR_CHECK_ARG(Nx >= 2 && Ny >= 2 && Nz >= 2, "nr_mc_count: the volume needs at least 2 samples per axis (%d,%d,%d)", Nx, Ny, Nz);
  const int64_t n = (int64_t)Nx * Ny * Nz;
  NR_CHECK_ARG(n < (1ll << 31), "nr_mc_count: more than 2^31 lattice points");
  NR_CHECK_ARG(vol && n_tris && flags && cases && vbase && fbase && totals, "nr_mc_count: null pointer");
  const size_t need = nr_mc_count_workspace(Nx, Ny, Nz);
  NR_CHECK_ARG(workspace && workspace_bytes >= need, "nr_mc_count: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
  const cudaStream_t st = (cudaStream_t)stream;
  mc_count_kernel<<<(unsigned)nr_cdiv(n, 256), 256, 0, st>>>(vol, Nx, Ny, Nz, level, flags, cases);
  NR_CHECK_LAUNCH("mc_count_kernel");
  size_t bytes = workspace_bytes;
  cub::TransformInputIterator<int, PopcLow3, const uint8_t*> it0(flags, PopcLow3());
  NR_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(workspace, bytes, it0, vbase, (int)n, st));
  bytes = workspace_bytes;
  cub::TransformInputIterator<int, TriCount, const uint8_t*> it1(cases, TriCount{n_tris});
  NR_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(workspace, bytes, it1, fbase, (int)n, st));
  mc_totals_kernel<<<1, 1, 0, st>>>(flags, cases, n_tris, vbase, fbase, n, totals);
  NR_CHECK_LAUNCH("mc_totals_kernel");
  return NR_OK;
}

extern "C" int nr_mc_generate(const float* vol, int32_t Nx, int32_t Ny, int32_t Nz, float level, float spacing_x, float spacing_y,
                              float spacing_z, int32_t ascent, const uint8_t* flags, const uint8_t* cases, const int32_t* vbase,
                              const int32_t* fbase, const int8_t* tri_table, float* verts, int32_t* faces, void* stream) {
  NR_CHECK_ARG(Nx >= 2 && Ny >= 2 && Nz >= 2, "nr_mc_generate: bad sizes");
  NR_CHECK_ARG(vol && flags && cases && vbase && fbase && tri_table, "nr_mc_generate: null pointer");
  const int64_t n = (int64_t)Nx * Ny * Nz;
  mc_generate_kernel<<<(unsigned)nr_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
      vol, Nx, Ny, Nz, level, spacing_x, spacing_y, spacing_z, ascent, flags, cases, vbase, fbase, tri_table, verts, faces);
  NR_CHECK_LAUNCH("mc_generate_kernel");
  return NR_OK;
}
