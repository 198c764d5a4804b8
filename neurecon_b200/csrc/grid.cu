// Lattice points of the dense SDF-grid query (utils/mesh_util.py:82-100 of the reference), generated on
// the device for an index range so that a 512^3 query never ships coordinates over PCIe.
#include "common.cuh"

namespace {
// The reference builds the lattice in float64 with TRUE division (`(idx / N) % N`), which makes the y and
// x indices fractional (a shear of the intended lattice); `faithful` reproduces that bit for bit, the
// alternative uses the integer lattice the code evidently intended.
__global__ void grid_points_kernel(int64_t i0, int64_t count, int N, double s, int faithful, float* __restrict__ pts) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int64_t i = i0 + k;
  const double n = (double)N, step = s / (double)(N - 1), org = -s / 2.0;
  double zi = (double)(i % N), yi, xi;
  if (faithful) {
    const double a = (double)i / n;
    yi = fmod(a, n);
    xi = fmod(a / n, n);
  } else {
    yi = (double)((i / N) % N);
    xi = (double)((i / ((int64_t)N * N)) % N);
  }
  pts[3 * k] = (float)(xi * step + org);
  pts[3 * k + 1] = (float)(yi * step + org);
  pts[3 * k + 2] = (float)(zi * step + org);
}
}  // namespace

extern "C" int nr_grid_points(int64_t i0, int64_t count, int32_t N, double volume_size, int32_t faithful, float* pts,
                              void* stream) {
  NR_CHECK_ARG(N >= 2 && i0 >= 0 && count >= 0 && i0 + count <= (int64_t)N * N * N, "nr_grid_points: bad range");
  if (count == 0) return NR_OK;
  NR_CHECK_ARG(pts, "nr_grid_points: null pointer");
  grid_points_kernel<<<(unsigned)nr_cdiv(count, 256), 256, 0, (cudaStream_t)stream>>>(i0, count, N, volume_size, faithful, pts);
  NR_CHECK_LAUNCH("grid_points_kernel");
  return NR_OK;
}
