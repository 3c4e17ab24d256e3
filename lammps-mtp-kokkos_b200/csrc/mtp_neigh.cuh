// Device build of the FULL neighbor list the pair style requests (pair_mtp.cpp:318 NeighConst::REQ_FULL; consumed as
// list->inum / ilist / numneigh / firstneigh, pair_mtp.cpp:77-85, or as the 2-D d_neighbors view of the Kokkos styles,
// pair_mtp_kokkos.cpp:236-240).  SURVEY.md section 8(f) row 1: the step before the hot path, upstream NeighborKokkos
// in a real LAMMPS run.
//
//   for every owned atom i < nlocal: all atoms j != i (owned or ghost) with |x_j - x_i|^2 <= cutneigh^2
//
// (ghosts carry their own shifted coordinates, as in LAMMPS, so there is no minimum-image arithmetic).  rsq is formed
// exactly like the CPU builds do -- three separately rounded squares added left to right, no FMA -- so the SET of
// neighbors of every atom is bit-identical to the host list; the order within a row is stencil order (deterministic).
//
// Pipeline: bounding box (two-stage reduction) -> bin index per atom -> stable radix sort of (bin, atom) (CUB) ->
// cell starts by binary search -> positions gathered in sorted order (32-byte records) -> warp per owned atom: the
// 3 x-adjacent bins of each of the 9 (dz, dy) stencil rows are ONE contiguous range of the sorted array, lanes
// stride over it (coalesced 32-byte records), ballot compaction into the row-major table neighbors[i][width].
#pragma once

#include "mtp_device.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace mtpb200 {

struct NeighGrid {
  double lo[3], inv[3];
  int n[3];
};

__global__ void neigh_bounds_kernel(int nall, const double *__restrict__ x, double *__restrict__ part)
{
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nall; i += gridDim.x * blockDim.x)
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const double v = x[3 * (size_t) i + a];
      lo[a] = fmin(lo[a], v);
      hi[a] = fmax(hi[a], v);
    }
  __shared__ double s[8][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int a = 0; a < 3; a++)
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fmin(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmax(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
  if (lane == 0)
    for (int a = 0; a < 3; a++) {
      s[warp][a] = lo[a];
      s[warp][3 + a] = hi[a];
    }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = s[0][threadIdx.x];
    for (int w = 1; w < (int) (blockDim.x >> 5); w++) v = threadIdx.x < 3 ? fmin(v, s[w][threadIdx.x]) : fmax(v, s[w][threadIdx.x]);
    part[(size_t) blockIdx.x * 6 + threadIdx.x] = v;
  }
}

__global__ void neigh_bounds_final_kernel(int nblocks, const double *__restrict__ part, double *__restrict__ out)
{
  if (threadIdx.x < 6) {
    double v = part[threadIdx.x];
    for (int b = 1; b < nblocks; b++)
      v = threadIdx.x < 3 ? fmin(v, part[(size_t) b * 6 + threadIdx.x]) : fmax(v, part[(size_t) b * 6 + threadIdx.x]);
    out[threadIdx.x] = v;
  }
}

__device__ __forceinline__ void neigh_cell_of(const NeighGrid &g, double px, double py, double pz, int c[3])
{
  const double p[3] = {px, py, pz};
#pragma unroll
  for (int a = 0; a < 3; a++) {
    int v = (int) floor((p[a] - g.lo[a]) * g.inv[a]);
    c[a] = min(max(v, 0), g.n[a] - 1);
  }
}

__global__ void neigh_bin_kernel(int nall, const double *__restrict__ x, NeighGrid g, int *__restrict__ keys,
                                 int *__restrict__ idx)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nall) return;
  int c[3];
  neigh_cell_of(g, x[3 * (size_t) i], x[3 * (size_t) i + 1], x[3 * (size_t) i + 2], c);
  keys[i] = (c[2] * g.n[1] + c[1]) * g.n[0] + c[0];
  idx[i] = i;
}

// cell_start[c] = first sorted position whose bin is >= c (c = 0 .. ncell)
__global__ void neigh_cellstart_kernel(int nall, const int *__restrict__ skeys, int ncell, int *__restrict__ cell_start)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncell) return;
  int lo = 0, hi = nall;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (skeys[mid] < c) lo = mid + 1;
    else
      hi = mid;
  }
  cell_start[c] = lo;
}

__global__ void neigh_gather_sorted_kernel(int nall, const double *__restrict__ x, const int *__restrict__ sidx,
                                           AtomRec *__restrict__ xs)
{
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nall) return;
  const int j = sidx[s];
  AtomRec r;
  r.x = x[3 * (size_t) j];
  r.y = x[3 * (size_t) j + 1];
  r.z = x[3 * (size_t) j + 2];
  r.t = j;
  xs[s] = r;
}

// warp per owned atom; neighbors row-major [nlocal][width]; numneigh[i] is the TRUE count (may exceed width: the
// caller compares the maximum with width and rebuilds with a wider table, like LAMMPS-KOKKOS's resize-and-retry)
__global__ void __launch_bounds__(256)
neigh_build_kernel(int nlocal, const double *__restrict__ x, NeighGrid g, const int *__restrict__ cell_start,
                   const AtomRec *__restrict__ xs, double cutsq, int width, int *__restrict__ numneigh,
                   int *__restrict__ neighbors, int *__restrict__ maxnn)
{
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  int wmax = 0;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < nlocal; i += gridDim.x * wpb) {
    const double xi = x[3 * (size_t) i], yi = x[3 * (size_t) i + 1], zi = x[3 * (size_t) i + 2];
    int c[3];
    neigh_cell_of(g, xi, yi, zi, c);
    const int cx0 = max(c[0] - 1, 0), cx1 = min(c[0] + 1, g.n[0] - 1);
    int *row = neighbors + (size_t) i * width;
    int n = 0;
    for (int dz = -1; dz <= 1; dz++) {
      const int cz = c[2] + dz;
      if (cz < 0 || cz >= g.n[2]) continue;
      for (int dy = -1; dy <= 1; dy++) {
        const int cy = c[1] + dy;
        if (cy < 0 || cy >= g.n[1]) continue;
        const int base = (cz * g.n[1] + cy) * g.n[0];
        const int s0 = cell_start[base + cx0], s1 = cell_start[base + cx1 + 1];
        for (int sb = s0; sb < s1; sb += 32) {
          const int s = sb + lane;
          bool in = false;
          int j = -1;
          if (s < s1) {
            const AtomRec r = xs[s];
            j = (int) r.t;
            const double d0 = r.x - xi, d1 = r.y - yi, d2 = r.z - zi;
            const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), __dmul_rn(d2, d2));
            in = j != i && rsq <= cutsq;
          }
          const unsigned m = __ballot_sync(0xffffffffu, in);
          if (in) {
            const int pos = n + __popc(m & ((1u << lane) - 1u));
            if (pos < width) row[pos] = j;
          }
          n += __popc(m);
        }
      }
    }
    if (lane == 0) numneigh[i] = n;
    wmax = max(wmax, n);
  }
  if (lane == 0 && wmax > 0) atomicMax(maxnn, wmax);
}

}    // namespace mtpb200
