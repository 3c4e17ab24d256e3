// Velocity-Verlet half steps on the device: what upstream LAMMPS's FixNVE::initial_integrate / final_integrate do
// around Pair::compute in the reference's own example deck (`velocity all create ...`, `fix 1 all nve`,
// README.md:148-149).  SURVEY.md section 8(f) row 3: the steps either side of the path, so that an MD step never
// leaves the device.  Arithmetic follows FixNVE line by line (dtfm = dtf / mass[type]; v += dtfm * f; x += dtv * v),
// each product and sum rounded separately like the CPU build (no FMA contraction), so a host replay is bit-exact.
#pragma once

#include <cuda_runtime.h>

namespace mtpb200 {

// x0 (optional): positions at the last neighbor-list build; *moved is set when an atom has moved farther than
// sqrt(trigger_sq) from them (LAMMPS's "half the skin" re-neighboring criterion, Neighbor::check_distance)
__global__ void nve_initial_kernel(int n, double *__restrict__ x, double *__restrict__ v, const double *__restrict__ f,
                                   const int *__restrict__ type, const double *__restrict__ mass, double dtf, double dtv,
                                   const double *__restrict__ x0, double trigger_sq, int *__restrict__ moved)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double dtfm = dtf / mass[type[i]];
  double d2 = 0.0;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const size_t k = 3 * (size_t) i + a;
    const double vn = __dadd_rn(v[k], __dmul_rn(dtfm, f[k]));
    const double xn = __dadd_rn(x[k], __dmul_rn(dtv, vn));
    v[k] = vn;
    x[k] = xn;
    if (x0) {
      const double d = xn - x0[k];
      d2 += d * d;
    }
  }
  if (x0 && d2 > trigger_sq) *moved = 1;
}

__global__ void nve_final_kernel(int n, double *__restrict__ v, const double *__restrict__ f, const int *__restrict__ type,
                                 const double *__restrict__ mass, double dtf)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double dtfm = dtf / mass[type[i]];
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const size_t k = 3 * (size_t) i + a;
    v[k] = __dadd_rn(v[k], __dmul_rn(dtfm, f[k]));
  }
}

}    // namespace mtpb200
