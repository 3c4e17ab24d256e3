// Device-side view of a loaded potential and of one compute call (plain structs passed by value).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace mtpb200 {

struct DevPass {
  const int *level_group_begin;    // [nlevels + 1]
  const int *group_term_base;      // [ngroups]
  const int *node;                 // [ngroups * 32]
  const int *nterms;               // [ngroups * 32]
  const uint2 *terms;              // [slot_rows * 32] {a | b << 16, float bits of mult}
  int nlevels;
};

struct DevChunkPass {
  const int *level_begin;      // [nlevels + 1]
  const int *node;             // [nslots]
  const int *term_begin;       // [nslots + 1]
  const uint32_t *term_idx;    // packed operands
  const double *term_coef;
  const double *init;          // [nslots]
  int nlevels;
};

// flat predicated term streams (mtp_potential.hpp: FlatPass)
struct DevFlatPass {
  const int *stream_begin;     // [nlevels * vw + 1]
  const uint4 *terms;          // {a_off, b_off, coef lo, coef hi}
  const unsigned *st;          // destination row byte offset | store flag
  int nlevels, vw, nterms;
};

struct DevPotential {
  int S, R, B, K, M, A, P, Q;
  int Mg;                    // rows of the adjoint table of the chunk-per-CTA program kernel (Program::adjoint_rows)
  double rmin, rmax, scaling, cutsq;
  const double *radial;      // [S][S][R][B]
  const uint32_t *basic;     // [K] mu | ax << 8 | ay << 16 | az << 24
  const double *species;     // [S]
  const double *lin;         // [A]
  const int *map;            // [A]
  const double *ginit;       // [M]
  DevPass fwd, rev;
  DevChunkPass cfwd, crev;
  DevFlatPass ffwd[2], frev[2];    // [0]: 32 atoms per CTA, [1]: 8 atoms per CTA
};

// packed position + species record: one 32-byte sector per gathered neighbor
struct __align__(32) AtomRec {
  double x, y, z;
  long long t;    // type - 1
};

struct SiteArgs {
  int inum, nall;
  const AtomRec *xt;
  const int *ilist;
  const int *numneigh;
  const int *neighbors;
  const long long *neigh_offsets;
  long long stride_i, stride_jj;
  int neighmask;
  int eflag_global, eflag_atom, vflag_any, vflag_atom, want_grade;
  double *f;
  double *eatom;
  double *vatom;
  double *grades;
  unsigned char *within;
  double *partials;       // [gridDim.x][8] per-CTA energy/virial/max-grade
  double *cand_rows;      // grade steps: [chunk rows][Qpad] candidate vectors
  int cand_ld;            // Qpad
  int first_ii;           // chunk offset into ilist (grade steps)
  int prog_shape;         // which flat-stream table set the program kernel uses (0 throughput, 1 latency)
  int prog_dsmem;         // program kernel: term streams staged in shared memory
  int prog_debug;         // timing experiments only (MTP_B200_PROG_DEBUG): bit0 skip forward pass, bit1 skip reverse pass, bit2 skip
                          // energy, bit3 skip the basic-moment fetch, bit4 skip the adjoint store (bits 3-4: 4-atoms-per-lane kernel)
  int prog_prefetch;      // program kernel: next chunk's basic moments prefetched (cp.async) into a staging buffer
  const short *slot_to_k; // rows of mb / gb are canonical slots (v2 pipeline) instead of basic-moment indices; or NULL
  int nslots;             // rows of mb / gb
  int *status;
};

// one 32-byte position record in ONE request (sm_100 256-bit load, LDG.E.ENL2.256): a gathered neighbor costs one
// sector lookup instead of two 16-byte loads of the same sector
__device__ __forceinline__ void ld_atomrec(const AtomRec *p, double &x, double &y, double &z, int &t)
{
  double tt;
  asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(tt) : "l"(p));
  t = (int) __double_as_longlong(tt);
}

// Ampere-style asynchronous copies global -> shared (LDGSTS): no registers, many in flight per thread
__device__ __forceinline__ void cp_async8_zfill(void *smem_dst, const void *gsrc, int src_bytes /*8 or 0*/)
{
  const unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
  const unsigned d = (unsigned) __cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

}    // namespace mtpb200
