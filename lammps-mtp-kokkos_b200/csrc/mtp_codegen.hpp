// Per-potential code generation for the contraction program (pair_mtp.cpp:196-233).
//
// The alpha_index_times table is a sparse polynomial program  m[a3] += mult * m[a0] * m[a1]  with a fixed
// structure per potential.  Instead of interpreting term descriptors at run time, the library emits one
// straight-line CUDA kernel per potential STRUCTURE (tables only -- every coefficient stays run-time data) and
// compiles it with NVRTC when the potential is loaded (cubins are cached on disk, keyed by a hash of the tables
// and of the generator parameters):
//   * CTA = NA atoms (lane = atom), W warps.  The program is cut into barrier-separated stages by dependency;
//     the tasks of a stage (forward evaluation of a target node / reverse-mode gather of a source node) are dealt
//     to the warps in contiguous runs of the node order, so the components of one tensor contraction stay in one
//     warp and share their operands;
//   * operand rows ([node][atom] FP64 in shared memory, only for nodes that feed a product) are read through a
//     software-managed register cache (Belady eviction, decided here, at generation time): offsets are immediates,
//     there are no descriptors, no predication, no per-term control flow;
//   * the reverse pass is in gather form (each adjoint is owned by one warp, no atomics, fixed summation order);
//     several adjoints of a warp are accumulated at once so that the consumer's adjoint g[a3] is read once for all of
//     them; adjoints of basic moments go straight from registers to global memory.
// The emitted text is plain C with a handful of macros (LD, ST, FMA, ...) so that the same file compiles for the
// host: tests/ run it lane by lane against the sequential program without a GPU.
#pragma once

#include "mtp_potential.hpp"

#include <string>
#include <vector>

namespace mtpb200 {

struct P4Params {
  int na = 32;          // atoms per CTA: 8, 16, 32 (one per lane) or 64 (two per lane)
  int warps = 4;        // warps per CTA
  int cache = 56;       // operand values the register cache of a warp may hold
  int acc_max = 16;     // adjoints a warp accumulates at once in the reverse pass
  int fn_cost = 4000;   // term steps per emitted function (bounds ptxas time and register pressure)
  size_t smem_budget = 0;    // > 0: dynamic shared memory a CTA may use; a program that needs more is cut into rounds
  int groups = 1;       // atom groups per CTA: every emitted function is executed `groups` times in a row, once per group
                        // of `na` atoms, so that an instruction fetched once serves groups * na atoms (the code is
                        // straight-line and otherwise never reused; small functions stay in the instruction cache)
  int sparse = 0;       // 1: a round keeps only the basic moments it reads in shared memory (implied by groups > 1)
  int spatial = 0;      // 1 (with groups > 1): the groups run SIDE BY SIDE on `groups` sets of `warps` warps instead of one
                        // after the other: warp w of every set executes the same function at the same time (the stage
                        // barriers keep them together), so one instruction stream feeds `groups` warps
  int rpar = 0;         // 1 (sparse, more than one round): ROUNDS IN PARALLEL -- blockIdx.y selects the round a CTA
                        // evaluates; every round adds its adjoint shares to a zeroed gb with RED.ADD.  For small systems
                        // (mtp/small/kk): the critical path of a chunk is one round instead of the whole program
};

struct P4Info {
  int rows = 0;             // shared-memory rows (m rows of operand nodes + g rows of non-basic operand nodes)
  int m_rows = 0, g_rows = 0;
  int stages = 0;
  int rounds = 1;           // > 1: the basis functions were dealt to several rounds that reuse the shared-memory rows
  size_t smem_bytes = 0;    // dynamic shared memory of the kernel
  long long terms = 0;      // multiply-add term steps per atom (forward T + reverse 2T, squares merged)
  long long loads = 0;      // shared-memory row loads per chunk summed over warps (after the register cache)
  long long stores = 0;
  long long crit_terms = 0; // sum over stages of the most loaded warp's term steps
  int threads = 0;
  int groups = 1;           // atom groups per CTA (atoms per CTA iteration = groups * na)
  int rpar_rounds = 0;      // > 0: rounds run in parallel, launch with gridDim.y = rpar_rounds on a zeroed gb
  unsigned long long hash = 0;
};

// shared-memory bytes the kernel needs for (p, prm) without generating it; 0 = structure not supported
// (rounds_out, optional: number of rounds the program is cut into)
size_t p4_smem_bytes(const Potential &p, const P4Params &prm, int *rounds_out = nullptr);

// Emits the kernel source for potential p.  slot_of_k maps basic moment k to its row of mb / gb (NULL = identity),
// nslots = rows of mb / gb.  Returns false (and a reason) when the table's structure is outside what the generator
// handles (a basic moment that is also a product target).
bool p4_generate(const Potential &p, const P4Params &prm, const short *slot_of_k, int nslots, std::string &source,
                 P4Info &info, std::string &why_not);

// Kernel argument block (must match the struct of the same name in the emitted source)
struct P4Args {
  const double *mb;
  double *gb;
  long long ld;
  int inum, first_ii;
  const int *ilist;
  const void *xt;          // AtomRec[nall]
  const double *lin, *species;
  int S;
  int eflag_global, eflag_atom, grade;
  double *eatom;
  double *cand_rows;
  long long cand_ld;
  int cand_col0;           // first column of the linear block of the candidate vector
  double *partials;        // [gridDim.y * gridDim.x][8]
  double *esite;           // rounds in parallel, eflag_atom: [round][ld] site-energy shares (summed by esite_sum_kernel)
};

}    // namespace mtpb200
