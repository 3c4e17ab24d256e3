// Host-side potential model: MLIP-3 .almtp parser and the "compiled" contraction program.
// Behavioural spec: PairMTP::read_file (pair_mtp.cpp:335-655), RadialMTPBasis::ReadBasisProperties
// (mtp_radial_basis.cpp:59-102), PairMTPExtrapolation::read_file (pair_mtp_extrapolation.cpp:528-619).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace mtpb200 {

struct Potential {
  int species_count = 0;
  int radial_func_count = 0;      // R
  int radial_basis_size = 0;      // B
  int alpha_moment_count = 0;     // M
  int alpha_index_basic_count = 0;
  int alpha_index_times_count = 0;
  int alpha_scalar_count = 0;
  int max_alpha_index_basic = 0;  // P
  double min_cutoff = 0, max_cutoff = 0, scaling = 1.0;
  std::string potential_name = "Untitled", potential_tag;
  std::vector<double> radial_basis_coeffs;   // [S][S][R][B]
  std::vector<int> alpha_index_basic;        // [K][4]
  std::vector<int> alpha_index_times;        // [T][4]
  std::vector<int> alpha_moment_mapping;     // [A]
  std::vector<double> species_coeffs, linear_coeffs;
  std::vector<int> setflag;                  // [(S+1)^2], pair_mtp.cpp:455
  // selection state
  bool has_selection_state = false;
  int configuration_mode = 0;
  int coeff_count = 0;                       // Q
  std::vector<double> active_set, inverse_active_set;
  std::string log;                           // what the reference prints through utils::logmesg

  int radial_coeff_count() const { return species_count * species_count * radial_func_count * radial_basis_size; }
};

// Throws std::runtime_error with the reference's message on any grammar violation.
void parse_almtp(const std::string &path, bool want_selection_state, Potential &out);
// Validates sizes / index ranges of tables handed in through mtp_create(); derives P and Q.
void finalize_tables(Potential &p);

// ---- contraction program ------------------------------------------------------------------------
// alpha_index_times is a sequential program  m[a3] += mult*m[a0]*m[a1]  (pair_mtp.cpp:196-201) whose
// reverse sweep gives dE/dm (pair_mtp.cpp:221-233).  At load it is re-expressed as atomic-free gather
// lists grouped by dependency level (the reference's "waves", pair_mtps_kokkos.cpp:179-200, for any
// depth): forward = per TARGET node the list of (a0,a1,mult) in file order; reverse = per SOURCE node
// the list of (a3,other,mult) in reverse file order.  Nodes of one level are packed 32 to a group and
// their lists are stored transposed (ELL, slot-major) so that lane l of a warp owns node l of the
// group and reads term t at  terms[(term_base + t) * 32 + l]  (coalesced).
struct ProgramTerm {
  uint16_t a;     // forward: a0      reverse: a3
  uint16_t b;     // forward: a1      reverse: the other factor
  float mult;     // small integer multiplicity (exact in fp32)
};

struct ProgramPass {
  std::vector<int> level_group_begin;   // [nlevels + 1] group index ranges per level (execution order)
  std::vector<int> group_term_base;     // [ngroups] first slot row of the group
  std::vector<int> group_max_terms;     // [ngroups]
  std::vector<int> node;                // [ngroups * 32] node id or -1
  std::vector<int> nterms;              // [ngroups * 32]
  std::vector<ProgramTerm> terms;       // [slot_rows * 32]
  int ngroups() const { return (int) group_term_base.size(); }
};

// The same program in the form used by the CTA-per-atom-chunk kernel: the atoms of a chunk sit in the lanes
// (moments stored [node][atom]), so every lane of a warp executes the SAME term and operand reads are
// conflict-free rows; nodes of one level are dealt to the warps longest-list-first.
struct ChunkPass {
  std::vector<int> level_begin;     // [nlevels + 1] node-slot ranges, execution order
  std::vector<int> node;            // [nslots]
  std::vector<int> term_begin;      // [nslots + 1]
  std::vector<uint32_t> term_idx;   // forward: a0 | a1 << 16;  reverse: a3 | other << 16, a3 == 0xFFFF: constant term
  std::vector<double> term_coef;    // forward: mult;  reverse: mult, or mult * ginit[a3] for a constant term
  std::vector<double> init;         // [nslots] reverse: ginit[node]; forward: unused (0)
};

// Flat predicated term streams for the CTA-per-atom-chunk program kernel (lane = atom).  The nodes of one
// dependency level are dealt to VW "virtual warps" (longest list first onto the least loaded one); each virtual
// warp then owns ONE flat stream of uniform terms per level
//        acc += coef * A[a] * B[b];   if (store) { dst[node] = acc; acc = 0; }
// (forward: A = B = dst = moments; reverse: A = dst = adjoints, B = moments).  Row M of both tables holds 1.0, so
// "acc = m[node]", "acc = ginit[node]" and terms whose adjoint is a constant are ordinary terms with a = M and/or
// b = M.  Streams are padded with no-ops (a = b = M, coef = 0) to a multiple of FLAT_UNROLL, so the kernel's inner
// loop is branch-free and all descriptor / operand loads of an unrolled group are independent.
constexpr int FLAT_UNROLL = 4;
struct FlatTerm {
  uint32_t a_off, b_off;    // byte offsets of the operand rows: row * NA * 8
  double coef;
};
static_assert(sizeof(FlatTerm) == 16, "FlatTerm packing");
struct FlatPass {
  int vw = 0, nlevels = 0, na = 0;
  std::vector<int> stream_begin;    // [nlevels * vw + 1]
  std::vector<FlatTerm> terms;
  std::vector<uint32_t> st;         // per term: byte offset of the destination row | 1 if the term ends its node
};

// Streams of the 4-atoms-per-lane program kernel (mtp_program_v3.cuh).  A "virtual warp" (NA/4 lanes) evaluates one
// node at a time; the VPW = 128/NA virtual warps of a physical warp work on a GROUP of VPW nodes of the same dependency
// level whose term lists were padded to one common length, so the end of a node is a warp-uniform event (no
// predication, no flags); a list longer than G3_SPLIT_ABOVE terms forms a group of its own, its terms dealt to all
// the virtual warps and the partial sums combined by shuffles (head.rows bit 31).  Per (level, physical warp) the stream is a flat sequence of term rows ([VPW] descriptors of
// 16 bytes, one per virtual warp) and a sequence of group heads ([VPW] x {destination, rows of the group, initial
// value}); every stream is a whole number of 4-row trips (a dummy group writing to the scratch row M+1 is appended
// otherwise) and the arrays end with padding rows / heads because the kernel prefetches ahead.
// Row M of both tables holds 1.0 (no-op terms: a = b = M, coef = 0), row M+1 is scratch.
struct G3Term {
  uint32_t a_off, b_off;    // byte offsets of the operand rows: row * NA * 8
  double coef;
};
struct G3Head {
  uint32_t dst_off, rows;
  double init;
};
static_assert(sizeof(G3Term) == 16 && sizeof(G3Head) == 16, "G3 packing");
constexpr int G3_WARPS = 8;        // physical warps per CTA of the kernel
constexpr int G3_SPLIT_ABOVE = 8;
constexpr int G3_PAD_ROWS = 8;     // rows readable past the end of the term array
struct Flat3Pass {
  int vpw = 0, nlevels = 0, na = 0;
  std::vector<int> row_begin;      // [nlevels * G3_WARPS + 1] term-row ranges
  std::vector<int> group_begin;    // [nlevels * G3_WARPS + 1] group ranges
  std::vector<G3Term> terms;       // [rows + G3_PAD_ROWS][vpw]
  std::vector<G3Head> heads;       // [groups + 2][vpw]
};

struct Program {
  int depth = 0;                  // number of waves
  ChunkPass cfwd, crev;
  FlatPass ffwd[2], frev[2];      // for the two atoms-per-CTA shapes of the program kernel (vw = 16 * 32 / NA)
  int f3_na = 0;                  // atoms per CTA of the v3 streams (32 / 16), 0 = tables not expressible
  Flat3Pass f3fwd, f3rev;
  std::vector<int> level;         // [M]
  // rows of the adjoint table of the chunk-per-CTA program kernel: only basic moments and nodes that feed a product
  // ever hold a variable adjoint (the others are constants folded into coefficients), so the table has
  // adjoint_rows = K + (non-basic sources) rows instead of M: grow[node] = its row, or -1
  int adjoint_rows = 0;
  std::vector<int> grow;
  ProgramPass fwd, rev;
  std::vector<double> ginit;      // [M]: dE/dm seed, g[map[s]] = xi_s (pair_mtp.cpp:217-218)
};

// K + number of non-basic nodes that are a factor of some product (Program::adjoint_rows, without compiling)
int count_adjoint_rows(const Potential &p);
// Throws std::runtime_error if the table is not a topologically ordered program.
void compile_program(const Potential &p, Program &out, int na_large = 32, int na_small = 8, int na_v3 = 0);

// Host interpreter of the grouped streams (Flat3Pass) for one atom with pseudo-random basic moments: runs the forward
// and reverse passes exactly as the kernel schedules them (stores of a level become visible at its barrier) and returns
// the largest relative deviation of all moments and basic-moment adjoints from the sequential program
// (pair_mtp.cpp:196-233).  Test infrastructure for the stream packer; no device involved.
double check_grouped_streams(const Potential &p, const Program &prog);

}    // namespace mtpb200
