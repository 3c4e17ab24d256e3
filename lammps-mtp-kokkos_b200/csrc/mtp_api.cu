// extern "C" layer of include/mtp_b200.h: handle, uploads, launch configuration, host-buffer path.
#include "../../include/mtp_b200.h"
#include "mtp_kernels.cu"
#include "mtp_potential.hpp"
#include "mtp_neigh.cuh"
#include "mtp_md.cuh"
#include "mtp_p4_runtime.hpp"
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <algorithm>
#include <atomic>
#include <ctime>
#include <cstdlib>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

using namespace mtpb200;

namespace {

thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};

struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define CUDA_CHECK(expr)                                                                           \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      throw CudaError(std::string(#expr) + ": " + cudaGetErrorString(e__));                        \
  } while (0)

template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;
  void ensure(size_t n)
  {
    if (n <= cap) return;
    if (p) cudaFree(p);
    p = nullptr;
    size_t want = n + n / 8 + 16;
    CUDA_CHECK(cudaMalloc((void **) &p, want * sizeof(T)));
    cap = want;
  }
  void upload(const T *h, size_t n, cudaStream_t s)
  {
    ensure(n);
    if (n) CUDA_CHECK(cudaMemcpyAsync(p, h, n * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  template <typename V> void upload(const std::vector<V> &v, cudaStream_t s)
  {
    static_assert(sizeof(V) == sizeof(T), "size mismatch");
    upload(reinterpret_cast<const T *>(v.data()), v.size(), s);
  }
  ~DevBuf()
  {
    if (p) cudaFree(p);
  }
};

struct PassBufs {
  DevBuf<int> lgb, gtb, node, nterms;
  DevBuf<uint2> terms;
  DevPass view(const ProgramPass &pp) const
  {
    DevPass d;
    d.level_group_begin = lgb.p;
    d.group_term_base = gtb.p;
    d.node = node.p;
    d.nterms = nterms.p;
    d.terms = terms.p;
    d.nlevels = (int) pp.level_group_begin.size() - 1;
    return d;
  }
  void upload(const ProgramPass &pp)
  {
    lgb.upload(pp.level_group_begin, 0);
    gtb.upload(pp.group_term_base, 0);
    node.upload(pp.node, 0);
    nterms.upload(pp.nterms, 0);
    static_assert(sizeof(ProgramTerm) == sizeof(uint2), "term packing");
    terms.upload(reinterpret_cast<const uint2 *>(pp.terms.data()), pp.terms.size(), 0);
  }
};

}    // namespace

typedef void (*V2GatherFn)(mtpb200::DevPotential, mtpb200::SiteArgs, mtpb200::PairBuf);
struct ChunkBufs {
  DevBuf<int> level_begin, node, term_begin;
  DevBuf<uint32_t> term_idx;
  DevBuf<double> term_coef, init;
  DevChunkPass upload(const ChunkPass &c)
  {
    level_begin.upload(c.level_begin, 0);
    node.upload(c.node, 0);
    term_begin.upload(c.term_begin, 0);
    term_idx.upload(c.term_idx, 0);
    term_coef.upload(c.term_coef, 0);
    init.upload(c.init, 0);
    DevChunkPass d;
    d.level_begin = level_begin.p;
    d.node = node.p;
    d.term_begin = term_begin.p;
    d.term_idx = term_idx.p;
    d.term_coef = term_coef.p;
    d.init = init.p;
    d.nlevels = (int) c.level_begin.size() - 1;
    return d;
  }
};

struct FlatBufs {
  DevBuf<int> begin;
  DevBuf<uint4> terms;
  DevBuf<unsigned> st;
  DevFlatPass upload(const FlatPass &f)
  {
    begin.upload(f.stream_begin, 0);
    static_assert(sizeof(FlatTerm) == sizeof(uint4), "FlatTerm packing");
    terms.upload(reinterpret_cast<const uint4 *>(f.terms.data()), f.terms.size(), 0);
    st.upload(f.st, 0);
    DevFlatPass d;
    d.stream_begin = begin.p;
    d.terms = terms.p;
    d.st = st.p;
    d.nlevels = f.nlevels;
    d.vw = f.vw;
    d.nterms = (int) f.terms.size();
    return d;
  }
};

static_assert(G3_WARPS == P3_WARPS && G3_PAD_ROWS == P3_PAD_ROWS, "stream packer and program kernel disagree");
struct Flat3Bufs {
  DevBuf<int> rbegin, gbegin;
  DevBuf<uint4> terms, heads;
  DevFlat3Pass upload(const Flat3Pass &f)
  {
    rbegin.upload(f.row_begin, 0);
    gbegin.upload(f.group_begin, 0);
    terms.upload(f.terms, 0);
    heads.upload(f.heads, 0);
    DevFlat3Pass d;
    d.row_begin = rbegin.p;
    d.group_begin = gbegin.p;
    d.terms = terms.p;
    d.heads = heads.p;
    d.nlevels = f.nlevels;
    d.nterms = (int) f.terms.size();
    d.nheads = (int) f.heads.size();
    return d;
  }
};

// 4-atoms-per-lane program kernel: [NA == 16][grade step][term streams in shared memory]
typedef void (*P3Kernel)(DevPotential, SiteArgs, DevFlat3Pass, DevFlat3Pass, const double *, double *, int, double *);
const P3Kernel kP3[2][2][2] = {
    {{mtp_program_v3<32, false, false>, mtp_program_v3<32, false, true>}, {mtp_program_v3<32, true, false>, mtp_program_v3<32, true, true>}},
    {{mtp_program_v3<16, false, false>, mtp_program_v3<16, false, true>}, {mtp_program_v3<16, true, false>, mtp_program_v3<16, true, true>}}};

struct mtp_handle {
  Potential pot;
  Program prog;
  int device = 0;
  int sm_count = 0;
  int chunksize = 32768;          // README.md:52-53 of the reference uses 32768 in every example
  int qpad = 0;
  // device copies of the potential
  DevBuf<double> d_radial, d_species, d_lin, d_ginit, d_ainv;
  DevBuf<uint32_t> d_basic;
  DevBuf<int> d_map;
  PassBufs d_fwd, d_rev;
  ChunkBufs d_cfwd, d_crev;
  FlatBufs d_ffwd[2], d_frev[2];
  Flat3Bufs d_f3fwd, d_f3rev;
  DevFlat3Pass f3f{}, f3r{};
  int last_path = 0;              // mtp_last_kernel_path()
  // mtp_neigh_build scratch
  DevBuf<double> nb_part, nb_bounds;
  DevBuf<int> nb_keys, nb_idx, nb_skeys, nb_sidx, nb_cell, nb_max;
  DevBuf<AtomRec> nb_xs;
  DevBuf<unsigned char> nb_tmp;
  // generated contraction-program kernel (mtp_codegen.hpp): [0] throughput shape, [1] latency shape (mtp/small/kk)
  P4Module p4[2];
  std::vector<short> slot_of_k;   // basic moment -> row of mb / gb
  int p4_nslots = 0;
  std::string p4_note;            // why the generated kernel is not in use (empty when it is)
  std::string p4_note_small;      // same for the latency shape
  int p3_na = 0;                  // atoms per CTA of the 4-atoms-per-lane program kernel, 0 = not usable for this potential
  size_t prog_max = 0;
  int pl_na[2] = {0, 0};          // atoms per CTA of the program kernel: throughput shape, latency shape
  DevPotential dpot{};
  // work buffers (grow-only)
  DevBuf<AtomRec> d_xt;
  DevBuf<double> d_partials, d_cand, d_blockmax, d_cfg;
  DevBuf<int> d_status;
  // host-buffer path
  DevBuf<double> h_f0;
  cudaEvent_t ev_f0 = nullptr;
  size_t h_nid = 0, h_ntype = 0, h_grades_n = 0;
  DevBuf<double> h_x, h_f, h_eatom, h_vatom, h_grades, h_ev, h_cfgc;
  DevBuf<int> h_type, h_ilist, h_numneigh, h_neigh, sel_ids;
  DevBuf<double> sel_val;
  DevBuf<long long> h_offsets;
  DevBuf<unsigned char> h_within;
  long long h_list_len = 0;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  int h_maxnn = 0;
  cudaStream_t hstream = nullptr;
  // register-resident kernel family (mtp_kernels_v1.cuh), -1 = generic kernel only
  int v1_entry = -1;
  DevBuf<short> d_fwd_slot, d_g_src;
  V1Tables v1tab{};
  // launch configuration per kernel flavour
  int warps[2] = {0, 0};
  int grid_cap[2] = {0, 0};
  // three-kernel pipeline (register-resident family)
  DevBuf<double> d_mb, d_gb;      // [K][ld] basic moments / their adjoints of the current super-chunk
  DevBuf<double> d_esite;         // rounds in parallel + eflag_atom: [lane][round][ld] site-energy shares
  int pl_na_fit = 0;              // largest atoms-per-CTA of the program kernel that fits shared memory
  int pl_grid_m = 0, pl_grid_p = 0, pl_grid_f[2] = {0, 0};
  size_t pl_smem_m = 0, pl_smem_f[2] = {0, 0};
  size_t smem[2] = {0, 0};
  // register-resident pair stages for standard basic-moment sets (mtp_kernels_v2.cuh), -1 = not applicable
  int v2_entry = -1;
  DevBuf<short> d_slot_to_k;
  DevBuf<int> d_maxnn;
  // super-chunks are dealt round-robin to `nlanes` internal streams, each with its own scratch, so that the
  // shared-memory-bound program kernel of one chunk overlaps the FP64-bound pair kernels of its neighbours
  struct Lane {
    DevBuf<double> pfld, mb, gb;
    DevBuf<int> pj, pjt, pcnt;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
  };
  static constexpr int kMaxLanes = 4;
  Lane lanes[kMaxLanes];
  cudaEvent_t ev_fork = nullptr;
  cudaStream_t phase_stream = nullptr;
  std::vector<cudaEvent_t> phase_events;
  int nlanes = 2;
  V2GatherFn v2_radial = nullptr;
  int v2_grid_g = 0, v2_grid_r = 0, v2_grid_m = 0, v2_grid_f = 0, v2_grid_fg[2] = {0, 0}, v2_ab = 0;
  size_t v2_smem_g = 0, v2_smem_f = 0, v2_smem_m = 0;
  int v2_chunk = 0;
  // optional per-kernel-class device timing (mtp_profile_enable): CUDA events recorded on the launch stream
  bool profile = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span {
    int cls;
    size_t e0, e1;
  };
  std::vector<Span> spans;
};

namespace {

void set_device(const mtp_handle *h) { CUDA_CHECK(cudaSetDevice(h->device)); }

// dynamic shared memory a kernel may request: opt-in limit minus its static allocation

size_t max_dynamic_smem(const void *fn, size_t optin)
{
  cudaFuncAttributes fa;
  CUDA_CHECK(cudaFuncGetAttributes(&fa, fn));
  return optin > fa.sharedSizeBytes ? optin - fa.sharedSizeBytes : 0;
}

// Kernel attributes are per-device state shared by every handle of the process: the dynamic shared-memory limit of a
// kernel is therefore always raised to the device's opt-in maximum (never to what one potential needs -- a second
// handle with a smaller table would lower it under the first); the per-potential sizes are launch arguments only.
void allow_max_dynamic_smem(const void *fn, size_t optin)
{
  CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) max_dynamic_smem(fn, optin)));
}

// ---- instantiations of the register-resident kernel: (max tensor rank, padded radial function count) ----
typedef void (*V1MomentsKernel)(DevPotential, V1Tables, SiteArgs, double *, int);
typedef void (*V1ForcesKernel)(DevPotential, V1Tables, SiteArgs, const double *, int, double *);
typedef V1Layout (*V1LayoutFn)(int, int, int, int, int, bool, int);
struct V1Entry {
  int deg, rp;
  V1MomentsKernel moments;
  V1ForcesKernel forces[2];
  V1LayoutFn layout;
};
#define V1_ENTRY(D, R) \
  {D, R, mtp_moments_kernel<D, R>, {mtp_forces_kernel<D, R, false>, mtp_forces_kernel<D, R, true>}, v1_layout<D, R>}
const V1Entry kV1[] = {V1_ENTRY(2, 2), V1_ENTRY(3, 4), V1_ENTRY(4, 4), V1_ENTRY(6, 4), V1_ENTRY(8, 6), V1_ENTRY(9, 6)};
constexpr int kV1Count = sizeof(kV1) / sizeof(kV1[0]);

// canonical-slot tables for entry e; returns false if the basic table cannot be mapped (duplicates)
bool build_v1_tables(mtp_handle *h, int e)
{
  const Potential &p = h->pot;
  const int deg = kV1[e].deg, rp = kV1[e].rp, smax = deg / 2;
  const int nb = tet(smax), nq = tet(deg);
  std::vector<short> fwd((size_t) nb * 64, (short) -1), gsrc((size_t) nq * rp, (short) -1);
  int rcnt[16] = {0};
  for (int k = 0; k < p.alpha_index_basic_count; k++) {
    const int *b = &p.alpha_index_basic[4 * (size_t) k];
    const int mu = b[0], ax = b[1], ay = b[2], az = b[3], d = ax + ay + az;
    if (mu >= rp || d > deg || k > 32767) return false;
    const int t = block_index(smax, ax / 2, ay / 2, az / 2);
    const int pcls = (ax & 1) | ((ay & 1) << 1) | ((az & 1) << 2);
    const int lane = mu * 4 + (pcls >> 1), j = pcls & 1;
    short &fs = fwd[((size_t) t * 32 + lane) * 2 + j];
    short &gs = gsrc[(size_t) canon_index(deg, ax, ay, az) * rp + mu];
    if (fs >= 0 || gs >= 0) return false;    // two basic moments with the same definition
    fs = (short) k;
    gs = (short) k;
    rcnt[d] = std::max(rcnt[d], mu + 1);
  }
  h->d_fwd_slot.upload(fwd, 0);
  h->d_g_src.upload(gsrc, 0);
  h->v1tab.fwd_slot = h->d_fwd_slot.p;
  h->v1tab.g_src = h->d_g_src.p;
  for (int d = 0; d < 16; d++) h->v1tab.rcnt[d] = rcnt[d];
  return true;
}

// ---- instantiations of the v2 pair stages, one per D0 (MLIP levels 2..24) ----
typedef void (*V2GatherKernel)(DevPotential, SiteArgs, PairBuf);
typedef void (*V2MomentsKernel)(SiteArgs, PairBuf, double *, int);
typedef void (*V2ForcesKernel)(SiteArgs, PairBuf, const double *, int, double *);
// atoms per CTA of the force kernel: three buffers of canonical adjoints (3 x KF x AB x 8 B) stay below ~110 KB so that
// two CTAs fit an SM; fixed per D0 at compile time so that only one variant per shape is instantiated
constexpr int v2_ab_for(int kf) { return 3 * kf * 64 * 8 <= 110 * 1024 ? 64 : 3 * kf * 32 * 8 <= 110 * 1024 ? 32 : 3 * kf * 16 * 8 <= 110 * 1024 ? 16 : 8; }
struct V2Entry {
  int d0, R, KF, NP, AB;
  V2GatherKernel radial_v[3];    // [0] 8 radial basis functions, unit row stride, no mask output; [1] 8 functions; [2] general
  V2MomentsKernel moments;
  V2ForcesKernel forces[2];    // [grade step]
};
#define V2_ENTRY(D)                                                                                                    \
  {D, V2Shape<D>::R, V2Shape<D>::KF, V2Shape<D>::NP, v2_ab_for(V2Shape<D>::KF), {mtp_gather_radial_kernel<V2Shape<D>::R, 2, 3, true, true>, mtp_gather_radial_kernel<V2Shape<D>::R, 2, 3, true, false>, \
    mtp_gather_radial_kernel<V2Shape<D>::R, 2, 3, false, false>},                                                       \
   mtp_moments_v2<D>,                                                                                                  \
   {mtp_forces_v2<D, v2_ab_for(V2Shape<D>::KF), false>, mtp_forces_v2<D, v2_ab_for(V2Shape<D>::KF), true>}}
const V2Entry kV2[] = {V2_ENTRY(0), V2_ENTRY(1), V2_ENTRY(2), V2_ENTRY(3), V2_ENTRY(4), V2_ENTRY(5),
                       V2_ENTRY(6), V2_ENTRY(7), V2_ENTRY(8), V2_ENTRY(9), V2_ENTRY(10)};
constexpr int kV2Count = sizeof(kV2) / sizeof(kV2[0]);

// canonical slot (q lexicographic in (a,b,c), then mu) -> basic moment index of the file, -1 = absent (host only)
bool v2_slot_table(const Potential &p, int e, std::vector<short> &s2k)
{
  const int D0 = kV2[e].d0, R = kV2[e].R;
  if (p.radial_func_count != R) return false;
  auto dmu = [&](int mu) { return std::max(D0 - 2 * mu, 0); };
  auto rcnt = [&](int d) { return d == 0 ? R : (d > D0 ? 0 : (D0 - d) / 2 + 1); };
  s2k.assign((size_t) kV2[e].KF, (short) -1);
  // slot prefix per monomial
  std::vector<int> prefix((size_t) tet(D0) + 1, 0);
  {
    int q = 0, s = 0;
    for (int a = 0; a <= D0; a++)
      for (int b = 0; b <= D0 - a; b++)
        for (int c = 0; c <= D0 - a - b; c++) {
          prefix[q++] = s;
          s += rcnt(a + b + c);
        }
    prefix[q] = s;
    if (s != kV2[e].KF) return false;
  }
  for (int k = 0; k < p.alpha_index_basic_count; k++) {
    const int *b = &p.alpha_index_basic[4 * (size_t) k];
    const int mu = b[0], d = b[1] + b[2] + b[3];
    if (mu >= R || d > dmu(mu) || k > 32767) return false;
    short &slot = s2k[(size_t) prefix[canon_index(D0, b[1], b[2], b[3])] + mu];
    if (slot >= 0) return false;    // two basic moments with the same definition
    slot = (short) k;
  }
  return true;
}

// v2 entry serving potential p (standard basic-moment set of degree P - 1), or -1
int v2_entry_for(const Potential &p, std::vector<short> &s2k)
{
  const int pmax = p.max_alpha_index_basic - 1;
  for (int e = 0; e < kV2Count; e++)
    if (kV2[e].d0 == pmax) return v2_slot_table(p, e, s2k) ? e : -1;
  return -1;
}

// basic moment -> row of the mb / gb arrays the program kernel exchanges with the pair kernels
void p4_slot_map(const Potential &p, int v2_entry, const std::vector<short> &s2k, std::vector<short> &slot_of_k, int &nslots)
{
  const int K = p.alpha_index_basic_count;
  slot_of_k.assign((size_t) K, (short) -1);
  if (v2_entry >= 0) {
    nslots = kV2[v2_entry].KF;
    for (int s = 0; s < nslots; s++)
      if (s2k[s] >= 0) slot_of_k[s2k[s]] = (short) s;
  } else {
    nslots = K;
    for (int k = 0; k < K; k++) slot_of_k[k] = (short) k;
  }
}

bool build_v2_tables(mtp_handle *h, int e)
{
  std::vector<short> s2k;
  if (!v2_slot_table(h->pot, e, s2k)) return false;
  h->d_slot_to_k.upload(s2k, 0);
  p4_slot_map(h->pot, e, s2k, h->slot_of_k, h->p4_nslots);
  return true;
}

void upload_potential(mtp_handle *h)
{
  Potential &p = h->pot;
  {
    // program kernel shape: largest power-of-two atoms-per-CTA whose moments + adjoints ((M+1) rows each) fit
    cudaDeviceProp prop0;
    CUDA_CHECK(cudaGetDeviceProperties(&prop0, h->device));
    const size_t prog_max = std::min(max_dynamic_smem((const void *) mtp_program_kernel<false>, prop0.sharedMemPerBlockOptin),
                                     max_dynamic_smem((const void *) mtp_program_kernel<true>, prop0.sharedMemPerBlockOptin));
    h->prog_max = prog_max;
    h->pl_na_fit = 0;
    for (int na = 32; na >= 2; na >>= 1)
      if (program_layout(p.alpha_moment_count, count_adjoint_rows(p), p.alpha_scalar_count, na, 2 * p.alpha_index_basic_count, 0, 0, false).total <= prog_max) {
        h->pl_na_fit = na;
        break;
      }
    h->pl_na[0] = std::max(2, h->pl_na_fit);
    if (const char *e = getenv("MTP_B200_PROG_NA")) h->pl_na[0] = std::max(2, std::min(h->pl_na[0], atoi(e)));
    h->pl_na[1] = std::min(h->pl_na[0], 8);
    // 4-atoms-per-lane kernel: 32 or 16 atoms per CTA, whichever fits with both tables resident
    int na_v3 = 0;
    if (!getenv("MTP_B200_NO_PROG_V3"))
      for (int na = getenv("MTP_B200_P3_NA") ? std::max(16, std::min(32, atoi(getenv("MTP_B200_P3_NA")))) : 32; na >= 16 && !na_v3; na >>= 1)
        if (program3_layout(p.alpha_moment_count, p.alpha_scalar_count, na, 2 * p.alpha_index_basic_count, 0, 0, 0, 0, false).total <= prog_max)
          na_v3 = na;
    compile_program(p, h->prog, h->pl_na[0], h->pl_na[1], na_v3);
    h->p3_na = h->prog.f3_na;
    for (int q = 0; q < 8; q++) {
      const void *k = (const void *) kP3[q >> 2][(q >> 1) & 1][q & 1];
      CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) prog_max));
      if (!getenv("MTP_B200_NO_CARVEOUT"))
        CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    CUDA_CHECK(cudaFuncSetAttribute((const void *) mtp_program_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) prog_max));
    CUDA_CHECK(cudaFuncSetAttribute((const void *) mtp_program_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) prog_max));
    // one shared-memory carve-out for every kernel of the pipeline: an SM cannot change its L1 / shared split while
    // CTAs are resident, so kernels of different lanes can only co-reside if they all ask for the same (maximum) split
    if (!getenv("MTP_B200_NO_CARVEOUT")) {
      CUDA_CHECK(cudaFuncSetAttribute((const void *) mtp_program_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      CUDA_CHECK(cudaFuncSetAttribute((const void *) mtp_program_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
  }
  std::vector<uint32_t> basic(p.alpha_index_basic_count);
  for (int k = 0; k < p.alpha_index_basic_count; k++) {
    const int *e = &p.alpha_index_basic[4 * (size_t) k];
    basic[k] = (uint32_t) e[0] | ((uint32_t) e[1] << 8) | ((uint32_t) e[2] << 16) | ((uint32_t) e[3] << 24);
  }
  if (p.radial_func_count > 255) throw std::runtime_error("radial_funcs_count above 255 is not supported.");
  h->d_radial.upload(p.radial_basis_coeffs, 0);
  h->d_species.upload(p.species_coeffs, 0);
  h->d_lin.upload(p.linear_coeffs, 0);
  h->d_ginit.upload(h->prog.ginit, 0);
  h->d_basic.upload(basic, 0);
  h->d_map.upload(p.alpha_moment_mapping, 0);
  h->d_fwd.upload(h->prog.fwd);
  h->d_rev.upload(h->prog.rev);
  if (p.has_selection_state) {
    const int Q = p.coeff_count;
    h->qpad = (Q + 7) / 8 * 8;    // zero padding cannot raise a max of absolute values
    std::vector<double> pad((size_t) h->qpad * h->qpad, 0.0);
    for (int i = 0; i < Q; i++)
      memcpy(&pad[(size_t) i * h->qpad], &p.inverse_active_set[(size_t) i * Q], sizeof(double) * Q);
    h->d_ainv.upload(pad, 0);
    h->d_cfg.ensure((size_t) h->qpad);
  }
  h->d_status.ensure(1);
  CUDA_CHECK(cudaMemset(h->d_status.p, 0, sizeof(int)));
  CUDA_CHECK(cudaDeviceSynchronize());

  DevPotential &d = h->dpot;
  d.S = p.species_count;
  d.R = p.radial_func_count;
  d.B = p.radial_basis_size;
  d.K = p.alpha_index_basic_count;
  d.M = p.alpha_moment_count;
  d.Mg = h->prog.adjoint_rows;
  d.A = p.alpha_scalar_count;
  d.P = p.max_alpha_index_basic;
  d.Q = p.has_selection_state ? p.coeff_count : 0;
  d.rmin = p.min_cutoff;
  d.rmax = p.max_cutoff;
  d.scaling = p.scaling;
  d.cutsq = p.max_cutoff * p.max_cutoff;    // pair_mtp.cpp:449
  d.radial = h->d_radial.p;
  d.basic = h->d_basic.p;
  d.species = h->d_species.p;
  d.lin = h->d_lin.p;
  d.map = h->d_map.p;
  d.ginit = h->d_ginit.p;
  d.fwd = h->d_fwd.view(h->prog.fwd);
  d.rev = h->d_rev.view(h->prog.rev);
  DevChunkPass cf = h->d_cfwd.upload(h->prog.cfwd), cr = h->d_crev.upload(h->prog.crev);
  CUDA_CHECK(cudaDeviceSynchronize());
  d.cfwd = cf;
  d.crev = cr;
  for (int v = 0; v < 2; v++) {
    d.ffwd[v] = h->d_ffwd[v].upload(h->prog.ffwd[v]);
    d.frev[v] = h->d_frev[v].upload(h->prog.frev[v]);
  }
  if (h->p3_na) {
    h->f3f = h->d_f3fwd.upload(h->prog.f3fwd);
    h->f3r = h->d_f3rev.upload(h->prog.f3rev);
  }
  CUDA_CHECK(cudaDeviceSynchronize());

  // kernel family: register-resident kernel for standard shapes, generic kernel otherwise
  h->v1_entry = -1;
  if (!getenv("MTP_B200_FORCE_GENERIC")) {
    for (int e = 0; e < kV1Count; e++)
      if (kV1[e].deg >= d.P - 1 && kV1[e].rp >= d.R && d.P - 1 >= 0) {
        if (build_v1_tables(h, e)) h->v1_entry = e;
        break;
      }
  }

  // launch configuration
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, h->device));
  h->sm_count = prop.multiProcessorCount;
  const size_t smem_max = prop.sharedMemPerBlockOptin;

  h->v2_entry = -1;
  if (!getenv("MTP_B200_FORCE_GENERIC") && !getenv("MTP_B200_NO_V2") && h->pl_na_fit >= 1) {
    int pmax = d.P - 1;
    for (int e = 0; e < kV2Count; e++)
      if (kV2[e].d0 == pmax) {
        if (build_v2_tables(h, e)) h->v2_entry = e;
        break;
      }
  }
  if (h->v2_entry >= 0) {
    const V2Entry &E = kV2[h->v2_entry];
    const int nrad = d.S * d.S * d.R * d.B;
    h->v2_smem_g = (size_t) ((nrad + 1) & ~1) * 8 + (size_t) 8 * V2_RING * 4 * 8;    // coefficients + one ring per warp
    // resident CTAs per SM the gather kernel is compiled for: the gather is bound by the latency of its L2 gathers, so
    // resident warps count more than registers (the radial phase spills a few values at 48 registers)
    h->v2_radial = E.radial_v[d.B == 8 ? 1 : 2];
    bool ok = h->v2_smem_g <= max_dynamic_smem((const void *) h->v2_radial, smem_max);
    if (ok) {
      for (int gv = 0; gv < 3; gv++) allow_max_dynamic_smem((const void *) E.radial_v[gv], smem_max);
      if (!getenv("MTP_B200_NO_CARVEOUT")) {
        // the gather lives on its L1 hit rate (see the kernel): it asks for the smallest shared-memory carve-out that holds
        // its rings, unlike the other kernels of the pipeline, which all ask for the maximum
        int carve = 25;
        if (const char *e = getenv("MTP_B200_GR_CARVEOUT")) carve = atoi(e);
        for (int gv = 0; gv < 3; gv++)
          CUDA_CHECK(cudaFuncSetAttribute((const void *) E.radial_v[gv], cudaFuncAttributePreferredSharedMemoryCarveout, carve));
        CUDA_CHECK(cudaFuncSetAttribute((const void *) E.moments, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        for (int gq = 0; gq < 2; gq++)
          CUDA_CHECK(cudaFuncSetAttribute((const void *) E.forces[gq], cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      }
      int per_sm = 0;
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *) h->v2_radial, 256, h->v2_smem_g));
      h->v2_grid_r = std::max(1, per_sm) * h->sm_count;
      h->v2_grid_g = h->v2_grid_r;
      h->v2_smem_m = (size_t) 2 * 32 * ((3 + E.R) * V2_NT + 4) * 8;
      ok = h->v2_smem_m <= max_dynamic_smem((const void *) E.moments, smem_max);
      if (ok) {
        allow_max_dynamic_smem((const void *) E.moments, smem_max);
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *) E.moments, 32 * E.NP, h->v2_smem_m));
        h->v2_grid_m = std::max(1, per_sm) * h->sm_count;
      }
      // forces: atoms per CTA fixed per shape (v2_ab_for)
      h->v2_ab = E.AB;
      h->v2_smem_f = (size_t) V2_FNB * E.KF * E.AB * 8 + (size_t) V2_FNB * (E.AB + 1) * 4;
      for (int gq = 0; gq < 2 && ok; gq++) {
        const void *fk = (const void *) E.forces[gq];
        ok = ok && h->v2_smem_f <= max_dynamic_smem(fk, smem_max);
        if (!ok) break;
        allow_max_dynamic_smem(fk, smem_max);
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fk, 256, h->v2_smem_f));
        h->v2_grid_fg[gq] = std::max(1, per_sm) * h->sm_count;
      }
      h->v2_grid_f = std::max(h->v2_grid_fg[0], h->v2_grid_fg[1]);
    }
    if (!ok) h->v2_entry = -1;
    if (const char *le = getenv("MTP_B200_LANES")) h->nlanes = std::max(1, std::min((int) mtp_handle::kMaxLanes, atoi(le)));
    const char *ce = getenv("MTP_B200_CHUNK");
    h->v2_chunk = ce ? std::max(1024, atoi(ce)) : 0;
    h->d_maxnn.ensure(1);
  }
  if (h->v1_entry >= 0) {
    const V1Entry &E = kV1[h->v1_entry];
    const int W = 8;
    bool ok = true;
    {
      const V1Layout L = E.layout(d.S, d.R, d.B, d.M, d.Q, false, 1);
      h->pl_smem_m = L.radial_bytes + W * L.warp_bytes_moments;
      ok = ok && h->pl_smem_m <= max_dynamic_smem((const void *) E.moments, smem_max);
      if (ok) {
        allow_max_dynamic_smem((const void *) E.moments, smem_max);
        int per_sm = 0;
        CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *) E.moments, W * 32, h->pl_smem_m));
        h->pl_grid_m = std::max(1, per_sm) * h->sm_count;
      }
    }
    for (int gflag = 0; gflag < 2 && ok; gflag++) {
      if (gflag == 1 && !p.has_selection_state) continue;
      const V1Layout L = E.layout(d.S, d.R, d.B, d.M, d.Q, gflag == 1, 1);
      h->pl_smem_f[gflag] = L.radial_bytes + W * L.warp_bytes_forces;
      ok = ok && h->pl_smem_f[gflag] <= max_dynamic_smem((const void *) E.forces[gflag], smem_max);
      if (!ok) break;
      allow_max_dynamic_smem((const void *) E.forces[gflag], smem_max);
      int per_sm = 0;
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *) E.forces[gflag], W * 32, h->pl_smem_f[gflag]));
      h->pl_grid_f[gflag] = std::max(1, per_sm) * h->sm_count;
    }
    ok = ok && h->pl_na_fit >= 1;
    if (!ok) h->v1_entry = -1;    // fall back to the generic kernel
  }
  for (int gflag = 0; gflag < 2; gflag++) {
    if (gflag == 1 && !p.has_selection_state) continue;
    const Layout L = make_layout(d.S, d.R, d.B, d.K, d.M, d.P, d.Q, gflag == 1);
    const void *fn = gflag ? (const void *) mtp_site_kernel<true> : (const void *) mtp_site_kernel<false>;
    int w = 4;
    const size_t gen_max = max_dynamic_smem(fn, smem_max);
    while (w > 1 && L.cta_bytes + (size_t) w * L.warp_bytes > gen_max) w--;
    const size_t bytes = L.cta_bytes + (size_t) w * L.warp_bytes;
    if (bytes > gen_max) {
      if (h->v1_entry >= 0 || h->v2_entry >= 0) continue;    // the pipelines serve this potential
      throw std::runtime_error("potential too large for on-chip per-atom state (alpha_moments_count)");
    }
    allow_max_dynamic_smem(fn, smem_max);
    int per_sm = 0;
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, w * 32, bytes));
    if (per_sm < 1) per_sm = 1;
    h->warps[gflag] = w;
    h->smem[gflag] = bytes;
    h->grid_cap[gflag] = per_sm * h->sm_count;
  }

  if (p.has_selection_state) {    // grade kernels (per device, like everything above)
    allow_max_dynamic_smem((const void *) grade_dmma_reg_kernel<64>, smem_max);
    allow_max_dynamic_smem((const void *) mtp_cand_radial_kernel<2>, smem_max);
    allow_max_dynamic_smem((const void *) mtp_cand_radial_kernel<4>, smem_max);
    allow_max_dynamic_smem((const void *) mtp_cand_radial_kernel<8>, smem_max);
  }
  // generated contraction-program kernel for the two-kernel pipelines (NVRTC, cached cubins)
  h->p4[0].unload();
  h->p4[1].unload();
  h->p4_note.clear();
  h->p4_note_small.clear();
  if ((h->v2_entry >= 0 || h->v1_entry >= 0) && !getenv("MTP_B200_NO_P4")) {
    if (h->v2_entry < 0) {
      std::vector<short> none;
      p4_slot_map(p, -1, none, h->slot_of_k, h->p4_nslots);
    }
    for (int shape = 0; shape < 2; shape++) {
      try {
        const P4Choice ch = p4_choose(p, smem_max, shape == 1);
        if (!ch.ok) throw std::runtime_error("potential structure or size outside the generator's range");
        if (shape == 1 && h->p4[0].loaded() && ch.prm.na == h->p4[0].choice.prm.na && ch.prm.groups == h->p4[0].choice.prm.groups) continue;    // same kernel
        P4Module &m = h->p4[shape];
        const std::vector<char> cubin = p4_cubin(p, ch, h->slot_of_k.data(), h->p4_nslots, m.info);
        m.choice = ch;
        m.load(cubin, h->device, h->sm_count);
      } catch (const std::exception &e) {
        h->p4[shape].unload();
        if (shape == 0) h->p4_note = e.what();
        else
          h->p4_note_small = e.what();
        if (getenv("MTP_B200_P4_REQUIRE")) throw;
      }
    }
  } else
    h->p4_note = "disabled";
}

// RAII span: records an event pair around the launches of one kernel class when profiling is on
struct ProfSpan {
  mtp_handle *h;
  cudaStream_t st;
  size_t e0 = 0;
  int cls;
  static size_t grab(mtp_handle *h)
  {
    if (h->ev_used == h->ev_pool.size()) {
      cudaEvent_t e;
      CUDA_CHECK(cudaEventCreate(&e));
      h->ev_pool.push_back(e);
    }
    return h->ev_used++;
  }
  ProfSpan(mtp_handle *h_, int cls_, cudaStream_t st_) : h(h_), st(st_), cls(cls_)
  {
    if (!h->profile) return;
    e0 = grab(h);
    CUDA_CHECK(cudaEventRecord(h->ev_pool[e0], st));
  }
  ~ProfSpan()
  {
    if (!h->profile) return;
    const size_t e1 = grab(h);
    cudaEventRecord(h->ev_pool[e1], st);
    h->spans.push_back({cls, e0, e1});
  }
};

// atoms per super-chunk for a call with inum centres (shared by launch_site and the host path's upload pipeline)
// grade steps ride the v2 pair stages too (candidate kernel keeps one accumulator per neighbor species: S <= 8)
bool v2_applies(const mtp_handle *h, bool grade)
{
  return h->v2_entry >= 0 && (!grade || (h->dpot.S <= 8 && h->dpot.B + h->dpot.R <= 22));
}

int plan_chunk(const mtp_handle *h, int inum, bool grade)
{
  const bool use_v2 = v2_applies(h, grade);
  const bool pipeline = use_v2 || h->v1_entry >= 0;
  int chunk = std::max(1, std::min(h->chunksize, inum > 0 ? inum : 1));
  if (pipeline) {
    long long fit = std::max(8192LL, (48LL << 20) / (16LL * h->dpot.K) / 1024 * 1024);
    if (use_v2) fit = h->v2_chunk > 0 ? h->v2_chunk : (1LL << 30);    // v2: the user's chunksize alone bounds the scratch
    chunk = (int) std::min<long long>(chunk, fit);
    // wave quantisation: the pair kernels run persistent grids (moments: v2_grid_m CTAs x 32 atoms, forces:
    // v2_grid_fg x AB, program: one CTA per SM x NA), so a super-chunk that is not a whole number of grid rounds of each
    // leaves SMs idle in the last round of EVERY launch (32768 atoms at level 16: 2.3 rounds of the moment kernel
    // cost 3).  Several super-chunks: round the size down to a common multiple of the three round sizes.
    if (use_v2 && inum > chunk && !getenv("MTP_B200_NO_WAVE_CHUNK")) {
      auto lcm = [](long long x, long long y) {
        long long a = x, b = y;
        while (b) {
          const long long t = a % b;
          a = b;
          b = t;
        }
        return x / a * y;
      };
      const long long um = (long long) h->v2_grid_m * 32, uf = (long long) h->v2_grid_fg[grade ? 1 : 0] * h->v2_ab;
      const long long up = h->p4[0].loaded() ? (long long) h->p4[0].grid_cap * h->p4[0].choice.atoms_per_cta()
                                             : (long long) h->sm_count * (h->p3_na ? h->p3_na : std::max(1, h->pl_na[0]));
      long long unit = 0;
      if (um > 0 && uf > 0) {
        unit = lcm(um, uf);
        if (unit <= chunk && lcm(unit, up) <= chunk) unit = lcm(unit, up);
      }
      if (unit > 0 && unit <= chunk) chunk = (int) (chunk / unit * unit);
    }
  } else if (!grade) {
    chunk = inum > 0 ? inum : 1;
  }
  return chunk;
}

// chunk_ready (optional): one event per super-chunk that must have completed before the chunk's kernels may read
// the neighbor list (the host path uploads the list slice by slice while earlier chunks compute)
// phases (optional): the centres are cut into consecutive runs of the list ("phases", mtp_compute_phased); super-chunks
// never straddle a phase, the kernels of a phase wait for its event, and an event is recorded when a phase is complete
struct PhaseSpec {
  int nphase = 0;
  const int *inum = nullptr;
  void *const *wait_events = nullptr;
  void *const *done_events = nullptr;
};

void launch_site(mtp_handle *h, const mtp_compute_args &a, cudaStream_t st,
                 const std::vector<cudaEvent_t> *chunk_ready = nullptr, const PhaseSpec *phases = nullptr)
{
  const DevPotential &d = h->dpot;
  const bool grade = a.want_grade != 0;
  if (grade && !h->pot.has_selection_state)
    throw std::invalid_argument("extrapolation grades requested but the potential has no selection state");
  if (a.inum < 0 || a.nall < a.inum) throw std::invalid_argument("bad inum / nall");

  // pack positions + species into 32-byte records
  h->d_xt.ensure((size_t) (a.nall > 0 ? a.nall : 1));
  if (a.nall > 0) {
    ProfSpan sp(h, MTP_PROF_PACK, st);
    pack_xt_kernel<<<(a.nall + 255) / 256, 256, 0, st>>>(a.nall, a.x, a.type, h->d_xt.p);
    g_launches++;
  }

  SiteArgs s{};
  s.nall = a.nall;
  s.xt = h->d_xt.p;
  s.ilist = a.ilist;
  s.numneigh = a.numneigh;
  s.neighbors = a.neighbors;
  s.neigh_offsets = a.neigh_offsets;
  s.stride_i = a.stride_i;
  s.stride_jj = a.stride_jj > 0 ? a.stride_jj : 1;
  s.neighmask = a.neighmask ? a.neighmask : 0x1FFFFFFF;
  s.eflag_global = (a.eflag & 1) != 0;
  s.eflag_atom = (a.eflag & 2) != 0 && a.eatom != nullptr;
  s.vflag_any = a.vflag != 0;
  s.vflag_atom = (a.vflag & 4) != 0 && a.vatom != nullptr;
  s.want_grade = grade;
  s.f = a.f;
  s.eatom = a.eatom;
  s.vatom = a.vatom;
  s.grades = a.grades;
  s.within = a.within_cutoff;
  s.status = h->d_status.p;

  const int gi = grade ? 1 : 0;
  const bool cfg = grade && h->pot.configuration_mode;
  const bool use_v2 = v2_applies(h, grade);
  const bool pipeline = use_v2 || h->v1_entry >= 0;
  // super-chunk: bounded by the user's "chunksize"; the pipeline additionally keeps its two [K][chunk]
  // intermediates within ~48 MB so that they stay L2-resident
  const int chunk = plan_chunk(h, a.inum, grade);
  PairBuf pb{};
  if (use_v2) {
    int maxnn = a.max_numneigh;
    if (maxnn <= 0) {    // unknown bound: one small reduction + a 4-byte read-back
      CUDA_CHECK(cudaMemsetAsync(h->d_maxnn.p, 0, sizeof(int), st));
      if (a.inum > 0) {
        max_numneigh_kernel<<<std::min(h->sm_count * 4, (a.inum + 255) / 256), 256, 0, st>>>(a.inum, a.ilist, a.numneigh,
                                                                                            h->d_maxnn.p);
        g_launches++;
      }
      CUDA_CHECK(cudaMemcpyAsync(&maxnn, h->d_maxnn.p, sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_CHECK(cudaStreamSynchronize(st));
    }
    const V2Entry &E = kV2[h->v2_entry];
    pb.ncap = std::max(4, (maxnn + 3) / 4 * 4);
    pb.cap = (long long) chunk * pb.ncap;
    (void) E;
  }
  const int ld = (chunk + 63) / 64 * 64;
  if (grade) h->d_cand.ensure((size_t) chunk * h->qpad);
  // super-chunks: {first centre, count, phase}
  struct Chunk {
    int first, n, phase;
  };
  std::vector<Chunk> chunks;
  if (phases && phases->nphase > 0) {
    int at = 0;
    for (int p = 0; p < phases->nphase; p++) {
      const int np = phases->inum[p];
      // a phase is cut into equal super-chunks no larger than `chunk`
      const int parts = np > 0 ? (np + chunk - 1) / chunk : 0;
      for (int q = 0; q < parts; q++) {
        const int b = (int) ((long long) np * q / parts), e = (int) ((long long) np * (q + 1) / parts);
        chunks.push_back({at + b, e - b, p});
      }
      at += np;
    }
    if (at != a.inum) throw std::invalid_argument("the phases do not add up to inum");
  } else
    for (int first = 0; first < std::max(a.inum, 1); first += chunk)
      chunks.push_back({first, std::max(0, std::min(chunk, a.inum - first)), 0});
  if (chunks.empty()) chunks.push_back({0, 0, 0});
  const int nsuper = (int) chunks.size();
  // grade scratch is not per lane; a program kernel that needs (nearly) all the shared memory of an SM cannot share it
  // with the kernels of another lane and only gets in their way (level 22: 215 ms serialised, 269 ms on three lanes)
  // (measured again with two 115 KB CTAs per SM at level 22: 8.7 ms on one lane, 12.1 ms on three; at level 16 the four
  // 57 KB CTAs are a fifth of the step and the lanes still gain 5 %: the rule looks at the program's share of the work too)
  const P4Module &pm0 = h->p4[0];
  const size_t p4_per_sm = pm0.loaded() ? pm0.info.smem_bytes * (size_t) std::max(1, pm0.grid_cap / std::max(1, h->sm_count)) : 0;
  const bool sm_filling_program = pm0.loaded() && !getenv("MTP_B200_FORCE_LANES") &&
                                  (pm0.info.smem_bytes > 160 * 1024 || (p4_per_sm > 200 * 1024 && pm0.info.terms >= 8000));
  const int nlanes = (use_v2 && !grade && !sm_filling_program) ? std::max(1, std::min(h->nlanes, nsuper)) : 1;
  if (use_v2) {
    const V2Entry &E = kV2[h->v2_entry];
    for (int l = 0; l < nlanes; l++) {
      mtp_handle::Lane &L = h->lanes[l];
      L.pfld.ensure((size_t) (4 + 2 * E.R) * pb.cap);
      L.pj.ensure((size_t) pb.cap);
      L.pjt.ensure((size_t) pb.cap);
      L.pcnt.ensure((size_t) chunk);
      L.mb.ensure((size_t) E.KF * ld);
      L.gb.ensure((size_t) E.KF * ld);
      if (nlanes > 1 && !L.stream) {
        CUDA_CHECK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
        CUDA_CHECK(cudaEventCreateWithFlags(&L.done, cudaEventDisableTiming));
      }
    }
    if (nlanes > 1 && !h->ev_fork) CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  } else if (pipeline) {
    h->d_mb.ensure((size_t) d.K * ld);
    h->d_gb.ensure((size_t) d.K * ld);
  }
  if (cfg) CUDA_CHECK(cudaMemsetAsync(h->d_cfg.p, 0, sizeof(double) * h->qpad, st));
  CUDA_CHECK(cudaMemsetAsync(a.ev_out, 0, sizeof(double) * 8, st));

  const int W = 8;
  // program kernel shape: atoms per CTA (power of two), smaller for small systems so that every SM gets work
  int na = 1, lna = 0, grid_p_cap = 1;
  size_t smem_p = 0;
  P3Kernel p3 = nullptr;
  const P4Module *p4m = nullptr;
  if (pipeline) {
    // throughput shape unless the system is too small to give every SM a chunk (or the latency variant is asked for)
    const int nfirst = std::min(chunk, std::max(a.inum, 1));
    const bool small = a.variant == MTP_VARIANT_SMALL || (nfirst + h->pl_na[0] - 1) / h->pl_na[0] < h->sm_count;
    na = h->pl_na[small ? 1 : 0];
    s.prog_shape = small ? 1 : 0;
    const int nslots = use_v2 ? kV2[h->v2_entry].KF : d.K;
    const int ntf = d.ffwd[s.prog_shape].nterms, ntr = d.frev[s.prog_shape].nterms;
    // shared-memory options in order of preference: prefetch staging + term streams, prefetch only, streams only
    auto fits = [&](bool ds, bool pf) { return program_layout(d.M, d.Mg, d.A, na, nslots, ntf, ntr, ds, pf).total <= h->prog_max; };
    s.prog_prefetch = 0;
    s.prog_dsmem = (nlanes == 1 && fits(true, s.prog_prefetch != 0)) ? 1 : 0;    // lanes: leave room for a co-resident CTA
    if (const char *e = getenv("MTP_B200_PROG_PREFETCH")) s.prog_prefetch = atoi(e) && fits(false, true);
    if (const char *e = getenv("MTP_B200_PROG_DSMEM")) s.prog_dsmem = atoi(e) && fits(true, s.prog_prefetch != 0);
    else if (!fits(true, s.prog_prefetch != 0)) s.prog_dsmem = 0;
    smem_p = program_layout(d.M, d.Mg, d.A, na, nslots, ntf, ntr, s.prog_dsmem != 0, s.prog_prefetch != 0).total;
    s.prog_debug = getenv("MTP_B200_PROG_DEBUG") ? atoi(getenv("MTP_B200_PROG_DEBUG")) : 0;
    for (lna = 0; (1 << lna) < na; lna++) {}
    s.slot_to_k = use_v2 ? h->d_slot_to_k.p : nullptr;
    s.nslots = nslots;
    int per_sm = 0;
    const void *pk = grade ? (const void *) mtp_program_kernel<true> : (const void *) mtp_program_kernel<false>;
    // throughput shape: the 4-atoms-per-lane kernel when the tables are expressible in its compact streams
    if (!small && h->p3_na) {
      auto l3 = [&](bool ds) {
        return program3_layout(d.M, d.A, h->p3_na, nslots, h->f3f.nterms, h->f3r.nterms, h->f3f.nheads, h->f3r.nheads, ds,
                               h->f3f.nlevels, h->f3r.nlevels).total;
      };
      if (l3(false) <= h->prog_max) {
        bool ds = l3(true) <= h->prog_max;    // term streams from L1/L2 cost far more than a co-resident CTA of another lane gains
        if (const char *e = getenv("MTP_B200_PROG_DSMEM")) ds = atoi(e) && l3(true) <= h->prog_max;
        p3 = kP3[h->p3_na == 16][grade ? 1 : 0][ds ? 1 : 0];
        na = h->p3_na;
        smem_p = l3(ds);
        pk = (const void *) p3;
      }
    }
    CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pk, p3 ? P3_THREADS : PROG_THREADS, smem_p));
    grid_p_cap = std::max(1, per_sm) * h->sm_count;
    // the generated kernel of this potential, when it was built: latency shape for small systems if there is one
    // (a system that cannot give every SM a 32-atom chunk is a small system whatever the style is called)
    const bool small_p4 = a.variant == MTP_VARIANT_SMALL || nfirst < 32 * h->sm_count;
    p4m = (small_p4 && h->p4[1].loaded()) ? &h->p4[1] : (h->p4[0].loaded() ? &h->p4[0] : nullptr);
    if (p4m) {
      p3 = nullptr;
      na = p4m->choice.atoms_per_cta();
      grid_p_cap = p4m->grid_cap;
    }
  }
  h->last_path = (use_v2 ? 2 : pipeline ? 1 : 0) | (p3 ? 16 : 0) | (p4m ? 32 : 0) | (na << 8);
  const int p4_rounds = (p4m && p4m->info.rpar_rounds > 1) ? p4m->info.rpar_rounds : 1;    // rounds in parallel: gridDim.y
  const int rows_per_super = pipeline ? grid_p_cap * p4_rounds + std::max(h->pl_grid_f[gi], h->v2_grid_f) : h->grid_cap[gi];
  h->d_partials.ensure((size_t) nsuper * rows_per_super * 8);
  int rows_used = 0;
  if (nlanes > 1) {    // fork: the lanes start after everything queued on the caller's stream so far
    CUDA_CHECK(cudaEventRecord(h->ev_fork, st));
    for (int l = 0; l < nlanes; l++) CUDA_CHECK(cudaStreamWaitEvent(h->lanes[l].stream, h->ev_fork, 0));
  }

  // returns the rows of `part` the launch fills
  auto launch_p4 = [&](const SiteArgs &sa, const double *mbp, double *gbp, double *part, int grid, cudaStream_t ls, int lane) -> int {
    P4Args pa{};
    pa.mb = mbp;
    pa.gb = gbp;
    pa.ld = ld;
    pa.inum = sa.inum;
    pa.first_ii = sa.first_ii;
    pa.ilist = sa.ilist;
    pa.xt = sa.xt;
    pa.lin = d.lin;
    pa.species = d.species;
    pa.S = d.S;
    pa.eflag_global = sa.eflag_global;
    pa.eflag_atom = sa.eflag_atom;
    pa.grade = grade ? 1 : 0;
    pa.eatom = sa.eatom;
    pa.cand_rows = sa.cand_rows;
    pa.cand_ld = sa.cand_ld;
    pa.cand_col0 = d.S * d.S * d.R * d.B + d.S;
    pa.partials = part;
    pa.esite = nullptr;
    if (p4_rounds > 1) {    // every round ADDS its shares: gb starts at zero; per-atom energies are summed in round order
      CUDA_CHECK(cudaMemsetAsync(gbp, 0, (size_t) h->p4_nslots * ld * sizeof(double), ls));
      if (sa.eflag_atom) {
        h->d_esite.ensure((size_t) mtp_handle::kMaxLanes * p4_rounds * ld);    // one region per lane
        pa.esite = h->d_esite.p + (size_t) lane * p4_rounds * ld;
      }
    }
    void *kargs[] = {&pa};
    CUDA_CHECK(cudaLaunchKernel((const void *) p4m->kernel, dim3(grid, p4_rounds), dim3(p4m->info.threads), kargs,
                                p4m->info.smem_bytes, ls));
    if (pa.esite && sa.inum > 0)
      esite_sum_kernel<<<(sa.inum + 255) / 256, 256, 0, ls>>>(sa.inum, sa.first_ii, sa.ilist, pa.esite, ld, p4_rounds, sa.eatom);
    return grid * p4_rounds;
  };

  // a phase is complete when every lane has drained the chunks dealt to it so far
  auto phase_done = [&](int p) {
    if (!phases || !phases->done_events || !phases->done_events[p]) return;
    cudaEvent_t done = (cudaEvent_t) phases->done_events[p];
    if (nlanes > 1) {
      if (!h->phase_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&h->phase_stream, cudaStreamNonBlocking));
      for (int l = 0; l < nlanes; l++) {
        if (h->phase_events.size() <= (size_t) l) {
          cudaEvent_t e;
          CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
          h->phase_events.push_back(e);
        }
        CUDA_CHECK(cudaEventRecord(h->phase_events[l], h->lanes[l].stream));
        CUDA_CHECK(cudaStreamWaitEvent(h->phase_stream, h->phase_events[l], 0));
      }
      CUDA_CHECK(cudaEventRecord(done, h->phase_stream));
    } else
      CUDA_CHECK(cudaEventRecord(done, st));
  };

  for (int sc = 0; sc < nsuper; sc++) {
    const int first = chunks[sc].first;
    const int n = chunks[sc].n;
    if (sc > 0 && chunks[sc].phase != chunks[sc - 1].phase)
      for (int p = chunks[sc - 1].phase; p < chunks[sc].phase; p++) phase_done(p);
    if (phases && phases->wait_events && phases->wait_events[chunks[sc].phase]) {
      cudaStream_t ws = (use_v2 && nlanes > 1) ? h->lanes[sc % nlanes].stream : st;
      CUDA_CHECK(cudaStreamWaitEvent(ws, (cudaEvent_t) phases->wait_events[chunks[sc].phase], 0));
    }
    s.inum = n;
    s.first_ii = first;
    s.cand_rows = grade ? h->d_cand.p : nullptr;
    s.cand_ld = h->qpad;
    if (chunk_ready && sc < (int) chunk_ready->size()) {
      cudaStream_t ws = (use_v2 && nlanes > 1) ? h->lanes[sc % nlanes].stream : st;
      CUDA_CHECK(cudaStreamWaitEvent(ws, (*chunk_ready)[sc], 0));
    }
    if (use_v2) {
      const V2Entry &E = kV2[h->v2_entry];
      mtp_handle::Lane &L = h->lanes[sc % nlanes];
      cudaStream_t ls = nlanes > 1 ? L.stream : st;
      pb.fld = L.pfld.p;
      pb.pj = L.pj.p;
      pb.pjt = L.pjt.p;
      pb.pcnt = L.pcnt.p;
      {
        ProfSpan sp(h, MTP_PROF_GATHER, ls);
        const int gr = std::max(1, std::min(h->v2_grid_r, (n + 7) / 8));
        // the plain instance when the call allows it (unit stride within a row, no cutoff-mask output)
        V2GatherFn gk = (d.B == 8 && s.stride_jj == 1 && !s.within) ? E.radial_v[0] : h->v2_radial;
        gk<<<gr, 256, h->v2_smem_g, ls>>>(d, s, pb);
      }
      {
        ProfSpan sp(h, MTP_PROF_MOMENTS, ls);
        const int gm = std::max(1, std::min(h->v2_grid_m, (n + 31) / 32));
        E.moments<<<gm, 32 * E.NP, h->v2_smem_m, ls>>>(s, pb, L.mb.p, ld);
      }
      const int gp = std::max(1, std::min(grid_p_cap, (n + na - 1) / na));
      double *part_p = h->d_partials.p + (size_t) rows_used * 8;
      int rows_p = gp;
      {
        ProfSpan sp(h, MTP_PROF_PROGRAM, ls);
        if (p4m) rows_p = launch_p4(s, L.mb.p, L.gb.p, part_p, gp, ls, sc % nlanes);
        else if (p3)
          p3<<<gp, P3_THREADS, smem_p, ls>>>(d, s, h->f3f, h->f3r, L.mb.p, L.gb.p, ld, part_p);
        else if (grade)
          mtp_program_kernel<true><<<gp, PROG_THREADS, smem_p, ls>>>(d, s, L.mb.p, L.gb.p, ld, na, lna, part_p);
        else
          mtp_program_kernel<false><<<gp, PROG_THREADS, smem_p, ls>>>(d, s, L.mb.p, L.gb.p, ld, na, lna, part_p);
      }
      rows_used += rows_p;
      const int ab = h->v2_ab;
      const int gf = std::max(1, std::min(h->v2_grid_fg[gi], (n + ab - 1) / ab));
      double *part_f = h->d_partials.p + (size_t) rows_used * 8;
      {
        ProfSpan sp(h, MTP_PROF_FORCES, ls);
        E.forces[gi]<<<gf, 256, h->v2_smem_f, ls>>>(s, pb, L.gb.p, ld, part_f);
        if (grade) {
          const size_t smem_c = (size_t) 8 * (CAND_TILE * (d.B + d.R) + CAND_TILE / 2) * 8;
          const int gc = std::max(1, std::min(4 * h->sm_count, (n + 7) / 8));
          if (d.S <= 2) mtp_cand_radial_kernel<2><<<gc, 256, smem_c, ls>>>(d, s, pb);
          else if (d.S <= 4)
            mtp_cand_radial_kernel<4><<<gc, 256, smem_c, ls>>>(d, s, pb);
          else
            mtp_cand_radial_kernel<8><<<gc, 256, smem_c, ls>>>(d, s, pb);
          g_launches++;
        }
      }
      rows_used += gf;
      g_launches += 4;
    } else if (pipeline) {
      const V1Entry &E = kV1[h->v1_entry];
      const int gm = std::max(1, std::min(h->pl_grid_m, (n + W - 1) / W));
      {
        ProfSpan sp(h, MTP_PROF_MOMENTS, st);
        E.moments<<<gm, W * 32, h->pl_smem_m, st>>>(d, h->v1tab, s, h->d_mb.p, ld);
      }
      const int gp = std::max(1, std::min(grid_p_cap, (n + na - 1) / na));
      double *part_p = h->d_partials.p + (size_t) rows_used * 8;
      int rows_p = gp;
      {
        ProfSpan sp(h, MTP_PROF_PROGRAM, st);
        if (p4m) rows_p = launch_p4(s, h->d_mb.p, h->d_gb.p, part_p, gp, st, 0);
        else if (p3)
          p3<<<gp, P3_THREADS, smem_p, st>>>(d, s, h->f3f, h->f3r, h->d_mb.p, h->d_gb.p, ld, part_p);
        else if (grade)
          mtp_program_kernel<true><<<gp, PROG_THREADS, smem_p, st>>>(d, s, h->d_mb.p, h->d_gb.p, ld, na, lna, part_p);
        else
          mtp_program_kernel<false><<<gp, PROG_THREADS, smem_p, st>>>(d, s, h->d_mb.p, h->d_gb.p, ld, na, lna, part_p);
      }
      rows_used += rows_p;
      const int gf = std::max(1, std::min(h->pl_grid_f[gi], (n + W - 1) / W));
      double *part_f = h->d_partials.p + (size_t) rows_used * 8;
      {
        ProfSpan sp(h, MTP_PROF_FORCES, st);
        E.forces[gi]<<<gf, W * 32, h->pl_smem_f[gi], st>>>(d, h->v1tab, s, h->d_gb.p, ld, part_f);
      }
      rows_used += gf;
      g_launches += 3;
    } else {
      const int w = h->warps[gi];
      const int grid = std::max(1, std::min(h->grid_cap[gi], (n + w - 1) / w));
      s.partials = h->d_partials.p + (size_t) rows_used * 8;
      ProfSpan sp(h, MTP_PROF_SITE, st);
      if (grade) mtp_site_kernel<true><<<grid, w * 32, h->smem[1], st>>>(d, s, w);
      else
        mtp_site_kernel<false><<<grid, w * 32, h->smem[0], st>>>(d, s, w);
      rows_used += grid;
      g_launches++;
    }
    if (grade && n > 0) {
      ProfSpan sp(h, MTP_PROF_GRADE, st);
      if (cfg) {
        cand_colsum_kernel<<<(d.Q + 127) / 128, 128, 0, st>>>(h->d_cand.p, n, h->qpad, d.Q, h->d_cfg.p, 1);
        g_launches++;
      } else {
        int gb = (n + GRADE_WARPS * 8 - 1) / (GRADE_WARPS * 8);
        const size_t smem_g = (size_t) 2 * GRT_COLS * (h->qpad + 4) * 8;
        if (h->qpad <= 256 && !getenv("MTP_B200_GRADE_SMEM")) {    // A operand of 8 rows fits one warp's registers
          gb = (n + GRT_WARPS * 8 - 1) / (GRT_WARPS * 8);
          h->d_blockmax.ensure((size_t) gb);
          grade_dmma_reg_kernel<64><<<std::min(gb, h->sm_count), GRT_WARPS * 32, smem_g, st>>>(
              h->d_cand.p, n, h->qpad, h->d_ainv.p, a.ilist, first, a.grades ? a.grades : nullptr, h->d_blockmax.p);
        } else {
          h->d_blockmax.ensure((size_t) gb);
          grade_dmma_kernel<<<gb, GRADE_WARPS * 32, 0, st>>>(h->d_cand.p, n, h->qpad, h->d_ainv.p, a.ilist, first,
                                                             a.grades ? a.grades : nullptr, h->d_blockmax.p);
        }
        g_launches++;
        finalize_max_kernel<<<1, 32, 0, st>>>(h->d_blockmax.p, gb, a.ev_out + 7, 1);
        g_launches++;
      }
    }
  }
  if (phases)
    for (int p = chunks.back().phase; p < phases->nphase; p++) phase_done(p);
  if (nlanes > 1)    // join
    for (int l = 0; l < nlanes; l++) {
      CUDA_CHECK(cudaEventRecord(h->lanes[l].done, h->lanes[l].stream));
      CUDA_CHECK(cudaStreamWaitEvent(st, h->lanes[l].done, 0));
    }
  {
    ProfSpan sp(h, MTP_PROF_FINALIZE, st);
    finalize_ev_kernel<<<1, 7 * 32, 0, st>>>(h->d_partials.p, rows_used, a.ev_out, 1);
    g_launches++;
  }
  if (cfg) {
    const long long nat = a.natoms_total > 0 ? a.natoms_total : a.inum;
    cfg_grade_kernel<<<1, 256, 0, st>>>(h->d_ainv.p, h->qpad, d.Q, h->d_cfg.p, nat > 0 ? 1.0 / (double) nat : 0.0,
                                        a.ev_out + 7);
    g_launches++;
    if (a.cfg_candidate)
      CUDA_CHECK(cudaMemcpyAsync(a.cfg_candidate, h->d_cfg.p, sizeof(double) * d.Q, cudaMemcpyDeviceToDevice, st));
  }
  CUDA_CHECK(cudaGetLastError());
}

int fail(int code, const std::string &msg)
{
  g_last_error = msg;
  return code;
}

template <typename F> int guarded(F &&fn)
{
  try {
    fn();
    return MTP_OK;
  } catch (const CudaError &e) {
    return fail(MTP_ERR_CUDA, e.what());
  } catch (const std::invalid_argument &e) {
    return fail(MTP_ERR_ARG, e.what());
  } catch (const std::exception &e) {
    return fail(MTP_ERR_FILE, e.what());
  }
}

mtp_handle *finish_create(mtp_handle *h, int device)
{
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    throw CudaError("no CUDA device available: the MTP B200 path has no CPU fallback");
  if (device < 0) CUDA_CHECK(cudaGetDevice(&device));
  h->device = device;
  set_device(h);
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    throw CudaError(std::string("device '") + prop.name + "' is not sm_100-class; this library is built for sm_100a only");
  upload_potential(h);
  return h;
}

}    // namespace

extern "C" {

const char *mtp_last_error(void) { return g_last_error.c_str(); }

long long mtp_kernel_launch_count(void) { return g_launches.load(); }

int mtp_program_check(const char *path, int atoms_per_cta, double *max_rel_err_out)
{
  if (!path || !max_rel_err_out) return fail(MTP_ERR_ARG, "null argument");
  if (atoms_per_cta != 32 && atoms_per_cta != 16) return fail(MTP_ERR_ARG, "atoms_per_cta must be 32 or 16");
  return guarded([&] {
    Potential p;
    Program prog;
    parse_almtp(path, false, p);
    compile_program(p, prog, 32, 8, atoms_per_cta);
    *max_rel_err_out = check_grouped_streams(p, prog);
  });
}

int mtp_last_kernel_path(const mtp_handle *h) { return h ? h->last_path : -1; }

const char *mtp_program_kernel_note(const mtp_handle *h) { return h ? h->p4_note.c_str() : ""; }
const char *mtp_program_kernel_note_small(const mtp_handle *h) { return h ? h->p4_note_small.c_str() : ""; }

namespace {
// host-only: the generator inputs mtp_create would use for the potential at `path`
void p4_host_plan(const char *path, int latency_shape, Potential &p, P4Choice &ch, std::vector<short> &slot_of_k, int &nslots)
{
  parse_almtp(path, false, p);
  Program prog;
  compile_program(p, prog);    // validates the table (topological order)
  std::vector<short> s2k;
  const int e = v2_entry_for(p, s2k);
  p4_slot_map(p, e, s2k, slot_of_k, nslots);
  ch = p4_choose(p, kSm100SmemOptin, latency_shape != 0);
  if (!ch.ok) throw std::runtime_error("potential structure or size outside the generator's range");
}
}    // namespace

int mtp_codegen_source(const char *path, int latency_shape, char *buf, long long cap, long long *needed, long long *info_out)
{
  if (!path || !needed) return fail(MTP_ERR_ARG, "null argument");
  return guarded([&] {
    Potential p;
    P4Choice ch;
    std::vector<short> slot_of_k;
    int nslots = 0;
    p4_host_plan(path, latency_shape, p, ch, slot_of_k, nslots);
    std::string src, why;
    P4Info info;
    if (!p4_generate(p, ch.prm, slot_of_k.data(), nslots, src, info, why)) throw std::runtime_error(why);
    *needed = (long long) src.size() + 1;
    if (buf && cap >= *needed) memcpy(buf, src.c_str(), src.size() + 1);
    if (info_out) {
      const long long v[13] = {ch.atoms_per_cta(), ch.prm.warps, ch.min_blocks, info.rows, info.stages, (long long) info.smem_bytes, info.terms,
                               info.loads, info.stores, info.crit_terms, nslots, (long long) info.hash, info.rounds};
      memcpy(info_out, v, sizeof(v));
    }
  });
}

int mtp_codegen_prebuild(const char *path, int latency_shape, int *compiled_out)
{
  if (!path) return fail(MTP_ERR_ARG, "null argument");
  return guarded([&] {
    Potential p;
    P4Choice ch;
    std::vector<short> slot_of_k;
    int nslots = 0;
    p4_host_plan(path, latency_shape, p, ch, slot_of_k, nslots);
    P4Info info;
    bool compiled = false;
    p4_cubin(p, ch, slot_of_k.data(), nslots, info, &compiled);
    if (compiled_out) *compiled_out = compiled ? 1 : 0;
  });
}

/* diagnostic builds only (-DMTP_PHASE_CLOCKS): per-phase SM clocks summed over warps, then reset */
int mtp_debug_phase_clocks(unsigned long long *out8)
{
#ifdef MTP_PHASE_CLOCKS
  unsigned long long z[8] = {0};
  if (cudaMemcpyFromSymbol(out8, g_phase_clocks, sizeof(z)) != cudaSuccess) return MTP_ERR_CUDA;
  if (cudaMemcpyToSymbol(g_phase_clocks, z, sizeof(z)) != cudaSuccess) return MTP_ERR_CUDA;
  return MTP_OK;
#else
  (void) out8;
  return MTP_ERR_ARG;
#endif
}

mtp_handle *mtp_create_from_file(const char *path, int want_selection_state, int device)
{
  if (!path) {
    fail(MTP_ERR_ARG, "null path");
    return nullptr;
  }
  mtp_handle *h = new (std::nothrow) mtp_handle;
  if (!h) return nullptr;
  int rc = guarded([&] {
    parse_almtp(path, want_selection_state != 0, h->pot);
    finish_create(h, device);
  });
  if (rc != MTP_OK) {
    delete h;
    return nullptr;
  }
  return h;
}

int mtp_potential_check(const char *path, int want_selection_state, mtp_info *o)
{
  if (!path) return fail(MTP_ERR_ARG, "null path");
  return guarded([&] {
    Potential p;
    Program prog;
    parse_almtp(path, want_selection_state != 0, p);
    compile_program(p, prog);
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->abi_version = MTP_B200_ABI_VERSION;
    o->species_count = p.species_count;
    o->radial_func_count = p.radial_func_count;
    o->radial_basis_size = p.radial_basis_size;
    o->alpha_moment_count = p.alpha_moment_count;
    o->alpha_index_basic_count = p.alpha_index_basic_count;
    o->alpha_index_times_count = p.alpha_index_times_count;
    o->alpha_scalar_count = p.alpha_scalar_count;
    o->max_alpha_index_basic = p.max_alpha_index_basic;
    o->coeff_count = p.has_selection_state ? p.coeff_count : 0;
    o->configuration_mode = p.configuration_mode;
    o->has_selection_state = p.has_selection_state ? 1 : 0;
    o->wave_count = prog.depth;
    o->chunksize = 0;
    o->device = -1;
    o->min_cutoff = p.min_cutoff;
    o->max_cutoff = p.max_cutoff;
    o->scaling = p.scaling;
  });
}

mtp_handle *mtp_create(const mtp_params_host *q, int device)
{
  if (!q) {
    fail(MTP_ERR_ARG, "null params");
    return nullptr;
  }
  mtp_handle *h = new (std::nothrow) mtp_handle;
  if (!h) return nullptr;
  int rc = guarded([&] {
    Potential &p = h->pot;
    p.species_count = q->species_count;
    p.radial_func_count = q->radial_func_count;
    p.radial_basis_size = q->radial_basis_size;
    p.alpha_moment_count = q->alpha_moment_count;
    p.alpha_index_basic_count = q->alpha_index_basic_count;
    p.alpha_index_times_count = q->alpha_index_times_count;
    p.alpha_scalar_count = q->alpha_scalar_count;
    p.min_cutoff = q->min_cutoff;
    p.max_cutoff = q->max_cutoff;
    p.scaling = q->scaling;
    const size_t S = (size_t) p.species_count, nr = S * S * p.radial_func_count * p.radial_basis_size;
    if (p.species_count < 1 || p.radial_func_count < 1 || p.radial_basis_size < 1 || p.alpha_index_basic_count < 1 ||
        p.alpha_index_times_count < 0 || p.alpha_scalar_count < 1)
      throw std::invalid_argument("bad table sizes");
    p.radial_basis_coeffs.assign(q->radial_basis_coeffs, q->radial_basis_coeffs + nr);
    p.alpha_index_basic.assign(q->alpha_index_basic, q->alpha_index_basic + 4 * (size_t) p.alpha_index_basic_count);
    p.alpha_index_times.assign(q->alpha_index_times, q->alpha_index_times + 4 * (size_t) p.alpha_index_times_count);
    p.alpha_moment_mapping.assign(q->alpha_moment_mapping, q->alpha_moment_mapping + p.alpha_scalar_count);
    p.species_coeffs.assign(q->species_coeffs, q->species_coeffs + S);
    p.linear_coeffs.assign(q->linear_coeffs, q->linear_coeffs + p.alpha_scalar_count);
    finalize_tables(p);
    if (q->inverse_active_set) {
      const size_t Q = (size_t) p.coeff_count;
      p.inverse_active_set.assign(q->inverse_active_set, q->inverse_active_set + Q * Q);
      p.configuration_mode = q->configuration_mode;
      p.has_selection_state = true;
    }
    finish_create(h, device);
  });
  if (rc != MTP_OK) {
    delete h;
    return nullptr;
  }
  return h;
}

void mtp_destroy(mtp_handle *h)
{
  if (!h) return;
  cudaSetDevice(h->device);
  h->p4[0].unload();
  h->p4[1].unload();
  if (h->hstream) cudaStreamDestroy(h->hstream);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (auto &L : h->lanes) {
    if (L.stream) cudaStreamDestroy(L.stream);
    if (L.done) cudaEventDestroy(L.done);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->phase_stream) cudaStreamDestroy(h->phase_stream);
  for (cudaEvent_t e : h->phase_events) cudaEventDestroy(e);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  if (h->ev_f0) cudaEventDestroy(h->ev_f0);
  for (cudaEvent_t e : h->copy_events) cudaEventDestroy(e);
  delete h;
}

int mtp_get_info(const mtp_handle *h, mtp_info *o)
{
  if (!h || !o) return fail(MTP_ERR_ARG, "null argument");
  const Potential &p = h->pot;
  o->abi_version = MTP_B200_ABI_VERSION;
  o->species_count = p.species_count;
  o->radial_func_count = p.radial_func_count;
  o->radial_basis_size = p.radial_basis_size;
  o->alpha_moment_count = p.alpha_moment_count;
  o->alpha_index_basic_count = p.alpha_index_basic_count;
  o->alpha_index_times_count = p.alpha_index_times_count;
  o->alpha_scalar_count = p.alpha_scalar_count;
  o->max_alpha_index_basic = p.max_alpha_index_basic;
  o->coeff_count = p.has_selection_state ? p.coeff_count : 0;
  o->configuration_mode = p.configuration_mode;
  o->has_selection_state = p.has_selection_state ? 1 : 0;
  o->wave_count = h->prog.depth;
  o->chunksize = h->chunksize;
  o->device = h->device;
  o->min_cutoff = p.min_cutoff;
  o->max_cutoff = p.max_cutoff;
  o->scaling = p.scaling;
  return MTP_OK;
}

int mtp_get_tables(const mtp_handle *h, double *radial, int *basic, int *times, int *mapping, double *species,
                   double *linear, double *inverse_active_set)
{
  if (!h) return fail(MTP_ERR_ARG, "null handle");
  const Potential &p = h->pot;
  if (radial) memcpy(radial, p.radial_basis_coeffs.data(), p.radial_basis_coeffs.size() * sizeof(double));
  if (basic) memcpy(basic, p.alpha_index_basic.data(), p.alpha_index_basic.size() * sizeof(int));
  if (times && !p.alpha_index_times.empty())
    memcpy(times, p.alpha_index_times.data(), p.alpha_index_times.size() * sizeof(int));
  if (mapping) memcpy(mapping, p.alpha_moment_mapping.data(), p.alpha_moment_mapping.size() * sizeof(int));
  if (species) memcpy(species, p.species_coeffs.data(), p.species_coeffs.size() * sizeof(double));
  if (linear) memcpy(linear, p.linear_coeffs.data(), p.linear_coeffs.size() * sizeof(double));
  if (inverse_active_set) {
    if (!p.has_selection_state) return fail(MTP_ERR_MODE, "no selection state loaded");
    memcpy(inverse_active_set, p.inverse_active_set.data(), p.inverse_active_set.size() * sizeof(double));
  }
  return MTP_OK;
}

int mtp_set_chunksize(mtp_handle *h, int chunksize)
{
  if (!h || chunksize < 1) return fail(MTP_ERR_ARG, "chunksize must be >= 1");
  h->chunksize = chunksize;
  return MTP_OK;
}

int mtp_profile_enable(mtp_handle *h, int on)
{
  if (!h) return fail(MTP_ERR_ARG, "null handle");
  h->profile = on != 0;
  return MTP_OK;
}

int mtp_profile_read(mtp_handle *h, double *ms, long long *count)
{
  if (!h || !ms || !count) return fail(MTP_ERR_ARG, "null argument");
  return guarded([&] {
    set_device(h);
    CUDA_CHECK(cudaDeviceSynchronize());
    for (int c = 0; c < MTP_PROF_CLASSES; c++) {
      ms[c] = 0.0;
      count[c] = 0;
    }
    for (const auto &sp : h->spans) {
      float t = 0;
      CUDA_CHECK(cudaEventElapsedTime(&t, h->ev_pool[sp.e0], h->ev_pool[sp.e1]));
      ms[sp.cls] += t;
      count[sp.cls]++;
    }
    h->spans.clear();
    h->ev_used = 0;
  });
}

int mtp_set_lanes(mtp_handle *h, int lanes)
{
  if (!h || lanes < 1 || lanes > mtp_handle::kMaxLanes) return fail(MTP_ERR_ARG, "lanes must be 1..4");
  h->nlanes = lanes;
  return MTP_OK;
}

int mtp_compute(mtp_handle *h, const mtp_compute_args *a)
{
  if (!h || !a) return fail(MTP_ERR_ARG, "null argument");
  if (!a->x || !a->type || !a->numneigh || !a->neighbors || !a->f || !a->ev_out)
    return fail(MTP_ERR_ARG, "x, type, numneigh, neighbors, f and ev_out are required");
  if (a->variant != MTP_VARIANT_LARGE && a->variant != MTP_VARIANT_SMALL) return fail(MTP_ERR_ARG, "unknown variant");
  return guarded([&] {
    set_device(h);
    launch_site(h, *a, (cudaStream_t) a->stream);
  });
}

int mtp_compute_phased(mtp_handle *h, const mtp_compute_args *a, int nphase, const int *phase_inum, void *const *wait_events,
                       void *const *done_events)
{
  if (!h || !a || nphase < 1 || !phase_inum) return fail(MTP_ERR_ARG, "null argument");
  if (!a->x || !a->type || !a->numneigh || !a->neighbors || !a->f || !a->ev_out)
    return fail(MTP_ERR_ARG, "x, type, numneigh, neighbors, f and ev_out are required");
  if (a->variant != MTP_VARIANT_LARGE && a->variant != MTP_VARIANT_SMALL) return fail(MTP_ERR_ARG, "unknown variant");
  return guarded([&] {
    set_device(h);
    PhaseSpec ph;
    ph.nphase = nphase;
    ph.inum = phase_inum;
    ph.wait_events = wait_events;
    ph.done_events = done_events;
    launch_site(h, *a, (cudaStream_t) a->stream, nullptr, &ph);
  });
}

int mtp_synchronize(mtp_handle *h)
{
  if (!h) return fail(MTP_ERR_ARG, "null handle");
  int status = 0;
  int rc = guarded([&] {
    set_device(h);
    CUDA_CHECK(cudaDeviceSynchronize());
    CUDA_CHECK(cudaMemcpy(&status, h->d_status.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (status) CUDA_CHECK(cudaMemset(h->d_status.p, 0, sizeof(int)));
  });
  if (rc != MTP_OK) return rc;
  if (status & 1) return fail(MTP_ERR_SPECIES, "Too few species count in the MTP potential!");
  return MTP_OK;
}

static double now_ms()
{
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

int mtp_compute_host(mtp_handle *h, const mtp_compute_args *a, int list_changed)
{
  static const bool timing = getenv("MTP_B200_HOST_TIMING") != nullptr;
  const double t_enter = now_ms();
  double t_up = 0, t_launch = 0, t_sync = 0;
  if (!h || !a) return fail(MTP_ERR_ARG, "null argument");
  if (!a->x || !a->type || !a->numneigh || !a->neighbors || !a->f || !a->ev_out)
    return fail(MTP_ERR_ARG, "x, type, numneigh, neighbors, f and ev_out are required");
  int rc = guarded([&] {
    set_device(h);
    if (!h->hstream) CUDA_CHECK(cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
    cudaStream_t st = h->hstream;
    const size_t nall = (size_t) a->nall, inum = (size_t) a->inum;
    // id range covered by numneigh / offsets: every listed atom
    const bool relist = list_changed || h->h_list_len == 0;
    size_t nid = inum;
    if (a->ilist) {
      if (relist) {
        for (size_t k = 0; k < inum; k++) nid = std::max(nid, (size_t) a->ilist[k] + 1);
        h->h_nid = nid;
      } else
        nid = h->h_nid;
    }
    mtp_compute_args d = *a;
    if (!h->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->ev_f0) CUDA_CHECK(cudaEventCreateWithFlags(&h->ev_f0, cudaEventDisableTiming));
    h->h_x.upload(a->x, 3 * nall, st);
    // types (like the list) only change on re-neighboring steps: atoms migrate / are sorted only then
    if (relist || h->h_type.cap < nall || h->h_ntype != nall) {
      h->h_type.upload(a->type, nall, st);
      h->h_ntype = nall;
    }
    // forces accumulate into the caller's f: the kernels add into a zeroed device array, the caller's values arrive
    // on the copy stream while they run and are added at the end (keeps 24 B/atom of H2D off the critical path);
    // f_overwrite: the caller's f is known to be zero, nothing to upload or add
    const bool add_f = !a->f_overwrite;
    h->h_f.ensure(3 * nall);
    if (add_f) h->h_f0.ensure(3 * nall);
    CUDA_CHECK(cudaMemsetAsync(h->h_f.p, 0, sizeof(double) * 3 * nall, st));
    std::vector<cudaEvent_t> ready;
    if (relist) {
      // one pass over the listed centres: extent of the list, max row length, and whether the CSR rows are laid out
      // in ilist order (then the list can be uploaded slice by slice, overlapped with the compute of earlier chunks)
      long long len = 0, prev_end = 0;
      int mx = 0;
      bool ordered = a->neigh_offsets != nullptr && a->stride_jj <= 1;
      if (a->neigh_offsets && a->stride_jj <= 1) {    // CSR: tight single pass
        const int *il = a->ilist, *nnp = a->numneigh;
        const long long *off = a->neigh_offsets;
        int notordered = 0;
        for (size_t k = 0; k < inum; k++) {
          const size_t i = il ? (size_t) il[k] : k;
          const int nn = nnp[i];
          const long long o = off[i];
          mx = nn > mx ? nn : mx;
          notordered |= (o < prev_end);
          prev_end = o + nn;
          len = prev_end > len ? prev_end : len;
        }
        ordered = !notordered;
      } else {
        ordered = false;
        for (size_t k = 0; k < inum; k++) {
          const size_t i = a->ilist ? (size_t) a->ilist[k] : k;
          const int nn = a->numneigh[i];
          mx = std::max(mx, nn);
          if (a->neigh_offsets)
            len = std::max(len, a->neigh_offsets[i] + (long long) nn * std::max(1LL, a->stride_jj));
          else
            len = std::max(len, (long long) i * a->stride_i + (long long) nn * a->stride_jj + 1);
        }
      }
      h->h_maxnn = mx;
      if (a->neigh_offsets) h->h_offsets.upload(a->neigh_offsets, nid, st);
      h->h_numneigh.upload(a->numneigh, nid, st);
      if (a->ilist) h->h_ilist.upload(a->ilist, inum, st);
      h->h_neigh.ensure((size_t) len);
      h->h_list_len = len;
      const int chunk = plan_chunk(h, a->inum, a->want_grade != 0);
      const int nsuper = a->inum > 0 ? (a->inum + chunk - 1) / chunk : 1;
      if (ordered && nsuper > 1 && !a->within_cutoff) {
        while ((int) h->copy_events.size() < nsuper) {
          cudaEvent_t e;
          CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
          h->copy_events.push_back(e);
        }
        for (int c = 0; c < nsuper; c++) {
          const size_t k0 = (size_t) c * chunk, k1 = std::min(inum, k0 + (size_t) chunk) - 1;
          const size_t i0 = a->ilist ? (size_t) a->ilist[k0] : k0, i1 = a->ilist ? (size_t) a->ilist[k1] : k1;
          const long long lo = a->neigh_offsets[i0], hi = a->neigh_offsets[i1] + a->numneigh[i1];
          if (hi > lo)
            CUDA_CHECK(cudaMemcpyAsync(h->h_neigh.p + lo, a->neighbors + lo, sizeof(int) * (size_t) (hi - lo),
                                       cudaMemcpyHostToDevice, h->copy_stream));
          CUDA_CHECK(cudaEventRecord(h->copy_events[c], h->copy_stream));
          ready.push_back(h->copy_events[c]);
        }
      } else if (len > 0)
        CUDA_CHECK(cudaMemcpyAsync(h->h_neigh.p, a->neighbors, sizeof(int) * (size_t) len, cudaMemcpyHostToDevice, st));
    }
    if (add_f) {
      if (3 * nall) CUDA_CHECK(cudaMemcpyAsync(h->h_f0.p, a->f, sizeof(double) * 3 * nall, cudaMemcpyHostToDevice, h->copy_stream));
      CUDA_CHECK(cudaEventRecord(h->ev_f0, h->copy_stream));
    }
    d.max_numneigh = h->h_maxnn;
    d.x = h->h_x.p;
    d.type = h->h_type.p;
    d.f = h->h_f.p;
    d.ilist = a->ilist ? h->h_ilist.p : nullptr;
    d.numneigh = h->h_numneigh.p;
    d.neighbors = h->h_neigh.p;
    d.neigh_offsets = a->neigh_offsets ? h->h_offsets.p : nullptr;
    h->h_ev.ensure(8);
    d.ev_out = h->h_ev.p;
    d.eatom = nullptr;
    d.vatom = nullptr;
    d.grades = nullptr;
    d.cfg_candidate = nullptr;
    d.within_cutoff = nullptr;
    if ((a->eflag & 2) && a->eatom) {
      h->h_eatom.upload(a->eatom, nall, st);
      d.eatom = h->h_eatom.p;
    }
    if ((a->vflag & 4) && a->vatom) {
      h->h_vatom.upload(a->vatom, 6 * nall, st);
      d.vatom = h->h_vatom.p;
    }
    if (a->want_grade && h->pot.has_selection_state && !h->pot.configuration_mode) {
      // neighbourhood grades stay resident on the device (mtp_fetch_grades / mtp_select_grades_host read them later);
      // they cross PCIe only when the caller hands in an array
      const bool fresh = h->h_grades.cap < nall || h->h_grades_n != nall;
      if (a->grades) h->h_grades.upload(a->grades, nall, st);
      else {
        h->h_grades.ensure(nall);
        if (fresh) CUDA_CHECK(cudaMemsetAsync(h->h_grades.p, 0, sizeof(double) * nall, st));
      }
      h->h_grades_n = nall;
      d.grades = h->h_grades.p;
    }
    if (a->want_grade && a->cfg_candidate && h->pot.has_selection_state) {
      h->h_cfgc.ensure((size_t) h->pot.coeff_count);
      d.cfg_candidate = h->h_cfgc.p;
    }
    if (a->within_cutoff) {
      h->h_within.ensure((size_t) h->h_list_len);
      CUDA_CHECK(cudaMemsetAsync(h->h_within.p, 0, (size_t) h->h_list_len, st));
      d.within_cutoff = h->h_within.p;
    }
    d.stream = st;
    t_up = now_ms();
    launch_site(h, d, st, ready.empty() ? nullptr : &ready);
    t_launch = now_ms();
    if (add_f) CUDA_CHECK(cudaStreamWaitEvent(st, h->ev_f0, 0));
    if (nall && add_f) {
      add_inplace_kernel<<<std::min<size_t>(4 * h->sm_count, (3 * nall + 255) / 256), 256, 0, st>>>(h->h_f.p, h->h_f0.p, 3 * nall);
      g_launches++;
    }
    CUDA_CHECK(cudaMemcpyAsync(a->f, d.f, sizeof(double) * 3 * nall, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(a->ev_out, d.ev_out, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
    if (d.eatom) CUDA_CHECK(cudaMemcpyAsync(a->eatom, d.eatom, sizeof(double) * nall, cudaMemcpyDeviceToHost, st));
    if (d.vatom) CUDA_CHECK(cudaMemcpyAsync(a->vatom, d.vatom, sizeof(double) * 6 * nall, cudaMemcpyDeviceToHost, st));
    if (d.grades && a->grades) CUDA_CHECK(cudaMemcpyAsync(a->grades, d.grades, sizeof(double) * nall, cudaMemcpyDeviceToHost, st));
    if (d.cfg_candidate)
      CUDA_CHECK(cudaMemcpyAsync(a->cfg_candidate, d.cfg_candidate, sizeof(double) * h->pot.coeff_count,
                                 cudaMemcpyDeviceToHost, st));
    if (d.within_cutoff)
      CUDA_CHECK(cudaMemcpyAsync(a->within_cutoff, d.within_cutoff, (size_t) h->h_list_len, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    t_sync = now_ms();
  });
  if (timing)
    fprintf(stderr, "[mtp host] uploads enqueued %.3f ms, kernels enqueued %.3f ms, stream drained %.3f ms\n", t_up - t_enter,
            t_launch - t_up, t_sync - t_launch);
  if (rc != MTP_OK) return rc;
  int status = 0;
  rc = guarded([&] {
    CUDA_CHECK(cudaMemcpy(&status, h->d_status.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (status) CUDA_CHECK(cudaMemset(h->d_status.p, 0, sizeof(int)));
  });
  if (rc != MTP_OK) return rc;
  if (status & 1) return fail(MTP_ERR_SPECIES, "Too few species count in the MTP potential!");
  return MTP_OK;
}

int mtp_neigh_build(mtp_handle *h, int nlocal, int nall, const double *x, double cutneigh, int *numneigh, int *neighbors,
                    int width, int *max_numneigh_out, void *stream)
{
  if (!h) return fail(MTP_ERR_ARG, "null handle");
  if (nlocal < 0 || nall < nlocal || width < 1 || !(cutneigh > 0.0)) return fail(MTP_ERR_ARG, "bad neighbor-build arguments");
  if (nlocal > 0 && (!x || !numneigh || !neighbors)) return fail(MTP_ERR_ARG, "null neighbor-build buffer");
  if (max_numneigh_out) *max_numneigh_out = 0;
  if (nlocal == 0) return MTP_OK;
  int maxnn = 0;
  const int rc = guarded([&] {
    CUDA_CHECK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t) stream;
    // bounding box of owned + ghost atoms
    const int nbb = std::max(1, std::min(4 * h->sm_count, (nall + 255) / 256));
    h->nb_part.ensure((size_t) nbb * 6);
    h->nb_bounds.ensure(6);
    neigh_bounds_kernel<<<nbb, 256, 0, st>>>(nall, x, h->nb_part.p);
    neigh_bounds_final_kernel<<<1, 32, 0, st>>>(nbb, h->nb_part.p, h->nb_bounds.p);
    double bb[6];
    CUDA_CHECK(cudaMemcpyAsync(bb, h->nb_bounds.p, sizeof(bb), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    // bins no smaller than cutneigh (27-bin stencil), and no more bins than a few per atom
    NeighGrid g;
    long long ncell = 1;
    for (int a = 0; a < 3; a++) {
      const double ext = std::max(bb[3 + a] - bb[a], 0.0);
      long long n = (long long) std::floor(ext / cutneigh);
      if (n > 1 && ext / (double) n < cutneigh * (1.0 + 1e-9)) n--;
      n = std::max(1LL, std::min(n, 1LL << 20));
      g.n[a] = (int) n;
      ncell *= n;
    }
    const long long cell_cap = std::max(4096LL, 4LL * nall);
    while (ncell > cell_cap) {    // sparse systems: coarsen the longest axis
      int a = g.n[0] >= g.n[1] && g.n[0] >= g.n[2] ? 0 : (g.n[1] >= g.n[2] ? 1 : 2);
      ncell /= g.n[a];
      g.n[a] = (g.n[a] + 1) / 2;
      ncell *= g.n[a];
    }
    for (int a = 0; a < 3; a++) {
      const double ext = std::max(bb[3 + a] - bb[a], 0.0);
      g.lo[a] = bb[a];
      g.inv[a] = ext > 0.0 ? (double) g.n[a] / ext : 0.0;
    }
    h->nb_keys.ensure(nall);
    h->nb_idx.ensure(nall);
    h->nb_skeys.ensure(nall);
    h->nb_sidx.ensure(nall);
    h->nb_cell.ensure((size_t) ncell + 1);
    h->nb_xs.ensure(nall);
    h->nb_max.ensure(1);
    neigh_bin_kernel<<<(nall + 255) / 256, 256, 0, st>>>(nall, x, g, h->nb_keys.p, h->nb_idx.p);
    int end_bit = 1;
    while ((1LL << end_bit) < ncell && end_bit < 31) end_bit++;
    size_t tmp_bytes = 0;
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, h->nb_keys.p, h->nb_skeys.p, h->nb_idx.p, h->nb_sidx.p, nall, 0,
                                               end_bit, st));
    h->nb_tmp.ensure(tmp_bytes);
    CUDA_CHECK(cub::DeviceRadixSort::SortPairs(h->nb_tmp.p, tmp_bytes, h->nb_keys.p, h->nb_skeys.p, h->nb_idx.p, h->nb_sidx.p, nall,
                                               0, end_bit, st));
    neigh_cellstart_kernel<<<(int) ((ncell + 1 + 255) / 256), 256, 0, st>>>(nall, h->nb_skeys.p, (int) ncell, h->nb_cell.p);
    neigh_gather_sorted_kernel<<<(nall + 255) / 256, 256, 0, st>>>(nall, x, h->nb_sidx.p, h->nb_xs.p);
    CUDA_CHECK(cudaMemsetAsync(h->nb_max.p, 0, sizeof(int), st));
    const int gb = std::max(1, std::min(16 * h->sm_count, (nlocal + 7) / 8));
    neigh_build_kernel<<<gb, 256, 0, st>>>(nlocal, x, g, h->nb_cell.p, h->nb_xs.p, cutneigh * cutneigh, width, numneigh, neighbors,
                                           h->nb_max.p);
    g_launches += 7;
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(&maxnn, h->nb_max.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
  if (rc != MTP_OK) return rc;
  if (max_numneigh_out) *max_numneigh_out = maxnn;
  if (maxnn > width)
    return fail(MTP_ERR_CAPACITY, "neighbor table too narrow: the longest row has " + std::to_string(maxnn) + " entries, width is " +
                                      std::to_string(width));
  return MTP_OK;
}

namespace {
struct GradeAtLeast {
  const double *grades;
  double threshold;
  __host__ __device__ bool operator()(const int &i) const { return grades[i] >= threshold; }
};
}    // namespace

int mtp_select_grades(mtp_handle *h, const double *grades, int n, double threshold, int *indices_out, int *count_out, void *stream)
{
  if (!h || !count_out) return fail(MTP_ERR_ARG, "null argument");
  *count_out = 0;
  if (n < 0 || (n > 0 && (!grades || !indices_out))) return fail(MTP_ERR_ARG, "bad selection arguments");
  if (n == 0) return MTP_OK;
  return guarded([&] {
    CUDA_CHECK(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t) stream;
    h->nb_max.ensure(1);
    cub::CountingInputIterator<int> ids(0);
    GradeAtLeast pred{grades, threshold};
    size_t tmp_bytes = 0;
    CUDA_CHECK(cub::DeviceSelect::If(nullptr, tmp_bytes, ids, indices_out, h->nb_max.p, n, pred, st));
    h->nb_tmp.ensure(tmp_bytes);
    CUDA_CHECK(cub::DeviceSelect::If(h->nb_tmp.p, tmp_bytes, ids, indices_out, h->nb_max.p, n, pred, st));
    g_launches += 2;
    CUDA_CHECK(cudaMemcpyAsync(count_out, h->nb_max.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
}

int mtp_fetch_grades(mtp_handle *h, double *grades_host, int n)
{
  if (!h || (n > 0 && !grades_host)) return fail(MTP_ERR_ARG, "null argument");
  if ((size_t) n > h->h_grades_n) return fail(MTP_ERR_MODE, "no neighbourhood grades of that size are resident (run a grade step through mtp_compute_host first)");
  if (n <= 0) return MTP_OK;
  return guarded([&] {
    set_device(h);
    CUDA_CHECK(cudaMemcpy(grades_host, h->h_grades.p, sizeof(double) * (size_t) n, cudaMemcpyDeviceToHost));
  });
}

int mtp_select_grades_host(mtp_handle *h, int n, double threshold, int *ids_out, double *grades_out, int cap, int *count_out)
{
  if (!h || !count_out) return fail(MTP_ERR_ARG, "null argument");
  *count_out = 0;
  if (n < 0 || cap < 0 || (cap > 0 && (!ids_out || !grades_out))) return fail(MTP_ERR_ARG, "bad selection arguments");
  if ((size_t) n > h->h_grades_n) return fail(MTP_ERR_MODE, "no neighbourhood grades of that size are resident (run a grade step through mtp_compute_host first)");
  if (n == 0) return MTP_OK;
  int count = 0;
  int rc = guarded([&] {
    set_device(h);
    if (!h->hstream) CUDA_CHECK(cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
    h->sel_ids.ensure((size_t) n);
  });
  if (rc != MTP_OK) return rc;
  rc = mtp_select_grades(h, h->h_grades.p, n, threshold, h->sel_ids.p, &count, h->hstream);
  if (rc != MTP_OK) return rc;
  *count_out = count;
  const int m = std::min(count, cap);
  if (m == 0) return MTP_OK;
  return guarded([&] {
    h->sel_val.ensure((size_t) m);
    gather_by_index_kernel<<<(m + 255) / 256, 256, 0, h->hstream>>>(h->h_grades.p, h->sel_ids.p, m, h->sel_val.p);
    g_launches++;
    CUDA_CHECK(cudaMemcpyAsync(ids_out, h->sel_ids.p, sizeof(int) * (size_t) m, cudaMemcpyDeviceToHost, h->hstream));
    CUDA_CHECK(cudaMemcpyAsync(grades_out, h->sel_val.p, sizeof(double) * (size_t) m, cudaMemcpyDeviceToHost, h->hstream));
    CUDA_CHECK(cudaStreamSynchronize(h->hstream));
  });
}

int mtp_cfg_grade(mtp_handle *h, const double *candidate_host, long long natoms_total, double *grade_out)
{
  if (!h || !candidate_host || !grade_out) return fail(MTP_ERR_ARG, "null argument");
  if (!h->pot.has_selection_state) return fail(MTP_ERR_MODE, "no selection state loaded");
  return guarded([&] {
    set_device(h);
    const int Q = h->pot.coeff_count;
    h->d_cfg.ensure((size_t) h->qpad);
    h->h_ev.ensure(8);
    CUDA_CHECK(cudaMemcpy(h->d_cfg.p, candidate_host, sizeof(double) * Q, cudaMemcpyHostToDevice));
    cfg_grade_kernel<<<1, 256>>>(h->d_ainv.p, h->qpad, Q, h->d_cfg.p, natoms_total > 0 ? 1.0 / (double) natoms_total : 0.0,
                                 h->h_ev.p + 7);
    g_launches++;
    CUDA_CHECK(cudaMemcpy(grade_out, h->h_ev.p + 7, sizeof(double), cudaMemcpyDeviceToHost));
  });
}

void *mtp_alloc_pinned(size_t bytes)
{
  void *p = nullptr;
  if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
    fail(MTP_ERR_CUDA, "cudaMallocHost failed");
    return nullptr;
  }
  return p;
}

void mtp_free_pinned(void *p)
{
  if (p) cudaFreeHost(p);
}

int mtp_nve_initial_integrate(int nlocal, double *x, double *v, const double *f, const int *type, const double *mass,
                              double dtf, double dtv, const double *x_at_build, double trigger_dist, int *moved_flag,
                              void *stream)
{
  if (nlocal < 0 || (nlocal > 0 && (!x || !v || !f || !type || !mass))) return fail(MTP_ERR_ARG, "bad integrator arguments");
  if ((x_at_build != nullptr) != (moved_flag != nullptr)) return fail(MTP_ERR_ARG, "x_at_build and moved_flag go together");
  if (nlocal == 0) return MTP_OK;
  return guarded([&] {
    nve_initial_kernel<<<(nlocal + 255) / 256, 256, 0, (cudaStream_t) stream>>>(nlocal, x, v, f, type, mass, dtf, dtv, x_at_build,
                                                                                 trigger_dist * trigger_dist, moved_flag);
    g_launches++;
    CUDA_CHECK(cudaGetLastError());
  });
}

int mtp_nve_final_integrate(int nlocal, double *v, const double *f, const int *type, const double *mass, double dtf, void *stream)
{
  if (nlocal < 0 || (nlocal > 0 && (!v || !f || !type || !mass))) return fail(MTP_ERR_ARG, "bad integrator arguments");
  if (nlocal == 0) return MTP_OK;
  return guarded([&] {
    nve_final_kernel<<<(nlocal + 255) / 256, 256, 0, (cudaStream_t) stream>>>(nlocal, v, f, type, mass, dtf);
    g_launches++;
    CUDA_CHECK(cudaGetLastError());
  });
}

int mtp_halo_pack_x(const double *x, const int *sendlist, int n, const double *shift, double *out, void *stream)
{
  if (n < 0 || (n > 0 && (!x || !sendlist || !out))) return fail(MTP_ERR_ARG, "bad halo arguments");
  if (n == 0) return MTP_OK;
  const double sx = shift ? shift[0] : 0.0, sy = shift ? shift[1] : 0.0, sz = shift ? shift[2] : 0.0;
  return guarded([&] {
    halo_pack_x_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>>(x, sendlist, n, sx, sy, sz, out);
    g_launches++;
    CUDA_CHECK(cudaGetLastError());
  });
}

int mtp_halo_pack_x_multi(const double *x, const int *sendlist, const unsigned char *seg, const double *shifts_dev, int n,
                          double *out, void *stream)
{
  if (n < 0 || (n > 0 && (!x || !sendlist || !seg || !shifts_dev || !out))) return fail(MTP_ERR_ARG, "bad halo arguments");
  if (n == 0) return MTP_OK;
  return guarded([&] {
    halo_pack_x_multi_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>>(x, sendlist, seg, shifts_dev, n, out);
    g_launches++;
    CUDA_CHECK(cudaGetLastError());
  });
}

int mtp_halo_unpack_add_f(double *f, const int *sendlist, int n, const double *buf, void *stream)
{
  if (n < 0 || (n > 0 && (!f || !sendlist || !buf))) return fail(MTP_ERR_ARG, "bad halo arguments");
  if (n == 0) return MTP_OK;
  return guarded([&] {
    halo_unpack_add_f_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t) stream>>>(f, sendlist, n, buf);
    g_launches++;
    CUDA_CHECK(cudaGetLastError());
  });
}

int mtp_fp64_peak(int device, double *dfma_tflops, double *dmma_tflops)
{
  return guarded([&] {
    if (device >= 0) CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_CHECK(cudaGetDevice(&dev));
    CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    double *out = nullptr;
    CUDA_CHECK(cudaMalloc((void **) &out, 8));
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
    for (int which = 0; which < 2; which++) {
      double best = 0.0;
      for (int rep = 0; rep < 4; rep++) {
        CUDA_CHECK(cudaEventRecord(e0));
        if (which == 0) dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        else
          dmma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
        g_launches++;
        CUDA_CHECK(cudaEventRecord(e1));
        CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        // DFMA: 8 fma/thread/iter = 16 flop; DMMA: 8 mma/warp/iter, each 8*8*4*2 = 512 flop
        const double flops = which == 0 ? (double) blocks * threads * iters * 16.0
                                        : (double) blocks * (threads / 32) * iters * 8.0 * 512.0;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
      }
      if (which == 0 && dfma_tflops) *dfma_tflops = best;
      if (which == 1 && dmma_tflops) *dmma_tflops = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
  });
}

}    // extern "C"
