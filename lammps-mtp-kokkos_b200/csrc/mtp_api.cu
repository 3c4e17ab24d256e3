// extern "C" layer of include/mtp_b200.h: handle, uploads, launch configuration, host-buffer path.
#include "../../include/mtp_b200.h"
#include "mtp_kernels.cu"
#include "mtp_potential.hpp"

#include <algorithm>
#include <atomic>
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

struct mtp_handle {
  Potential pot;
  Program prog;
  int device = 0;
  int sm_count = 0;
  int chunksize = 1 << 30;
  int qpad = 0;
  // device copies of the potential
  DevBuf<double> d_radial, d_species, d_lin, d_ginit, d_ainv;
  DevBuf<uint32_t> d_basic;
  DevBuf<int> d_map;
  PassBufs d_fwd, d_rev;
  DevPotential dpot{};
  // work buffers (grow-only)
  DevBuf<AtomRec> d_xt;
  DevBuf<double> d_partials, d_cand, d_blockmax, d_cfg;
  DevBuf<int> d_status;
  // host-buffer path
  DevBuf<double> h_x, h_f, h_eatom, h_vatom, h_grades, h_ev, h_cfgc;
  DevBuf<int> h_type, h_ilist, h_numneigh, h_neigh;
  DevBuf<long long> h_offsets;
  DevBuf<unsigned char> h_within;
  long long h_list_len = 0;
  cudaStream_t hstream = nullptr;
  // launch configuration per kernel flavour
  int warps[2] = {0, 0};
  int grid_cap[2] = {0, 0};
  size_t smem[2] = {0, 0};
};

namespace {

void set_device(const mtp_handle *h) { CUDA_CHECK(cudaSetDevice(h->device)); }

void upload_potential(mtp_handle *h)
{
  Potential &p = h->pot;
  compile_program(p, h->prog);
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

  // launch configuration: as many warps per CTA (<= 8) as shared memory allows, then occupancy
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, h->device));
  h->sm_count = prop.multiProcessorCount;
  const size_t smem_max = prop.sharedMemPerBlockOptin;
  for (int gflag = 0; gflag < 2; gflag++) {
    if (gflag == 1 && !p.has_selection_state) continue;
    const Layout L = make_layout(d.S, d.R, d.B, d.K, d.M, d.P, d.Q, gflag == 1);
    int w = 4;
    while (w > 1 && L.cta_bytes + (size_t) w * L.warp_bytes > smem_max) w--;
    const size_t bytes = L.cta_bytes + (size_t) w * L.warp_bytes;
    if (bytes > smem_max)
      throw std::runtime_error("potential too large for on-chip per-atom state (alpha_moments_count)");
    if (gflag == 0) {
      CUDA_CHECK(cudaFuncSetAttribute(mtp_site_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    } else {
      CUDA_CHECK(cudaFuncSetAttribute(mtp_site_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
    }
    int per_sm = 0;
    if (gflag == 0) {
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mtp_site_kernel<false>, w * 32, bytes));
    } else {
      CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mtp_site_kernel<true>, w * 32, bytes));
    }
    if (per_sm < 1) per_sm = 1;
    h->warps[gflag] = w;
    h->smem[gflag] = bytes;
    h->grid_cap[gflag] = per_sm * h->sm_count;
  }
}

void launch_site(mtp_handle *h, const mtp_compute_args &a, cudaStream_t st)
{
  const DevPotential &d = h->dpot;
  const bool grade = a.want_grade != 0;
  if (grade && !h->pot.has_selection_state)
    throw std::invalid_argument("extrapolation grades requested but the potential has no selection state");
  if (a.inum < 0 || a.nall < a.inum) throw std::invalid_argument("bad inum / nall");

  // pack positions + species into 32-byte records
  h->d_xt.ensure((size_t) (a.nall > 0 ? a.nall : 1));
  if (a.nall > 0) {
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
  const int w = h->warps[gi];
  const bool cfg = grade && h->pot.configuration_mode;
  // chunking only bounds the candidate-row scratch of grade steps (README.md:44 "chunksize")
  const int chunk = grade ? std::max(1, std::min(h->chunksize, a.inum > 0 ? a.inum : 1)) : (a.inum > 0 ? a.inum : 1);
  if (grade) h->d_cand.ensure((size_t) chunk * h->qpad);
  if (cfg) CUDA_CHECK(cudaMemsetAsync(h->d_cfg.p, 0, sizeof(double) * h->qpad, st));
  CUDA_CHECK(cudaMemsetAsync(a.ev_out, 0, sizeof(double) * 8, st));

  int nchunk = 0;
  for (int first = 0; first < a.inum || (first == 0 && nchunk == 0); first += chunk, nchunk++) {
    const int n = std::max(0, std::min(chunk, a.inum - first));
    int grid = std::min(h->grid_cap[gi], (n + w - 1) / w);
    if (grid < 1) grid = 1;
    h->d_partials.ensure((size_t) h->grid_cap[gi] * 8);
    s.inum = n;
    s.first_ii = first;
    s.partials = h->d_partials.p;
    s.cand_rows = grade ? h->d_cand.p : nullptr;
    s.cand_ld = h->qpad;
    if (grade) mtp_site_kernel<true><<<grid, w * 32, h->smem[1], st>>>(d, s, w);
    else
      mtp_site_kernel<false><<<grid, w * 32, h->smem[0], st>>>(d, s, w);
    g_launches++;
    finalize_ev_kernel<<<1, 32, 0, st>>>(h->d_partials.p, grid, a.ev_out, 1);
    g_launches++;
    if (grade && n > 0) {
      if (cfg) {
        cand_colsum_kernel<<<(d.Q + 127) / 128, 128, 0, st>>>(h->d_cand.p, n, h->qpad, d.Q, h->d_cfg.p, 1);
        g_launches++;
      } else {
        const int gb = (n + GRADE_WARPS * 8 - 1) / (GRADE_WARPS * 8);
        h->d_blockmax.ensure((size_t) gb);
        grade_dmma_kernel<<<gb, GRADE_WARPS * 32, 0, st>>>(h->d_cand.p, n, h->qpad, h->d_ainv.p, a.ilist, first,
                                                           a.grades ? a.grades : nullptr, h->d_blockmax.p);
        g_launches++;
        finalize_max_kernel<<<1, 32, 0, st>>>(h->d_blockmax.p, gb, a.ev_out + 7, 1);
        g_launches++;
      }
    }
    if (a.inum == 0) break;
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
  if (h->hstream) cudaStreamDestroy(h->hstream);
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

int mtp_compute_host(mtp_handle *h, const mtp_compute_args *a, int list_changed)
{
  if (!h || !a) return fail(MTP_ERR_ARG, "null argument");
  if (!a->x || !a->type || !a->numneigh || !a->neighbors || !a->f || !a->ev_out)
    return fail(MTP_ERR_ARG, "x, type, numneigh, neighbors, f and ev_out are required");
  int rc = guarded([&] {
    set_device(h);
    if (!h->hstream) CUDA_CHECK(cudaStreamCreateWithFlags(&h->hstream, cudaStreamNonBlocking));
    cudaStream_t st = h->hstream;
    const size_t nall = (size_t) a->nall, inum = (size_t) a->inum;
    // id range covered by numneigh / offsets: every listed atom
    size_t nid = inum;
    if (a->ilist)
      for (size_t k = 0; k < inum; k++) nid = std::max(nid, (size_t) a->ilist[k] + 1);
    mtp_compute_args d = *a;
    h->h_x.upload(a->x, 3 * nall, st);
    h->h_type.upload(a->type, nall, st);
    h->h_f.upload(a->f, 3 * nall, st);
    if (list_changed || h->h_list_len == 0) {
      long long len = 0;
      if (a->neigh_offsets) {
        for (size_t k = 0; k < inum; k++) {
          const size_t i = a->ilist ? (size_t) a->ilist[k] : k;
          len = std::max(len, a->neigh_offsets[i] + (long long) a->numneigh[i] * std::max(1LL, a->stride_jj));
        }
        h->h_offsets.upload(a->neigh_offsets, nid, st);
      } else {
        long long mx = 0;
        for (size_t k = 0; k < inum; k++) {
          const size_t i = a->ilist ? (size_t) a->ilist[k] : k;
          mx = std::max(mx, (long long) i * a->stride_i + (long long) a->numneigh[i] * a->stride_jj);
        }
        len = mx + 1;
      }
      h->h_neigh.upload(a->neighbors, (size_t) len, st);
      h->h_numneigh.upload(a->numneigh, nid, st);
      if (a->ilist) h->h_ilist.upload(a->ilist, inum, st);
      h->h_list_len = len;
    }
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
    if (a->want_grade && a->grades) {
      h->h_grades.upload(a->grades, nall, st);
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
    launch_site(h, d, st);
    CUDA_CHECK(cudaMemcpyAsync(a->f, d.f, sizeof(double) * 3 * nall, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(a->ev_out, d.ev_out, sizeof(double) * 8, cudaMemcpyDeviceToHost, st));
    if (d.eatom) CUDA_CHECK(cudaMemcpyAsync(a->eatom, d.eatom, sizeof(double) * nall, cudaMemcpyDeviceToHost, st));
    if (d.vatom) CUDA_CHECK(cudaMemcpyAsync(a->vatom, d.vatom, sizeof(double) * 6 * nall, cudaMemcpyDeviceToHost, st));
    if (d.grades) CUDA_CHECK(cudaMemcpyAsync(a->grades, d.grades, sizeof(double) * nall, cudaMemcpyDeviceToHost, st));
    if (d.cfg_candidate)
      CUDA_CHECK(cudaMemcpyAsync(a->cfg_candidate, d.cfg_candidate, sizeof(double) * h->pot.coeff_count,
                                 cudaMemcpyDeviceToHost, st));
    if (d.within_cutoff)
      CUDA_CHECK(cudaMemcpyAsync(a->within_cutoff, d.within_cutoff, (size_t) h->h_list_len, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
  });
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
