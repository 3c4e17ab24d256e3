// TEST INFRASTRUCTURE -- plays the role of upstream LAMMPS (oracle/lammps_shim/) around the product's own
// PairStyle classes (lammps-mtp-kokkos_b200/lammps/pair_mtp_b200.cpp): creates the style by its pair_style
// string, calls settings()/coeff()/init_style()/init_one()/compute() in LAMMPS's order on caller-provided
// atoms + full neighbor list, and hands the results back as flat arrays.  Same shape as oracle/ref_driver.cpp,
// so that tests drive the reference's CPU styles and these styles with identical inputs.
// With -DLMP_KOKKOS (second shim library) the same driver plays LAMMPS-KOKKOS instead: atom data and the neighbor list
// are put on the device behind the stub views of tests/shim/kokkos_stub/ (2-D neighbor view in LayoutLeft, as on CUDA),
// and the style's compute_device_views() path runs.
#include "pair_mtp_b200.h"
#ifdef LMP_KOKKOS
#include "atom_kokkos.h"
#include "neigh_list_kokkos.h"
#endif

#include <algorithm>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

using namespace LAMMPS_NS;

namespace {
struct Handle {
  LAMMPS *lmp = nullptr;
#ifdef LMP_KOKKOS
  NeighListKokkos<LMPDeviceType> list;
  std::vector<int> table;
  int *d_ilist = nullptr, *d_numneigh = nullptr, *d_table = nullptr;
#else
  NeighList list;
#endif
  Pair *pair = nullptr;
  std::vector<double *> xrows, frows;
  std::vector<double> xbuf, fbuf;
  std::vector<int> typebuf, ilistbuf, numneighbuf;
  std::vector<int *> firstneigh;
  ~Handle()
  {
    delete pair;
    delete lmp;
  }
};
void set_err(char *err, int errlen, const std::string &m)
{
  if (!err || errlen <= 0) return;
  strncpy(err, m.c_str(), errlen - 1);
  err[errlen - 1] = '\0';
}
}    // namespace

extern "C" {

void *b200drv_create(const char *style, int narg, const char **args, int species, char *err, int errlen)
{
  auto *h = new Handle;
  try {
    h->lmp = new LAMMPS;
#ifdef LMP_KOKKOS
    delete h->lmp->atom;
    h->lmp->atom = new AtomKokkos;
#endif
    std::vector<std::string> keep(args, args + narg);
    std::vector<char *> argv;
    for (auto &s : keep) argv.push_back(s.data());
    const std::string st(style);
    if (st == "mtp/kk" || st == "mtp/kk/device") h->pair = new PairMTPB200Large(h->lmp);
    else if (st == "mtp/small/kk" || st == "mtp/small/kk/device") h->pair = new PairMTPB200Small(h->lmp);
    else if (st == "mtp/extrapolation/kk" || st == "mtp/extrapolation/kk/device")
      h->pair = new PairMTPB200ExtrapolationLarge(h->lmp);
    else if (st == "mtp/extrapolation/small/kk" || st == "mtp/extrapolation/small/kk/device")
      h->pair = new PairMTPB200ExtrapolationSmall(h->lmp);
    else {
      set_err(err, errlen, "Unrecognized pair style '" + st + "'");
      delete h;
      return nullptr;
    }
    h->pair->settings(narg, argv.data());
    char star[] = "*";
    char *cargs[2] = {star, star};
    h->pair->coeff(2, cargs);
    h->pair->init_style();
    h->pair->init_list(0, &h->list);
    for (int i = 1; i <= species; i++)
      for (int j = 1; j <= species; j++) h->pair->init_one(i, j);
    h->lmp->atom->ntypes = species;
  } catch (std::exception &e) {
    set_err(err, errlen, e.what());
    delete h;
    return nullptr;
  }
  return h;
}

void b200drv_destroy(void *hv) { delete (Handle *) hv; }

const char *b200drv_log(void *hv) { return ((Handle *) hv)->lmp->log.c_str(); }

void b200drv_set_domain(void *hv, const double *prd, long natoms)
{
  auto *h = (Handle *) hv;
  Domain *d = h->lmp->domain;
  d->xprd = prd[0];
  d->yprd = prd[1];
  d->zprd = prd[2];
  d->xy = prd[3];
  d->xz = prd[4];
  d->yz = prd[5];
  h->lmp->atom->natoms = natoms;
}

void b200drv_set_newton(void *hv, int newton) { ((Handle *) hv)->lmp->force->newton_pair = newton; }

// lone != 0: the style is the top-level pair style (force->pair), as with a plain `pair_style mtp/kk ...` line;
// 0 (default): it sits below pair_style hybrid, so forces may already hold other contributions
void b200drv_set_lone_pair(void *hv, int lone)
{
  auto *h = (Handle *) hv;
  h->lmp->force->pair = lone ? h->pair : nullptr;
}

int b200drv_is_kokkos(void)
{
#ifdef LMP_KOKKOS
  return 1;
#else
  return 0;
#endif
}

// ago = neighbor->ago (0 on re-neighboring steps).  ev[0] = eng_vdwl, ev[1..6] = virial, ev[7] = pvector[0].
int b200drv_compute(void *hv, int nlocal, int nghost, const double *x, const int *type, int inum, const int *ilist,
                    const int *numneigh, const int *neigh_flat, const long *neigh_offsets, int eflag, int vflag, int ago,
                    int extrapolation_flag, double *f, double *eatom, double *vatom, double *ev, double *grades,
                    char *err, int errlen)
{
  auto *h = (Handle *) hv;
  const int nall = nlocal + nghost;
  try {
    Atom *atom = h->lmp->atom;
    h->xbuf.assign(x, x + 3 * (size_t) nall);
    h->fbuf.assign(f, f + 3 * (size_t) nall);
    h->typebuf.assign(type, type + nall);
    h->xrows.resize(nall);
    h->frows.resize(nall);
    for (int i = 0; i < nall; i++) {
      h->xrows[i] = &h->xbuf[3 * (size_t) i];
      h->frows[i] = &h->fbuf[3 * (size_t) i];
    }
    atom->x = h->xrows.data();
    atom->f = h->frows.data();
    atom->type = h->typebuf.data();
    atom->nlocal = nlocal;
    atom->nghost = nghost;
    atom->nmax = nall;
    if (atom->natoms == 0) atom->natoms = nlocal;
    h->lmp->neighbor->ago = ago;

    h->ilistbuf.assign(ilist, ilist + inum);
    h->numneighbuf.assign(numneigh, numneigh + nall);
    h->firstneigh.assign(nall, nullptr);
    for (int ii = 0; ii < inum; ii++) h->firstneigh[ilist[ii]] = const_cast<int *>(neigh_flat) + neigh_offsets[ilist[ii]];
    h->list.inum = inum;
    h->list.ilist = h->ilistbuf.data();
    h->list.numneigh = h->numneighbuf.data();
    h->list.firstneigh = h->firstneigh.data();

#ifdef LMP_KOKKOS
    // device side of LAMMPS-KOKKOS: AtomKokkos DualViews + NeighListKokkos views (d_neighbors(i, jj) LayoutLeft)
    auto *akk = (AtomKokkos *) atom;
    akk->k_x.allocate(h->xbuf.data(), nall, 3);
    akk->k_f.allocate(h->fbuf.data(), nall, 3);
    akk->k_type.allocate(h->typebuf.data(), nall);
    akk->k_x.modify<LMPHostType>();
    akk->k_f.modify<LMPHostType>();
    akk->k_type.modify<LMPHostType>();
    int maxn = 1;
    for (int ii = 0; ii < inum; ii++) maxn = std::max(maxn, numneigh[ilist[ii]]);
    h->table.assign((size_t) nall * maxn, 0);
    for (int ii = 0; ii < inum; ii++) {
      const int i = ilist[ii];
      for (int jj = 0; jj < numneigh[i]; jj++) h->table[(size_t) jj * nall + i] = neigh_flat[neigh_offsets[i] + jj];
    }
    auto up = [](int *&dst, const int *src, size_t n) {
      if (dst) cudaFree(dst);
      dst = nullptr;
      kk_check(cudaMalloc((void **) &dst, sizeof(int) * std::max<size_t>(n, 1)), "cudaMalloc");
      if (n) kk_check(cudaMemcpy(dst, src, sizeof(int) * n, cudaMemcpyHostToDevice), "cudaMemcpy");
    };
    up(h->d_ilist, h->ilistbuf.data(), (size_t) inum);
    up(h->d_numneigh, h->numneighbuf.data(), (size_t) nall);
    up(h->d_table, h->table.data(), h->table.size());
    h->list.d_ilist.ptr = h->d_ilist;
    h->list.d_ilist.ext[0] = inum;
    h->list.d_numneigh.ptr = h->d_numneigh;
    h->list.d_numneigh.ext[0] = nall;
    h->list.d_neighbors.ptr = h->d_table;
    h->list.d_neighbors.ext[0] = nall;
    h->list.d_neighbors.ext[1] = maxn;
    h->list.d_neighbors.str[0] = 1;
    h->list.d_neighbors.str[1] = nall;
    h->list.maxneighs = maxn;
#endif
    int dim = 0;
    if (int *flag = (int *) h->pair->extract("extrapolation_flag", dim)) *flag = extrapolation_flag;
    h->pair->compute(eflag, vflag);
#ifdef LMP_KOKKOS
    akk->k_f.sync<LMPHostType>();    // what a host-side fix would trigger through atomKK->sync(Host, F_MASK)
    if (akk->nsync < 1 || akk->nmodified < 1) throw std::runtime_error("the style did not sync / mark atom data");
#endif
    memcpy(f, h->fbuf.data(), sizeof(double) * 3 * (size_t) nall);
    ev[0] = h->pair->eng_vdwl;
    for (int k = 0; k < 6; k++) ev[1 + k] = h->pair->virial[k];
    ev[7] = h->pair->pvector ? h->pair->pvector[0] : 0.0;
    if (eatom && h->pair->eflag_atom) memcpy(eatom, h->pair->eatom, sizeof(double) * nall);
    if (vatom && h->pair->vflag_atom)
      for (int i = 0; i < nall; i++)
        for (int k = 0; k < 6; k++) vatom[6 * (size_t) i + k] = h->pair->vatom[i][k];
    if (grades && extrapolation_flag) {
      int ncol = 0;
      if (double *g = (double *) h->pair->extract_peratom("extrapolation", ncol))
        for (int ii = 0; ii < inum; ii++) grades[ilist[ii]] = g[ilist[ii]];
    }
  } catch (std::exception &e) {
    set_err(err, errlen, e.what());
    return -1;
  }
  return 0;
}
}
