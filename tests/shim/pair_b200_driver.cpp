// TEST INFRASTRUCTURE -- plays the role of upstream LAMMPS (oracle/lammps_shim/) around the product's own
// PairStyle classes (lammps-mtp-kokkos_b200/lammps/pair_mtp_b200.cpp): creates the style by its pair_style
// string, calls settings()/coeff()/init_style()/init_one()/compute() in LAMMPS's order on caller-provided
// atoms + full neighbor list, and hands the results back as flat arrays.  Same shape as oracle/ref_driver.cpp,
// so that tests drive the reference's CPU styles and these styles with identical inputs.
#include "pair_mtp_b200.h"

#include <cstring>
#include <string>
#include <vector>

using namespace LAMMPS_NS;

namespace {
struct Handle {
  LAMMPS *lmp = nullptr;
  NeighList list;
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

    int dim = 0;
    if (int *flag = (int *) h->pair->extract("extrapolation_flag", dim)) *flag = extrapolation_flag;
    h->pair->compute(eflag, vflag);

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
