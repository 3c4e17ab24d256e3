// TEST INFRASTRUCTURE -- not part of the shipped product.
//
// extern "C" driver around the UNMODIFIED reference CPU pair styles
// (PairMTP, /root/reference/LAMMPS/ML-MTP/pair_mtp.cpp:72-280 and
// PairMTPExtrapolation, pair_mtp_extrapolation.cpp:68-342), compiled from where
// they lie against oracle/lammps_shim/.  Built by oracle/Makefile into
// oracle/_ref/libmtp_ref.so.  It plays the role of upstream LAMMPS for one
// force evaluation: owns atom->x/f/type, the full neighbor list, calls
// settings()/coeff()/init_style()/init_one()/compute() in LAMMPS's order and
// hands the results back as flat arrays.  Used (a) to generate and check the
// golden vectors under tests/golden/, (b) to pin the C restatement
// (oracle/mtp_oracle.c), (c) as the "reference" CPU baseline of bench.py.
#include "pair_mtp.h"
#include "pair_mtp_extrapolation.h"

#include <cstring>
#include <string>
#include <vector>

using namespace LAMMPS_NS;

namespace {

// accessor subclasses: expose the protected potential tables of the reference classes
struct RefPlain : public PairMTP {
  explicit RefPlain(LAMMPS *l) : PairMTP(l) {}
  friend struct RefHandle;
  void info(int *iv, double *dv)
  {
    iv[0] = species_count;
    iv[1] = alpha_index_basic_count;
    iv[2] = alpha_index_times_count;
    iv[3] = alpha_moment_count;
    iv[4] = alpha_scalar_count;
    iv[5] = radial_func_count;
    iv[6] = radial_basis_size;
    iv[7] = max_alpha_index_basic;
    iv[8] = 0;
    iv[9] = 0;
    dv[0] = min_cutoff;
    dv[1] = max_cutoff;
    dv[2] = scaling;
  }
  const bool *mask() const { return within_cutoff; }
};

struct RefExtrap : public PairMTPExtrapolation {
  explicit RefExtrap(LAMMPS *l) : PairMTPExtrapolation(l) { extrapolation_flag = 0; }
  void info(int *iv, double *dv)
  {
    iv[0] = species_count;
    iv[1] = alpha_index_basic_count;
    iv[2] = alpha_index_times_count;
    iv[3] = alpha_moment_count;
    iv[4] = alpha_scalar_count;
    iv[5] = radial_func_count;
    iv[6] = radial_basis_size;
    iv[7] = max_alpha_index_basic;
    iv[8] = coeff_count;
    iv[9] = configuration_mode;
    dv[0] = min_cutoff;
    dv[1] = max_cutoff;
    dv[2] = scaling;
  }
  const bool *mask() const { return within_cutoff; }
  void close_out()
  {
    if (mlip3_style && preselected_file) {
      fflush(preselected_file);
    }
  }
  const double *candidate() const { return energy_ders_wrt_coeffs; }
  double maxgrade() const { return max_grade; }
};

struct RefHandle {
  LAMMPS *lmp = nullptr;
  NeighList list;
  RefPlain *plain = nullptr;
  RefExtrap *extrap = nullptr;
  Pair *pair = nullptr;
  std::vector<double *> xrows, frows;
  std::vector<double> xbuf, fbuf;
  std::vector<int> typebuf;
  std::vector<int *> firstneigh;
  std::vector<int> neighbuf, ilistbuf, numneighbuf;
  ~RefHandle()
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

// style: "mtp" or "mtp/extrapolation"; args: the pair_style arguments after the style name
void *mtpref_create(const char *style, int narg, const char **args, char *err, int errlen)
{
  auto *h = new RefHandle;
  try {
    h->lmp = new LAMMPS;
    std::vector<char *> argv;
    std::vector<std::string> keep(args, args + narg);
    for (auto &s : keep) argv.push_back(s.data());
    if (std::string(style) == "mtp") {
      h->plain = new RefPlain(h->lmp);
      h->pair = h->plain;
    } else if (std::string(style) == "mtp/extrapolation") {
      h->extrap = new RefExtrap(h->lmp);
      h->pair = h->extrap;
    } else {
      set_err(err, errlen, "unknown style");
      delete h;
      return nullptr;
    }
    h->pair->settings(narg, argv.data());
    char star[] = "*";
    char *cargs[2] = {star, star};
    h->pair->coeff(2, cargs);
    h->pair->init_style();
    h->pair->init_list(0, &h->list);
    int iv[10];
    double dv[3];
    if (h->plain) h->plain->info(iv, dv);
    else
      h->extrap->info(iv, dv);
    for (int i = 1; i <= iv[0]; i++)
      for (int j = 1; j <= iv[0]; j++) h->pair->init_one(i, j);
    h->lmp->atom->ntypes = iv[0];
  } catch (std::exception &e) {
    set_err(err, errlen, e.what());
    // the reference leaves partially-built state behind on a fatal error; leak rather than crash
    h->pair = nullptr;
    h->lmp = nullptr;
    delete h;
    return nullptr;
  }
  return h;
}

void mtpref_destroy(void *hv)
{
  delete (RefHandle *) hv;
}

int mtpref_info(void *hv, int *iv, double *dv)
{
  auto *h = (RefHandle *) hv;
  if (h->plain) h->plain->info(iv, dv);
  else
    h->extrap->info(iv, dv);
  return 0;
}

const char *mtpref_log(void *hv)
{
  return ((RefHandle *) hv)->lmp->log.c_str();
}

void mtpref_set_domain(void *hv, const double *prd /*xprd,yprd,zprd,xy,xz,yz*/, long natoms)
{
  auto *h = (RefHandle *) hv;
  Domain *d = h->lmp->domain;
  d->xprd = prd[0];
  d->yprd = prd[1];
  d->zprd = prd[2];
  d->xy = prd[3];
  d->xz = prd[4];
  d->yz = prd[5];
  h->lmp->atom->natoms = natoms;
}

// One force evaluation.  neigh_offsets[i] (i < nall+1... only entries for listed atoms are read)
// gives the start of atom i's row inside neigh_flat.  f is accumulated into (+=), as LAMMPS does.
// ev[0]=eng_vdwl, ev[1..6]=virial, ev[7]=pvector[0] (extrapolation styles).
// grades (optional) receives nbh_extrapolation_grades[0..nlocal) after a grade step.
// mask (optional, inum==1 only) receives within_cutoff[0..numneigh) of the single listed atom.
int mtpref_compute(void *hv, int nlocal, int nghost, const double *x, const int *type, int inum,
                   const int *ilist, const int *numneigh, const int *neigh_flat,
                   const long *neigh_offsets, int eflag, int vflag, int extrapolation_flag,
                   double *f, double *eatom, double *vatom, double *ev, double *grades,
                   unsigned char *mask, char *err, int errlen)
{
  auto *h = (RefHandle *) hv;
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

    h->ilistbuf.assign(ilist, ilist + inum);
    h->numneighbuf.assign(numneigh, numneigh + nall);
    h->firstneigh.assign(nall, nullptr);
    for (int ii = 0; ii < inum; ii++) {
      int i = ilist[ii];
      h->firstneigh[i] = const_cast<int *>(neigh_flat) + neigh_offsets[i];
    }
    h->list.inum = inum;
    h->list.ilist = h->ilistbuf.data();
    h->list.numneigh = h->numneighbuf.data();
    h->list.firstneigh = h->firstneigh.data();

    if (h->extrap) {
      int dim;
      int *flag = (int *) h->extrap->extract("extrapolation_flag", dim);
      *flag = extrapolation_flag;
    }
    h->pair->compute(eflag, vflag);

    memcpy(f, h->fbuf.data(), sizeof(double) * 3 * (size_t) nall);
    ev[0] = h->pair->eng_vdwl;
    for (int k = 0; k < 6; k++) ev[1 + k] = h->pair->virial[k];
    // pvector[0] is only written on grade steps (pair_mtp_extrapolation.cpp:381); report 0 otherwise
    ev[7] = (h->extrap && extrapolation_flag) ? h->pair->pvector[0] : 0.0;
    if (eatom && h->pair->eflag_atom) memcpy(eatom, h->pair->eatom, sizeof(double) * nall);
    if (vatom && h->pair->vflag_atom)
      for (int i = 0; i < nall; i++)
        for (int k = 0; k < 6; k++) vatom[6 * (size_t) i + k] = h->pair->vatom[i][k];
    int iv_[10];
    double dv_[3];
    if (h->extrap) h->extrap->info(iv_, dv_);
    if (grades && h->extrap && !iv_[9]) {    // extract_peratom is a fatal error in configuration mode (:644-645)
      int ncol;
      double *g = (double *) h->extrap->extract_peratom("extrapolation", ncol);
      if (g)
        for (int ii = 0; ii < inum; ii++) grades[ilist[ii]] = g[ilist[ii]];
    }
    if (mask && inum == 1) {
      const bool *m = h->plain ? h->plain->mask() : h->extrap->mask();
      for (int jj = 0; jj < numneigh[ilist[0]]; jj++) mask[jj] = m[jj] ? 1 : 0;
    }
    if (h->extrap) h->extrap->close_out();
  } catch (std::exception &e) {
    set_err(err, errlen, e.what());
    return -1;
  }
  return 0;
}

// configuration-mode candidate vector (energy_ders_wrt_coeffs, Q doubles) after a grade step
int mtpref_candidate(void *hv, double *out, int q)
{
  auto *h = (RefHandle *) hv;
  if (!h->extrap) return -1;
  memcpy(out, h->extrap->candidate(), sizeof(double) * q);
  return 0;
}
}
