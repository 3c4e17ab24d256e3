/* ----------------------------------------------------------------------
   LAMMPS pair styles mtp/kk, mtp/small/kk, mtp/extrapolation/kk, mtp/extrapolation/small/kk on top of
   the B200-native MTP library (extern "C" API of include/mtp_b200.h).

   This file is the HOST side of the drop-in boundary.  It keeps what the reference keeps on the host --
   argument grammar and messages (pair_mtp_kokkos.cpp:104-117, pair_mtp_extrapolation_kokkos.cpp:116-138,
   pair_mtp.cpp:286-329), the grade reduction / threshold logic / MLIP-3 .cfg writer
   (pair_mtp_extrapolation.cpp:363-479), extract() / extract_peratom() (:624-652), pvector -- and hands
   positions, types and the full neighbor list to the CUDA layer, which owns everything per-atom.

   Two build flavours:
     * plain LAMMPS (default): atom->x / f / type and the paged list->firstneigh live on the host; the style
       calls mtp_compute_host(), re-uploading the neighbor list only on re-neighboring steps (neighbor->ago == 0).
     * LAMMPS-KOKKOS (-DLMP_KOKKOS): device views are passed straight to mtp_compute().
------------------------------------------------------------------------- */

#include "pair_mtp_b200.h"

#include "atom.h"
#include "comm.h"
#include "domain.h"
#include "error.h"
#include "force.h"
#include "memory.h"
#include "neigh_list.h"
#include "neighbor.h"
#include "utils.h"
#ifdef LMP_KOKKOS
#include "atom_kokkos.h"
#include "atom_masks.h"
#include "memory_kokkos.h"
#include "neigh_list_kokkos.h"
#endif

#include "mtp_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

using namespace LAMMPS_NS;

/* ---------------------------------------------------------------------- */

PairMTPB200::PairMTPB200(LAMMPS *lmp, int variant_, bool extrapolation_) :
    Pair(lmp), variant(variant_), extrapolation(extrapolation_)
{
  // pair_mtp.cpp:37-40
  single_enable = 0;
  restartinfo = 0;
  one_coeff = 1;
  manybody_flag = 1;
  respa_enable = 0;
#ifdef LMP_KOKKOS
  // pair_mtp_kokkos.cpp:37-45: the style reads and writes atom data on the device only
  kokkosable = 1;
  execution_space = Device;
  datamask_read = EMPTY_MASK;
  datamask_modify = EMPTY_MASK;
#endif
  if (extrapolation) {    // pair_mtp_extrapolation.cpp:42-44
    nextra = 1;
    pvector = new double[nextra];
    pvector[0] = 0.0;
  }
}

PairMTPB200::~PairMTPB200()
{
  if (copymode) return;
  if (handle) mtp_destroy(handle);
  memory->destroy(nbh_extrapolation_grades);
  if (preselected_file) fclose(preselected_file);
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
  }
#ifdef LMP_KOKKOS
  MemoryKokkos mk;
  mk.destroy_kokkos(k_eatom, eatom);
  mk.destroy_kokkos(k_vatom, vatom_rows_kk);
  k_grades.release();
  mtp_free_pinned(ev_pinned);
  vatom = nullptr;
#endif
  delete[] pvector;
  pvector = nullptr;
}

// The library reports, LAMMPS aborts (the reference's fatal-error convention, pair_mtp.cpp:92,288,306,315,327).
// A failure of one force evaluation (the species bound on this rank's atoms, a CUDA error on this rank's GPU) is this
// rank's alone -- error->one, like the reference's per-atom check (pair_mtp.cpp:91-93,116-118); error->all would wait
// for ranks that never get here.
void PairMTPB200::fatal_one(const char *file, int line, int rc)
{
  error->one(file, line, "{} (mtp_b200 error {})", mtp_last_error(), rc);
  throw std::runtime_error("unreachable");
}

void PairMTPB200::fatal_all(const char *file, int line, int rc)
{
  error->all(file, line, "{} (mtp_b200 error {})", mtp_last_error(), rc);
  throw std::runtime_error("unreachable");
}

/* ----------------------------------------------------------------------
   global settings: same grammar as the reference's KOKKOS styles
------------------------------------------------------------------------- */

void PairMTPB200::settings(int narg, char **arg)
{
  if (!extrapolation) {
    // pair_mtp_kokkos.cpp:109-113
    if (narg != 3 || utils::lowercase(arg[1]) != "chunksize")
      error->all(FLERR, "Pair mtp/kk requires 3 arguments {{potential_file}} \"chunksize\" {{chunksize}}.");
    chunksize = utils::inumeric(FLERR, arg[2], true, lmp);
  } else {
    // pair_mtp_extrapolation_kokkos.cpp:120-138
    if (narg != 3 && narg != 6)
      error->all(FLERR,
                 "Pair mtp/extrapolation/kk/s requires 3 : {{potential_file}} \"chunksize\" {{chunksize}} "
                 "Or 6 arguments: {{potential_file}} {{output_file}} {{selection_threshold}} "
                 "{{break_threshold}} \"chunksize\" {{chunksize}}.");
    const int kw = narg == 3 ? 1 : 4;
    if (utils::lowercase(arg[kw]) != "chunksize")
      error->all(FLERR, "Chunksize not found, please specify \"chunksize\" {{chunksize}}.");
    chunksize = utils::inumeric(FLERR, arg[kw + 1], true, lmp);
    if (narg == 6) {    // MLIP-3 style thresholds, pair_mtp_extrapolation.cpp:495-499
      mlip3_style = true;
      select_threshold = utils::numeric(FLERR, arg[2], true, lmp);
      break_threshold = utils::numeric(FLERR, arg[3], true, lmp);
    }
  }
  if (chunksize < 1) error->all(FLERR, "Illegal chunksize {}", chunksize);

  if (handle) mtp_destroy(handle);
  handle = mtp_create_from_file(arg[0], extrapolation ? 1 : 0, -1);
  if (!handle) error->all(FLERR, "{}", mtp_last_error());
  int rc = mtp_set_chunksize(handle, chunksize);
  if (rc) fatal_all(FLERR, rc);

  mtp_info info;
  mtp_get_info(handle, &info);
  species_count = info.species_count;
  coeff_count = info.coeff_count;
  configuration_mode = info.configuration_mode;
  max_cutoff = info.max_cutoff;
  if (comm->me == 0) {
    utils::logmesg(lmp, "The scaling is : {:.2e}.\n", info.scaling);
    utils::logmesg(lmp, "There are {} species.\n", species_count);
  }

  // setflag / cutsq exactly like pair_mtp.cpp:392-393,448-449,455: every pair the file lists is set
  if (allocated) {
    memory->destroy(setflag);
    memory->destroy(cutsq);
  }
  memory->create(setflag, species_count + 1, species_count + 1, "pair:setflag");
  memory->create(cutsq, species_count + 1, species_count + 1, "pair:cutsq");
  for (int i = 0; i <= species_count; i++)
    for (int j = 0; j <= species_count; j++) {
      setflag[i][j] = (i > 0 && j > 0) ? 1 : 0;
      cutsq[i][j] = max_cutoff * max_cutoff;
    }
  allocated = 1;

  if (extrapolation) {
    if (comm->me == 0) {    // pair_mtp_extrapolation.cpp:505-517
      if (mlip3_style)
        utils::logmesg(lmp,
                       "Extrapolation Scheme: {} mode, with a selection threshold of {} and break threshold of {}.\n",
                       (configuration_mode ? "Configuration" : "Neighborhood"), select_threshold, break_threshold);
      else
        utils::logmesg(lmp, "Extrapolation Mode: {} mode.\n", (configuration_mode ? "Configuration" : "Neighborhood"));
    }
    if (mlip3_style && comm->me == 0) {    // :519-522
      if (preselected_file) fclose(preselected_file);
      preselected_file = std::fopen(arg[1], "w");
      if (!preselected_file) error->one(FLERR, "Cannot open preselected configuration file {}", arg[1]);
    }
    cfg_candidate.assign((size_t) std::max(coeff_count, 1), 0.0);
  }
}

/* ----------------------------------------------------------------------
   set coeffs for one or more type pairs (pair_mtp.cpp:303-307)
------------------------------------------------------------------------- */

void PairMTPB200::coeff(int narg, char ** /*arg*/)
{
  // The potential file is specified in the setting function instead.
  if (narg != 2) error->all(FLERR, "Only \"pair_coeff * *\" is permitted");
}

/* ----------------------------------------------------------------------
   init specific to this pair style (pair_mtp.cpp:313-319, pair_mtp_kokkos.cpp:62-83)
------------------------------------------------------------------------- */

void PairMTPB200::init_style()
{
  if (force->newton_pair != 1) error->all(FLERR, "Pair style MTP requires Newton Pair on");
  // Request a full neighbourhood list which is needed for MTP
  neighbor->add_request(this, NeighConst::REQ_FULL);
}

double PairMTPB200::init_one(int i, int j)
{
  if (setflag[i][j] == 0) error->all(FLERR, "Not all pair coeffs are set. See types {}-{}.", i, j);
  return max_cutoff;
}

/* ----------------------------------------------------------------------
   one force evaluation (pair_mtp.cpp:72-280, pair_mtp_extrapolation.cpp:68-342)
------------------------------------------------------------------------- */

void PairMTPB200::compute(int eflag, int vflag)
{
  const bool want_grade = extrapolation && (extrapolation_flag || mlip3_style);    // pair_mtp_extrapolation.cpp:71
#ifdef LMP_KOKKOS
  ev_init(eflag, vflag, 0);    // per-atom arrays are DualViews of this class (pair_mtp_kokkos.cpp:212-224)
#else
  ev_init(eflag, vflag);
#endif

  double ev[8] = {0, 0, 0, 0, 0, 0, 0, 0};    // E, virial xx yy zz xy xz yz, max grade of this rank
  if (atom->nlocal + atom->nghost > 0 && list->inum > 0) {
#ifdef LMP_KOKKOS
    compute_device_views(eflag, vflag, want_grade, ev);
#else
    compute_host_buffers(eflag, vflag, want_grade, ev);
#endif
  }
  if (eflag_global) eng_vdwl += ev[0];
  if (vflag)
    for (int k = 0; k < 6; k++) virial[k] += ev[1 + k];    // pair_mtp.cpp:257-266: -sym(F (x) r), never fdotr

  if (want_grade) {
    max_grade = ev[7];
    host_grades_stale = true;
    grade_rows = atom->nlocal + atom->nghost;
    reduce_max_grade();
    if (mlip3_style) act_on_thresholds();
  }
}

/* ----------------------------------------------------------------------
   plain LAMMPS: atom data and the paged neighbor list live on the host
------------------------------------------------------------------------- */

void PairMTPB200::compute_host_buffers(int /*eflag*/, int vflag, bool want_grade, double *ev)
{
  const int nall = atom->nlocal + atom->nghost;
  const int inum = list->inum;

  // The list only changes on re-neighboring steps: it is flattened (one block copy per run of consecutive pages) and
  // uploaded then, and stays resident on the device in between.
  const bool list_changed = neighbor->ago == 0 || flat_offsets.empty();
  if (list_changed) {
    flat_offsets.assign((size_t) nall + 1, 0);
    long long total = 0;
    for (int ii = 0; ii < inum; ii++) total += list->numneigh[list->ilist[ii]];
    flat_neigh.resize((size_t) std::max<long long>(total, 1));
    long long at = 0;
    int ii = 0;
    while (ii < inum) {
      // rows that follow each other in a page of LAMMPS's neighbor pool are copied as one block
      const int *src = list->firstneigh[list->ilist[ii]];
      long long run = 0;
      int jj = ii;
      while (jj < inum && list->firstneigh[list->ilist[jj]] == src + run) {
        flat_offsets[list->ilist[jj]] = at + run;
        run += list->numneigh[list->ilist[jj]];
        jj++;
      }
      if (run) std::memcpy(flat_neigh.data() + at, src, sizeof(int) * (size_t) run);
      at += run;
      ii = jj;
    }
  }

  mtp_compute_args a;
  std::memset(&a, 0, sizeof(a));
  a.variant = variant;
  a.inum = inum;
  a.nall = nall;
  a.x = &atom->x[0][0];
  a.type = atom->type;
  a.ilist = list->ilist;
  a.numneigh = list->numneigh;
  a.neighbors = flat_neigh.data();
  a.neigh_offsets = flat_offsets.data();
  a.stride_jj = 1;
  a.neighmask = NEIGHMASK;
  a.eflag = (eflag_global ? 1 : 0) | (eflag_atom ? 2 : 0);
  a.vflag = vflag ? ((vflag_atom ? 4 : 0) | 1) : 0;    // the CPU style tallies the pairwise virial whenever vflag != 0
  a.want_grade = want_grade ? 1 : 0;
  a.natoms_total = (long long) atom->natoms;
  a.f = &atom->f[0][0];
  // LAMMPS clears f before Pair::compute; a lone pair style is the first to add to it, so nothing needs uploading
  a.f_overwrite = (force->pair == this) ? 1 : 0;
  a.eatom = eflag_atom ? eatom : nullptr;
  a.vatom = vflag_atom ? &vatom[0][0] : nullptr;
  a.ev_out = ev;
  a.grades = nullptr;    // neighbourhood grades stay on the device until somebody asks (host_grades_current)
  a.cfg_candidate = (want_grade && configuration_mode) ? cfg_candidate.data() : nullptr;
  const int rc = mtp_compute_host(handle, &a, list_changed ? 1 : 0);
  if (rc) fatal_one(FLERR, rc);
}

#ifdef LMP_KOKKOS
/* ----------------------------------------------------------------------
   LAMMPS-KOKKOS: device views of atom data and of the neighbor list go straight to the CUDA layer
   (what pair_mtp_kokkos.cpp:231-240 hands its functors)
------------------------------------------------------------------------- */

void PairMTPB200::compute_device_views(int /*eflag*/, int vflag, bool want_grade, double *ev)
{
  AtomKokkos *atomKK = (AtomKokkos *) atom;
  MemoryKokkos memoryKK;
  const int nall = atom->nlocal + atom->nghost;

  // reallocate per-atom arrays if necessary (pair_mtp_kokkos.cpp:215-224)
  if (eflag_atom) {
    memoryKK.destroy_kokkos(k_eatom, eatom);
    memoryKK.create_kokkos(k_eatom, eatom, maxeatom, "pair:eatom");
  }
  if (vflag_atom) {
    memoryKK.destroy_kokkos(k_vatom, vatom_rows_kk);
    memoryKK.create_kokkos(k_vatom, vatom_rows_kk, maxvatom, 6, "pair:vatom");
    vatom = vatom_rows_kk;
  }
  if (want_grade && !configuration_mode && k_grades.d_view.extent(0) < nall) {
    memory->grow(nbh_extrapolation_grades, nall, "nbh_extrapolation_grades");
    nbh_count = nall;
    k_grades.allocate(nbh_extrapolation_grades, nall);
  }
  if (!ev_pinned) ev_pinned = (double *) mtp_alloc_pinned(8 * sizeof(double));
  if (!ev_pinned) fatal_one(FLERR, MTP_ERR_CUDA);

  atomKK->sync((ExecutionSpace) execution_space, X_MASK | F_MASK | TYPE_MASK);
  auto x = atomKK->k_x.view<LMPDeviceType>();
  auto f = atomKK->k_f.view<LMPDeviceType>();
  auto type = atomKK->k_type.view<LMPDeviceType>();
  auto *k_list = static_cast<NeighListKokkos<LMPDeviceType> *>(list);

  mtp_compute_args a;
  std::memset(&a, 0, sizeof(a));
  a.variant = variant;
  a.inum = list->inum;
  a.nall = nall;
  a.x = x.data();
  a.type = type.data();
  a.ilist = k_list->d_ilist.data();
  a.numneigh = k_list->d_numneigh.data();
  a.neighbors = k_list->d_neighbors.data();         // d_neighbors(i, jj), either layout
  a.stride_i = k_list->d_neighbors.stride(0);
  a.stride_jj = k_list->d_neighbors.stride(1);
  a.max_numneigh = (int) k_list->d_neighbors.extent(1);
  a.neighmask = NEIGHMASK;
  a.eflag = (eflag_global ? 1 : 0) | (eflag_atom ? 2 : 0);
  a.vflag = vflag ? ((vflag_atom ? 4 : 0) | 1) : 0;
  a.want_grade = want_grade ? 1 : 0;
  a.natoms_total = (long long) atom->natoms;
  a.f = f.data();
  a.eatom = eflag_atom ? k_eatom.view<LMPDeviceType>().data() : nullptr;
  a.vatom = vflag_atom ? k_vatom.view<LMPDeviceType>().data() : nullptr;
  a.ev_out = ev_pinned;
  a.grades = (want_grade && !configuration_mode) ? k_grades.view<LMPDeviceType>().data() : nullptr;
  DAT::tdual_efloat_1d k_cand;
  if (want_grade && configuration_mode) {
    k_cand.allocate(cfg_candidate.data(), coeff_count);
    a.cfg_candidate = k_cand.view<LMPDeviceType>().data();
  }
  int rc = mtp_compute(handle, &a);
  if (!rc) rc = mtp_synchronize(handle);
  if (rc) {
    k_cand.release();
    fatal_one(FLERR, rc);
  }
  for (int k = 0; k < 8; k++) ev[k] = ev_pinned[k];
  atomKK->modified((ExecutionSpace) execution_space, F_MASK);

  // per-atom results back to the host side of their DualViews (pair_mtp_kokkos.cpp:379-390)
  if (eflag_atom) {
    k_eatom.modify<LMPDeviceType>();
    k_eatom.sync<LMPHostType>();
  }
  if (vflag_atom) {
    k_vatom.modify<LMPDeviceType>();
    k_vatom.sync<LMPHostType>();
  }
  if (a.grades) k_grades.modify<LMPDeviceType>();    // synced on demand (host_grades_current)
  if (a.cfg_candidate) {
    k_cand.modify<LMPDeviceType>();
    k_cand.sync<LMPHostType>();
    k_cand.release();
  }
}
#endif

/* ----------------------------------------------------------------------
   The neighbourhood grades of a grade step stay on the device; the host array that extract_peratom() and the
   .cfg writer expose (indexed by atom id, pair_mtp_extrapolation.cpp:335,641-652) is refreshed the first time it
   is asked for after the step.  Only the 8-double record crosses PCIe on a step nobody looks at the array.
------------------------------------------------------------------------- */

void PairMTPB200::host_grades_current()
{
  if (!host_grades_stale || configuration_mode) return;
#ifdef LMP_KOKKOS
  k_grades.sync<LMPHostType>();
#else
  if (nbh_count < grade_rows) {    // pair_mtp_extrapolation.cpp:91-94
    memory->grow(nbh_extrapolation_grades, grade_rows, "nbh_extrapolation_grades");
    nbh_count = grade_rows;
  }
  const int rc = mtp_fetch_grades(handle, nbh_extrapolation_grades, grade_rows);
  if (rc) fatal_one(FLERR, rc);
#endif
  host_grades_stale = false;
}

/* ----------------------------------------------------------------------
   Grade of the whole system from the per-rank results (what pair_mtp_extrapolation.cpp:363-382 computes).
   Neighbourhood mode: the largest per-atom grade of any rank.  Configuration mode: the candidate vectors of the
   ranks add up to the configuration's, whose grade max|Ainv . b| / natoms the device evaluates against the
   resident inverse active set (one rank: mtp_compute already did, with natoms_total).
   pvector[0] is written on rank 0 only, because `compute pair` SUMs it over ranks.
------------------------------------------------------------------------- */

void PairMTPB200::reduce_max_grade()
{
  if (comm->nprocs > 1) {
    if (configuration_mode) {
      MPI_Allreduce(MPI_IN_PLACE, cfg_candidate.data(), coeff_count, MPI_DOUBLE, MPI_SUM, world);
      const int rc = mtp_cfg_grade(handle, cfg_candidate.data(), (long long) atom->natoms, &max_grade);
      if (rc) fatal_one(FLERR, rc);
    } else {
      double mine = max_grade;
      MPI_Allreduce(&mine, &max_grade, 1, MPI_DOUBLE, MPI_MAX, world);
    }
  }
  if (comm->me == 0) pvector[0] = max_grade;
}

/* ----------------------------------------------------------------------
   MLIP-3 style thresholds (pair_mtp_extrapolation.cpp:387-397): a configuration whose grade reaches the selection
   threshold is appended to the preselected file; at the break threshold the run stops with the file closed.
------------------------------------------------------------------------- */

void PairMTPB200::act_on_thresholds()
{
  if (max_grade >= select_threshold) append_selected_configuration();
  if (max_grade < break_threshold || comm->me != 0) return;
  if (preselected_file) {
    std::fclose(preselected_file);    // fclose flushes: everything selected so far is on disk before the abort
    preselected_file = nullptr;
  }
  error->one(FLERR, "Exceeded Break Threshold: {:.5f}. Terminating simulation.\n", max_grade);
}

/* ----------------------------------------------------------------------
   One BEGIN_CFG ... END_CFG block of the MLIP-3 preselected-configuration format, byte for byte what
   pair_mtp_extrapolation.cpp:401-479 writes (tests compare with the reference writer's own output): header from
   the domain, one line per listed atom in rank order -- id = running 1-based index over ranks, 0-based type, raw
   coordinates, and in neighbourhood mode the atom's grade; like the reference the rows are indexed by ii itself,
   SURVEY.md App. B9 -- then the grade of the configuration.  Rank 0 writes; the other ranks' lines reach it
   through one gather of sizes and one gather of text.
------------------------------------------------------------------------- */

void PairMTPB200::append_selected_configuration()
{
  const int inum = list->inum;
  if (!configuration_mode) host_grades_current();

  // running atom index over ranks: inclusive prefix sum minus my own count
  int before_me = 0;
  MPI_Scan(&inum, &before_me, 1, MPI_INT, MPI_SUM, world);
  before_me -= inum;

  std::string mine;
  mine.reserve((size_t) inum * 64);
  char row[256];
  for (int ii = 0; ii < inum; ii++) {
    const double *xi = atom->x[ii];
    int n = snprintf(row, sizeof(row), "%d\t%d\t%.6f\t%.6f\t%.6f", before_me + ii + 1, atom->type[ii] - 1, xi[0], xi[1], xi[2]);
    if (!configuration_mode) n += snprintf(row + n, sizeof(row) - n, "\t%.5f", nbh_extrapolation_grades[ii]);
    row[n++] = '\n';
    mine.append(row, (size_t) n);
  }

  // rank-ordered concatenation on rank 0
  const int nprocs = comm->nprocs;
  int my_len = (int) mine.size();
  std::vector<int> lens((size_t) nprocs, 0), at((size_t) nprocs, 0);
  MPI_Gather(&my_len, 1, MPI_INT, lens.data(), 1, MPI_INT, 0, world);
  std::string all;
  if (comm->me == 0) {
    long long total = 0;
    for (int p = 0; p < nprocs; p++) {
      at[p] = (int) total;
      total += lens[p];
    }
    all.resize((size_t) total);
  }
  MPI_Gatherv(mine.data(), my_len, MPI_CHAR, all.empty() ? nullptr : &all[0], lens.data(), at.data(), MPI_CHAR, 0, world);
  if (comm->me != 0) return;

  FILE *out = preselected_file;
  std::fprintf(out, "BEGIN_CFG\nSize\n%ld\nSupercell\n", (long) atom->natoms);
  std::fprintf(out, "%.6f %.6f %.6f\n", domain->xprd, 0.0, 0.0);
  std::fprintf(out, "%.6f %.6f %.6f\n", domain->xy, domain->yprd, 0.0);
  std::fprintf(out, "%.6f %.6f %.6f\n", domain->xz, domain->yz, domain->zprd);
  std::fprintf(out, "AtomData:  id type       cartes_x      cartes_y      cartes_z%s\n", configuration_mode ? "" : "       nbh_grades");
  std::fwrite(all.data(), 1, all.size(), out);
  std::fprintf(out, "Feature   MV_grade\t%.6f\nEND_CFG\n\n", max_grade);
}

/* ----------------------------------------------------------------------
   fix pair / compute pair hooks (pair_mtp_extrapolation.cpp:624-652)
------------------------------------------------------------------------- */

void *PairMTPB200::extract(const char *str, int &dim)
{
  dim = 0;
  if (extrapolation && strcmp(str, "extrapolation_flag") == 0) return (void *) &extrapolation_flag;
  return nullptr;
}

void *PairMTPB200::extract_peratom(const char *str, int &ncol)
{
  if (extrapolation && strcmp(str, "extrapolation") == 0) {
    if (configuration_mode)
      error->one(FLERR, "Please use the MLIP-3 style extrapolation for configuration mode MTPs!");
    ncol = 0;
    host_grades_current();
    return (void *) nbh_extrapolation_grades;
  }
  return nullptr;
}
