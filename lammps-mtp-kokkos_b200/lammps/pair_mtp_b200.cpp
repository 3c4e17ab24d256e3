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

#include "mtp_b200.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

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
  // pair_mtp_kokkos.cpp:37-45 minus the Kokkos bookkeeping
  respa_enable = 0;
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
  delete[] pvector;
  pvector = nullptr;
}

void PairMTPB200::fatal(const char *file, int line, int rc)
{
  // the library reports, LAMMPS aborts: the reference's fatal-error convention (pair_mtp.cpp:92,288,306,315,327)
  error->all(file, line, "{} (mtp_b200 error {})", mtp_last_error(), rc);
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
  if (rc) fatal(FLERR, rc);

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
  if (extrapolation && want_grade) max_grade = 0;
  ev_init(eflag, vflag);

  const int nlocal = atom->nlocal, nall = atom->nlocal + atom->nghost;
  const int inum = list->inum;
  if (want_grade && !configuration_mode && nbh_count < std::max(inum, nall)) {    // :91-94, indexed by atom id
    memory->grow(nbh_extrapolation_grades, std::max(inum, nall), "nbh_extrapolation_grades");
    nbh_count = std::max(inum, nall);
  }

  // flatten the paged host list on re-neighboring steps only
  const bool list_changed = neighbor->ago == 0 || flat_offsets.empty();
  if (list_changed) {
    flat_offsets.assign((size_t) nall + 1, 0);
    long long total = 0;
    for (int ii = 0; ii < inum; ii++) total += list->numneigh[list->ilist[ii]];
    flat_neigh.resize((size_t) std::max<long long>(total, 1));
    long long at = 0;
    for (int ii = 0; ii < inum; ii++) {
      const int i = list->ilist[ii];
      flat_offsets[i] = at;
      std::memcpy(flat_neigh.data() + at, list->firstneigh[i], sizeof(int) * (size_t) list->numneigh[i]);
      at += list->numneigh[i];
    }
  }

  double ev[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  mtp_compute_args a;
  std::memset(&a, 0, sizeof(a));
  a.variant = variant;
  a.inum = inum;
  a.nall = nall;
  a.x = nall ? &atom->x[0][0] : nullptr;
  a.type = atom->type;
  a.ilist = list->ilist;
  a.numneigh = list->numneigh;
  a.neighbors = flat_neigh.data();
  a.neigh_offsets = flat_offsets.data();
  a.stride_i = 0;
  a.stride_jj = 1;
  a.neighmask = NEIGHMASK;
  a.eflag = (eflag_global ? 1 : 0) | (eflag_atom ? 2 : 0);
  a.vflag = vflag ? ((vflag_atom ? 4 : 0) | 1) : 0;    // the CPU style tallies the pairwise virial whenever vflag != 0
  a.want_grade = want_grade ? 1 : 0;
  a.natoms_total = (long long) atom->natoms;
  a.f = nall ? &atom->f[0][0] : nullptr;
  a.eatom = eflag_atom ? eatom : nullptr;
  a.vatom = vflag_atom ? &vatom[0][0] : nullptr;
  a.ev_out = ev;
  a.grades = (want_grade && !configuration_mode) ? nbh_extrapolation_grades : nullptr;
  a.cfg_candidate = (want_grade && configuration_mode) ? cfg_candidate.data() : nullptr;
  if (nall > 0 && inum > 0) {
    const int rc = mtp_compute_host(handle, &a, list_changed ? 1 : 0);
    if (rc) fatal(FLERR, rc);
  }

  if (eflag_global) eng_vdwl += ev[0];
  if (vflag)
    for (int k = 0; k < 6; k++) virial[k] += ev[1 + k];    // pair_mtp.cpp:257-266: -sym(F (x) r), never fdotr

  if (want_grade) {
    max_grade = ev[7];
    compile_grades();
    if (mlip3_style) evaluate_grades();
  }
  (void) nlocal;
}

/* ----------------------------------------------------------------------
   collective reduction (pair_mtp_extrapolation.cpp:363-382).  In configuration mode the candidate vector
   is summed over ranks and the grade is re-evaluated from the sum.
------------------------------------------------------------------------- */

void PairMTPB200::compile_grades()
{
  if (configuration_mode) {
    if (comm->nprocs > 1) {
      MPI_Allreduce(MPI_IN_PLACE, cfg_candidate.data(), coeff_count, MPI_DOUBLE, MPI_SUM, world);
      // max_i |Ainv[i,:] . b| / natoms on the summed vector
      std::vector<double> ainv((size_t) coeff_count * coeff_count);
      mtp_get_tables(handle, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, ainv.data());
      double g = 0.0;
      for (int i = 0; i < coeff_count; i++) {
        double s = 0.0;
        for (int j = 0; j < coeff_count; j++) s += ainv[(size_t) i * coeff_count + j] * cfg_candidate[j];
        g = std::max(g, std::fabs(s));
      }
      max_grade = atom->natoms > 0 ? g / (double) atom->natoms : 0.0;
    }
    // single rank: the library already normalised by natoms_total
  } else {
    MPI_Allreduce(MPI_IN_PLACE, &max_grade, 1, MPI_DOUBLE, MPI_MAX, world);
  }
  if (comm->me == 0) pvector[0] = max_grade;    // Expose the max grade (rank 0 only: compute pair SUMs pvector)
}

void PairMTPB200::evaluate_grades()
{
  if (max_grade >= select_threshold) write_config();
  if (max_grade >= break_threshold && comm->me == 0) {
    std::fflush(preselected_file);    // Ensure the writing buffers are flushed before breaking.
    std::fclose(preselected_file);
    preselected_file = nullptr;
    error->one(FLERR, "Exceeded Break Threshold: {:.5f}. Terminating simulation.\n", max_grade);
  }
}

/* ----------------------------------------------------------------------
   MLIP-3 preselected-configuration block (pair_mtp_extrapolation.cpp:401-479), same text byte for byte
------------------------------------------------------------------------- */

void PairMTPB200::write_config()
{
  write_buffer.clear();
  const int inum = list->inum;
  int *type = atom->type;
  double **x = atom->x;
  int index_offset = 0;
  MPI_Scan(&inum, &index_offset, 1, MPI_INT, MPI_SUM, MPI_COMM_WORLD);
  index_offset -= inum;

  char line[256];
  for (int ii = 0; ii < inum; ii++) {
    const int i = ii;    // (sic) the reference indexes by ii, SURVEY.md App. B9
    const int itype = type[i] - 1;
    const int global_i = i + index_offset + 1;
    int n;
    if (!configuration_mode)
      n = snprintf(line, sizeof(line), "%d\t%d\t%.6f\t%.6f\t%.6f\t%.5f\n", global_i, itype, x[i][0], x[i][1], x[i][2],
                   nbh_extrapolation_grades[i]);
    else
      n = snprintf(line, sizeof(line), "%d\t%d\t%.6f\t%.6f\t%.6f\n", global_i, itype, x[i][0], x[i][1], x[i][2]);
    write_buffer.append(line, (size_t) n);
  }

  bigint char_buffer_size = (bigint) write_buffer.size();
  bigint max_char_buffer_size = char_buffer_size;
  MPI_Reduce(&char_buffer_size, &max_char_buffer_size, 1, MPI_LMP_BIGINT, MPI_MAX, 0, world);

  if (comm->me == 0) {
    std::fprintf(preselected_file, "BEGIN_CFG\n");
    std::fprintf(preselected_file, "Size\n");
    std::fprintf(preselected_file, "%ld\n", (long) atom->natoms);
    std::fprintf(preselected_file, "Supercell\n");
    std::fprintf(preselected_file, "%.6f %.6f %.6f\n", domain->xprd, 0.0, 0.0);
    std::fprintf(preselected_file, "%.6f %.6f %.6f\n", domain->xy, domain->yprd, 0.0);
    std::fprintf(preselected_file, "%.6f %.6f %.6f\n", domain->xz, domain->yz, domain->zprd);
    if (!configuration_mode)
      std::fprintf(preselected_file,
                   "AtomData:  id type       cartes_x      cartes_y      cartes_z       nbh_grades\n");
    else
      std::fprintf(preselected_file, "AtomData:  id type       cartes_x      cartes_y      cartes_z\n");
    std::fwrite(write_buffer.data(), 1, (size_t) char_buffer_size, preselected_file);
  }

  if (comm->me != 0) {
    MPI_Send(write_buffer.data(), (int) char_buffer_size, MPI_CHAR, 0, 0, world);
  } else {
    std::vector<char> recv((size_t) std::max<bigint>(max_char_buffer_size, 1));
    for (int p = 1; p < comm->nprocs; p++) {
      MPI_Status status;
      int n_chars = 0;
      MPI_Recv(recv.data(), (int) max_char_buffer_size, MPI_CHAR, p, 0, world, &status);
      MPI_Get_count(&status, MPI_CHAR, &n_chars);
      std::fwrite(recv.data(), 1, (size_t) n_chars, preselected_file);
    }
    std::fprintf(preselected_file, "Feature   MV_grade\t%.6f\n", max_grade);
    std::fprintf(preselected_file, "END_CFG\n\n");
  }
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
    return (void *) nbh_extrapolation_grades;
  }
  return nullptr;
}
