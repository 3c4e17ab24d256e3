/* -*- c++ -*- ----------------------------------------------------------
   LAMMPS pair styles backed by the B200-native MTP library (include/mtp_b200.h).

   Drop-in for the reference's KOKKOS styles (same style strings, same argument grammar, same
   extract() / extract_peratom() / pvector contract), see
     LAMMPS/KOKKOS/pair_mtp_kokkos.h:18-23,           pair_mtps_kokkos.h
     LAMMPS/KOKKOS/pair_mtp_extrapolation_kokkos.h,   pair_mtps_extrapolation_kokkos.h
   of RichardZJM/lammps-mtp-kokkos.  There is no "/host" alias: this path has no CPU fallback
   (use the reference's own `mtp` / `mtp/extrapolation` CPU styles for that).
------------------------------------------------------------------------- */

#ifdef PAIR_CLASS
// clang-format off
PairStyle(mtp/kk,PairMTPB200Large);
PairStyle(mtp/kk/device,PairMTPB200Large);
PairStyle(mtp/small/kk,PairMTPB200Small);
PairStyle(mtp/small/kk/device,PairMTPB200Small);
PairStyle(mtp/extrapolation/kk,PairMTPB200ExtrapolationLarge);
PairStyle(mtp/extrapolation/kk/device,PairMTPB200ExtrapolationLarge);
PairStyle(mtp/extrapolation/small/kk,PairMTPB200ExtrapolationSmall);
PairStyle(mtp/extrapolation/small/kk/device,PairMTPB200ExtrapolationSmall);
// clang-format on
#else

#ifndef LMP_PAIR_MTP_B200_H
#define LMP_PAIR_MTP_B200_H

#include "pair.h"
#ifdef LMP_KOKKOS
#include "kokkos_type.h"
#endif

#include <string>
#include <vector>

struct mtp_handle;

namespace LAMMPS_NS {

class PairMTPB200 : public Pair {
 public:
  PairMTPB200(class LAMMPS *, int variant, bool extrapolation);
  ~PairMTPB200() override;
  void compute(int, int) override;
  void settings(int, char **) override;
  void coeff(int, char **) override;
  void init_style() override;
  double init_one(int, int) override;
  void *extract(const char *, int &) override;
  void *extract_peratom(const char *, int &) override;

 protected:
  // grade bookkeeping after a grade step (semantics of pair_mtp_extrapolation.cpp:363-397)
  void reduce_max_grade();            // across ranks: MAX (neighbourhood) or grade of the SUMMED candidate (configuration)
  void act_on_thresholds();           // select_threshold -> append the configuration, break_threshold -> abort
  void append_selected_configuration();   // MLIP-3 .cfg block (:401-479), rank-ordered text gathered on rank 0
  void host_grades_current();         // fetches the device-resident neighbourhood grades once per grade step
  void compute_host_buffers(int eflag, int vflag, bool want_grade, double *ev);
#ifdef LMP_KOKKOS
  void compute_device_views(int eflag, int vflag, bool want_grade, double *ev);
#endif
  [[noreturn]] void fatal_one(const char *file, int line, int rc);    // this rank only (a device or species error)
  [[noreturn]] void fatal_all(const char *file, int line, int rc);    // every rank fails alike (settings, file parsing)

  mtp_handle *handle = nullptr;
  int variant;               // MTP_VARIANT_LARGE / MTP_VARIANT_SMALL
  bool extrapolation;        // one of the mtp/extrapolation styles
  int chunksize = 32768;

  // mirrors of the potential's shape (pair_mtp.h:47-70)
  int species_count = 0, coeff_count = 0, configuration_mode = 0;
  double max_cutoff = 0.0;

  // extrapolation state (pair_mtp_extrapolation.h:46-59)
  int extrapolation_flag = 0;    // set by fix pair through extract() (MUST BE INT)
  bool mlip3_style = false;
  double select_threshold = 0.0, break_threshold = 0.0, max_grade = 0.0;
  int nbh_count = 0;
  double *nbh_extrapolation_grades = nullptr;    // host copy, by atom id; filled lazily (host_grades_current)
  bool host_grades_stale = true;
  int grade_rows = 0;                            // rows the last grade step covered
  std::vector<double> cfg_candidate;
  FILE *preselected_file = nullptr;

  // host-list staging for a non-KOKKOS LAMMPS (list->firstneigh is paged): flattened on re-neighboring
  std::vector<int> flat_neigh;
  std::vector<long long> flat_offsets;
#ifdef LMP_KOKKOS
  // LAMMPS-KOKKOS flavour: per-atom outputs are DualViews like the reference's (pair_mtp_kokkos.h:128-133)
  DAT::tdual_efloat_1d k_eatom, k_grades;
  DAT::tdual_virial_array k_vatom;
  double **vatom_rows_kk = nullptr;
  double *ev_pinned = nullptr;                   // [8] written by the device, read after mtp_synchronize
#endif
};

class PairMTPB200Large : public PairMTPB200 {
 public:
  explicit PairMTPB200Large(class LAMMPS *lmp) : PairMTPB200(lmp, 0, false) {}
};
class PairMTPB200Small : public PairMTPB200 {
 public:
  explicit PairMTPB200Small(class LAMMPS *lmp) : PairMTPB200(lmp, 1, false) {}
};
class PairMTPB200ExtrapolationLarge : public PairMTPB200 {
 public:
  explicit PairMTPB200ExtrapolationLarge(class LAMMPS *lmp) : PairMTPB200(lmp, 0, true) {}
};
class PairMTPB200ExtrapolationSmall : public PairMTPB200 {
 public:
  explicit PairMTPB200ExtrapolationSmall(class LAMMPS *lmp) : PairMTPB200(lmp, 1, true) {}
};

}    // namespace LAMMPS_NS

#endif
#endif
