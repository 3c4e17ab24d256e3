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
  void compile_grades();     // pair_mtp_extrapolation.cpp:363-382
  void evaluate_grades();    // :387-396
  void write_config();       // :401-479
  void fatal(const char *file, int line, int rc);

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
  double *nbh_extrapolation_grades = nullptr;
  std::vector<double> cfg_candidate;
  FILE *preselected_file = nullptr;
  std::string write_buffer;

  // host-list staging for a non-KOKKOS LAMMPS (list->firstneigh is paged): flattened on re-neighboring
  std::vector<int> flat_neigh;
  std::vector<long long> flat_offsets;
  std::vector<double> fbuf;
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
