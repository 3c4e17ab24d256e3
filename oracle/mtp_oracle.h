/* TEST INFRASTRUCTURE -- parity checker, not part of the shipped product.
 *
 * Plain-C restatement of the reference's CPU Moment Tensor Potential pair styles
 * (`mtp`: /root/reference/LAMMPS/ML-MTP/pair_mtp.cpp:72-280; `mtp/extrapolation`:
 * pair_mtp_extrapolation.cpp:68-382; Chebyshev basis: mtp_rb_chevbyshev_basis.cpp:29-54).
 *
 * PARITY PIN: the reference ships no tests or golden vectors, so this oracle is pinned
 * against the reference ITSELF: oracle/_ref/libmtp_ref.so is the reference's unmodified
 * sources compiled against oracle/lammps_shim/, and tests/golden/ holds outputs generated
 * from it (tests/golden/make_golden.py).  tests/test_oracle.py checks this restatement
 * against those fixtures and, when the .so is present, against the reference directly
 * (bit-exact: same expression order, -O2 -ffp-contract=off).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 */
#ifndef MTP_ORACLE_H
#define MTP_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int species_count;
  int radial_func_count;        /* R */
  int radial_basis_size;        /* B */
  int alpha_moment_count;       /* M */
  int alpha_index_basic_count;  /* K */
  int alpha_index_times_count;  /* T */
  int alpha_scalar_count;       /* A */
  int max_alpha_index_basic;    /* P = 1 + max rank */
  double min_cutoff, max_cutoff, scaling;
  const double *radial_basis_coeffs; /* [S][S][R][B] */
  const int *alpha_index_basic;      /* [K][4] mu,ax,ay,az */
  const int *alpha_index_times;      /* [T][4] a0,a1,mult,a3 */
  const int *alpha_moment_mapping;   /* [A] */
  const double *species_coeffs;      /* [S] */
  const double *linear_coeffs;       /* [A] */
  /* extrapolation (may be 0/NULL for plain mtp) */
  int coeff_count;                   /* Q = S*S*R*B + S + A */
  int configuration_mode;
  const double *inverse_active_set;  /* [Q][Q] row-major */
} mtp_oracle_params;

/* Chebyshev radial basis values and derivatives (mtp_rb_chevbyshev_basis.cpp:29-54). */
void mtp_oracle_chebyshev(double dist, double min_cutoff, double max_cutoff, double scaling, int size,
                          double *vals, double *ders);

/* One force evaluation with the reference's semantics.
 *   neigh_flat/neigh_offsets: row of atom i starts at neigh_flat[neigh_offsets[i]], numneigh[i] entries,
 *   entries are masked with NEIGHMASK (0x1FFFFFFF) like pair_mtp.cpp:114.
 *   f [nall][3] is accumulated into.  ev[0] += energy, ev[1..6] += virial (xx,yy,zz,xy,xz,yz);
 *   ev[7] = max grade (grade steps; cfg mode: grade/natoms with natoms = natoms_total).
 *   eatom[i] assigned (eflag&2), vatom[i][6] accumulated (vflag&4); either may be NULL.
 *   grade_flag != 0 selects the mtp/extrapolation path; grades[i] (by atom index) receives the
 *   neighbourhood grade; candidate[Q] receives the configuration-mode candidate vector (cfg mode)
 *   or the LAST atom's candidate vector (nbh mode).
 *   mask_flat (optional, same indexing as neigh_flat) receives within_cutoff (1/0).
 * Returns 0, or -1 if a type exceeds species_count (pair_mtp.cpp:91-93,116-118). */
int mtp_oracle_compute(const mtp_oracle_params *p, int nall, const double *x, const int *type, int inum,
                       const int *ilist, const int *numneigh, const int *neigh_flat,
                       const long *neigh_offsets, int eflag, int vflag, int grade_flag,
                       long natoms_total, double *f, double *eatom, double *vatom, double *ev,
                       double *grades, double *candidate, unsigned char *mask_flat);

/* max_i |sum_j inv[i][j] b[j]|  (pair_mtp_extrapolation.cpp:347-358) */
double mtp_oracle_grade(const double *inverse_active_set, const double *b, int q);

#ifdef __cplusplus
}
#endif
#endif
