/* TEST INFRASTRUCTURE -- parity checker, not part of the shipped product.
 * See mtp_oracle.h for scope, provenance and the parity pin.
 *
 * The arithmetic keeps the reference's expression order so that, compiled with the same
 * flags (-O2 -ffp-contract=off), results are bit-identical to oracle/_ref/libmtp_ref.so.
 */
#include "mtp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_NEIGHMASK 0x1FFFFFFF

/* mtp_rb_chevbyshev_basis.cpp:29-38 (values) and :40-54 (derivatives) */
void mtp_oracle_chebyshev(double dist, double rmin, double rmax, double scaling, int size, double *vals,
                          double *ders)
{
  const double ksi = (2 * dist - (rmin + rmax)) / (rmax - rmin);
  const double mult = 2.0 / (rmax - rmin);
  const double t = dist - rmax;
  vals[0] = scaling * (1 * t * t);
  if (size > 1) vals[1] = scaling * (ksi * t * t);
  for (int i = 2; i < size; i++) vals[i] = 2 * ksi * vals[i - 1] - vals[i - 2];
  ders[0] = scaling * 2 * t;
  if (size > 1) ders[1] = scaling * (mult * t * t + 2 * ksi * t);
  for (int i = 2; i < size; i++) ders[i] = 2 * (mult * vals[i - 1] + ksi * ders[i - 1]) - ders[i - 2];
}

/* pair_mtp_extrapolation.cpp:347-358 */
double mtp_oracle_grade(const double *inv, const double *b, int q)
{
  double best = 0;
  for (int i = 0; i < q; i++) {
    double acc = 0;
    for (int j = 0; j < q; j++) acc += b[j] * inv[(size_t) i * q + j];
    best = fmax(fabs(acc), best);
  }
  return best;
}

typedef struct {
  double *m, *g;         /* moments and adjoints, [M] */
  double *jac;           /* [jcap][K][3] */
  unsigned char *within; /* [jcap] */
  int jcap;
  double *dpow, *cpow;   /* [P], [P][3] */
  double *rvals, *rders; /* [R] */
  double *bvals, *bders; /* [B] */
  double *rjac;          /* [K][S][R*B] (grade steps) */
  double *cand;          /* [Q] */
} work_t;

static int work_init(work_t *w, const mtp_oracle_params *p, int grade)
{
  memset(w, 0, sizeof(*w));
  const int M = p->alpha_moment_count, P = p->max_alpha_index_basic;
  w->m = (double *) calloc((size_t) (M > 0 ? M : 1), sizeof(double));
  w->g = (double *) calloc((size_t) (M > 0 ? M : 1), sizeof(double));
  w->dpow = (double *) calloc((size_t) P, sizeof(double));
  w->cpow = (double *) calloc((size_t) P * 3, sizeof(double));
  w->rvals = (double *) calloc((size_t) p->radial_func_count, sizeof(double));
  w->rders = (double *) calloc((size_t) p->radial_func_count, sizeof(double));
  w->bvals = (double *) calloc((size_t) p->radial_basis_size, sizeof(double));
  w->bders = (double *) calloc((size_t) p->radial_basis_size, sizeof(double));
  if (grade) {
    size_t n = (size_t) p->alpha_index_basic_count * p->species_count * p->radial_func_count *
        p->radial_basis_size;
    w->rjac = (double *) calloc(n ? n : 1, sizeof(double));
    w->cand = (double *) calloc((size_t) (p->coeff_count > 0 ? p->coeff_count : 1), sizeof(double));
  }
  /* pair_mtp.cpp:647: zeroth powers are the constant 1 */
  w->dpow[0] = w->cpow[0] = w->cpow[1] = w->cpow[2] = 1;
  return 0;
}

static void work_free(work_t *w)
{
  free(w->m); free(w->g); free(w->jac); free(w->within); free(w->dpow); free(w->cpow);
  free(w->rvals); free(w->rders); free(w->bvals); free(w->bders); free(w->rjac); free(w->cand);
}

int mtp_oracle_compute(const mtp_oracle_params *p, int nall, const double *x, const int *type, int inum,
                       const int *ilist, const int *numneigh, const int *neigh_flat,
                       const long *neigh_offsets, int eflag, int vflag, int grade_flag,
                       long natoms_total, double *f, double *eatom, double *vatom, double *ev,
                       double *grades, double *candidate, unsigned char *mask_flat)
{
  (void) nall;
  const int S = p->species_count, R = p->radial_func_count, B = p->radial_basis_size;
  const int M = p->alpha_moment_count, K = p->alpha_index_basic_count, T = p->alpha_index_times_count;
  const int A = p->alpha_scalar_count, P = p->max_alpha_index_basic, Q = p->coeff_count;
  const int RB = R * B;                 /* radial_coeff_count_per_pair */
  const int radial_coeff_count = S * S * RB;
  const double cutsq = p->max_cutoff * p->max_cutoff;   /* pair_mtp.cpp:449,456: one cutoff for all pairs */
  const int (*basic)[4] = (const int (*)[4]) p->alpha_index_basic;
  const int (*times)[4] = (const int (*)[4]) p->alpha_index_times;
  const int eflag_global = eflag & 1, eflag_atom = eflag & 2, vflag_atom = vflag & 4;
  const int cfg = grade_flag && p->configuration_mode;
  double max_grade = 0;

  work_t w;
  work_init(&w, p, grade_flag);
  if (cfg) memset(w.cand, 0, sizeof(double) * (size_t) Q);   /* pair_mtp_extrapolation.cpp:97-98 */

  int status = 0;
  for (int ii = 0; ii < inum && status == 0; ii++) {          /* pair_mtp.cpp:88 */
    const int i = ilist[ii];
    const int itype = type[i] - 1;
    if (itype >= S) { status = -1; break; }
    const int jnum = numneigh[i];
    const int *row = neigh_flat + neigh_offsets[i];
    const double xi[3] = {x[3 * (size_t) i], x[3 * (size_t) i + 1], x[3 * (size_t) i + 2]};

    if (w.jcap < jnum) {                                       /* pair_mtp.cpp:99-104 */
      w.jac = (double *) realloc(w.jac, sizeof(double) * (size_t) jnum * (K > 0 ? K : 1) * 3);
      w.within = (unsigned char *) realloc(w.within, (size_t) jnum);
      w.jcap = jnum;
    }
    memset(w.m, 0, sizeof(double) * (size_t) M);
    memset(w.g, 0, sizeof(double) * (size_t) M);
    if (grade_flag) {
      memset(w.rjac, 0, sizeof(double) * (size_t) K * S * RB);
      if (!cfg) memset(w.cand, 0, sizeof(double) * (size_t) Q);
    }

    /* ---- basic moments and their Jacobian: pair_mtp.cpp:112-193 ---- */
    for (int jj = 0; jj < jnum; jj++) {
      const int j = row[jj] & ORACLE_NEIGHMASK;
      const int jtype = type[j] - 1;
      if (jtype >= S) { status = -1; break; }
      const double r[3] = {x[3 * (size_t) j] - xi[0], x[3 * (size_t) j + 1] - xi[1],
                           x[3 * (size_t) j + 2] - xi[2]};
      const double rsq = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      if (rsq > cutsq) {
        w.within[jj] = 0;
        if (mask_flat) mask_flat[neigh_offsets[i] + jj] = 0;
        continue;
      }
      w.within[jj] = 1;
      if (mask_flat) mask_flat[neigh_offsets[i] + jj] = 1;

      const double dist = sqrt(rsq);
      mtp_oracle_chebyshev(dist, p->min_cutoff, p->max_cutoff, p->scaling, B, w.bvals, w.bders);

      for (int k = 1; k < P; k++) {                              /* :133-136 */
        w.dpow[k] = w.dpow[k - 1] * dist;
        for (int a = 0; a < 3; a++) w.cpow[3 * k + a] = w.cpow[3 * (k - 1) + a] * r[a];
      }
      for (int mu = 0; mu < R; mu++) {                           /* :139-151 */
        double val = 0, der = 0;
        const int offset = ((itype * S + jtype) * RB) + mu * B;
        for (int ri = 0; ri < B; ri++) {
          val += p->radial_basis_coeffs[offset + ri] * w.bvals[ri];
          der += p->radial_basis_coeffs[offset + ri] * w.bders[ri];
        }
        w.rvals[mu] = val;
        w.rders[mu] = der;
      }
      double *J = w.jac + (size_t) jj * K * 3;
      for (int k = 0; k < K; k++) {                              /* :154-192 */
        const int mu = basic[k][0], a0 = basic[k][1], a1 = basic[k][2], a2 = basic[k][3];
        double val = w.rvals[mu];
        double der = w.rders[mu];
        const int rank = a0 + a1 + a2;
        const double norm_fac = 1.0 / w.dpow[rank];
        const double pow0 = w.cpow[3 * a0 + 0];
        const double pow1 = w.cpow[3 * a1 + 1];
        const double pow2 = w.cpow[3 * a2 + 2];
        double pw = pow0 * pow1 * pow2;
        if (grade_flag) {                                        /* pair_mtp_extrapolation.cpp:193-198 */
          double *rj = w.rjac + ((size_t) k * S + jtype) * RB + mu * B;
          for (int ri = 0; ri < B; ri++) rj[ri] += w.bvals[ri] * norm_fac * pw;
        }
        val *= norm_fac;
        der = der * norm_fac - rank * val / dist;
        w.m[k] += val * pw;
        pw *= der / dist;
        J[3 * k + 0] = pw * r[0];
        J[3 * k + 1] = pw * r[1];
        J[3 * k + 2] = pw * r[2];
        if (a0 != 0) J[3 * k + 0] += val * a0 * w.cpow[3 * (a0 - 1) + 0] * pow1 * pow2;
        if (a1 != 0) J[3 * k + 1] += val * a1 * pow0 * w.cpow[3 * (a1 - 1) + 1] * pow2;
        if (a2 != 0) J[3 * k + 2] += val * a2 * pow0 * pow1 * w.cpow[3 * (a2 - 1) + 2];
      }
    }
    if (status) break;

    /* ---- contraction program: pair_mtp.cpp:196-201 ---- */
    for (int t = 0; t < T; t++) {
      const double v0 = w.m[times[t][0]];
      const double v1 = w.m[times[t][1]];
      const int mult = times[t][2];
      w.m[times[t][3]] += mult * v0 * v1;
    }

    /* ---- site energy (+ linear part of the candidate vector): pair_mtp.cpp:204-212,
     *      pair_mtp_extrapolation.cpp:235-252 ---- */
    if (grade_flag) {
      const int lin = radial_coeff_count + S;
      if (eflag) {
        double e = p->species_coeffs[itype];
        for (int k = 0; k < A; k++) {
          const double bm = w.m[p->alpha_moment_mapping[k]];
          w.cand[lin + k] += bm;
          e += p->linear_coeffs[k] * bm;
        }
        if (eflag_atom && eatom) eatom[i] = e;
        if (eflag_global) ev[0] += e;
      } else {
        for (int k = 0; k < A; k++) w.cand[lin + k] += w.m[p->alpha_moment_mapping[k]];
      }
      w.cand[radial_coeff_count + itype] += 1;
    } else if (eflag_atom || eflag_global) {
      double e = p->species_coeffs[itype];
      for (int k = 0; k < A; k++) e += p->linear_coeffs[k] * w.m[p->alpha_moment_mapping[k]];
      if (eflag_atom && eatom) eatom[i] = e;
      if (eflag_global) ev[0] += e;
    }

    /* ---- reverse mode through the program: pair_mtp.cpp:217-233 ---- */
    for (int k = 0; k < A; k++) w.g[p->alpha_moment_mapping[k]] = p->linear_coeffs[k];
    for (int t = T - 1; t >= 0; t--) {
      const int a0 = times[t][0], a1 = times[t][1], mult = times[t][2], a3 = times[t][3];
      const double v0 = w.m[a0], v1 = w.m[a1], v3 = w.g[a3];
      w.g[a1] += v3 * mult * v0;
      w.g[a0] += v3 * mult * v1;
    }

    /* ---- forces and virial: pair_mtp.cpp:236-278 ---- */
    for (int jj = 0; jj < jnum; jj++) {
      const int j = row[jj] & ORACLE_NEIGHMASK;
      if (!w.within[jj]) continue;
      const double *J = w.jac + (size_t) jj * K * 3;
      double tf[3] = {0, 0, 0};
      for (int k = 0; k < K; k++)
        for (int a = 0; a < 3; a++) tf[a] += w.g[k] * J[3 * k + a];
      for (int a = 0; a < 3; a++) {
        f[3 * (size_t) i + a] += tf[a];
        f[3 * (size_t) j + a] -= tf[a];
      }
      if (vflag) {
        const double r[3] = {x[3 * (size_t) j] - xi[0], x[3 * (size_t) j + 1] - xi[1],
                             x[3 * (size_t) j + 2] - xi[2]};
        double v[6];
        v[0] = tf[0] * r[0];
        v[1] = tf[1] * r[1];
        v[2] = tf[2] * r[2];
        v[3] = (tf[0] * r[1] + tf[1] * r[0]) / 2;
        v[4] = (tf[0] * r[2] + tf[2] * r[0]) / 2;
        v[5] = (tf[1] * r[2] + tf[2] * r[1]) / 2;
        for (int c = 0; c < 6; c++) ev[1 + c] -= v[c];
        if (vflag_atom && vatom)
          for (int c = 0; c < 6; c++) vatom[6 * (size_t) i + c] -= v[c];
      }
    }

    if (grade_flag) {
      /* radial part of the candidate vector: pair_mtp_extrapolation.cpp:322-329 */
      for (int k = 0; k < K; k++)
        for (int jt = 0; jt < S; jt++) {
          const int offset = (itype * S + jt) * RB;
          const double *rj = w.rjac + ((size_t) k * S + jt) * RB;
          for (int ri = 0; ri < RB; ri++) w.cand[offset + ri] += w.g[k] * rj[ri];
        }
      if (!cfg) {                                                /* :331-336 */
        const double grade = mtp_oracle_grade(p->inverse_active_set, w.cand, Q);
        max_grade = fmax(grade, max_grade);
        if (grades) grades[i] = grade;
      }
    }
  }

  if (grade_flag && status == 0) {                               /* compile_grades, :363-382 */
    if (cfg) {
      max_grade = mtp_oracle_grade(p->inverse_active_set, w.cand, Q);
      if (natoms_total > 0) max_grade /= natoms_total;
      else
        max_grade = 0.0;
    }
    ev[7] = max_grade;
    if (candidate) memcpy(candidate, w.cand, sizeof(double) * (size_t) Q);
  }
  work_free(&w);
  return status;
}
