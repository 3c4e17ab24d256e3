/* mtp_b200.h -- C ABI of the B200-native Moment Tensor Potential pair-style compute.
 *
 * This is the drop-in boundary for ONE path of RichardZJM/lammps-mtp-kokkos: the per-step
 * energy / force / virial / extrapolation-grade evaluation that the reference performs in
 *   PairMTP::compute                      LAMMPS/ML-MTP/pair_mtp.cpp:72-280          (semantics / oracle)
 *   PairMTPExtrapolation::compute         LAMMPS/ML-MTP/pair_mtp_extrapolation.cpp:68-382
 *   PairMTPKokkos<Dev>::compute           LAMMPS/KOKKOS/pair_mtp_kokkos.cpp:197-399   (mtp/kk)
 *   PairMTPsKokkos<Dev>::compute          LAMMPS/KOKKOS/pair_mtps_kokkos.cpp:223-423  (mtp/small/kk)
 *   PairMTP*ExtrapolationKokkos::compute  LAMMPS/KOKKOS/pair_mtp_extrapolation_kokkos.cpp:275-610,
 *                                         pair_mtps_extrapolation_kokkos.cpp:305-639
 * and the potential-file load the reference performs in
 *   PairMTP::read_file                    pair_mtp.cpp:335-655
 *   PairMTPExtrapolation::read_file       pair_mtp_extrapolation.cpp:528-619.
 *
 * The host side that binds these entry points is the LAMMPS PairStyle in
 * lammps-mtp-kokkos_b200/lammps/pair_mtp_b200.{h,cpp}.
 *
 * Conventions: plain C, no torch / Kokkos / LAMMPS types.  Every function that returns int
 * returns 0 on success or a negative MTP_ERR_* code; mtp_last_error() gives the text of the
 * calling thread's last failure (the PairStyle turns it into error->all, the reference's
 * fatal-error convention, pair_mtp.cpp:92,288,306,315,327).  All arithmetic is IEEE FP64.
 * There is no CPU fallback: every compute entry point requires a CUDA device (sm_100a).
 */
#ifndef MTP_B200_H
#define MTP_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MTP_B200_ABI_VERSION 4 /* 4: + mtp_codegen_source, mtp_codegen_prebuild, mtp_program_kernel_note, mtp_fetch_grades, mtp_select_grades_host, mtp_cfg_grade, mtp_alloc_pinned; mtp_compute_args.f_overwrite */

#define MTP_OK 0
#define MTP_ERR_ARG (-1)      /* bad argument */
#define MTP_ERR_FILE (-2)     /* potential file cannot be opened / parsed (message = reference wording) */
#define MTP_ERR_CUDA (-3)     /* CUDA runtime failure or no sm_100 device */
#define MTP_ERR_SPECIES (-4)  /* "Too few species count in the MTP potential!" (pair_mtp.cpp:91-93) */
#define MTP_ERR_MODE (-5)     /* request incompatible with the potential's selection mode */
#define MTP_ERR_TABLE (-6)    /* alpha_index_times is not a topologically ordered program */
#define MTP_ERR_CAPACITY (-7) /* mtp_neigh_build: a row needs more than `width` entries (see max_numneigh_out) */

/* pair_style variants (README.md:36-40 of the reference) */
#define MTP_VARIANT_LARGE 0   /* mtp/kk, mtp/extrapolation/kk              : throughput path */
#define MTP_VARIANT_SMALL 1   /* mtp/small/kk, mtp/extrapolation/small/kk  : latency path   */

typedef struct mtp_handle mtp_handle;

/* Host-side description of a potential: the tables of pair_mtp.h:47-70 (+ pair_mtp_extrapolation.h:46-59).
 * Used by mtp_create(); mtp_create_from_file() fills the same structure from an MLIP-3 .almtp file. */
typedef struct {
  int species_count;
  int radial_func_count;          /* R */
  int radial_basis_size;          /* B */
  int alpha_moment_count;         /* M */
  int alpha_index_basic_count;    /* K */
  int alpha_index_times_count;    /* T */
  int alpha_scalar_count;         /* A */
  double min_cutoff, max_cutoff, scaling;
  const double *radial_basis_coeffs;   /* [S][S][R][B]  index ((it*S+jt)*R+mu)*B+ri, pair_mtp.cpp:142-147 */
  const int *alpha_index_basic;        /* [K][4] mu,ax,ay,az */
  const int *alpha_index_times;        /* [T][4] a0,a1,mult,a3 */
  const int *alpha_moment_mapping;     /* [A] */
  const double *species_coeffs;        /* [S] */
  const double *linear_coeffs;         /* [A] (moment_coeffs) */
  /* selection state; inverse_active_set == NULL means "no extrapolation" */
  int configuration_mode;              /* energy_weight == 1, pair_mtp_extrapolation.cpp:605 */
  const double *inverse_active_set;    /* [Q][Q] row-major, Q = S*S*R*B + S + A */
} mtp_params_host;

typedef struct {
  int abi_version;
  int species_count, radial_func_count, radial_basis_size;
  int alpha_moment_count, alpha_index_basic_count, alpha_index_times_count, alpha_scalar_count;
  int max_alpha_index_basic;      /* P = 1 + max rank, pair_mtp.cpp:510-515 */
  int coeff_count;                /* Q, 0 when no selection state is loaded */
  int configuration_mode;
  int has_selection_state;
  int wave_count;                 /* dependency depth of the contraction program */
  int chunksize;
  int device;
  double min_cutoff, max_cutoff, scaling;
} mtp_info;

/* One force evaluation = one call of Pair::compute(eflag, vflag).  All pointers are DEVICE pointers
 * for mtp_compute() and HOST pointers for mtp_compute_host().
 *
 * Neighbor list (full list, pair_mtp.cpp:318): entry jj of atom i is
 *     neighbors[(neigh_offsets ? neigh_offsets[i] : i * stride_i) + jj * stride_jj] & neighmask
 * which covers both LAMMPS-KOKKOS's 2-D d_neighbors(i,jj) view (pair_mtp_kokkos.cpp:236-240,438) in
 * either layout and a CSR / paged host list (list->firstneigh, pair_mtp.cpp:84-85,113).  numneigh is
 * indexed by atom id, ilist (NULL = identity) lists the inum owned centres (pair_mtp.cpp:81-89). */
typedef struct {
  int variant;                    /* MTP_VARIANT_* */
  int inum, nall;
  const double *x;                /* [nall][3]  atom->x */
  const int *type;                /* [nall]     atom->type, 1-based */
  const int *ilist;               /* [inum] or NULL */
  const int *numneigh;            /* [>= max listed id + 1] */
  const int *neighbors;
  const long long *neigh_offsets; /* [>= max listed id + 1] or NULL */
  long long stride_i, stride_jj;
  int neighmask;                  /* NEIGHMASK = 0x1FFFFFFF; 0 means "use 0x1FFFFFFF" */
  int eflag;                      /* bit0 global energy, bit1 per-atom (LAMMPS ENERGY_GLOBAL/ATOM) */
  int vflag;                      /* !=0: pairwise virial -sym(F (x) r) like pair_mtp.cpp:257-266; bit2: per-atom */
  int want_grade;                 /* extrapolation_flag || mlip3_style (pair_mtp_extrapolation.cpp:71) */
  long long natoms_total;         /* atom->natoms, configuration-mode normalisation (:373-376); 0 = inum */
  double *f;                      /* [nall][3] accumulated into (ghost rows included) */
  double *eatom;                  /* [nall] or NULL; eatom[i] ASSIGNED for listed i (pair_mtp.cpp:210) */
  double *vatom;                  /* [nall][6] or NULL; accumulated on the centre atom (:268-276) */
  double *ev_out;                 /* [8] overwritten: E, virial xx,yy,zz,xy,xz,yz, max grade */
  double *grades;                 /* [nall] by atom id or NULL: neighbourhood grades (:335) */
  double *cfg_candidate;          /* [Q] or NULL: configuration-mode candidate vector, overwritten */
  unsigned char *within_cutoff;   /* optional: 1/0 per neighbor entry, same indexing as `neighbors` */
  void *stream;                   /* cudaStream_t (mtp_compute only) */
  int max_numneigh;               /* upper bound of numneigh[] over the listed centres, e.g. d_neighbors.extent(1)
                                     of the LAMMPS-KOKKOS list; 0 = unknown (mtp_compute then reduces numneigh on the
                                     device and waits for that one integer) */
  int f_overwrite;                /* mtp_compute_host only: != 0 = the caller's f holds zeros on entry (LAMMPS's
                                     force_clear precedes Pair::compute and this style is the first contributor), so the
                                     result is STORED into f instead of added and f is never uploaded */
} mtp_compute_args;

/* ---- lifetime ------------------------------------------------------------------------------ */
/* Parse an MLIP-3 .almtp potential (grammar of pair_mtp.cpp:345-570 + mtp_radial_basis.cpp:59-102);
 * with want_selection_state != 0 also the MaxVol selection state that must follow it
 * (pair_mtp_extrapolation.cpp:545-612).  device < 0 = current device.  NULL on failure. */
mtp_handle *mtp_create_from_file(const char *path, int want_selection_state, int device);
mtp_handle *mtp_create(const mtp_params_host *params, int device);
void mtp_destroy(mtp_handle *h);
const char *mtp_last_error(void);
int mtp_get_info(const mtp_handle *h, mtp_info *out);
/* Host-only validation of a potential file: the same parser and contraction-program compiler as
 * mtp_create_from_file(), no CUDA device touched (what PairMTP::coeff -> read_file does before any
 * compute, pair_mtp.cpp:310,335-570).  out may be NULL; out->device = -1, out->chunksize = 0. */
int mtp_potential_check(const char *path, int want_selection_state, mtp_info *out);
/* Copy the parsed tables back out (sizes from mtp_get_info); any pointer may be NULL. */
int mtp_get_tables(const mtp_handle *h, double *radial_basis_coeffs, int *alpha_index_basic,
                   int *alpha_index_times, int *alpha_moment_mapping, double *species_coeffs,
                   double *linear_coeffs, double *inverse_active_set);

/* "chunksize n" keyword (pair_mtp_kokkos.cpp:113-117, README.md:44): upper bound on the rows of
 * per-atom scratch that live in HBM at once (only the grade path keeps any; the force path keeps
 * all per-atom state on chip).  Must be >= 1. */
int mtp_set_chunksize(mtp_handle *h, int chunksize);

/* Number of internal streams ("lanes") the super-chunks of one mtp_compute() are dealt to (1..4, default 2): with
 * more than one lane the shared-memory-bound contraction-program kernel of one chunk overlaps the FP64-bound pair
 * kernels of its neighbours.  1 = all kernels serialised on the caller's stream (used for per-kernel timing). */
int mtp_set_lanes(mtp_handle *h, int lanes);

/* ---- the hot path ---------------------------------------------------------------------------- */
/* Asynchronous on args->stream; results are valid after the stream is synchronised. */
int mtp_compute(mtp_handle *h, const mtp_compute_args *device_args);
/* The same evaluation with the listed centres cut into `nphase` consecutive runs of ilist (phase_inum[p] centres each,
 * adding up to args->inum).  The kernels of phase p start only after wait_events[p] (a cudaEvent_t, or NULL) has
 * completed, and done_events[p] (cudaEvent_t or NULL) is recorded when all work of phases 0..p is complete.  This is
 * how a multi-GPU caller hides the ghost halo (LAMMPS Comm::forward_comm / reverse_comm around pair_mtp.cpp:72-280)
 * behind the pair style: list the interior centres -- those without a ghost in their neighbor list -- first and last,
 * the boundary centres in between with wait = "ghost positions arrived"; when the boundary phase is done the ghost
 * forces can travel back while the remaining interior centres are evaluated.  One energy / virial record, one set of
 * lanes: unlike separate calls per run, the phases share the library's internal streams. */
int mtp_compute_phased(mtp_handle *h, const mtp_compute_args *device_args, int nphase, const int *phase_inum,
                       void *const *wait_events, void *const *done_events);
/* Waits for the device and reports deferred errors (species bound check, pair_mtp.cpp:91-93). */
int mtp_synchronize(mtp_handle *h);
/* Same evaluation with HOST buffers: copies x (every call) and, when list_changed != 0 -- LAMMPS re-neighboring
 * steps, neighbor->ago == 0, the only steps on which atoms migrate, are sorted or change type -- also type and the
 * neighbor list to the device, runs mtp_compute, adds the result into the host f / eatom / vatom / grades and returns
 * after the copies complete.  Between re-neighboring steps the list and the types stay resident on the device. */
int mtp_compute_host(mtp_handle *h, const mtp_compute_args *host_args, int list_changed);

/* ---- ghost-atom halo helpers (device) ---------------------------------------------------------- */
/* forward: out[k][:] = x[sendlist[k]][:] + shift[:]   (LAMMPS Comm::forward_comm pack, on device) */
int mtp_halo_pack_x(const double *x, const int *sendlist, int n, const double *shift3_host, double *out,
                    void *stream);
/* forward, several swaps at once: out[k][:] = x[sendlist[k]][:] + shifts[seg[k]][:]   (shifts: DEVICE array [nseg][3]) */
int mtp_halo_pack_x_multi(const double *x, const int *sendlist, const unsigned char *seg, const double *shifts_dev, int n,
                          double *out, void *stream);
/* reverse: f[sendlist[k]][:] += buf[k][:]             (Comm::reverse_comm unpack, newton on) */
int mtp_halo_unpack_add_f(double *f, const int *sendlist, int n, const double *buf, void *stream);

/* ---- device neighbor-list build (the step before the path) ---------------------- */
/* FULL neighbor list as the pair style requests it (pair_mtp.cpp:318 NeighConst::REQ_FULL; consumed through
 * list->ilist / numneigh / firstneigh, pair_mtp.cpp:77-85, or the 2-D d_neighbors view, pair_mtp_kokkos.cpp:236-240):
 * for every owned atom i < nlocal all atoms j != i, owned or ghost (j < nall), with |x_j - x_i|^2 <= cutneigh^2,
 * cutneigh = cutoff + skin.  rsq is rounded exactly as the CPU builds round it, so the neighbor SET of every atom is
 * bit-identical to a host-built list; the order within a row is deterministic (stencil order).
 * x: DEVICE [nall][3].  Outputs (DEVICE): numneigh[nlocal] = true count of every row, neighbors = row-major table
 * [nlocal][width]; hand it to mtp_compute with neighbors / neigh_offsets = NULL / stride_i = width / stride_jj = 1.
 * max_numneigh_out (HOST, optional) receives the longest row.  Returns MTP_ERR_CAPACITY when that exceeds width
 * (rows are truncated; numneigh and *max_numneigh_out are valid): rebuild with a wider table, as LAMMPS-KOKKOS does.
 * Blocks until the list is complete (two small device -> host reads: bounding box, longest row). */
int mtp_neigh_build(mtp_handle *h, int nlocal, int nall, const double *x, double cutneigh, int *numneigh, int *neighbors,
                    int width, int *max_numneigh_out, void *stream);

/* ---- device-side selection of extrapolating atoms (caller side of the grade path) ---- */
/* indices_out[0 .. *count_out) = ascending ids i < n with grades[i] >= threshold -- the comparison of
 * PairMTPExtrapolation::evaluate_grades (pair_mtp_extrapolation.cpp:389-390) applied per atom to the neighborhood
 * grades that `fix pair` exports (extract_peratom "extrapolation", :641-652) -- so that a caller copies only the
 * selected atoms to the host instead of the whole per-atom array.  grades, indices_out: DEVICE ([n]); count_out: HOST.
 * Blocks until the count is known. */
int mtp_select_grades(mtp_handle *h, const double *grades, int n, double threshold, int *indices_out, int *count_out,
                      void *stream);

/* Host-buffer flavour of the same (plain LAMMPS): the neighbourhood grades of the last grade step of mtp_compute_host()
 * stay resident on the device (they are copied to the host only when args->grades is given).
 *   mtp_fetch_grades        copies the first n of them to the host -- what extract_peratom("extrapolation") hands to
 *                           `fix pair` and what the .cfg writer needs when a configuration is selected
 *                           (pair_mtp_extrapolation.cpp:641-652, 418-425);
 *   mtp_select_grades_host  ids (ascending) and grades of the atoms i < n with grade >= threshold: only those cross PCIe.
 *                           *count_out = number selected; at most cap of them are written. */
int mtp_fetch_grades(mtp_handle *h, double *grades_host, int n);
int mtp_select_grades_host(mtp_handle *h, int n, double threshold, int *ids_out, double *grades_out, int cap, int *count_out);
/* Configuration-mode grade of a candidate vector that was summed over ranks on the host (MPI_Allreduce,
 * pair_mtp_extrapolation.cpp:366-376): max_i |Ainv[i,:] . b| / natoms_total, evaluated on the device against the
 * resident inverse active set. */
int mtp_cfg_grade(mtp_handle *h, const double *candidate_host, long long natoms_total, double *grade_out);
/* page-locked host memory for buffers the device writes directly (ev_out of mtp_compute in the LAMMPS-KOKKOS flavour) */
void *mtp_alloc_pinned(size_t bytes);
void mtp_free_pinned(void *p);

/* ---- velocity-Verlet half steps on the device (the steps either side of the path) ---- */
/* Upstream FixNVE::initial_integrate / final_integrate, which the reference's example deck runs around the pair style
 * (`fix 1 all nve`, README.md:148-149): dtfm = dtf / mass[type[i]]; v += dtfm * f; x += dtv * v (initial) and
 * v += dtfm * f (final) for the nlocal owned atoms.  All pointers DEVICE; mass is indexed by the 1-based type
 * ([ntypes + 1]); dtf = 0.5 * dt * force->ftm2v, dtv = dt.  x_at_build / moved_flag (optional, both or neither):
 * *moved_flag is set to 1 when an atom is farther than trigger_dist from x_at_build (LAMMPS's re-neighboring
 * criterion, half the skin).  Products and sums are rounded separately, so a host replay is bit-exact. */
int mtp_nve_initial_integrate(int nlocal, double *x, double *v, const double *f, const int *type, const double *mass,
                              double dtf, double dtv, const double *x_at_build, double trigger_dist, int *moved_flag,
                              void *stream);
int mtp_nve_final_integrate(int nlocal, double *v, const double *f, const int *type, const double *mass, double dtf,
                            void *stream);

/* ---- measured roofs ----------------------------------------------------------------------------- */
/* FP64 vector (DFMA) and FP64 tensor (mma.sync.m8n8k4.f64, DMMA) peak of the device, in TFLOP/s. */
int mtp_fp64_peak(int device, double *dfma_tflops, double *dmma_tflops);

/* Per-kernel-class device timing for bench.py's roofline: when enabled, mtp_compute() brackets the launches
 * of each class with CUDA events on the launch stream; mtp_profile_read() synchronises the device, returns
 * the accumulated milliseconds and span counts per class since the previous read, and resets them. */
#define MTP_PROF_PACK 0      /* x/type -> 32-byte records */
#define MTP_PROF_GATHER 1    /* neighbor gather + cutoff mask + radial basis */
#define MTP_PROF_MOMENTS 2   /* basic moments */
#define MTP_PROF_PROGRAM 3   /* contraction program forward, site energy, reverse mode */
#define MTP_PROF_FORCES 4    /* per-pair forces, scatter, virial, candidate vector */
#define MTP_PROF_GRADE 5     /* extrapolation grade (DMMA) */
#define MTP_PROF_FINALIZE 6  /* energy / virial reduction */
#define MTP_PROF_SITE 7      /* generic fused site kernel (non-standard potentials) */
#define MTP_PROF_CLASSES 8
int mtp_profile_enable(mtp_handle *h, int on);
int mtp_profile_read(mtp_handle *h, double *ms /*[MTP_PROF_CLASSES]*/, long long *count /*[MTP_PROF_CLASSES]*/);

/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
long long mtp_kernel_launch_count(void);

/* Host-only self-check of the contraction-program stream packer: parses the potential, builds the grouped term
 * streams for `atoms_per_cta` (32 or 16) and interprets them on the CPU against the sequential program
 * (pair_mtp.cpp:196-233).  *max_rel_err_out receives the largest relative deviation (moments and adjoints). */
int mtp_program_check(const char *path, int atoms_per_cta, double *max_rel_err_out);

/* ---- generated contraction-program kernel ------------------------------------------------------------------------
 * The alpha_index_times program of a potential (pair_mtp.cpp:196-233; executed by the reference as a serial loop per
 * thread, pair_mtp_kokkos.cpp:550-592, or in three atomic waves, pair_mtps_kokkos.cpp:572-639) is compiled into one
 * straight-line sm_100a kernel per potential STRUCTURE when the potential is loaded (NVRTC; cubins are cached under
 * $MTP_B200_KCACHE or <library directory>/kcache).  The two entry points below run the generator without a device:
 * mtp_codegen_source returns the emitted CUDA source (tests compile it for the host and run it against the sequential
 * program), mtp_codegen_prebuild compiles it and stores the cubin in the cache (build step).
 * info_out[13] = atoms per CTA, warps, CTAs per SM, shared-memory rows, stages, shared-memory bytes, term steps per
 * atom, row loads per chunk, row stores per chunk, critical-path term steps, rows of mb/gb, structure hash, rounds
 * (> 1: a program whose rows exceed a CTA's shared memory is evaluated in rounds of basis functions that reuse them). */
int mtp_codegen_source(const char *path, int latency_shape, char *buf, long long cap, long long *needed, long long *info_out);
int mtp_codegen_prebuild(const char *path, int latency_shape, int *compiled_out);
/* empty when the generated kernel serves this handle, else the reason it does not (the interpreting kernels run) */
const char *mtp_program_kernel_note(const mtp_handle *h);
const char *mtp_program_kernel_note_small(const mtp_handle *h);    /* same for the latency shape (mtp/small/kk) */

/* Which kernels the last mtp_compute() of this handle launched (tests assert that the intended path ran):
 * bits 0-3 kernel family (0 generic fused site kernel, 1 DMMA-moment pipeline, 2 register-resident pair kernels),
 * bit 4 set when the contraction program ran in its 4-atoms-per-lane interpreting form, bit 5 when the generated
 * per-potential kernel ran, bits 8-15 atoms per CTA of the program kernel. */
int mtp_last_kernel_path(const mtp_handle *h);

#ifdef __cplusplus
}
#endif
#endif
