// Register-resident pair stages for standard MTP shapes (template <DEG = max tensor rank, RP = padded R>).
//
// Same per-atom pipeline as mtp_site_kernel (mtp_kernels.cu) -- gather/mask, Chebyshev x cutoff, radial
// contraction, basic moments, contraction program forward / reverse, force + virial scatter -- but the two
// O(N_c * K) stages no longer touch shared memory per (neighbor, moment):
//
//  * FORWARD  m[mu][q] = sum_n f_mu(d_n) u_n^q is a small GEMM whose reduction runs over the neighbors.  It is
//    issued on the FP64 tensor pipe: mma.sync.m8n8k4.f64 (DMMA.8x8x4, the only native FP64 MMA shape on
//    sm_100a) with A = f^T (rows = mu, k = 4 neighbors) and B = monomials (k = 4 neighbors, cols = 8
//    monomials).  Monomials are grouped in blocks of 8 that share their even part:
//        x^a y^b z^c = (x^2)^a' (y^2)^b' (z^2)^c'  *  x^pa y^pb z^pc ,   (pa,pb,pc) in {0,1}^3 = column,
//    so every lane builds its B element with two multiplies from registers (block factor E_t, identical code
//    in all lanes, times its own parity factor w) -- no shared-memory operand traffic, no cross-lane
//    reduction: the tensor core does the sum over neighbors.
//  * BACKWARD  lane = neighbor.  With W(q) = sum_mu (f_mu/d) g[mu][q] and W'(q) = sum_mu f'_mu g[mu][q]
//        F = u (P_W'(u) - u . grad P_W(u)) + grad P_W(u)
//    (Euler's identity folds the -rho f_mu/d term of dm_k/dr into the gradient), and P_W', grad P_W are
//    evaluated by a fully unrolled three-level Horner scheme in registers; g is read as warp-wide
//    broadcasts in canonical (q, mu) order.
#pragma once

#include "mtp_device.cuh"

namespace mtpb200 {

// ---- compile-time monomial enumerations -------------------------------------------------------------
__host__ __device__ constexpr int tri(int n) { return n < 0 ? 0 : (n + 1) * (n + 2) / 2; }
__host__ __device__ constexpr int tet(int n) { return n < 0 ? 0 : (n + 1) * (n + 2) * (n + 3) / 6; }

// canonical index of x^a y^b z^c among all monomials of total degree <= deg, (a, b, c) lexicographic
__host__ __device__ constexpr int canon_index(int deg, int a, int b, int c)
{
  int idx = 0;
  for (int aa = 0; aa < a; aa++) idx += tri(deg - aa);
  for (int bb = 0; bb < b; bb++) idx += deg - a - bb + 1;
  return idx + c;
}
// index of the even-part block (a', b', c'), a'+b'+c' <= smax, same lexicographic order
__host__ __device__ constexpr int block_index(int smax, int a, int b, int c) { return canon_index(smax, a, b, c); }

struct V1Tables {
  const short *fwd_slot;    // [NB][32][2]  basic moment index held by (block, lane, c-register) or -1
  const short *g_src;       // [NQ][RP]     basic moment index feeding canonical slot (q, mu) or -1
  int rcnt[16];             // rcnt[d] = number of leading mu that own a basic moment of degree d
};

constexpr int V1_PEND = 64;

template <int DEG, int RP> struct V1Shape {
  static constexpr int SMAX = DEG / 2;
  static constexpr int NB = tet(SMAX);       // DMMA column blocks
  static constexpr int NQ = tet(DEG);        // canonical monomials
  static constexpr int GC = NQ * RP;         // canonical adjoint table (doubles)
  static constexpr int STAGE = (3 + RP) * 32;
};

// Shared-memory plans.  Pair kernels (moments / forces): [radial coeffs][per-warp scratch x W].
// Program kernel: [cm: M x (NA+1)][cg: M x (NA+1)] -- moments and adjoints of the NA atoms of the current
// chunk, node-major with the atom index fastest (row stride NA+1).
struct V1Layout {
  size_t radial_bytes, warp_bytes_moments, warp_bytes_forces, node_bytes;
};

template <int DEG, int RP>
__host__ __device__ inline V1Layout v1_layout(int S, int R, int B, int M, int Q, bool grade, int na)
{
  using Sh = V1Shape<DEG, RP>;
  V1Layout L;
  L.radial_bytes = ((size_t) S * S * R * B * 8 + 15) & ~(size_t) 15;
  L.node_bytes = ((size_t) M * (na + 1) * 8 + 15) & ~(size_t) 15;
  const size_t pend = (size_t) 3 * V1_PEND * 8 + (size_t) 2 * V1_PEND * 4;
  L.warp_bytes_moments = ((size_t) Sh::STAGE * 8 + pend + 15) & ~(size_t) 15;
  L.warp_bytes_forces = ((size_t) (Sh::GC + (grade ? Q : 0)) * 8 + pend + 15) & ~(size_t) 15;
  return L;
}

struct V1Warp {
  double *gc, *stage, *pr, *cand;    // canonical adjoints, DMMA operand staging, pending pairs, candidate vector
  int *pj, *pt;
};

__device__ __forceinline__ void dmma884_v1(double &c0, double &c1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// Chebyshev x cutoff and the radial contraction for one neighbor, everything in registers.
// F[mu] = f_mu(d), Fd[mu] = f'_mu(d)   (pair_mtp.cpp:139-151, mtp_rb_chevbyshev_basis.cpp:29-54)
template <int RP>
__device__ __forceinline__ void radial_functions(const DevPotential &pot, const double *c /*[R][B] for (it,jt)*/,
                                                 double d, double (&F)[RP], double (&Fd)[RP])
{
  const double t = d - pot.rmax;
  const double ksi = (2 * d - (pot.rmin + pot.rmax)) / (pot.rmax - pot.rmin);
  const double mult = 2.0 / (pot.rmax - pot.rmin);
  double v_prev = 0, v_cur = pot.scaling * (1 * t * t);
  double d_prev = 0, d_cur = pot.scaling * 2 * t;
#pragma unroll
  for (int mu = 0; mu < RP; mu++) F[mu] = Fd[mu] = 0.0;
  const int B = pot.B, R = pot.R;
  for (int ri = 0; ri < B; ri++) {
    if (ri == 1) {
      v_prev = v_cur;
      d_prev = d_cur;
      v_cur = pot.scaling * (ksi * t * t);
      d_cur = pot.scaling * (mult * t * t + 2 * ksi * t);
    } else if (ri > 1) {
      const double vn = 2 * ksi * v_cur - v_prev;
      const double dn = 2 * (mult * v_cur + ksi * d_cur) - d_prev;
      v_prev = v_cur;
      d_prev = d_cur;
      v_cur = vn;
      d_cur = dn;
    }
#pragma unroll
    for (int mu = 0; mu < RP; mu++)
      if (mu < R) {
        const double cc = c[mu * B + ri];
        F[mu] += cc * v_cur;
        Fd[mu] += cc * d_cur;
      }
  }
}

// ---- forward: one batch of nb <= 32 staged neighbors into the DMMA accumulators ------------------------
template <int DEG, int RP>
__device__ __forceinline__ void v1_forward_batch(const V1Warp &ws, int nb, int lane,
                                                 double (&acc)[V1Shape<DEG, RP>::NB][2])
{
  constexpr int SMAX = DEG / 2;
  const int mu = lane >> 2;          // A row / C row
  const int pcls = lane >> 2;        // B column: parity class (pa | pb << 1 | pc << 2)
  const double *su = ws.stage;       // [3][32] unit vectors, [RP][32] radial values
  for (int n0 = 0; n0 < nb; n0 += 4) {
    const int n = n0 + (lane & 3);
    const bool ok = n < nb;
    const double ux = ok ? su[n] : 0.0, uy = ok ? su[32 + n] : 0.0, uz = ok ? su[64 + n] : 0.0;
    const double fa = (ok && mu < RP) ? su[(3 + mu) * 32 + n] : 0.0;
    const double w = ((pcls & 1) ? ux : 1.0) * ((pcls & 2) ? uy : 1.0) * ((pcls & 4) ? uz : 1.0);
    const double x2 = ux * ux, y2 = uy * uy, z2 = uz * uz;
    double ea = w;
    int t = 0;    // == block_index(SMAX, a, b, c): the loops run in the same lexicographic order
#pragma unroll
    for (int a = 0; a <= SMAX; a++) {
      double eab = ea;
#pragma unroll
      for (int b = 0; b <= SMAX - a; b++) {
        double eabc = eab;
#pragma unroll
        for (int c = 0; c <= SMAX - a - b; c++) {
          dmma884_v1(acc[t][0], acc[t][1], fa, eabc);
          t++;
          eabc *= z2;
        }
        eab *= y2;
      }
      ea *= x2;
    }
  }
}

// ---- backward: lane = neighbor, Horner evaluation of P_W'(u) and grad P_W(u) -----------------------------
// gc: canonical adjoints [NQ][RP] (shared memory, broadcast reads).  Returns F (force on the centre from this pair).
template <int DEG, int RP, bool GRADE>
__device__ __forceinline__ void v1_backward_neighbor(const double *gc, const int (&rcnt)[16], double ux, double uy,
                                                     double uz, const double (&fvi)[RP], const double (&fder)[RP],
                                                     double &Fx, double &Fy, double &Fz, double (&pm)[RP])
{
  double Pv = 0, Px = 0, Py = 0, Pz = 0, Pd = 0;
  int q = V1Shape<DEG, RP>::NQ - 1;    // == canon_index(DEG, a, b, c): reverse lexicographic sweep
  double PM[RP];
#pragma unroll
  for (int mu = 0; mu < RP; mu++) PM[mu] = 0.0;
#pragma unroll
  for (int a = DEG; a >= 0; a--) {
    double Q = 0, Qy = 0, Qz = 0, Qd = 0;
    double QM[RP];
#pragma unroll
    for (int mu = 0; mu < RP; mu++) QM[mu] = 0.0;
#pragma unroll
    for (int b = DEG - a; b >= 0; b--) {
      double T = 0, Tz = 0, Td = 0;
      double TM[RP];
#pragma unroll
      for (int mu = 0; mu < RP; mu++) TM[mu] = 0.0;
#pragma unroll
      for (int c = DEG - a - b; c >= 0; c--) {
        const int rc = rcnt[a + b + c];
        double W = 0, Wd = 0;
        const double2 *g2 = reinterpret_cast<const double2 *>(gc + q * RP);
#pragma unroll
        for (int m2 = 0; m2 < RP / 2; m2++) {
          if (2 * m2 < rc) {
            const double2 gg = g2[m2];
            W += fvi[2 * m2] * gg.x;
            Wd += fder[2 * m2] * gg.x;
            if (GRADE) TM[2 * m2] = TM[2 * m2] * uz + gg.x;
            if (2 * m2 + 1 < rc) {
              W += fvi[2 * m2 + 1] * gg.y;
              Wd += fder[2 * m2 + 1] * gg.y;
            }
            if (GRADE) TM[2 * m2 + 1] = TM[2 * m2 + 1] * uz + gg.y;
          } else if (GRADE) {
            TM[2 * m2] *= uz;
            TM[2 * m2 + 1] *= uz;
          }
        }
        Tz = Tz * uz + T;
        T = T * uz + W;
        Td = Td * uz + Wd;
        q--;
      }
      Qy = Qy * uy + Q;
      Q = Q * uy + T;
      Qz = Qz * uy + Tz;
      Qd = Qd * uy + Td;
      if (GRADE) {
#pragma unroll
        for (int mu = 0; mu < RP; mu++) QM[mu] = QM[mu] * uy + TM[mu];
      }
    }
    Px = Px * ux + Pv;
    Pv = Pv * ux + Q;
    Py = Py * ux + Qy;
    Pz = Pz * ux + Qz;
    Pd = Pd * ux + Qd;
    if (GRADE) {
#pragma unroll
      for (int mu = 0; mu < RP; mu++) PM[mu] = PM[mu] * ux + QM[mu];
    }
  }
  const double S = Pd - (ux * Px + uy * Py + uz * Pz);
  Fx = ux * S + Px;
  Fy = uy * S + Py;
  Fz = uz * S + Pz;
#pragma unroll
  for (int mu = 0; mu < RP; mu++) pm[mu] = PM[mu];
}


// ---- one sweep over the neighbor list of atom i -------------------------------------------------------------
// PHASE 0: basic moments (DMMA accumulators);  PHASE 1: forces / virial (and the radial candidate block).
struct V1Atom {
  int i, itype, jnum;
  long long row0;
  double xi0, xi1, xi2;
  double fx, fy, fz, v[6];
};

template <int DEG, int RP, bool GRADE, int PHASE>
__device__ __forceinline__ void v1_sweep(const DevPotential &pot, const V1Tables &tb, const SiteArgs &a,
                                         const V1Warp &ws, const double *s_radial, V1Atom &at, int lane,
                                         double (&acc)[V1Shape<DEG, RP>::NB][2])
{
  int cnt = 0;
  for (int base = 0; base < at.jnum || cnt > 0; base += 32) {
    // ---- gather + cutoff mask + compaction (pair_mtp.cpp:112-129) ----
    if (base < at.jnum) {
      const int jj = base + lane;
      bool within = false;
      int j = 0, jt = 0;
      double r0 = 0, r1 = 0, r2 = 0;
      if (jj < at.jnum) {
        const long long pos = at.row0 + (long long) jj * a.stride_jj;
        j = a.neighbors[pos] & a.neighmask;
        const double2 *nrec = reinterpret_cast<const double2 *>(a.xt + j);
        const double2 nxy = __ldg(nrec);
        const double2 nzt = __ldg(nrec + 1);
        jt = (int) __double_as_longlong(nzt.y);
        r0 = nxy.x - at.xi0;
        r1 = nxy.y - at.xi1;
        r2 = nzt.x - at.xi2;
        // separately rounded, left to right, exactly pair_mtp.cpp:121-123 (no FMA contraction)
        const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2));
        within = !(rsq > pot.cutsq);
        if (jt < 0 || jt >= pot.S) {    // pair_mtp.cpp:116-118
          atomicOr(a.status, 1);
          within = false;
        }
        if (PHASE == 0 && a.within) a.within[pos] = within ? 1 : 0;
      }
      const unsigned bal = __ballot_sync(FULL, within);
      if (within) {
        const int slot = cnt + __popc(bal & ((1u << lane) - 1u));
        ws.pr[slot] = r0;
        ws.pr[V1_PEND + slot] = r1;
        ws.pr[2 * V1_PEND + slot] = r2;
        ws.pj[slot] = j;
        ws.pt[slot] = jt;
      }
      cnt += __popc(bal);
      __syncwarp();
      if (cnt < 32 && base + 32 < at.jnum) continue;    // keep filling the batch
    }
    const int nb = cnt < 32 ? cnt : 32;
    if (nb == 0) break;

    // ---- per-neighbor radial part, lane = neighbor ----
    double ux = 0, uy = 0, uz = 0, r0 = 0, r1 = 0, r2 = 0, dist = 1.0, invd = 1.0;
    double F[RP], Fd[RP];
#pragma unroll
    for (int mu = 0; mu < RP; mu++) F[mu] = Fd[mu] = 0.0;
    int jt = 0;
    if (lane < nb) {
      r0 = ws.pr[lane];
      r1 = ws.pr[V1_PEND + lane];
      r2 = ws.pr[2 * V1_PEND + lane];
      jt = ws.pt[lane];
      dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2)));
      invd = 1.0 / dist;
      ux = r0 * invd;
      uy = r1 * invd;
      uz = r2 * invd;
      radial_functions<RP>(pot, s_radial + (size_t) (at.itype * pot.S + jt) * pot.R * pot.B, dist, F, Fd);
    }
    if (PHASE == 0) {
      ws.stage[lane] = ux;
      ws.stage[32 + lane] = uy;
      ws.stage[64 + lane] = uz;
#pragma unroll
      for (int mu = 0; mu < RP; mu++) ws.stage[(3 + mu) * 32 + lane] = F[mu];
      __syncwarp();
      v1_forward_batch<DEG, RP>(ws, nb, lane, acc);
      __syncwarp();
    } else {
      double pm[RP];
#pragma unroll
      for (int mu = 0; mu < RP; mu++) pm[mu] = 0.0;
      if (lane < nb) {
        double Fx, Fy, Fz;
        double fvi[RP];
#pragma unroll
        for (int mu = 0; mu < RP; mu++) fvi[mu] = F[mu] * invd;
        v1_backward_neighbor<DEG, RP, GRADE>(ws.gc, tb.rcnt, ux, uy, uz, fvi, Fd, Fx, Fy, Fz, pm);
        const int j = ws.pj[lane];
        atomicAdd(&a.f[3 * (size_t) j], -Fx);
        atomicAdd(&a.f[3 * (size_t) j + 1], -Fy);
        atomicAdd(&a.f[3 * (size_t) j + 2], -Fz);
        at.fx += Fx;
        at.fy += Fy;
        at.fz += Fz;
        if (a.vflag_any) {
          at.v[0] -= Fx * r0;
          at.v[1] -= Fy * r1;
          at.v[2] -= Fz * r2;
          at.v[3] -= (Fx * r1 + Fy * r0) / 2;
          at.v[4] -= (Fx * r2 + Fz * r0) / 2;
          at.v[5] -= (Fy * r2 + Fz * r1) / 2;
        }
      }
      if (GRADE) {
        // radial block of the candidate vector: b[(it*S+jt)*RB + mu*B + ri] += phi_ri(d_n) * P_mu(u_n)
        const double t = dist - pot.rmax;
        const double ksi = (2 * dist - (pot.rmin + pot.rmax)) / (pot.rmax - pot.rmin);
        double v_prev = 0, v_cur = pot.scaling * (1 * t * t);
        for (int ri = 0; ri < pot.B; ri++) {
          if (ri == 1) {
            v_prev = v_cur;
            v_cur = pot.scaling * (ksi * t * t);
          } else if (ri > 1) {
            const double vn = 2 * ksi * v_cur - v_prev;
            v_prev = v_cur;
            v_cur = vn;
          }
          for (int s = 0; s < pot.S; s++) {
#pragma unroll
            for (int mu = 0; mu < RP; mu++) {
              if (mu < pot.R) {
                double c = (lane < nb && jt == s) ? v_cur * pm[mu] : 0.0;
                c = warp_sum(c);
                if (lane == 0) ws.cand[(size_t) (at.itype * pot.S + s) * pot.R * pot.B + mu * pot.B + ri] += c;
              }
            }
          }
        }
      }
      __syncwarp();
    }

    // ---- drop the consumed batch, keep the remainder (< 32 entries) ----
    const int rem = cnt - nb;
    double t0 = 0, t1 = 0, t2 = 0;
    int tj = 0, tt = 0;
    if (lane < rem) {
      t0 = ws.pr[32 + lane];
      t1 = ws.pr[V1_PEND + 32 + lane];
      t2 = ws.pr[2 * V1_PEND + 32 + lane];
      tj = ws.pj[32 + lane];
      tt = ws.pt[32 + lane];
    }
    __syncwarp();
    if (lane < rem) {
      ws.pr[lane] = t0;
      ws.pr[V1_PEND + lane] = t1;
      ws.pr[2 * V1_PEND + lane] = t2;
      ws.pj[lane] = tj;
      ws.pt[lane] = tt;
    }
    cnt = rem;
    __syncwarp();
  }
}

#ifdef MTP_PHASE_CLOCKS
__device__ unsigned long long g_phase_clocks[8];
#define PHASE_T0() long long pc__ = clock64()
#define PHASE_MARK(k)                                                   \
  do {                                                                  \
    const long long now__ = clock64();                                  \
    if (lane == 0) atomicAdd(&g_phase_clocks[k], (unsigned long long) (now__ - pc__)); \
    pc__ = now__;                                                       \
  } while (0)
#else
#define PHASE_T0()
#define PHASE_MARK(k)
#endif

__device__ __forceinline__ void v1_load_atom(const DevPotential &pot, const SiteArgs &a, int ii, int lane, V1Atom &at)
{
  at.i = a.ilist ? a.ilist[a.first_ii + ii] : a.first_ii + ii;
  const double2 *rec = reinterpret_cast<const double2 *>(a.xt + at.i);
  const double2 xy = __ldg(rec);
  const double2 zt = __ldg(rec + 1);
  at.xi0 = xy.x;
  at.xi1 = xy.y;
  at.xi2 = zt.x;
  at.itype = (int) __double_as_longlong(zt.y);
  if (at.itype < 0 || at.itype >= pot.S) {    // pair_mtp.cpp:91-93
    if (lane == 0) atomicOr(a.status, 1);
    at.itype = 0;
  }
  at.jnum = a.numneigh[at.i];
  at.row0 = a.neigh_offsets ? a.neigh_offsets[at.i] : (long long) at.i * a.stride_i;
  at.fx = at.fy = at.fz = 0.0;
#pragma unroll
  for (int c = 0; c < 6; c++) at.v[c] = 0.0;
}

// =====================================================================================================================
// The per-step pipeline over one super-chunk of centre atoms [first_ii, first_ii + inum) (the "chunksize" of the
// pair_style line bounds it so that the two intermediates below stay L2-resident):
//
//   mtp_moments_kernel   warp per atom   gather/mask, radial functions, basic moments (DMMA)   -> mb[k][atom]
//   mtp_program_kernel   CTA per NA atoms, lane = atom: contraction program forward, site energy,
//                        reverse mode                                                          -> gb[k][atom]
//   mtp_forces_kernel    warp per atom   canonical adjoints, per-pair forces, scatter, virial, candidate vector
//
// mb / gb are [K][ld] FP64 (atom index fastest, ld = padded super-chunk size): 2 x K x 8 bytes per atom that are
// written and read once through L2; nothing per (atom, neighbor) ever leaves the SM.
// (warp_sum and FULL come from mtp_kernels.cu, which includes this header.)
// =====================================================================================================================
template <int DEG, int RP>
__global__ void __launch_bounds__(256)
mtp_moments_kernel(DevPotential pot, V1Tables tb, SiteArgs a, double *__restrict__ mb, int ld)
{
  using Sh = V1Shape<DEG, RP>;
  extern __shared__ __align__(16) unsigned char smem[];
  const V1Layout L = v1_layout<DEG, RP>(pot.S, pot.R, pot.B, pot.M, pot.Q, false, 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  double *s_radial = reinterpret_cast<double *>(smem);
  for (int t = threadIdx.x; t < pot.S * pot.S * pot.R * pot.B; t += blockDim.x) s_radial[t] = pot.radial[t];
  V1Warp ws;
  {
    double *d = reinterpret_cast<double *>(smem + L.radial_bytes + (size_t) warp * L.warp_bytes_moments);
    ws.gc = nullptr;
    ws.cand = nullptr;
    ws.stage = d;
    ws.pr = ws.stage + Sh::STAGE;
    ws.pj = reinterpret_cast<int *>(ws.pr + 3 * V1_PEND);
    ws.pt = ws.pj + V1_PEND;
  }
  __syncthreads();
  for (int ii = blockIdx.x * W + warp; ii < a.inum; ii += gridDim.x * W) {
    V1Atom at;
    v1_load_atom(pot, a, ii, lane, at);
    double acc[Sh::NB][2];
#pragma unroll
    for (int t = 0; t < Sh::NB; t++) acc[t][0] = acc[t][1] = 0.0;
    v1_sweep<DEG, RP, false, 0>(pot, tb, a, ws, s_radial, at, lane, acc);
#pragma unroll
    for (int t = 0; t < Sh::NB; t++) {
      const short2 k = reinterpret_cast<const short2 *>(tb.fwd_slot)[t * 32 + lane];
      if (k.x >= 0) mb[(size_t) k.x * ld + ii] = acc[t][0];
      if (k.y >= 0) mb[(size_t) k.y * ld + ii] = acc[t][1];
    }
  }
}

// contraction program forward (pair_mtp.cpp:196-201), site energy (:204-212), reverse mode (:217-233)
//
// CTA per chunk of NA atoms (NA = 32 or 8), lane = atom: moments cm[node][atom] and adjoints cg[node][atom] of the
// chunk live in shared memory (row stride NA+1, row M = 1.0).  The program is executed as flat predicated term
// streams (mtp_potential.hpp: FlatPass): one stream per (dependency level, virtual warp); every term is
//     acc += coef * A[a] * B[b];  if (store) { dst[node] = acc; acc = 0; }
// so the inner loop has no data-dependent branch and the descriptor + operand loads of FLAT_UNROLL terms are all
// independent (the only serial chain is the accumulator).
// shared-memory plan of the program kernel (host and device agree through this one function)
constexpr int PROG_THREADS = 256;    // 8 warps; every thread owns TWO adjacent atoms of the chunk (double2 operands)
struct ProgLayout {
  size_t node_bytes, off_cg, off_epart, off_s2k, off_lin, off_map, off_stage, off_terms[2], off_st[2], total;
};
__host__ __device__ inline ProgLayout program_layout(int M, int Mg, int A, int na, int nslots, int nterms_f, int nterms_r,
                                                     bool dsmem, bool prefetch = false)
{
  ProgLayout L;
  const int vw = 16 * 32 / na;
  L.node_bytes = ((size_t) (M + 1) * na * 8 + 15) & ~(size_t) 15;    // moments: M rows + the row of ones
  L.off_cg = L.node_bytes;
  L.off_epart = L.node_bytes + (((size_t) Mg * na * 8 + 15) & ~(size_t) 15);    // adjoints: the compact table
  L.off_s2k = L.off_epart + (size_t) vw * na * 8;
  size_t o = (L.off_s2k + (size_t) nslots * 2 + 15) & ~(size_t) 15;
  L.off_lin = o;
  o += (size_t) A * 8;
  L.off_map = o;
  o = (o + (size_t) A * 4 + 15) & ~(size_t) 15;
  L.off_stage = o;
  if (prefetch) o += (size_t) nslots * na * 8;
  L.off_terms[0] = o;
  if (dsmem) o += (size_t) nterms_f * 16;
  L.off_terms[1] = o;
  if (dsmem) o += (size_t) nterms_r * 16;
  L.off_st[0] = o;
  if (dsmem) o += (size_t) nterms_f * 4;
  L.off_st[1] = o;
  if (dsmem) o += (size_t) nterms_r * 4;
  L.total = (o + 15) & ~(size_t) 15;
  return L;
}

template <int U>
__device__ __forceinline__ void flat_terms(const uint4 *__restrict__ terms, const unsigned *__restrict__ st, int t,
                                           const unsigned char *A, const unsigned char *B, unsigned char *dst,
                                           const unsigned char *one, double2 &acc)
{
  uint4 d[U];
  unsigned sw[U];
#pragma unroll
  for (int q = 0; q < U / 4; q++) {
    const uint4 s4 = *reinterpret_cast<const uint4 *>(st + t + 4 * q);
    sw[4 * q] = s4.x, sw[4 * q + 1] = s4.y, sw[4 * q + 2] = s4.z, sw[4 * q + 3] = s4.w;
  }
#pragma unroll
  for (int u = 0; u < U; u++) d[u] = terms[t + u];
  // operand flag bit 0: the operand is the constant 1.0 -> every lane reads the same 16-byte word (one broadcast
  // wavefront instead of four); selected by address so that the loads stay unconditional and batched
  double2 va[U], vb[U];
#pragma unroll
  for (int u = 0; u < U; u++) {
    const unsigned char *pa = (d[u].x & 1u) ? one : A + d[u].x;
    const unsigned char *pb = (d[u].y & 1u) ? one : B + d[u].y;
    va[u] = *reinterpret_cast<const double2 *>(pa);
    vb[u] = *reinterpret_cast<const double2 *>(pb);
  }
#pragma unroll
  for (int u = 0; u < U; u++) {
    const double coef = __hiloint2double((int) d[u].w, (int) d[u].z);
    acc.x = fma(coef * va[u].x, vb[u].x, acc.x);
    acc.y = fma(coef * va[u].y, vb[u].y, acc.y);
    if (sw[u] & 1u) {
      *reinterpret_cast<double2 *>(dst + (sw[u] & ~7u)) = acc;
      acc = make_double2(0.0, 0.0);
    }
  }
}

template <bool REVERSE>
__device__ __forceinline__ void run_flat_pass(const DevFlatPass &ps, const uint4 *__restrict__ terms,
                                              const unsigned *__restrict__ st, unsigned char *cm_lane,
                                              unsigned char *cg_lane, const unsigned char *one, int vwarp)
{
  const unsigned char *A = REVERSE ? cg_lane : cm_lane;
  unsigned char *dst = REVERSE ? cg_lane : cm_lane;
  for (int lv = 0; lv < ps.nlevels; lv++) {
    const int t0 = ps.stream_begin[lv * ps.vw + vwarp], t1 = ps.stream_begin[lv * ps.vw + vwarp + 1];
    double2 acc = make_double2(0.0, 0.0);
    int t = t0;
    for (; t + 8 <= t1; t += 8) flat_terms<8>(terms, st, t, A, cm_lane, dst, one, acc);
    if (t < t1) flat_terms<4>(terms, st, t, A, cm_lane, dst, one, acc);    // streams are padded to a multiple of 4
    __syncthreads();
  }
}

// NA = atoms per CTA (power of two, >= 2), lna = log2(NA)
template <bool GRADE>
__global__ void __launch_bounds__(PROG_THREADS)
mtp_program_kernel(DevPotential pot, SiteArgs a, const double *__restrict__ mb, double *__restrict__ gb, int ld,
                   int NA, int lna, double *__restrict__ partials)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int fi = a.prog_shape;
  const DevFlatPass &pf = pot.ffwd[fi], &pr = pot.frev[fi];
  const int nslots = a.slot_to_k ? a.nslots : pot.K;
  const ProgLayout L = program_layout(pot.M, pot.Mg, pot.A, NA, nslots, pf.nterms, pr.nterms, a.prog_dsmem != 0, a.prog_prefetch != 0);
  double *s_lin = reinterpret_cast<double *>(smem + L.off_lin);
  int *s_map = reinterpret_cast<int *>(smem + L.off_map);
  double *stage = reinterpret_cast<double *>(smem + L.off_stage);
  double *cm = reinterpret_cast<double *>(smem);
  double *cg = reinterpret_cast<double *>(smem + L.off_cg);
  double *epart = reinterpret_cast<double *>(smem + L.off_epart);
  short *s2k = reinterpret_cast<short *>(smem + L.off_s2k);
  const int lpv = NA >> 1;                      // lanes per virtual warp (each lane = 2 atoms)
  const int al2 = (lane & (lpv - 1)) * 2;       // first atom of this lane
  const int VW = (PROG_THREADS * 2) >> lna;
  const int vwarp = threadIdx.x >> (lna - 1);
  const int radial_count = pot.S * pot.S * pot.R * pot.B;
  double e_thread = 0.0;

  // one-time setup: constant row, slot map, term streams
  for (int t = threadIdx.x; t < NA; t += blockDim.x) cm[pot.M * NA + t] = 1.0;
  for (int t = threadIdx.x; t < nslots; t += blockDim.x) s2k[t] = a.slot_to_k ? a.slot_to_k[t] : (short) t;
  for (int t = threadIdx.x; t < pot.A; t += blockDim.x) {
    s_lin[t] = pot.lin[t];
    s_map[t] = pot.map[t] << lna;    // row offset in doubles
  }
  const uint4 *terms_f = pf.terms, *terms_r = pr.terms;
  const unsigned *st_f = pf.st, *st_r = pr.st;
  if (a.prog_dsmem) {
    uint4 *tf = reinterpret_cast<uint4 *>(smem + L.off_terms[0]), *tr = reinterpret_cast<uint4 *>(smem + L.off_terms[1]);
    unsigned *sf = reinterpret_cast<unsigned *>(smem + L.off_st[0]), *sr = reinterpret_cast<unsigned *>(smem + L.off_st[1]);
    for (int t = threadIdx.x; t < pf.nterms; t += blockDim.x) {
      tf[t] = pf.terms[t];
      sf[t] = pf.st[t];
    }
    for (int t = threadIdx.x; t < pr.nterms; t += blockDim.x) {
      tr[t] = pr.terms[t];
      sr[t] = pr.st[t];
    }
    terms_f = tf;
    terms_r = tr;
    st_f = sf;
    st_r = sr;
  }
  unsigned char *cm_lane = smem + (size_t) al2 * 8, *cg_lane = smem + L.off_cg + (size_t) al2 * 8;
  const unsigned char *one_ptr = smem + (size_t) pot.M * NA * 8;    // row M of cm holds 1.0
  __syncthreads();

  // basic moments of chunk c0 -> dst rows (dst = cm directly, or the staging buffer [slot][atom] when prefetching)
  auto fetch_basic = [&](int c0, bool to_stage) {
    const int nac = min(NA, a.inum - c0);
    for (int t = threadIdx.x; t < (nslots << lna); t += blockDim.x) {
      const int s = t >> lna, al = t & (NA - 1);
      const int k = s2k[s];
      double *dstp = to_stage ? stage + t : cm + (k << lna) + al;
      if (to_stage || k >= 0)
        cp_async8_zfill(dstp, mb + (size_t) s * ld + c0 + (al < nac ? al : 0), al < nac ? 8 : 0);
    }
    cp_async_commit();
  };
  if (a.prog_prefetch && (int) (blockIdx.x * NA) < a.inum) fetch_basic(blockIdx.x * NA, true);

  for (int chunk0 = blockIdx.x * NA; chunk0 < a.inum; chunk0 += gridDim.x * NA) {
    const int na = min(NA, a.inum - chunk0);
    // this thread's atom for the energy epilogue: id and species requested now, consumed after the forward pass
    int my_i = 0, my_type = 0;
    if ((int) threadIdx.x < na && (a.eflag_global || a.eflag_atom)) {
      my_i = a.ilist ? a.ilist[a.first_ii + chunk0 + threadIdx.x] : a.first_ii + chunk0 + threadIdx.x;
      my_type = (int) a.xt[my_i].t;
    }
    if (a.prog_prefetch) {
      cp_async_wait<0>();
      __syncthreads();
      for (int t = threadIdx.x; t < (nslots << lna); t += blockDim.x) {
        const int k = s2k[t >> lna];
        if (k >= 0) cm[(k << lna) + (t & (NA - 1))] = stage[t];
      }
      __syncthreads();
      const int next0 = chunk0 + gridDim.x * NA;
      if (next0 < a.inum) fetch_basic(next0, true);    // in flight during this chunk's passes
    } else {
      fetch_basic(chunk0, false);
      cp_async_wait<0>();
      __syncthreads();
    }
    if (!(a.prog_debug & 1)) run_flat_pass<false>(pf, terms_f, st_f, cm_lane, cg_lane, one_ptr, vwarp);
    // site energies: the virtual warps split the basis functions; fixed-order reduction
    if ((a.eflag_global || a.eflag_atom || GRADE) && !(a.prog_debug & 4)) {
      double e0 = 0.0, e1 = 0.0;
      for (int s = vwarp; s < pot.A; s += VW) {
        const double2 bm = *reinterpret_cast<const double2 *>(cm + s_map[s] + al2);
        const double c = s_lin[s];
        e0 = fma(c, bm.x, e0);
        e1 = fma(c, bm.y, e1);
        if (GRADE) {
          if (al2 < na) a.cand_rows[(size_t) (chunk0 + al2) * a.cand_ld + radial_count + pot.S + s] = bm.x;
          if (al2 + 1 < na) a.cand_rows[(size_t) (chunk0 + al2 + 1) * a.cand_ld + radial_count + pot.S + s] = bm.y;
        }
      }
      *reinterpret_cast<double2 *>(epart + (size_t) vwarp * NA + al2) = make_double2(e0, e1);
      __syncthreads();
      if (threadIdx.x < na) {
        const int al = threadIdx.x;
        const int i = my_i;
        int itype = my_type;
        if (itype < 0 || itype >= pot.S) itype = 0;
        double es = 0.0;
        for (int v = 0; v < VW; v++) es += epart[v * NA + al];
        es += pot.species[itype];
        if (a.eflag_atom) a.eatom[i] = es;
        if (a.eflag_global) e_thread += es;
      }
    }
    if (!(a.prog_debug & 2)) run_flat_pass<true>(pr, terms_r, st_r, cm_lane, cg_lane, one_ptr, vwarp);
    // adjoints of the basic moments -> gb
#pragma unroll 4
    for (int t = threadIdx.x; t < (nslots << lna); t += blockDim.x) {
      const int s = t >> lna, al = t & (NA - 1);
      const int k = s2k[s];
      if (al < na) gb[(size_t) s * ld + chunk0 + al] = k >= 0 ? cg[(k << lna) + al] : 0.0;
    }
    __syncthreads();
  }

  // per-CTA energy partial (fixed order): only warp 0 holds per-atom energies
  if (warp == 0) {
    const double s = warp_sum(e_thread);
    if (lane < 8) partials[(size_t) blockIdx.x * 8 + lane] = lane == 0 ? s : 0.0;
  }
}

template <int DEG, int RP, bool GRADE>
__global__ void __launch_bounds__(256)
mtp_forces_kernel(DevPotential pot, V1Tables tb, SiteArgs a, const double *__restrict__ gb, int ld,
                  double *__restrict__ partials)
{
  using Sh = V1Shape<DEG, RP>;
  extern __shared__ __align__(16) unsigned char smem[];
  const V1Layout L = v1_layout<DEG, RP>(pot.S, pot.R, pot.B, pot.M, pot.Q, GRADE, 1);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  double *s_radial = reinterpret_cast<double *>(smem);
  for (int t = threadIdx.x; t < pot.S * pot.S * pot.R * pot.B; t += blockDim.x) s_radial[t] = pot.radial[t];
  V1Warp ws;
  {
    double *d = reinterpret_cast<double *>(smem + L.radial_bytes + (size_t) warp * L.warp_bytes_forces);
    ws.stage = nullptr;
    ws.gc = d;
    ws.cand = ws.gc + Sh::GC;
    ws.pr = ws.cand + (GRADE ? pot.Q : 0);
    ws.pj = reinterpret_cast<int *>(ws.pr + 3 * V1_PEND);
    ws.pt = ws.pj + V1_PEND;
  }
  __syncthreads();
  const int radial_count = pot.S * pot.S * pot.R * pot.B;
  double v_warp[6] = {0, 0, 0, 0, 0, 0};

  for (int ii = blockIdx.x * W + warp; ii < a.inum; ii += gridDim.x * W) {
    V1Atom at;
    v1_load_atom(pot, a, ii, lane, at);
    for (int s = lane; s < Sh::GC; s += 32) {
      const short k = tb.g_src[s];
      ws.gc[s] = k >= 0 ? gb[(size_t) k * ld + ii] : 0.0;
    }
    if (GRADE)
      for (int q = lane; q < radial_count + pot.S; q += 32) ws.cand[q] = 0.0;
    __syncwarp();
    double none[Sh::NB][2];
    v1_sweep<DEG, RP, GRADE, 1>(pot, tb, a, ws, s_radial, at, lane, none);

    const double fx = warp_sum(at.fx), fy = warp_sum(at.fy), fz = warp_sum(at.fz);
    if (lane == 0) {
      atomicAdd(&a.f[3 * (size_t) at.i], fx);
      atomicAdd(&a.f[3 * (size_t) at.i + 1], fy);
      atomicAdd(&a.f[3 * (size_t) at.i + 2], fz);
    }
    if (a.vflag_any) {
#pragma unroll
      for (int c = 0; c < 6; c++) {
        const double vc = warp_sum(at.v[c]);
        v_warp[c] += vc;
        if (a.vflag_atom && lane == 0) a.vatom[6 * (size_t) at.i + c] += vc;
      }
    }
    if (GRADE) {
      // radial block + species one-hot; the linear block was written by the program kernel
      if (lane == 0) ws.cand[radial_count + at.itype] += 1.0;
      __syncwarp();
      double *dst = a.cand_rows + (size_t) ii * a.cand_ld;
      for (int q = lane; q < radial_count + pot.S; q += 32) dst[q] = ws.cand[q];
      for (int q = pot.Q + lane; q < a.cand_ld; q += 32) dst[q] = 0.0;
    }
    __syncwarp();
  }

  __shared__ double s_part[8][8];
  if (lane == 0) {
    s_part[warp][0] = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) s_part[warp][1 + c] = v_warp[c];
    s_part[warp][7] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int w = 0; w < W; w++) s += s_part[w][threadIdx.x];
    partials[(size_t) blockIdx.x * 8 + threadIdx.x] = s;
  }
}

}    // namespace mtpb200
