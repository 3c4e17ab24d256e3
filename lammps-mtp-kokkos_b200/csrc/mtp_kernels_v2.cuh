// Pair stages of the per-step pipeline for standard MTP basic-moment sets (template <D0>).
//
// A level-L MLIP potential uses the basic moments M_{mu,nu} with nu <= D_mu = D0 - 2 mu (D0 = (L-4)/2, plus a
// trailing scalar M_{mu,0} when D0 is odd), i.e. for every monomial q = x^a y^b z^c of degree d the radial
// indices mu < rcnt(d).  With D0 a compile-time constant every loop over (a, b, c, mu) below unrolls into
// straight-line FP64 code on registers.  Canonical slot order: q lexicographic in (a, b, c), then mu.
//
//   mtp_gather_radial_kernel   warp per centre atom: neighbor gather (32-byte position records), cutoff mask exactly as
//                        pair_mtp.cpp:112-129, ballot compaction into a ring in shared memory; then lane = in-cutoff
//                        pair: Chebyshev x cutoff and the radial contraction (mtp_rb_chevbyshev_basis.cpp:29-54,
//                        pair_mtp.cpp:139-151) -> record {u, d, f_mu, f'_mu, j} in the pair buffer (field-major,
//                        atom ii owning the slots [ii * ncap, ii * ncap + pcnt[ii])).
//   mtp_moments_v2       CTA = 32 atoms x NP warps, lane = atom, warp = "pass" (a contiguous range of canonical
//                        slots, <= ~48 accumulators held in registers).  Pair records are staged through shared
//                        memory (coalesced read, transposed so that lane = atom reads are conflict free);
//                        m[mu][q] += f_mu u^q is one DFMA per (pair, moment), monomials by running products.
//   mtp_forces_v2        CTA = AB atoms, lane = pair (all lanes busy whatever the neighbor counts): the adjoints
//                        dE/dm of the CTA's atoms sit in shared memory in canonical order, the per-pair force is
//                        the gradient of sum_mu f_mu(d) P_mu(u) evaluated by a fully unrolled three-level Horner
//                        scheme (pair_mtp.cpp:175-191,236-254 without ever forming the Jacobian), red.f64 scatter to
//                        the neighbor, segmented warp reduction for the centre atom, virial -sym(F (x) r) (:257-276).
#pragma once

#include "mtp_device.cuh"

namespace mtpb200 {

struct PairBuf {
  double *fld;        // [4 + 2R][cap]: ux, uy, uz, d, f_0..f_{R-1}, f'_0..f'_{R-1}
  int *pj;            // [cap] neighbor atom index (ghosts included)
  int *pjt;           // [cap] neighbor species (0-based)
  int *pcnt;          // [chunk] in-cutoff pairs per centre
  long long cap;      // chunk * ncap
  int ncap;           // slots per centre (>= max numneigh)
};

#ifndef MTP_V2_ACC
#define MTP_V2_ACC 0       // > 0: force the accumulator budget of the moment kernel (experiments); 0 = per-level choice
#endif
template <int D0> struct V2Shape {
  static constexpr int R = D0 / 2 + 1 + (D0 & 1);
  __host__ __device__ static constexpr int dmu(int mu) { return D0 - 2 * mu > 0 ? D0 - 2 * mu : 0; }
  __host__ __device__ static constexpr int rcnt(int d) { return d == 0 ? R : (d > D0 ? 0 : (D0 - d) / 2 + 1); }
  __host__ __device__ static constexpr int kfull()
  {
    int s = 0;
    for (int mu = 0; mu < R; mu++) s += tet(dmu(mu));
    return s;
  }
  static constexpr int NQ = tet(D0);
  static constexpr int KF = kfull();
  static constexpr int NF = 3 + R;                 // fields the moment kernel stages: u, f
  static constexpr int NFLD = 4 + 2 * R;           // fields of a pair record
  // ---- forward passes: columns (a, b) in lexicographic order, greedily packed under a register budget
  __host__ __device__ static constexpr int col_count(int a, int b)
  {
    int s = 0;
    for (int c = 0; c <= D0 - a - b; c++) s += rcnt(a + b + c);
    return s;
  }
  // basic-moment accumulators a thread keeps in registers (one pass): fewer and therefore more passes (warps) per CTA
  // up to level 18 -- the kernel is latency bound and wants resident warps -- 48 above, where the passes are many anyway
  static constexpr int ACC = MTP_V2_ACC > 0 ? MTP_V2_ACC : (D0 <= 7 ? 32 : 48);
  static constexpr int NPASS0 = (KF + ACC - 1) / ACC;    // passes if the columns packed perfectly
  static constexpr int BUDGET = (KF + NPASS0 - 1) / NPASS0 + 2;
  // pass of column (a, b); with (a, b) = (D0 + 1, 0) returns the number of passes
  __host__ __device__ static constexpr int pass_of(int qa, int qb)
  {
    int pass = 0, fill = 0;
    for (int a = 0; a <= D0; a++)
      for (int b = 0; b <= D0 - a; b++) {
        const int n = col_count(a, b);
        if (fill + n > BUDGET && fill > 0) {
          pass++;
          fill = 0;
        }
        if (a == qa && b == qb) return pass;
        fill += n;
      }
    return pass + 1;
  }
  static constexpr int NP = pass_of(D0 + 1, 0);
  __host__ __device__ static constexpr int col_begin(int qa, int qb)    // canonical slot of (qa, qb, c = 0)
  {
    int s = 0;
    for (int a = 0; a <= D0; a++)
      for (int b = 0; b <= D0 - a; b++) {
        if (a == qa && b == qb) return s;
        s += col_count(a, b);
      }
    return s;
  }
  __host__ __device__ static constexpr int pass_begin(int p)    // first canonical slot of pass p
  {
    int s = 0;
    for (int a = 0; a <= D0; a++)
      for (int b = 0; b <= D0 - a; b++) {
        if (pass_of(a, b) >= p) return s;
        s += col_count(a, b);
      }
    return s;
  }
  __host__ __device__ static constexpr int amax()
  {
    int m = 0;
    for (int p = 0; p < NP; p++) {
      const int n = pass_begin(p + 1) - pass_begin(p);
      m = n > m ? n : m;
    }
    return m;
  }
  static constexpr int AMAX = amax();
};

constexpr int V2_PEND = 64;
#ifndef MTP_V2_NT
#define MTP_V2_NT 16
#endif
constexpr int V2_NT = MTP_V2_NT;      // pairs per staged tile of the moment kernel (power of two)


// ===================================================================================================== gather + radial
// Chebyshev x cutoff (mtp_rb_chevbyshev_basis.cpp:29-54) contracted with the radial coefficients of one species
// pair (pair_mtp.cpp:139-151), fully unrolled.  ct = coefficient table in shared memory, element (ri, mu) of species
// pair pt at ct[(ri * R + mu) * SP + pt]: the lanes of a warp that differ in their species pair read adjacent words
// (no bank conflict), lanes with the same pair read one word (broadcast).
struct RadialConsts {
  double rmax, ka, kb, mult, scaling;    // xi = ka d + kb,  mult = d xi / d d
};
__device__ __forceinline__ RadialConsts radial_consts(const DevPotential &pot)
{
  RadialConsts c;
  const double inv = 1.0 / (pot.rmax - pot.rmin);    // one division per thread instead of one per pair
  c.rmax = pot.rmax;
  c.ka = 2.0 * inv;
  c.kb = -(pot.rmin + pot.rmax) * inv;
  c.mult = 2.0 * inv;
  c.scaling = pot.scaling;
  return c;
}
template <int R, int B>
__device__ __forceinline__ void radial_unrolled(const RadialConsts &rc, const double *__restrict__ ct, int SP, double d,
                                                double (&F)[R], double (&Fd)[R])
{
  const double t = d - rc.rmax;
  const double ksi = fma(d, rc.ka, rc.kb);
  const double mult = rc.mult;
  double v_prev = rc.scaling * (1 * t * t), d_prev = rc.scaling * 2 * t;
#pragma unroll
  for (int mu = 0; mu < R; mu++) {
    F[mu] = ct[mu * SP] * v_prev;
    Fd[mu] = ct[mu * SP] * d_prev;
  }
  if (B == 1) return;
  double v_cur = rc.scaling * (ksi * t * t), d_cur = rc.scaling * (mult * t * t + 2 * ksi * t);
#pragma unroll
  for (int ri = 1; ri < B; ri++) {
    if (ri > 1) {
      const double vn = 2 * ksi * v_cur - v_prev;
      const double dn = 2 * (mult * v_cur + ksi * d_cur) - d_prev;
      v_prev = v_cur;
      d_prev = d_cur;
      v_cur = vn;
      d_cur = dn;
    }
#pragma unroll
    for (int mu = 0; mu < R; mu++) {
      const double cc = ct[(ri * R + mu) * SP];
      F[mu] = fma(cc, v_cur, F[mu]);
      Fd[mu] = fma(cc, d_cur, Fd[mu]);
    }
  }
}

// stages 1 + 2 in one kernel, warp per centre atom.
//   gather: lane = listed neighbor; 32-byte position records, cutoff mask exactly as pair_mtp.cpp:112-129; the
//           in-cutoff displacements are compacted, in list order, into a 64-entry ring of the warp in shared memory;
//   radial: whenever the ring holds 32 entries (and once more at the end of the list) lane = in-cutoff pair: distance,
//           unit vector, Chebyshev x cutoff, radial contraction -> the pair record {u, d, f_mu, f'_mu, j, jt}.
// The displacement never travels through global memory, every lane of the radial phase is busy, and the FP64 work of
// one warp runs under the gather latency of the others.
constexpr int V2_RING = 64;
// B8: the radial basis has 8 functions (every MLIP template) -> fully unrolled recurrence; PLAIN: unit stride within a
// neighbor row and no cutoff-mask output (what a force evaluation inside an MD run asks for)
template <int R, int MINB, int V2_GB, bool B8, bool PLAIN>
__global__ void __launch_bounds__(256, MINB)
mtp_gather_radial_kernel(DevPotential pot, SiteArgs a, PairBuf pb)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int SS = pot.S * pot.S, RB = pot.R * pot.B, nrad = SS * RB;
  double *s_ct = reinterpret_cast<double *>(smem);                                     // [B][R][S*S]
  double *ring = s_ct + ((nrad + 1) & ~1) + (size_t) warp * (V2_RING * 4);             // per warp: r0, r1, r2, {j, jt}
  for (int t = threadIdx.x; t < nrad; t += blockDim.x) {
    const int pt = t / RB, r = t - pt * RB, mu = r / pot.B, ri = r - mu * pot.B;
    s_ct[(ri * pot.R + mu) * SS + pt] = pot.radial[t];
  }
  __syncthreads();
  // The chain ilist -> {position, numneigh, row offset} -> neighbor ids -> neighbor records is four dependent memory
  // latencies per centre.  It is software-pipelined: the id of the centre after next and the header of the next centre
  // are requested before the current centre is processed, and the ids / records of up to V2_GB batches of 32 listed
  // neighbors are requested back to back, so that a centre costs about two exposed latencies instead of eight.
  struct Hdr {
    int i, jnum, itype;
    long long row0;
    double x0, x1, x2;
  };
  double *const fld0 = pb.fld;
  int *const pj0 = pb.pj, *const pjt0 = pb.pjt;
  const long long cap = pb.cap;
  const RadialConsts rc = radial_consts(pot);
  const long long stride_jj = PLAIN ? 1 : a.stride_jj;
  // every warp walks a CONTIGUOUS run of the list: consecutive centres share most of their neighbors (the list is in
  // spatial order after LAMMPS's atom sort), so the gathered records of a run stay in L1 (the kernel asks for a large
  // L1 carve-out) and only ~10 % of the gathers go to L2
  const int nwarps = gridDim.x * W, gw = blockIdx.x * W + warp;
  const int per = (a.inum + nwarps - 1) / nwarps;
  const int ii_begin = min(a.inum, gw * per), ii_end = min(a.inum, ii_begin + per);
  const int stride = 1;
  auto load_id = [&](int ii) { return ii < ii_end ? (a.ilist ? a.ilist[a.first_ii + ii] : a.first_ii + ii) : 0; };
  auto load_hdr = [&](int i, Hdr &h) {
    h.i = i;
    ld_atomrec(a.xt + i, h.x0, h.x1, h.x2, h.itype);
    h.jnum = a.numneigh[i];
    h.row0 = a.neigh_offsets ? a.neigh_offsets[i] : (long long) i * a.stride_i;
  };
  int ii = ii_begin;
  Hdr nxt;
  nxt.i = nxt.jnum = nxt.itype = 0;
  nxt.row0 = 0;
  nxt.x0 = nxt.x1 = nxt.x2 = 0.0;
  if (ii < ii_end) load_hdr(load_id(ii), nxt);
  int id_after = load_id(ii + stride);
  for (; ii < ii_end; ii += stride) {
    V1Atom at;
    at.i = nxt.i;
    at.itype = nxt.itype;
    at.jnum = nxt.jnum;
    at.row0 = nxt.row0;
    at.xi0 = nxt.x0;
    at.xi1 = nxt.x1;
    at.xi2 = nxt.x2;
    if (at.itype < 0 || at.itype >= pot.S) {    // pair_mtp.cpp:91-93
      if (lane == 0) atomicOr(a.status, 1);
      at.itype = 0;
    }
    if (ii + stride < ii_end) load_hdr(id_after, nxt);
    id_after = load_id(ii + 2 * stride);
    const long long slot0 = (long long) ii * pb.ncap;
    int head = 0, count = 0, done = 0;
    // lane = entry `head + lane` of the ring -> record `done + lane` of this centre
    auto flush = [&](int n) {
      __syncwarp();
      if (lane < n) {
        const int e = (head + lane) & (V2_RING - 1);
        const double r0 = ring[e], r1 = ring[V2_RING + e], r2 = ring[2 * V2_RING + e];
        const int2 jj = reinterpret_cast<const int2 *>(ring + 3 * V2_RING)[e];
        const int jt = jj.y;
        // one reciprocal square root gives both d and 1/d (the cutoff mask above used rsq itself, bit-exactly)
        const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2));
        const double invd = rsqrt(rsq);
        const double dist = rsq * invd;
        double F[R], Fd[R];
        const double *ct = s_ct + (at.itype * pot.S + jt);
        if (B8) radial_unrolled<R, 8>(rc, ct, SS, dist, F, Fd);
        else {    // any other basis size: same recurrence with a run-time trip count
          const double t = dist - pot.rmax;
          const double ksi = (2 * dist - (pot.rmin + pot.rmax)) / (pot.rmax - pot.rmin);
          const double mult = 2.0 / (pot.rmax - pot.rmin);
          double v_prev = 0, v_cur = pot.scaling * (1 * t * t), d_prev = 0, d_cur = pot.scaling * 2 * t;
#pragma unroll
          for (int mu = 0; mu < R; mu++) F[mu] = Fd[mu] = 0.0;
          for (int ri = 0; ri < pot.B; ri++) {
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
            for (int mu = 0; mu < R; mu++) {
              const double cc = ct[(ri * R + mu) * SS];
              F[mu] += cc * v_cur;
              Fd[mu] += cc * d_cur;
            }
          }
        }
        const long long s = slot0 + done + lane;
        double *rec = fld0 + s;    // field f of this record: rec[f * cap]
        rec[0] = r0 * invd;
        rec[cap] = r1 * invd;
        rec[2 * cap] = r2 * invd;
        rec[3 * cap] = dist;
#pragma unroll
        for (int mu = 0; mu < R; mu++) {
          rec[(4 + mu) * cap] = F[mu];
          rec[(4 + R + mu) * cap] = Fd[mu];
        }
        pj0[s] = jj.x;
        pjt0[s] = jt | (at.itype << 16);    // neighbor species | centre species
      }
      __syncwarp();
      head = (head + n) & (V2_RING - 1);
      count -= n;
      done += n;
    };
    for (int base0 = 0; base0 < at.jnum; base0 += 32 * V2_GB) {
      int jv[V2_GB], jtv[V2_GB];
      double nx[V2_GB], ny[V2_GB], nz[V2_GB];
#pragma unroll
      for (int b = 0; b < V2_GB; b++) {
        const int jj = base0 + 32 * b + lane;
        jv[b] = jj < at.jnum ? (a.neighbors[at.row0 + (long long) jj * stride_jj] & a.neighmask) : at.i;
      }
#pragma unroll
      for (int b = 0; b < V2_GB; b++)
        if (base0 + 32 * b < at.jnum) ld_atomrec(a.xt + jv[b], nx[b], ny[b], nz[b], jtv[b]);    // warp-uniform condition
#pragma unroll
      for (int b = 0; b < V2_GB; b++) {
        if (base0 + 32 * b >= at.jnum) break;
        const int jj = base0 + 32 * b + lane;
        bool within = false;
        const int j = jv[b], jt = jtv[b];
        double r0 = 0, r1 = 0, r2 = 0;
        if (jj < at.jnum) {
          r0 = nx[b] - at.xi0;
          r1 = ny[b] - at.xi1;
          r2 = nz[b] - at.xi2;
          // separately rounded, left to right, exactly pair_mtp.cpp:121-123 (no FMA contraction)
          const double rsq = __dadd_rn(__dadd_rn(__dmul_rn(r0, r0), __dmul_rn(r1, r1)), __dmul_rn(r2, r2));
          within = !(rsq > pot.cutsq);
          if (jt < 0 || jt >= pot.S) {    // pair_mtp.cpp:116-118
            atomicOr(a.status, 1);
            within = false;
          }
          if (!PLAIN && a.within) a.within[at.row0 + (long long) jj * stride_jj] = within ? 1 : 0;
        }
        const unsigned bal = __ballot_sync(FULL, within);
        if (within) {
          const int e = (head + count + __popc(bal & ((1u << lane) - 1u))) & (V2_RING - 1);
          ring[e] = r0;
          ring[V2_RING + e] = r1;
          ring[2 * V2_RING + e] = r2;
          reinterpret_cast<int2 *>(ring + 3 * V2_RING)[e] = make_int2(j, jt);
        }
        count += __popc(bal);
        if (count >= 32) flush(32);
      }
    }
    if (count > 0) flush(count);
    if (lane == 0) pb.pcnt[ii] = done;
  }
}

// ===================================================================================================== moments
// accumulate one pair into the accumulators of pass P:  acc[slot - pass_begin(P)] += f[mu] * x^a y^b z^c.
// The (a, b) loops are template recursion so that pass_of / col_begin are evaluated by the constexpr evaluator.
template <int D0, int P, int A, int B> struct V2FwdB {
  __device__ __forceinline__ static void run(double xy, double uy, double uz, const double (&f)[V2Shape<D0>::R],
                                             double (&acc)[V2Shape<D0>::AMAX])
  {
    using Sh = V2Shape<D0>;
    if constexpr (B <= D0 - A) {
      if constexpr (Sh::pass_of(A, B) == P) {
        constexpr int base = Sh::col_begin(A, B) - Sh::pass_begin(P);
        int idx = base;
        double m = xy;
#pragma unroll
        for (int c = 0; c <= D0 - A - B; c++) {
#pragma unroll
          for (int mu = 0; mu < Sh::rcnt(A + B + c); mu++) {
            acc[idx] = fma(f[mu], m, acc[idx]);
            idx++;
          }
          m *= uz;
        }
      }
      V2FwdB<D0, P, A, B + 1>::run(xy * uy, uy, uz, f, acc);
    }
  }
};
template <int D0, int P, int A> struct V2FwdA {
  __device__ __forceinline__ static void run(double xa, double ux, double uy, double uz,
                                             const double (&f)[V2Shape<D0>::R], double (&acc)[V2Shape<D0>::AMAX])
  {
    if constexpr (A <= D0) {
      V2FwdB<D0, P, A, 0>::run(xa, uy, uz, f, acc);
      V2FwdA<D0, P, A + 1>::run(xa * ux, ux, uy, uz, f, acc);
    }
  }
};
template <int D0, int P>
__device__ __forceinline__ void v2_fwd_accumulate(double ux, double uy, double uz,
                                                  const double (&f)[V2Shape<D0>::R], double (&acc)[V2Shape<D0>::AMAX])
{
  V2FwdA<D0, P, 0>::run(1.0, ux, uy, uz, f, acc);
}

// Work items of a CTA are (atom block, tile of V2_NT pairs) in order; the tile of item i + 1 is in flight
// (cp.async into the other buffer) while item i is consumed, across block boundaries too.
template <int D0, int P>
__device__ __forceinline__ void v2_moments_body(const SiteArgs &a, const PairBuf &pb, double *__restrict__ mb, int ld,
                                                double *tiles)
{
  using Sh = V2Shape<D0>;
  // A tile holds V2_NT pair records of each of the 32 atoms of a block, atom-major: row al = [field][pair], ROW doubles
  // long.  It is filled by straight 16-byte copies (two pairs of one field are contiguous in the pair buffer and in the
  // row).  Lane al reads its row; it takes the pairs of every aligned group of four in the order n ^ ((al >> 2) & 3):
  // with ROW = 4 (mod 16) the 16 lanes of a half warp then hit 16 different 8-byte bank pairs (4 (al & 3) + (n ^ q)).
  constexpr int R = Sh::R, NF = Sh::NF, NTH = 32 * Sh::NP, ROW = NF * V2_NT + 4, TILE = 32 * ROW;
  static_assert(V2_NT % 4 == 0 && (NF * V2_NT) % 16 == 0, "tile shape");
  const int lane = threadIdx.x & 31;
  const int nblk = (a.inum + 31) >> 5;

  auto load_cnt = [&](int blk) { return (blk < nblk && blk * 32 + lane < a.inum) ? pb.pcnt[blk * 32 + lane] : 0; };
  auto warp_max_i = [&](int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL, v, o));
    return v;
  };
  // stage tile (blk, n0) into buffer buf; cnt_l = this lane's pair count for atom `lane` of blk
  auto stage = [&](int buf, int blk, int n0, int cnt_l) {
    double *tile = tiles + (size_t) buf * TILE;
    // warps split the atoms; the lanes cover the NF x V2_NT / 2 16-byte chunks of an atom's row
    constexpr int CH = V2_NT / 2;
    const int warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int al = warp; al < 32; al += Sh::NP) {
      const int cnt_al = __shfl_sync(FULL, cnt_l, al);
      const double *abase = pb.fld + (size_t) (blk * 32 + al) * pb.ncap + n0;
      const unsigned trow = (unsigned) __cvta_generic_to_shared(tile + al * ROW);
#pragma unroll
      for (int c0 = 0; c0 < NF * CH; c0 += 32) {
        const int c = c0 + lane;
        if (c < NF * CH) {
          const int fi = c / CH, ch = c % CH;       // CH is a power of two
          const int gf = fi < 3 ? fi : fi + 1;      // skip the distance field
          const int live = min(2, max(0, cnt_al - (n0 + 2 * ch)));
          const double *src = abase + (size_t) gf * pb.cap + (live ? 2 * ch : 0);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(trow + (unsigned) (fi * V2_NT + 2 * ch) * 8u), "l"(src),
                       "r"(live * 8));
        }
      }
    }
    cp_async_commit();
  };

  int blk = blockIdx.x;
  if (blk >= nblk) return;
  int cnt_l = load_cnt(blk), nmax = warp_max_i(cnt_l), n0 = 0, buf = 0;
  int cnt_nb = load_cnt(blk + gridDim.x);    // next block's counts, one block ahead
  stage(0, blk, 0, cnt_l);
  double acc[Sh::AMAX];
#pragma unroll
  for (int t = 0; t < Sh::AMAX; t++) acc[t] = 0.0;

  while (true) {
    // next item
    int nblk_i = blk, nn0 = n0 + V2_NT, ncnt = cnt_l;
    const bool last_tile = nn0 >= nmax;
    if (last_tile) {
      nblk_i = blk + gridDim.x;
      nn0 = 0;
      ncnt = cnt_nb;
    }
    const bool have_next = nblk_i < nblk;
    if (have_next) {
      stage(buf ^ 1, nblk_i, nn0, ncnt);
      cp_async_wait<1>();
    } else
      cp_async_wait<0>();
    __syncthreads();
    {
      const double *row = tiles + (size_t) buf * TILE + lane * ROW;
      const int nt = (min(V2_NT, nmax - n0) + 3) & ~3, q = (lane >> 2) & 3;
      for (int n = 0; n < nt; n++) {
        const double *rec = row + (n ^ q);
        const double ux = rec[0], uy = rec[V2_NT], uz = rec[2 * V2_NT];
        double f[R];
#pragma unroll
        for (int mu = 0; mu < R; mu++) f[mu] = rec[(3 + mu) * V2_NT];
        v2_fwd_accumulate<D0, P>(ux, uy, uz, f, acc);    // padded records are all zero
      }
    }
    if (last_tile) {
      const int ii = blk * 32 + lane;
      if (ii < a.inum) {
        constexpr int beg = Sh::pass_begin(P), cnt = Sh::pass_begin(P + 1) - Sh::pass_begin(P);
#pragma unroll
        for (int t = 0; t < cnt; t++) mb[(size_t) (beg + t) * ld + ii] = acc[t];    // rows = canonical slots
      }
#pragma unroll
      for (int t = 0; t < Sh::AMAX; t++) acc[t] = 0.0;
    }
    __syncthreads();    // everyone is done with `buf` before it is refilled
    if (!have_next) break;
    if (last_tile) {
      blk = nblk_i;
      cnt_l = cnt_nb;
      nmax = warp_max_i(cnt_l);
      cnt_nb = load_cnt(blk + gridDim.x);
    }
    n0 = nn0;
    buf ^= 1;
  }
}

template <int D0, int P> struct V2PassDispatch {
  __device__ __forceinline__ static void run(int pass, const SiteArgs &a, const PairBuf &pb, double *mb, int ld,
                                             double *tiles)
  {
    if (pass == P) v2_moments_body<D0, P>(a, pb, mb, ld, tiles);
    else
      V2PassDispatch<D0, P + 1>::run(pass, a, pb, mb, ld, tiles);
  }
};
template <int D0> struct V2PassDispatch<D0, V2Shape<D0>::NP> {
  __device__ __forceinline__ static void run(int, const SiteArgs &, const PairBuf &, double *, int, double *) {}
};

// resident CTAs the kernel is compiled for: shared memory allows three; with many passes (warps) per CTA two are enough
// to keep ~16 warps on an SM, and the register cap stays above what a pass of MTP_V2_ACC accumulators needs
template <int D0> struct V2MomentsBounds {
  static constexpr int NP = V2Shape<D0>::NP;
  static constexpr int MINB = NP <= 5 ? 3 : (NP <= 8 ? 2 : 1);
};
template <int D0>
__global__ void __launch_bounds__(32 * V2Shape<D0>::NP, V2MomentsBounds<D0>::MINB)
mtp_moments_v2(SiteArgs a, PairBuf pb, double *__restrict__ mb, int ld)
{
  extern __shared__ __align__(16) unsigned char smem[];    // two tiles of 32 x (NF x V2_NT + 4) doubles
  V2PassDispatch<D0, 0>::run(threadIdx.x >> 5, a, pb, mb, ld, reinterpret_cast<double *>(smem));
}

// ===================================================================================================== forces
// gradient of sum_mu f_mu(d) P_mu(u), P_mu(u) = sum_q g[q][mu] u^q, by three nested Horner sweeps; g = canonical row
template <int D0, int AB, bool GRADE>
__device__ __forceinline__ void v2_pair_force(const double *__restrict__ gr /* g + al, row stride AB */, double ux, double uy, double uz,
                                              const double (&fvi)[V2Shape<D0>::R], const double (&fder)[V2Shape<D0>::R],
                                              double &Fx, double &Fy, double &Fz, double (&pm)[V2Shape<D0>::R])
{
  using Sh = V2Shape<D0>;
  constexpr int R = Sh::R;
  double Pv = 0, Px = 0, Py = 0, Pz = 0, Pd = 0;
  double PM[R];    // GRADE: P_mu(u) = sum_q g[q][mu] u^q, one Horner chain per radial index
#pragma unroll
  for (int mu = 0; mu < R; mu++) PM[mu] = 0.0;
  int sl = Sh::KF;
#pragma unroll
  for (int a = D0; a >= 0; a--) {
    double Q = 0, Qy = 0, Qz = 0, Qd = 0;
    double QM[R];
#pragma unroll
    for (int mu = 0; mu < R; mu++) QM[mu] = 0.0;
#pragma unroll
    for (int b = D0 - a; b >= 0; b--) {
      double T = 0, Tz = 0, Td = 0;
      double TM[R];
#pragma unroll
      for (int mu = 0; mu < R; mu++) TM[mu] = 0.0;
#pragma unroll
      for (int c = D0 - a - b; c >= 0; c--) {
        const int rc = Sh::rcnt(a + b + c);
        sl -= rc;
        double W = 0, Wd = 0;
#pragma unroll
        for (int mu = 0; mu < R; mu++) {
          if (mu < rc) {
            const double g = gr[(sl + mu) * AB];
            if (mu == 0) {
              W = fvi[0] * g;
              Wd = fder[0] * g;
            } else {
              W = fma(fvi[mu], g, W);
              Wd = fma(fder[mu], g, Wd);
            }
            if (GRADE) TM[mu] = (c == D0 - a - b) ? g : fma(TM[mu], uz, g);
          } else if (GRADE && c != D0 - a - b)
            TM[mu] *= uz;
        }
        if (c == D0 - a - b) {    // top of the z sweep: T = Tz = Td = 0
          T = W;
          Td = Wd;
        } else {
          Tz = fma(Tz, uz, T);
          T = fma(T, uz, W);
          Td = fma(Td, uz, Wd);
        }
      }
      if (b == D0 - a) {
        Q = T;
        Qz = Tz;
        Qd = Td;
        if (GRADE) {
#pragma unroll
          for (int mu = 0; mu < R; mu++) QM[mu] = TM[mu];
        }
      } else {
        Qy = fma(Qy, uy, Q);
        Q = fma(Q, uy, T);
        Qz = fma(Qz, uy, Tz);
        Qd = fma(Qd, uy, Td);
        if (GRADE) {
#pragma unroll
          for (int mu = 0; mu < R; mu++) QM[mu] = fma(QM[mu], uy, TM[mu]);
        }
      }
    }
    if (a == D0) {
      Pv = Q;
      Py = Qy;
      Pz = Qz;
      Pd = Qd;
      if (GRADE) {
#pragma unroll
        for (int mu = 0; mu < R; mu++) PM[mu] = QM[mu];
      }
    } else {
      Px = fma(Px, ux, Pv);
      Pv = fma(Pv, ux, Q);
      Py = fma(Py, ux, Qy);
      Pz = fma(Pz, ux, Qz);
      Pd = fma(Pd, ux, Qd);
      if (GRADE) {
#pragma unroll
        for (int mu = 0; mu < R; mu++) PM[mu] = fma(PM[mu], ux, QM[mu]);
      }
    }
  }
  const double S = Pd - (ux * Px + uy * Py + uz * Pz);
  Fx = fma(ux, S, Px);
  Fy = fma(uy, S, Py);
  Fz = fma(uz, S, Pz);
#pragma unroll
  for (int mu = 0; mu < R; mu++) pm[mu] = PM[mu];
}

// The pairs of a CTA's atom blocks form ONE stream: lane l of iteration t takes pair 256 t + l of the stream, whatever
// block it belongs to, so no lane idles at the end of a block (a block of 32 atoms x 26 pairs fills 3.25 iterations of
// 256 lanes).  Three blocks of adjoints are resident: the current one, the next one (an iteration may straddle the two)
// and the one after, in flight (cp.async).
constexpr int V2_FNB = 3;
template <int D0, int AB, bool GRADE>
__global__ void __launch_bounds__(256, 2)
mtp_forces_v2(SiteArgs a, PairBuf pb, const double *__restrict__ gb, int ld, double *__restrict__ partials)
{
  using Sh = V2Shape<D0>;
  constexpr int R = Sh::R, KF = Sh::KF, GSZ = KF * AB;
  extern __shared__ __align__(16) unsigned char smem[];
  double *gbuf = reinterpret_cast<double *>(smem);                            // [3][KF][AB] adjoints, rows = canonical slots
  int *prebuf = reinterpret_cast<int *>(gbuf + V2_FNB * (size_t) GSZ);        // [3][AB + 1] exclusive prefix of pcnt
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double vloc[6] = {0, 0, 0, 0, 0, 0};
  const int nblk = (a.inum + AB - 1) / AB, G = gridDim.x;

  // asynchronous fill of one block's adjoints (straight 16-byte copies of the rows of gb) + pair-count prefix
  auto prefetch = [&](int buf, int blk) {
    double *g = gbuf + (size_t) buf * GSZ;
    const int ii0 = blk * AB;
    for (int c = threadIdx.x; c < KF * (AB / 2); c += 256) {
      const int s = c / (AB / 2), c2 = c - s * (AB / 2);
      cp_async16(g + s * AB + 2 * c2, gb + (size_t) s * ld + ii0 + 2 * c2);
    }
    cp_async_commit();
    if (warp == 0) {
      int *pre = prebuf + buf * (AB + 1);
      const int na = min(AB, a.inum - ii0);
      int run = 0;
#pragma unroll
      for (int b0 = 0; b0 < AB; b0 += 32) {
        const int al = b0 + lane;
        const int c = (al < na) ? pb.pcnt[ii0 + al] : 0;
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += v;
        }
        if (al < AB) pre[al] = run + inc - c;
        run += __shfl_sync(FULL, inc, 31);
      }
      if (lane == 0) pre[AB] = run;
    }
  };

  int b_cur = blockIdx.x, buf_cur = 0, buf_nxt = 1, buf_nn = 2;
  if (b_cur < nblk) {
    prefetch(buf_cur, b_cur);
    if (b_cur + G < nblk) prefetch(buf_nxt, b_cur + G);
    cp_async_wait<0>();
    __syncthreads();
    if (b_cur + 2 * G < nblk) prefetch(buf_nn, b_cur + 2 * G);

    // a pair record, located in the stream: state 0 = past the end of everything this CTA owns, 1 = live,
    // 2 = beyond the two resident blocks (located again once the blocks have moved on)
    struct Rec {
      double ux, uy, uz, d, f[R], fd[R];
      size_t s;
      int j, al, which, state;
    };
    auto load_rec = [&](int q, Rec &r) {
      const int *pre_c = prebuf + buf_cur * (AB + 1), *pre_n = prebuf + buf_nxt * (AB + 1);
      const int tot_c = pre_c[AB];
      const bool has_n = b_cur + G < nblk;
      const int tot_n = has_n ? pre_n[AB] : 0;
      r.al = 0;
      r.j = 0;
      r.which = 0;
      r.s = 0;
      const int *pre = pre_c;
      int p = q;
      if (q >= tot_c) {
        p = q - tot_c;
        pre = pre_n;
        r.which = 1;
        if (!has_n) {
          r.state = 0;
          return;
        }
        if (p >= tot_n) {
          r.state = (b_cur + 2 * G < nblk) ? 2 : 0;
          return;
        }
      }
      r.state = 1;
      int lo = 0, hi = AB;    // largest al with pre[al] <= p
#pragma unroll
      for (int it = 0; (1 << it) < AB; it++) {
        const int mid = (lo + hi) >> 1;
        if (pre[mid] <= p) lo = mid;
        else
          hi = mid;
      }
      r.al = lo;
      const int ii0 = (r.which ? b_cur + G : b_cur) * AB;
      const size_t s = (size_t) (ii0 + lo) * pb.ncap + (p - pre[lo]);
      r.s = s;
      r.ux = pb.fld[s];
      r.uy = pb.fld[pb.cap + s];
      r.uz = pb.fld[2 * pb.cap + s];
      r.d = pb.fld[3 * pb.cap + s];
#pragma unroll
      for (int mu = 0; mu < R; mu++) {
        r.f[mu] = pb.fld[(4 + mu) * pb.cap + s];
        r.fd[mu] = pb.fld[(4 + R + mu) * pb.cap + s];
      }
      r.j = pb.pj[s];
    };

    int base = 0;
    Rec cur;
    load_rec(threadIdx.x, cur);
    while (true) {
      const int tot_c = prebuf[buf_cur * (AB + 1) + AB];
      if (base >= tot_c) {    // the current block is finished: the next becomes current, the one in flight next
        if (b_cur + G >= nblk) break;
        __syncthreads();      // everyone is done with the current buffer
        base -= tot_c;
        const int t = buf_cur;
        buf_cur = buf_nxt;
        buf_nxt = buf_nn;
        buf_nn = t;
        b_cur += G;
        cp_async_wait<0>();
        __syncthreads();      // the new next block has landed
        if (b_cur + 2 * G < nblk) prefetch(buf_nn, b_cur + 2 * G);
        if (cur.state == 1) cur.which = 0;    // (a located record of the old next block)
        else if (cur.state == 2)
          load_rec(base + threadIdx.x, cur);
        continue;
      }
      // an iteration serves the pairs [base, base + adv): all 256 lanes unless the two resident blocks end earlier
      // (tiny blocks) while another block is still to come -- those pairs are taken up after the blocks have moved on
      const int resident = tot_c - base + ((b_cur + G < nblk) ? prebuf[buf_nxt * (AB + 1) + AB] : 0);
      const int adv = (b_cur + 2 * G < nblk) ? min(256, resident) : 256;
      // pair records are requested one iteration ahead of their use
      Rec nxt;
      load_rec(base + adv + threadIdx.x, nxt);
      const bool live = cur.state == 1 && (int) threadIdx.x < adv;
      const int al = cur.al, which = cur.which;
      double Fx = 0, Fy = 0, Fz = 0;
      const int ii_l = (which ? b_cur + G : b_cur) * AB + al;
      const int i = a.ilist ? a.ilist[a.first_ii + ii_l] : a.first_ii + ii_l;
      if (live) {
        const double *g = gbuf + (size_t) (which ? buf_nxt : buf_cur) * GSZ;
        const double ux = cur.ux, uy = cur.uy, uz = cur.uz, d = cur.d;
        const double invd = 1.0 / d;
        double fvi[R], fder[R];
#pragma unroll
        for (int mu = 0; mu < R; mu++) {
          fvi[mu] = cur.f[mu] * invd;
          fder[mu] = cur.fd[mu];
        }
        double pm[R];
        v2_pair_force<D0, AB, GRADE>(g + al, ux, uy, uz, fvi, fder, Fx, Fy, Fz, pm);
        if (GRADE) {    // P_mu(u_n) replaces f_mu in the pair record (consumed by mtp_cand_radial_kernel)
#pragma unroll
          for (int mu = 0; mu < R; mu++) pb.fld[(4 + mu) * pb.cap + cur.s] = pm[mu];
        }
        const int j = cur.j;
        atomicAdd(&a.f[3 * (size_t) j], -Fx);
        atomicAdd(&a.f[3 * (size_t) j + 1], -Fy);
        atomicAdd(&a.f[3 * (size_t) j + 2], -Fz);
        if (a.vflag_any) {
          const double r0 = ux * d, r1 = uy * d, r2 = uz * d;
          const double v0 = Fx * r0, v1 = Fy * r1, v2 = Fz * r2;
          const double v3 = (Fx * r1 + Fy * r0) / 2, v4 = (Fx * r2 + Fz * r0) / 2, v5 = (Fy * r2 + Fz * r1) / 2;
          vloc[0] -= v0;
          vloc[1] -= v1;
          vloc[2] -= v2;
          vloc[3] -= v3;
          vloc[4] -= v4;
          vloc[5] -= v5;
          if (a.vflag_atom) {    // pair_mtp.cpp:268-276: all on the centre atom
            atomicAdd(&a.vatom[6 * (size_t) i], -v0);
            atomicAdd(&a.vatom[6 * (size_t) i + 1], -v1);
            atomicAdd(&a.vatom[6 * (size_t) i + 2], -v2);
            atomicAdd(&a.vatom[6 * (size_t) i + 3], -v3);
            atomicAdd(&a.vatom[6 * (size_t) i + 4], -v4);
            atomicAdd(&a.vatom[6 * (size_t) i + 5], -v5);
          }
        }
      }
      // centre atom: segmented sum over the lanes that share it (lanes are sorted by block, then atom)
      const int key = live ? which * AB + al : -1 - lane;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int ko = __shfl_down_sync(FULL, key, o);
        const double x = __shfl_down_sync(FULL, Fx, o), y = __shfl_down_sync(FULL, Fy, o), z = __shfl_down_sync(FULL, Fz, o);
        if (lane + o < 32 && ko == key) {
          Fx += x;
          Fy += y;
          Fz += z;
        }
      }
      const int kprev = __shfl_up_sync(FULL, key, 1);
      if (live && (lane == 0 || kprev != key)) {
        atomicAdd(&a.f[3 * (size_t) i], Fx);
        atomicAdd(&a.f[3 * (size_t) i + 1], Fy);
        atomicAdd(&a.f[3 * (size_t) i + 2], Fz);
      }
      cur = nxt;
      base += adv;
    }
  }

  // per-CTA virial partial, fixed order
  __shared__ double s_part[8][8];
#pragma unroll
  for (int c = 0; c < 6; c++) vloc[c] = warp_sum(vloc[c]);
  if (lane == 0) {
    s_part[warp][0] = 0.0;
#pragma unroll
    for (int c = 0; c < 6; c++) s_part[warp][1 + c] = vloc[c];
    s_part[warp][7] = 0.0;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int w = 0; w < (int) (blockDim.x >> 5); w++) s += s_part[w][threadIdx.x];
    partials[(size_t) blockIdx.x * 8 + threadIdx.x] = s;
  }
}

// Grade steps: radial block + species one-hot of the candidate vector of every centre atom
// (pair_mtp_extrapolation.cpp:193-198,235-252,322-329):
//     b[(it*S + jt)*R*B + mu*B + ri] = sum_n [jt_n == jt] phi_ri(d_n) P_mu(u_n),   b[S*S*R*B + it] = 1
// Warp per atom, lane = (mu, ri): no atomics, fixed summation order.  P_mu(u_n) was left in the f_mu fields of
// the pair records by mtp_forces_v2<GRADE>; the linear block was written by the program kernel.
constexpr int CAND_TILE = 32;    // pairs staged per warp
template <int SMAX>
__global__ void __launch_bounds__(256)
mtp_cand_radial_kernel(DevPotential pot, SiteArgs a, PairBuf pb)
{
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int R = pot.R, B = pot.B, RB = R * B, nrad = pot.S * pot.S * RB;
  // per warp: phi[B][TILE] (lane = pair computes the Chebyshev values once), P_mu[R][TILE], neighbor species[TILE]
  double *sphi = reinterpret_cast<double *>(smem) + (size_t) warp * (CAND_TILE * (B + R) + CAND_TILE / 2);
  double *spm = sphi + (size_t) B * CAND_TILE;
  int *sjt = reinterpret_cast<int *>(spm + (size_t) R * CAND_TILE);
  for (int ii = blockIdx.x * W + warp; ii < a.inum; ii += gridDim.x * W) {
    const int i = a.ilist ? a.ilist[a.first_ii + ii] : a.first_ii + ii;
    int itype = (int) a.xt[i].t;
    if (itype < 0 || itype >= pot.S) itype = 0;
    const int cnt = pb.pcnt[ii];
    double *row = a.cand_rows + (size_t) ii * a.cand_ld;
    for (int q = lane; q < nrad + pot.S; q += 32) row[q] = (q == nrad + itype) ? 1.0 : 0.0;
    for (int q = pot.Q + lane; q < a.cand_ld; q += 32) row[q] = 0.0;
    for (int c0 = 0; c0 < RB; c0 += 32) {
      const int c = c0 + lane;
      const int mu = c < RB ? c / B : 0, ri = c < RB ? c - mu * B : 0;
      double acc[SMAX];
#pragma unroll
      for (int s = 0; s < SMAX; s++) acc[s] = 0.0;
      for (int n0 = 0; n0 < cnt; n0 += CAND_TILE) {
        const int nt = min(CAND_TILE, cnt - n0);
        __syncwarp();
        if (lane < nt) {    // lane = pair: one coalesced round of loads, the Chebyshev recurrence once per pair
          const size_t sp = (size_t) ii * pb.ncap + n0 + lane;
          const double d = pb.fld[3 * pb.cap + sp];
          sjt[lane] = pb.pjt[sp] & 0xffff;
          for (int m = 0; m < R; m++) spm[m * CAND_TILE + lane] = pb.fld[(size_t) (4 + m) * pb.cap + sp];
          const double t = d - pot.rmax;
          const double ksi = (2 * d - (pot.rmin + pot.rmax)) / (pot.rmax - pot.rmin);
          double v_prev = pot.scaling * (1 * t * t), v_cur = pot.scaling * (ksi * t * t);
          sphi[lane] = v_prev;
          if (B > 1) sphi[CAND_TILE + lane] = v_cur;
          for (int k = 2; k < B; k++) {
            const double vn = 2 * ksi * v_cur - v_prev;
            v_prev = v_cur;
            v_cur = vn;
            sphi[k * CAND_TILE + lane] = vn;
          }
        }
        __syncwarp();
        const double *ph = sphi + ri * CAND_TILE, *pmv = spm + mu * CAND_TILE;
        for (int n = 0; n < nt; n++) {
          const double val = ph[n] * pmv[n];
          const int jt = sjt[n];
#pragma unroll
          for (int s = 0; s < SMAX; s++) acc[s] += (jt == s) ? val : 0.0;
        }
      }
      if (c < RB) {
#pragma unroll
        for (int s = 0; s < SMAX; s++)
          if (s < pot.S) row[(size_t) (itype * pot.S + s) * RB + c] = acc[s];
      }
    }
    __syncwarp();
  }
}

}    // namespace mtpb200
