// Contraction program, 4-atoms-per-lane form: forward pass (pair_mtp.cpp:196-201), site energy (:204-212) and
// reverse mode (:217-233) for a chunk of NA atoms per CTA.
//
// The 2-atoms-per-lane kernel (mtp_program_kernel) spends ~33 instructions per term step (descriptor unpacking,
// predicated end-of-node work for every term) and runs at one third of the issue rate with 8 warps per SM: it is
// bound by instruction latency, not by the shared-memory pipe.  Here
//   * a lane owns FOUR atoms (two 16-byte columns of a row); a "virtual warp" of NA/4 lanes evaluates one node;
//   * the VPW = 128/NA virtual warps of a physical warp work on a GROUP of nodes whose term lists were padded to one
//     common length at load (mtp_potential.hpp: Flat3Pass), so the end of a node is a warp-uniform branch: no
//     per-term flags, no predicated stores, and the accumulator starts from the node's seed (no "init" terms); a long
//     list is dealt to all the virtual warps of its group and the partial sums are combined by shuffles;
//   * a term descriptor is {byte offset a, byte offset b, FP64 coefficient}: one IADD per operand, nothing to unpack;
//   * operands are software-pipelined across term rows, groups and nodes: while the DFMAs of one 2-row iteration run,
//     the 8 LDS.128 of the next and the descriptors of the one after are in flight (two register stages, no copies).
// Row layout: [node][atom] FP64, NA*8 bytes per row, row M = 1.0, row M+1 = scratch.  A lane's two columns are l*16
// and NA*4 + l*16 (NA = 32: the 8 lanes of a virtual warp read 128 contiguous bytes; NA = 16: the two virtual warps of
// a quarter warp read opposite halves of their rows, so any two rows are conflict-free).
#pragma once

#include "mtp_device.cuh"

namespace mtpb200 {

constexpr int P3_THREADS = 256;    // 8 warps == G3_WARPS (mtp_potential.hpp); 16 warps (128 registers) measured no faster
constexpr int P3_WARPS = P3_THREADS / 32;
constexpr int P3_PAD_ROWS = 8;     // == G3_PAD_ROWS

struct DevFlat3Pass {
  const int *row_begin;       // [nlevels * P3_WARPS + 1]
  const int *group_begin;     // [nlevels * P3_WARPS + 1]
  const uint4 *terms;         // G3Term  [rows + pad][VPW]
  const uint4 *heads;         // G3Head  [groups + 2][VPW]
  int nlevels, nterms, nheads;    // array lengths in 16-byte words, padding included
};

struct Prog3Layout {
  size_t table_bytes, off_cg, off_epart, off_s2k, off_lin, off_map, off_begin, off_terms[2], off_heads[2], total;
};
__host__ __device__ inline Prog3Layout program3_layout(int M, int A, int na, int nslots, int ntf, int ntr, int nhf, int nhr,
                                                       bool dsmem, int nlevels_f = 8, int nlevels_r = 8)
{
  Prog3Layout L;
  L.table_bytes = (size_t) (M + 2) * na * 8;
  L.off_cg = L.table_bytes;
  L.off_epart = 2 * L.table_bytes;
  L.off_s2k = L.off_epart + (size_t) P3_WARPS * na * 8;
  size_t o = (L.off_s2k + (size_t) nslots * 2 + 15) & ~(size_t) 15;
  L.off_lin = o;
  o += (size_t) A * 8;
  L.off_map = o;
  o = (o + (size_t) A * 4 + 15) & ~(size_t) 15;
  L.off_begin = o;    // row_begin / group_begin of both passes
  o = (o + (size_t) 2 * ((nlevels_f + nlevels_r) * P3_WARPS + 2) * 4 + 15) & ~(size_t) 15;
  L.off_heads[0] = o;
  if (dsmem) o += (size_t) nhf * 16;
  L.off_heads[1] = o;
  if (dsmem) o += (size_t) nhr * 16;
  L.off_terms[0] = o;
  if (dsmem) o += (size_t) ntf * 16;
  L.off_terms[1] = o;
  if (dsmem) o += (size_t) ntr * 16;
  L.total = (o + 15) & ~(size_t) 15;
  return L;
}

template <bool DSMEM> __device__ __forceinline__ uint4 p3_ld(const uint4 *p)
{
  if (DSMEM) return *p;
  return __ldg(p);
}

struct P3Ops {
  double2 a0[2], a1[2], b0[2], b1[2];    // [term row of the iteration]: first / second column of operands a and b
  double c[2];
};
// 16-byte shared-memory load that is skipped (the registers keep their values) when on == 0.  Written as predicated
// PTX: as C++ the compiler turns the condition into a branch, which diverges between the virtual warps of a warp and
// serialises their loads.
__device__ __forceinline__ void p3_lds_if(double2 &v, unsigned addr, unsigned on)
{
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p ld.shared.v2.f64 {%0, %1}, [%2];\n\t}"
               : "+d"(v.x), "+d"(v.y)
               : "r"(addr), "r"(on));
}
__device__ __forceinline__ void p3_load(P3Ops &o, const uint4 &d0, const uint4 &d1, unsigned A0, unsigned A1, unsigned B0,
                                        unsigned B1)
{
  // padding slots (coef == +0.0) keep whatever operands the registers hold: no shared-memory traffic for them
  const unsigned on0 = d0.z | d0.w, on1 = d1.z | d1.w;
  p3_lds_if(o.a0[0], A0 + d0.x, on0);
  p3_lds_if(o.a1[0], A1 + d0.x, on0);
  p3_lds_if(o.b0[0], B0 + d0.y, on0);
  p3_lds_if(o.b1[0], B1 + d0.y, on0);
  p3_lds_if(o.a0[1], A0 + d1.x, on1);
  p3_lds_if(o.a1[1], A1 + d1.x, on1);
  p3_lds_if(o.b0[1], B0 + d1.y, on1);
  p3_lds_if(o.b1[1], B1 + d1.y, on1);
  o.c[0] = __hiloint2double((int) d0.w, (int) d0.z);
  o.c[1] = __hiloint2double((int) d1.w, (int) d1.z);
}
__device__ __forceinline__ void p3_math(const P3Ops &o, int u, double2 &acc0, double2 &acc1)
{
  acc0.x = fma(o.c[u] * o.a0[u].x, o.b0[u].x, acc0.x);
  acc0.y = fma(o.c[u] * o.a0[u].y, o.b0[u].y, acc0.y);
  acc1.x = fma(o.c[u] * o.a1[u].x, o.b1[u].x, acc1.x);
  acc1.y = fma(o.c[u] * o.a1[u].y, o.b1[u].y, acc1.y);
}
// sum over the VPW virtual warps of a physical warp (lanes that differ in the bits above log2(32 / VPW)), fixed order
template <int VPW> __device__ __forceinline__ double p3_vsum(double v)
{
#pragma unroll
  for (int o = 32 / VPW; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one pass (all dependency levels) of this warp's streams.  A0/A1: the lane's two columns of the table the first
// operand and the destination live in (forward: moments, reverse: adjoints); B0/B1: columns of the moments.
template <int VPW, bool DSMEM>
__device__ __forceinline__ void p3_run_pass(int nlevels, const int *row_begin, const int *group_begin, const uint4 *__restrict__ terms,
                                            const uint4 *__restrict__ heads, unsigned char *A0p, unsigned char *A1p,
                                            const unsigned char *B0p, const unsigned char *B1p, int warp, int vq)
{
  const unsigned A0 = (unsigned) __cvta_generic_to_shared(A0p), A1 = (unsigned) __cvta_generic_to_shared(A1p);
  const unsigned B0 = (unsigned) __cvta_generic_to_shared(B0p), B1 = (unsigned) __cvta_generic_to_shared(B1p);
  for (int lv = 0; lv < nlevels; lv++) {
    const int sidx = lv * P3_WARPS + warp;
    const int r0 = row_begin[sidx], r1 = row_begin[sidx + 1];    // r1 - r0 is a multiple of 4
    if (r0 < r1) {
      const uint4 *hp = heads + (size_t) group_begin[sidx] * VPW + vq;
      uint4 hn = p3_ld<DSMEM>(hp);
      unsigned dst = hn.x;
      int remaining = (int) (hn.y & 0x7FFFFFFFu);
      bool split = (hn.y >> 31) != 0;
      double ini = __hiloint2double((int) hn.w, (int) hn.z);
      double2 acc0 = make_double2(ini, ini), acc1 = acc0;
      hn = p3_ld<DSMEM>(hp + VPW);
      hp += 2 * VPW;
      const uint4 *tp = terms + (size_t) r0 * VPW + vq;
      uint4 dA0 = p3_ld<DSMEM>(tp), dA1 = p3_ld<DSMEM>(tp + VPW);
      uint4 dB0 = p3_ld<DSMEM>(tp + 2 * VPW), dB1 = p3_ld<DSMEM>(tp + 3 * VPW);
      P3Ops oa, ob;
#pragma unroll
      for (int u = 0; u < 2; u++) {
        oa.a0[u] = oa.a1[u] = oa.b0[u] = oa.b1[u] = make_double2(0.0, 0.0);
        ob.a0[u] = ob.a1[u] = ob.b0[u] = ob.b1[u] = make_double2(0.0, 0.0);
      }
      p3_load(oa, dA0, dA1, A0, A1, B0, B1);
      dA0 = p3_ld<DSMEM>(tp + 4 * VPW);
      dA1 = p3_ld<DSMEM>(tp + 5 * VPW);
      tp += 6 * VPW;
      auto end_of_group = [&]() {
        if (split) {    // warp-uniform: one node, partial sums in the virtual warps
          if (VPW == 8 && (vq & 1)) {    // 16 atoms per CTA: odd virtual warps hold the two columns swapped
            const double2 t = acc0;
            acc0 = acc1;
            acc1 = t;
          }
          acc0.x = p3_vsum<VPW>(acc0.x);
          acc0.y = p3_vsum<VPW>(acc0.y);
          acc1.x = p3_vsum<VPW>(acc1.x);
          acc1.y = p3_vsum<VPW>(acc1.y);
        }
        if (!split || vq == 0) {
          *reinterpret_cast<double2 *>(A0p + dst) = acc0;
          *reinterpret_cast<double2 *>(A1p + dst) = acc1;
        }
        dst = hn.x;
        remaining = (int) (hn.y & 0x7FFFFFFFu);
        split = (hn.y >> 31) != 0;
        ini = __hiloint2double((int) hn.w, (int) hn.z);
        acc0 = make_double2(ini, ini);
        acc1 = acc0;
        hn = p3_ld<DSMEM>(hp);
        hp += VPW;
      };
#pragma unroll 1
      for (int r = r0; r < r1; r += 4) {
        p3_load(ob, dB0, dB1, A0, A1, B0, B1);
        dB0 = p3_ld<DSMEM>(tp);
        dB1 = p3_ld<DSMEM>(tp + VPW);
        p3_math(oa, 0, acc0, acc1);
        if (--remaining == 0) end_of_group();
        p3_math(oa, 1, acc0, acc1);
        if (--remaining == 0) end_of_group();
        p3_load(oa, dA0, dA1, A0, A1, B0, B1);
        dA0 = p3_ld<DSMEM>(tp + 2 * VPW);
        dA1 = p3_ld<DSMEM>(tp + 3 * VPW);
        tp += 4 * VPW;
        p3_math(ob, 0, acc0, acc1);
        if (--remaining == 0) end_of_group();
        p3_math(ob, 1, acc0, acc1);
        if (--remaining == 0) end_of_group();
      }
    }
    __syncthreads();
  }
}

template <int NA, bool GRADE, bool DSMEM>
__global__ void __launch_bounds__(P3_THREADS, 1)
mtp_program_v3(DevPotential pot, SiteArgs a, DevFlat3Pass pf, DevFlat3Pass pr, const double *__restrict__ mb,
               double *__restrict__ gb, int ld, double *__restrict__ partials)
{
  static_assert(NA == 32 || NA == 16, "atoms per CTA");
  constexpr int LPV = NA / 4, VW = P3_THREADS / LPV, VPW = 32 / LPV, HALF = NA * 4, NP2 = NA / 2;
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nslots = a.slot_to_k ? a.nslots : pot.K;
  const Prog3Layout L = program3_layout(pot.M, pot.A, NA, nslots, pf.nterms, pr.nterms, pf.nheads, pr.nheads, DSMEM,
                                        pf.nlevels, pr.nlevels);
  double *cm = reinterpret_cast<double *>(smem);
  double *cg = reinterpret_cast<double *>(smem + L.off_cg);
  double *epart = reinterpret_cast<double *>(smem + L.off_epart);
  short *s2k = reinterpret_cast<short *>(smem + L.off_s2k);
  double *s_lin = reinterpret_cast<double *>(smem + L.off_lin);
  int *s_map = reinterpret_cast<int *>(smem + L.off_map);
  const int l = threadIdx.x & (LPV - 1), vwarp = threadIdx.x / LPV;
  const int swz = NA == 16 ? (vwarp & 1) : 0;
  const unsigned off0 = (swz ? HALF : 0) + l * 16, off1 = (swz ? 0 : HALF) + l * 16;    // byte offsets == atom * 8
  const int radial_count = pot.S * pot.S * pot.R * pot.B;
  double e_thread = 0.0;

  // one-time setup: constant rows, slot map, energy tables, term streams
  for (int t = threadIdx.x; t < 2 * NA; t += blockDim.x) {
    cm[pot.M * NA + t] = t < NA ? 1.0 : 0.0;
    cg[pot.M * NA + t] = t < NA ? 1.0 : 0.0;
  }
  for (int t = threadIdx.x; t < nslots; t += blockDim.x) s2k[t] = a.slot_to_k ? a.slot_to_k[t] : (short) t;
  for (int t = threadIdx.x; t < pot.A; t += blockDim.x) {
    s_lin[t] = pot.lin[t];
    s_map[t] = pot.map[t] * NA;    // row offset in doubles
  }
  // stream tables of both passes: {row_begin, group_begin} forward, then reverse
  int *s_begin = reinterpret_cast<int *>(smem + L.off_begin);
  const int nbf = pf.nlevels * P3_WARPS + 1, nbr = pr.nlevels * P3_WARPS + 1;
  for (int t = threadIdx.x; t < nbf; t += blockDim.x) {
    s_begin[t] = pf.row_begin[t];
    s_begin[nbf + t] = pf.group_begin[t];
  }
  for (int t = threadIdx.x; t < nbr; t += blockDim.x) {
    s_begin[2 * nbf + t] = pr.row_begin[t];
    s_begin[2 * nbf + nbr + t] = pr.group_begin[t];
  }
  const uint4 *terms_f = pf.terms, *terms_r = pr.terms;
  const uint4 *heads_f = pf.heads, *heads_r = pr.heads;
  if (DSMEM) {
    uint4 *tf = reinterpret_cast<uint4 *>(smem + L.off_terms[0]), *tr = reinterpret_cast<uint4 *>(smem + L.off_terms[1]);
    uint4 *hf = reinterpret_cast<uint4 *>(smem + L.off_heads[0]), *hr = reinterpret_cast<uint4 *>(smem + L.off_heads[1]);
    for (int t = threadIdx.x; t < pf.nterms; t += blockDim.x) tf[t] = pf.terms[t];
    for (int t = threadIdx.x; t < pr.nterms; t += blockDim.x) tr[t] = pr.terms[t];
    for (int t = threadIdx.x; t < pf.nheads; t += blockDim.x) hf[t] = pf.heads[t];
    for (int t = threadIdx.x; t < pr.nheads; t += blockDim.x) hr[t] = pr.heads[t];
    terms_f = tf;
    terms_r = tr;
    heads_f = hf;
    heads_r = hr;
  }
  const int vq = lane / LPV;    // virtual warp within the physical warp
  unsigned char *cm0 = smem + off0, *cm1 = smem + off1;
  unsigned char *cg0 = smem + L.off_cg + off0, *cg1 = smem + L.off_cg + off1;
  __syncthreads();

  for (int chunk0 = blockIdx.x * NA; chunk0 < a.inum; chunk0 += gridDim.x * NA) {
    const int na = min(NA, a.inum - chunk0);
    int my_i = 0, my_type = 0;
    if ((int) threadIdx.x < na && (a.eflag_global || a.eflag_atom)) {
      my_i = a.ilist ? a.ilist[a.first_ii + chunk0 + threadIdx.x] : a.first_ii + chunk0 + threadIdx.x;
      my_type = (int) a.xt[my_i].t;
    }
    // basic moments of the chunk -> their rows (16-byte cp.async, zero fill past the end of the list)
    for (int t = threadIdx.x; t < ((a.prog_debug & 8) ? 0 : nslots * NP2); t += blockDim.x) {
      const int s = t / NP2, al = (t % NP2) * 2;
      const int k = s2k[s];
      if (k >= 0) {
        const int nb = max(0, min(2, na - al)) * 8;
        const unsigned dsm = (unsigned) __cvta_generic_to_shared(cm + k * NA + al);
        const double *src = mb + (size_t) s * ld + chunk0 + (nb ? al : 0);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dsm), "l"(src), "r"(nb));
      }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    if (!(a.prog_debug & 1)) p3_run_pass<VPW, DSMEM>(pf.nlevels, s_begin, s_begin + nbf, terms_f, heads_f, cm0, cm1, cm0, cm1, warp, vq);

    // site energies: the virtual warps split the basis functions; fixed-order reduction
    if ((a.eflag_global || a.eflag_atom || GRADE) && !(a.prog_debug & 4)) {
      double2 e0 = make_double2(0.0, 0.0), e1 = e0;
      for (int s = vwarp; s < pot.A; s += VW) {
        const double2 b0 = *reinterpret_cast<const double2 *>(cm0 + (size_t) s_map[s] * 8);
        const double2 b1 = *reinterpret_cast<const double2 *>(cm1 + (size_t) s_map[s] * 8);
        const double c = s_lin[s];
        e0.x = fma(c, b0.x, e0.x);
        e0.y = fma(c, b0.y, e0.y);
        e1.x = fma(c, b1.x, e1.x);
        e1.y = fma(c, b1.y, e1.y);
        if (GRADE) {
          const int c0 = off0 >> 3, c1 = off1 >> 3;
          const size_t col = (size_t) radial_count + pot.S + s;
          if (c0 < na) a.cand_rows[(size_t) (chunk0 + c0) * a.cand_ld + col] = b0.x;
          if (c0 + 1 < na) a.cand_rows[(size_t) (chunk0 + c0 + 1) * a.cand_ld + col] = b0.y;
          if (c1 < na) a.cand_rows[(size_t) (chunk0 + c1) * a.cand_ld + col] = b1.x;
          if (c1 + 1 < na) a.cand_rows[(size_t) (chunk0 + c1 + 1) * a.cand_ld + col] = b1.y;
        }
      }
      // sum over the virtual warps of this warp (fixed order), then one row of partials per warp
      if (VPW == 8 && (vq & 1)) {    // 16 atoms per CTA: odd virtual warps hold the two columns swapped
        const double2 t = e0;
        e0 = e1;
        e1 = t;
      }
      e0.x = p3_vsum<VPW>(e0.x);
      e0.y = p3_vsum<VPW>(e0.y);
      e1.x = p3_vsum<VPW>(e1.x);
      e1.y = p3_vsum<VPW>(e1.y);
      if (vq == 0) {
        unsigned char *ep = reinterpret_cast<unsigned char *>(epart) + (size_t) warp * NA * 8;
        *reinterpret_cast<double2 *>(ep + off0) = e0;
        *reinterpret_cast<double2 *>(ep + off1) = e1;
      }
      __syncthreads();
      if ((int) threadIdx.x < na) {
        int itype = my_type;
        if (itype < 0 || itype >= pot.S) itype = 0;
        double es = 0.0;
        for (int v = 0; v < P3_WARPS; v++) es += epart[v * NA + threadIdx.x];
        es += pot.species[itype];
        if (a.eflag_atom) a.eatom[my_i] = es;
        if (a.eflag_global) e_thread += es;
      }
    }

    if (!(a.prog_debug & 2)) p3_run_pass<VPW, DSMEM>(pr.nlevels, s_begin + 2 * nbf, s_begin + 2 * nbf + nbr, terms_r, heads_r, cg0, cg1, cm0, cm1, warp, vq);

    // adjoints of the basic moments -> gb
    for (int t = threadIdx.x; t < ((a.prog_debug & 16) ? 0 : nslots * NP2); t += blockDim.x) {
      const int s = t / NP2, al = (t % NP2) * 2;
      const int k = s2k[s];
      const double2 g = k >= 0 ? *reinterpret_cast<const double2 *>(cg + k * NA + al) : make_double2(0.0, 0.0);
      double *dstp = gb + (size_t) s * ld + chunk0 + al;
      if (al + 1 < na) *reinterpret_cast<double2 *>(dstp) = g;
      else if (al < na)
        *dstp = g.x;
    }
    __syncthreads();
  }

  // per-CTA energy partial (fixed order): only the first NA threads hold per-atom energies
  if (warp == 0) {
    double s = e_thread;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane < 8) partials[(size_t) blockIdx.x * 8 + lane] = lane == 0 ? s : 0.0;
  }
}

}    // namespace mtpb200
