// Host emulation of the generated contraction-program kernel (test infrastructure, no GPU).
//
// The source emitted by mtp_codegen_source() is plain C under a few macros; this harness compiles it for the host
// (-DP4_SOURCE="<file>", -DP4_TWO when the kernel keeps two atoms per lane) and executes it the way the kernel
// schedules it: stage by stage, every warp's share lane by lane, with the shared-memory stores of a stage becoming
// visible only at the stage's barrier (a read of a row written in the same stage -- a missing dependency -- therefore
// shows up as a wrong value).  The result is compared with the reference's sequential program, restated here from
// pair_mtp.cpp:196-233.
//
// usage: p4_host_check tables.txt     (text: K M T A, then T x {a0 a1 mult a3}, A x map, A x lin)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define P4_HOST
#define P4_FN static inline
#define P4_STAGE_FN static
#define P4_TABLE static const
#define P4_PARAMS P4Ctx &x
#define P4_RET void
#define P4_RETURN return
#define P4_CALL(f) f(x)
#define P4_GCALL(f) f(x)    // sparse / grouped form: the harness plays ONE atom group (the group loop is kernel plumbing)

#ifdef P4_TWO
struct T_ {
  double x, y;
};
static inline T_ t2(double a, double b)
{
  T_ r;
  r.x = a;
  r.y = b;
  return r;
}
#else
typedef double T_;
#endif

struct P4Ctx {
  const double *S;      // rows visible in this stage, + atom offset
  double *Snext;        // rows as they will be after the barrier
  const double *lin;
  double *gb, *cand;
  long long ld, cand_ld;
  int grade;
  T_ e;
};

#ifdef P4_TWO
#define LD(r) t2(x.S[(r) * P4_NA], x.S[(r) * P4_NA + 1])
#define ST(r, v) do { const T_ v_ = (v); x.Snext[(r) * P4_NA] = v_.x; x.Snext[(r) * P4_NA + 1] = v_.y; } while (0)
#define MUL(a, b) t2((a).x * (b).x, (a).y * (b).y)
#define ADD(a, b) t2((a).x + (b).x, (a).y + (b).y)
#define FMA(a, b, c) t2(std::fma((a).x, (b).x, (c).x), std::fma((a).y, (b).y, (c).y))
#define MULK(k, a) t2((k) * (a).x, (k) * (a).y)
#define FMAK(k, a, c) t2(std::fma((k), (a).x, (c).x), std::fma((k), (a).y, (c).y))
#define MULU(u, a) MULK(u, a)
#define FMAU(u, a, c) FMAK(u, a, c)
#define SPLAT(u) t2((u), (u))
#define ZERO t2(0.0, 0.0)
#define GBST(slot, v) do { const T_ v_ = (v); x.gb[(long long) (slot) * x.ld] = v_.x; x.gb[(long long) (slot) * x.ld + 1] = v_.y; } while (0)
#define GBACC(slot, v) do { const T_ v_ = (v); x.gb[(long long) (slot) * x.ld] += v_.x; x.gb[(long long) (slot) * x.ld + 1] += v_.y; } while (0)
#define ESC(s, v) do { const T_ v_ = (v); x.e = FMAK(x.lin[s], v_, x.e); \
    if (x.grade) { x.cand[s] = v_.x; x.cand[x.cand_ld + (s)] = v_.y; } } while (0)
#else
#define LD(r) (x.S[(r) * P4_NA])
#define ST(r, v) (x.Snext[(r) * P4_NA] = (v))
#define MUL(a, b) ((a) * (b))
#define ADD(a, b) ((a) + (b))
#define FMA(a, b, c) std::fma((a), (b), (c))
#define MULK(k, a) ((k) * (a))
#define FMAK(k, a, c) std::fma((k), (a), (c))
#define MULU(u, a) ((u) * (a))
#define FMAU(u, a, c) std::fma((u), (a), (c))
#define SPLAT(u) (u)
#define ZERO 0.0
#define GBST(slot, v) (x.gb[(long long) (slot) * x.ld] = (v))
#define GBACC(slot, v) (x.gb[(long long) (slot) * x.ld] += (v))
#define ESC(s, v) do { x.e = std::fma(x.lin[s], (v), x.e); if (x.grade) x.cand[s] = (v); } while (0)
#endif
#define LIN(s) (x.lin[s])

#include P4_SOURCE

int main(int argc, char **argv)
{
  if (argc < 2) return 2;
  FILE *f = fopen(argv[1], "r");
  if (!f) return 2;
  int K, M, T, A;
  if (fscanf(f, "%d %d %d %d", &K, &M, &T, &A) != 4) return 2;
  std::vector<int> times(4 * (size_t) T), map(A);
  std::vector<double> lin(A);
  for (int &v : times)
    if (fscanf(f, "%d", &v) != 1) return 2;
  for (int &v : map)
    if (fscanf(f, "%d", &v) != 1) return 2;
  for (double &v : lin)
    if (fscanf(f, "%lf", &v) != 1) return 2;
  fclose(f);
  if (K != P4_K || M != P4_M || A != P4_A) {
    printf("table sizes do not match the generated source\n");
    return 1;
  }
  const int NA = P4_NA, ld = NA;
  // pseudo-random basic moments per atom
  std::vector<double> basic((size_t) K * NA);
  unsigned long long seed = 88172645463325252ULL;
  for (double &v : basic) {
    seed ^= seed << 13;
    seed ^= seed >> 7;
    seed ^= seed << 17;
    v = ((double) (seed >> 11) / 9007199254740992.0 - 0.5) * 1.6;
  }
  // ---- sequential program, pair_mtp.cpp:196-233 ----
  std::vector<double> mref((size_t) M * NA, 0.0), gref((size_t) M * NA, 0.0), eref(NA, 0.0);
  for (int at = 0; at < NA; at++) {
    std::vector<double> m(M, 0.0), g(M, 0.0);
    for (int k = 0; k < K; k++) m[k] = basic[(size_t) k * NA + at];
    for (int e = 0; e < T; e++) m[times[4 * e + 3]] += times[4 * e + 2] * m[times[4 * e]] * m[times[4 * e + 1]];    // :196-201
    for (int s = 0; s < A; s++) eref[at] += lin[s] * m[map[s]];                                                       // :207-209
    for (int s = 0; s < A; s++) g[map[s]] = lin[s];                                                                   // :217-218
    for (int e = T - 1; e >= 0; e--) {                                                                                // :221-233
      const int a0 = times[4 * e], a1 = times[4 * e + 1], a3 = times[4 * e + 3];
      const double mult = times[4 * e + 2];
      g[a1] += g[a3] * mult * m[a0];
      g[a0] += g[a3] * mult * m[a1];
    }
    for (int n = 0; n < M; n++) {
      mref[(size_t) n * NA + at] = m[n];
      gref[(size_t) n * NA + at] = g[n];
    }
  }
  // ---- generated program, scheduled like the kernel ----
  std::vector<double> S((size_t) P4_ROWS * NA, 1e300), Snext;    // poison: reading a row that was never written is visible
#ifndef P4_SPARSE
  for (int k = 0; k < K; k++)
    for (int at = 0; at < NA; at++) S[(size_t) k * NA + at] = basic[(size_t) k * NA + at];
#endif
  Snext = S;
#if defined(P4_RPAR) && P4_RPAR
  const double gb0 = 0.0;      // rounds in parallel: every round ADDS to a zeroed gb (the rounds are played in order here)
#else
  const double gb0 = 1e300;    // poison: the first round must write every row
#endif
  std::vector<double> gb((size_t) P4_NSLOTS * ld, gb0), cand((size_t) NA * A, 1e300);
  std::vector<T_> e((size_t) P4_W * 32, ZERO);
  const int lanes = NA / P4_APL;
  int round = 0;
  for (int st = 0; st < P4_NSTAGE; st++) {
#ifdef P4_SPARSE
    if (st == p4_round_stage0[round]) {    // a round starts: every row is stale, then the basic moments it reads are staged
      S.assign(S.size(), 1e300);
      std::vector<int> slot2k(P4_NSLOTS, -1);
      for (int k = 0; k < K; k++) slot2k[p4_slot_of_k[k]] = k;
      for (int i = p4_stage_off[round]; i < p4_stage_off[round + 1]; i++)
        for (int at = 0; at < NA; at++) S[(size_t) p4_stage_row[i] * NA + at] = basic[(size_t) slot2k[p4_stage_slot[i]] * NA + at];
      Snext = S;
      round++;
    }
#endif
    (void) round;
    for (int w = 0; w < P4_W; w++)
      for (int lane = 0; lane < lanes; lane++) {
        const int al = lane * P4_APL;
        P4Ctx x;
        x.S = S.data() + al;
        x.Snext = Snext.data() + al;
        x.lin = lin.data();
        x.gb = gb.data() + al;
        x.ld = ld;
        x.cand = cand.data() + (size_t) al * A;
        x.cand_ld = A;
        x.grade = 1;
        x.e = e[(size_t) w * 32 + lane];
        p4_run_stage(st, w, x);
        e[(size_t) w * 32 + lane] = x.e;
      }
    S = Snext;    // the barrier
  }
  // ---- compare ----
  double worst = 0.0;
  auto rel = [&](double got, double want, double scale) {
    const double d = std::fabs(got - want) / scale;
    if (!(d <= worst)) worst = std::isnan(d) ? 1e300 : d;
  };
  double mscale = 0.0, gscale = 0.0;
  for (int n = 0; n < M; n++)
    for (int at = 0; at < NA; at++) {
      mscale = std::fmax(mscale, std::fabs(mref[(size_t) n * NA + at]));
      if (n < K) gscale = std::fmax(gscale, std::fabs(gref[(size_t) n * NA + at]));
    }
  for (int n = 0; n < M; n++)
    if (p4_mrow[n] >= 0)
      for (int at = 0; at < NA; at++) {
        const double want = mref[(size_t) n * NA + at];
        rel(S[(size_t) p4_mrow[n] * NA + at], want, std::fmax(std::fabs(want), 1e-6 * mscale));
      }
  const double m_err = worst;
  worst = 0.0;
  for (int k = 0; k < K; k++)
    for (int at = 0; at < NA; at++) rel(gb[(size_t) p4_slot_of_k[k] * ld + at], gref[(size_t) k * NA + at], gscale);
  const double g_err = worst;
  worst = 0.0;
  for (int at = 0; at < NA; at++) {
    double es = 0.0;
    for (int w = 0; w < P4_W; w++) {
#ifdef P4_TWO
      const T_ v = e[(size_t) w * 32 + at / 2];
      es += (at & 1) ? v.y : v.x;
#else
      es += e[(size_t) w * 32 + at];
#endif
    }
    rel(es, eref[at], std::fabs(eref[at]));
    for (int s = 0; s < A; s++) {
      const double want = mref[(size_t) map[s] * NA + at];
      rel(cand[(size_t) at * A + s], want, std::fmax(std::fabs(want), 1e-6 * mscale));
    }
  }
  printf("moments %.3e adjoints %.3e energy+candidate %.3e\n", m_err, g_err, worst);
  return (m_err < 1e-11 && g_err < 1e-11 && worst < 1e-11) ? 0 : 1;
}
