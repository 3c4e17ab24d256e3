// Generator of the per-potential contraction-program kernel (see mtp_codegen.hpp).
#include "mtp_codegen.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <unordered_map>

namespace mtpb200 {

namespace {

struct Edge {
  int a0, a1, c;
};
struct RevTerm {
  int other, coef;    // g[n] += coef * g[t] * m[other]
};
struct RevPair {
  int t;
  std::vector<RevTerm> terms;
};

struct Analysis {
  int K = 0, M = 0, T = 0, A = 0;
  std::vector<std::vector<Edge>> in;        // per target node: its products, file order
  std::vector<std::vector<RevPair>> out;    // per node: consumers grouped by target, ascending target
  std::vector<int> scalar;                  // index into linear_coeffs, or -1
  std::vector<char> operand, target;
  std::vector<int> mrow, grow;              // shared-memory row of m[n] / g[n], or -1
  std::vector<int> fstage, rstage;          // stage of the forward / reverse task of node n, or -1
  int m_rows = 0, g_rows = 0, nstages = 0;
  long long terms = 0;
  bool first_round = true;                  // every basic moment gets its adjoint written (later rounds: only the touched)
};

// active (optional, [A]): the basis functions this round evaluates; the program is then restricted to their ancestors,
// and the other basis functions that are ancestors themselves count as plain intermediates (adjoint seed 0)
bool analyse(const Potential &p, Analysis &an, std::string &why, const std::vector<char> *active = nullptr,
             bool first_round = true, bool sparse_basics = false)
{
  an.K = p.alpha_index_basic_count;
  an.M = p.alpha_moment_count;
  an.T = p.alpha_index_times_count;
  an.A = p.alpha_scalar_count;
  const int K = an.K, M = an.M, T = an.T;
  an.in.assign(M, {});
  an.out.assign(M, {});
  an.scalar.assign(M, -1);
  an.operand.assign(M, 0);
  an.target.assign(M, 0);
  an.first_round = first_round;
  const int *tm = p.alpha_index_times.data();
  for (int e = 0; e < T; e++)
    if (tm[4 * e + 3] < K) {
      why = "a basic moment is the target of a product";
      return false;
    }
  // nodes this round needs: ancestors of its basis functions (the file is topologically ordered: one backward sweep)
  std::vector<char> needed(M, active ? 0 : 1);
  if (active) {
    for (int s = 0; s < an.A; s++)
      if ((*active)[s]) needed[p.alpha_moment_mapping[s]] = 1;
    for (int e = T - 1; e >= 0; e--)
      if (needed[tm[4 * e + 3]]) needed[tm[4 * e]] = needed[tm[4 * e + 1]] = 1;
  }
  for (int e = 0; e < T; e++) {
    const int a0 = tm[4 * e], a1 = tm[4 * e + 1], c = tm[4 * e + 2], t = tm[4 * e + 3];
    if (!needed[t]) continue;
    an.in[t].push_back({a0, a1, c});
    an.operand[a0] = an.operand[a1] = 1;
    an.target[t] = 1;
  }
  for (int s = 0; s < an.A; s++) {
    const int n = p.alpha_moment_mapping[s];
    if (active && !(*active)[s]) continue;
    if (an.scalar[n] >= 0) {
      why = "alpha_moment_mapping lists a moment twice";
      return false;
    }
    an.scalar[n] = s;
  }
  // consumers of every node, grouped by target; a square m[a]*m[a] feeds its source twice (pair_mtp.cpp:229-232)
  for (int t = K; t < M; t++)
    for (const Edge &e : an.in[t]) {
      auto add = [&](int n, int other, int coef) {
        auto &v = an.out[n];
        if (v.empty() || v.back().t != t) v.push_back({t, {}});
        for (auto &rt : v.back().terms)
          if (rt.other == other) {
            rt.coef += coef;
            return;
          }
        v.back().terms.push_back({other, coef});
      };
      if (e.a0 == e.a1) add(e.a0, e.a0, 2 * e.c);
      else {
        add(e.a0, e.a1, e.c);
        add(e.a1, e.a0, e.c);
      }
    }
  for (int n = 0; n < M; n++)
    std::sort(an.out[n].begin(), an.out[n].end(), [](const RevPair &x, const RevPair &y) { return x.t < y.t; });
  // (targets are visited in ascending order above, so the sort is a no-op for sorted files; kept for safety)
  // rows
  an.mrow.assign(M, -1);
  an.grow.assign(M, -1);
  int r = 0;
  for (int n = 0; n < K; n++)    // sparse: only the basic moments this round reads (as a factor or as a basis function)
    if (!sparse_basics || an.operand[n] || an.scalar[n] >= 0) an.mrow[n] = r++;
  for (int n = K; n < M; n++)
    if (an.operand[n]) an.mrow[n] = r++;
  an.m_rows = r;
  int g = 0;
  for (int n = K; n < M; n++)
    if (an.operand[n]) an.grow[n] = r + g++;
  an.g_rows = g;
  // stages.  m of a basic moment is available in stage 0; a value stored in stage s is visible from stage s + 1
  an.fstage.assign(M, -1);
  an.rstage.assign(M, -1);
  for (int pass = 0; pass < M + 2; pass++) {
    bool changed = false;
    for (int t = K; t < M; t++) {
      if (!an.target[t]) continue;
      int s = 0;
      for (const Edge &e : an.in[t])
        for (int a : {e.a0, e.a1})
          if (a >= K) {
            if (!an.target[a]) continue;    // never written: stays 0
            s = std::max(s, an.fstage[a] < 0 ? 0 : an.fstage[a] + 1);
          }
      if (s != an.fstage[t]) {
        an.fstage[t] = s;
        changed = true;
      }
    }
    if (!changed) break;
    if (pass == M + 1) {
      why = "alpha_index_times has a dependency cycle";
      return false;
    }
  }
  for (int pass = 0; pass < M + 2; pass++) {
    bool changed = false;
    for (int n = M - 1; n >= 0; n--) {
      if (n >= K && !an.operand[n]) continue;
      int s = 0;
      for (const RevPair &rp : an.out[n]) {
        if (an.operand[rp.t]) s = std::max(s, an.rstage[rp.t] < 0 ? 0 : an.rstage[rp.t] + 1);
        for (const RevTerm &rt : rp.terms)
          if (rt.other >= K && an.target[rt.other]) s = std::max(s, an.fstage[rt.other] + 1);
      }
      if (s != an.rstage[n]) {
        an.rstage[n] = s;
        changed = true;
      }
    }
    if (!changed) break;
    if (pass == M + 1) {
      why = "alpha_index_times has a dependency cycle";
      return false;
    }
  }
  an.nstages = 1;
  for (int n = 0; n < M; n++) an.nstages = std::max(an.nstages, std::max(an.fstage[n], an.rstage[n]) + 1);
  an.terms = 0;
  for (int t = K; t < M; t++) an.terms += (long long) an.in[t].size();
  for (int n = 0; n < M; n++)
    for (const RevPair &rp : an.out[n]) an.terms += (long long) rp.terms.size();
  return true;
}

struct Task {
  int kind;    // 0 forward, 1 reverse, 2 energy of a basic moment that is a basis function
  int node;
  int cost;
};

// ---- emission with a register cache decided at generation time -------------------------------------------
struct Emitter {
  bool record = true;
  int capacity = 48;
  std::vector<int> seq;                 // recorded key sequence
  std::vector<int> next_use;            // per position: next position of the same key, or INT_MAX
  size_t pos = 0;
  std::unordered_map<int, std::pair<std::string, int>> live;    // key -> (variable, next use)
  std::string *out = nullptr;
  int nvar = 0;
  int uniform_base = 0;                 // keys >= uniform_base are uniform scalars (LIN)
  long long loads = 0;

  void begin_emit()
  {
    record = false;
    next_use.assign(seq.size(), 0x7fffffff);
    std::unordered_map<int, int> last;
    for (int i = (int) seq.size() - 1; i >= 0; i--) {
      auto it = last.find(seq[i]);
      if (it != last.end()) next_use[i] = it->second;
      last[seq[i]] = i;
    }
    pos = 0;
    live.clear();
  }
  // returns the expression (a variable name) holding row / uniform `key`
  std::string use(int key)
  {
    if (record) {
      seq.push_back(key);
      return "";
    }
    const int nu = next_use[pos++];
    auto it = live.find(key);
    if (it != live.end()) {
      it->second.second = nu;
      std::string name = it->second.first;
      if (nu == 0x7fffffff) live.erase(it);    // dead after this use
      return name;
    }
    char name[32];
    if (key >= uniform_base) {
      snprintf(name, sizeof(name), "u%d", nvar++);
      *out += std::string("  const double ") + name + " = LIN(" + std::to_string(key - uniform_base) + ");\n";
    } else {
      snprintf(name, sizeof(name), "v%d", nvar++);
      *out += std::string("  const T_ ") + name + " = LD(" + std::to_string(key) + ");\n";
      loads++;
    }
    if (nu != 0x7fffffff) {
      if ((int) live.size() >= capacity) {    // Belady: drop the value whose next use is farthest away
        auto worst = live.begin();
        for (auto j = live.begin(); j != live.end(); ++j)
          if (j->second.second > worst->second.second) worst = j;
        live.erase(worst);
      }
      live[key] = {name, nu};
    }
    return name;
  }
  void line(const std::string &s)
  {
    if (!record) *out += s;
  }
};

std::string kdouble(int c)
{
  char b[32];
  snprintf(b, sizeof(b), "%d.0", c);
  return b;
}

// forward evaluation of target t (pair_mtp.cpp:196-201)
void emit_forward(const Analysis &an, int t, Emitter &E, int &ntmp, long long &stores)
{
  std::map<int, std::vector<Edge>> by_coef;
  for (const Edge &e : an.in[t]) by_coef[e.c].push_back(e);
  const std::string acc = "a" + std::to_string(ntmp++);
  bool have = false;
  E.line("  T_ " + acc + ";\n");
  auto group = [&](int c, const std::vector<Edge> &es) {
    if (c == 1) {
      for (const Edge &e : es) {
        const std::string x0 = E.use(an.mrow[e.a0]), x1 = E.use(an.mrow[e.a1]);
        E.line("  " + acc + " = " + (have ? "FMA(" + x0 + ", " + x1 + ", " + acc + ")" : "MUL(" + x0 + ", " + x1 + ")") + ";\n");
        have = true;
      }
    } else if (es.size() == 1) {
      const Edge &e = es[0];
      const std::string x0 = E.use(an.mrow[e.a0]), x1 = E.use(an.mrow[e.a1]);
      E.line("  " + acc + " = " + (have ? "FMA(MULK(" + kdouble(c) + ", " + x0 + "), " + x1 + ", " + acc + ")"
                                        : "MUL(MULK(" + kdouble(c) + ", " + x0 + "), " + x1 + ")") + ";\n");
      have = true;
    } else {
      const std::string pv = "p" + std::to_string(ntmp++);
      bool hp = false;
      E.line("  T_ " + pv + ";\n");
      for (const Edge &e : es) {
        const std::string x0 = E.use(an.mrow[e.a0]), x1 = E.use(an.mrow[e.a1]);
        E.line("  " + pv + " = " + (hp ? "FMA(" + x0 + ", " + x1 + ", " + pv + ")" : "MUL(" + x0 + ", " + x1 + ")") + ";\n");
        hp = true;
      }
      E.line("  " + acc + " = " + (have ? "FMAK(" + kdouble(c) + ", " + pv + ", " + acc + ")" : "MULK(" + kdouble(c) + ", " + pv + ")") + ";\n");
      have = true;
    }
  };
  auto one = by_coef.find(1);
  if (one != by_coef.end()) group(1, one->second);
  for (const auto &kv : by_coef)
    if (kv.first != 1) group(kv.first, kv.second);
  if (!have) E.line("  " + acc + " = ZERO;\n");
  if (an.mrow[t] >= 0) {
    E.line("  ST(" + std::to_string(an.mrow[t]) + ", " + acc + ");\n");
    if (!E.record) stores++;
  }
  if (an.scalar[t] >= 0) E.line("  ESC(" + std::to_string(an.scalar[t]) + ", " + acc + ");\n");
}

// reverse-mode gather of the adjoints of `nodes` (pair_mtp.cpp:217-233 regrouped by source node)
void emit_reverse_block(const Analysis &an, const std::vector<int> &nodes, const short *slot_of_k, Emitter &E, int &ntmp,
                        long long &stores)
{
  struct Item {
    int t, i;
    const RevPair *rp;
  };
  std::vector<Item> items;
  std::vector<std::string> acc(nodes.size());
  std::vector<char> have(nodes.size(), 0);
  for (size_t i = 0; i < nodes.size(); i++) {
    const int n = nodes[i];
    acc[i] = "a" + std::to_string(ntmp++);
    if (an.scalar[n] >= 0) {    // g[map[s]] = xi_s  (pair_mtp.cpp:217-218)
      const std::string u = E.use(E.uniform_base + an.scalar[n]);
      E.line("  T_ " + acc[i] + " = SPLAT(" + u + ");\n");
      have[i] = 1;
    } else
      E.line("  T_ " + acc[i] + ";\n");
    for (const RevPair &rp : an.out[n]) {
      const bool variable = an.operand[rp.t] != 0;
      if (!variable && an.scalar[rp.t] < 0) continue;    // the consumer's adjoint is identically zero
      items.push_back({rp.t, (int) i, &rp});
    }
  }
  std::stable_sort(items.begin(), items.end(), [](const Item &x, const Item &y) { return x.t < y.t; });
  for (const Item &it : items) {
    const bool variable = an.operand[it.t] != 0;
    const std::string g = variable ? E.use(an.grow[it.t]) : E.use(E.uniform_base + an.scalar[it.t]);
    const std::string fma = variable ? "FMA(" : "FMAU(", mul = variable ? "MUL(" : "MULU(";
    std::vector<RevTerm> ts = it.rp->terms;
    std::stable_sort(ts.begin(), ts.end(), [](const RevTerm &x, const RevTerm &y) { return (x.coef != 1) < (y.coef != 1); });
    std::string s;
    if (ts.size() == 1) {
      const std::string m = E.use(an.mrow[ts[0].other]);
      s = ts[0].coef == 1 ? m : "MULK(" + kdouble(ts[0].coef) + ", " + m + ")";
    } else {
      s = "p" + std::to_string(ntmp++);
      bool hs = false;
      E.line("  T_ " + s + ";\n");
      for (const RevTerm &rt : ts) {
        const std::string m = E.use(an.mrow[rt.other]);
        if (!hs) E.line("  " + s + " = " + (rt.coef == 1 ? m : "MULK(" + kdouble(rt.coef) + ", " + m + ")") + ";\n");
        else if (rt.coef == 1)
          E.line("  " + s + " = ADD(" + s + ", " + m + ");\n");
        else
          E.line("  " + s + " = FMAK(" + kdouble(rt.coef) + ", " + m + ", " + s + ");\n");
        hs = true;
      }
    }
    const std::string &a = acc[it.i];
    E.line("  " + a + " = " + (have[it.i] ? fma + g + ", " + s + ", " + a + ")" : mul + g + ", " + s + ")") + ";\n");
    have[it.i] = 1;
  }
  for (size_t i = 0; i < nodes.size(); i++) {
    const int n = nodes[i];
    if (!have[i]) E.line("  " + acc[i] + " = ZERO;\n");
    if (n < an.K)
      E.line(std::string(an.first_round ? "  GBST(" : "  GBACC(") + std::to_string(slot_of_k ? (int) slot_of_k[n] : n) + ", " + acc[i] + ");\n");
    else
      E.line("  ST(" + std::to_string(an.grow[n]) + ", " + acc[i] + ");\n");
    if (!E.record) stores++;
  }
}

unsigned long long fnv(unsigned long long h, const void *data, size_t n)
{
  const unsigned char *b = (const unsigned char *) data;
  for (size_t i = 0; i < n; i++) {
    h ^= b[i];
    h *= 1099511628211ULL;
  }
  return h;
}

const char *kGeneratorVersion = "p4-r2-11";

// ---- fixed text: device prelude and kernel skeleton ------------------------------------------------------
const char *kDevicePrelude = R"P4(
#ifndef P4_HOST
struct P4Args {
  const double *mb;
  double *gb;
  long long ld;
  int inum, first_ii;
  const int *ilist;
  const void *xt;
  const double *lin, *species;
  int S;
  int eflag_global, eflag_atom, grade;
  double *eatom;
  double *cand_rows;
  long long cand_ld;
  int cand_col0;
  double *partials;
  double *esite;
};
#define P4_FN __device__ __forceinline__
#define P4_STAGE_FN __device__ __noinline__
#define P4_TABLE __device__ const
// Stage functions are compiled separately (noinline): everything they need arrives in registers, and shared memory is
// addressed explicitly (32-bit shared-window address + immediate offset), never through generic pointers.
#if P4_APL == 1
typedef double T_;
#else
struct T_ {
  double x, y;
};
#endif
#ifdef P4_SPARSE
struct P4Ctx {           // one context for the P4_G atom groups of the CTA iteration; group g: rows at sb + g * P4_GROUP_BYTES
  unsigned sb, lb;
  double *gb, *cand;
  long long ld, cand_ld;
  int na, al, fbase;     // atoms of this iteration, first atom of the lane within a group, bit 2: grade step, bit 3: owner lane
  int grp;               // side-by-side groups: the group of this warp
  T_ e[P4_GE];
};
#define P4_GROUP_BYTES (P4_ROWS * P4_NA * 8)
// every emitted function runs P4_G times in a row, once per atom group: the second to last executions find it in the
// instruction cache
#if P4_SPATIAL
#define P4_GLOOP for (int g_ = x.grp, e_ = 0; e_ < 1; e_++)
#else
#define P4_GLOOP _Pragma("unroll") for (int g_ = 0, e_ = 0; g_ < P4_G; g_++, e_++)
#endif
#define P4_GCALL(f) \
  P4_GLOOP { \
    const int c_ = g_ * P4_NA + x.al; \
    const int fl_ = (x.fbase & 4) | (((x.fbase & 8) && c_ < x.na) ? 1 : 0) | ((P4_APL == 2 && (x.fbase & 8) && c_ + 1 < x.na) ? 2 : 0); \
    x.e[e_] = f(x.sb + g_ * P4_GROUP_BYTES, x.lb, x.gb + g_ * P4_NA, x.ld, \
                x.cand ? x.cand + (long long) g_ * P4_NA * x.cand_ld : x.cand, x.cand_ld, fl_, x.e[e_]); \
  }
#else
struct P4Ctx {
  unsigned sb, lb;       // shared-window byte addresses: rows + this lane's atom offset, linear coefficients
  double *gb, *cand;
  long long ld, cand_ld;
  int flags;             // bit 0: first atom of the lane is a listed centre, bit 1: second atom, bit 2: grade step
  T_ e;
};
#endif
#define P4_PARAMS const unsigned sb, const unsigned lb, double *const gb, const long long ld, double *const cand, \
                  const long long cand_ld, const int flags, T_ e
#define P4_RET T_
#define P4_RETURN return e
#define P4_CALL(f) x.e = f(x.sb, x.lb, x.gb, x.ld, x.cand, x.cand_ld, x.flags, x.e)
template <int OFF> P4_FN double p4_lds1(unsigned a)
{
  double v;
  asm("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
  return v;
}
P4_FN void p4_red(double *q, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(q), "d"(v) : "memory"); }
#define LIN(s) p4_lds1<(s) * 8>(lb)
#if P4_APL == 1
template <int OFF> P4_FN void p4_sts(unsigned a, double v) { asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v) : "memory"); }
#define LD(r) p4_lds1<(r) * P4_NA * 8>(sb)
#define ST(r, v) p4_sts<(r) * P4_NA * 8>(sb, (v))
#define MUL(a, b) ((a) * (b))
#define ADD(a, b) ((a) + (b))
#define FMA(a, b, c) fma((a), (b), (c))
#define MULK(k, a) ((k) * (a))
#define FMAK(k, a, c) fma((k), (a), (c))
#define MULU(u, a) ((u) * (a))
#define FMAU(u, a, c) fma((u), (a), (c))
#define SPLAT(u) (u)
#define ZERO 0.0
#define GBST(slot, v) do { if (flags & 1) __stcg(gb + (long long) (slot) * ld, (v)); } while (0)
// later rounds ADD their share: a reduction without return value (RED.ADD.F64), so the round never waits for the old value;
// one lane owns an address and its updates are issued in program order, so the sum has a fixed order
#define GBACC(slot, v) do { if (flags & 1) p4_red(gb + (long long) (slot) * ld, (v)); } while (0)
#define ESC(s, v) do { e = fma(LIN(s), (v), e); if ((flags & 5) == 5) cand[s] = (v); } while (0)
#else
template <int OFF> P4_FN T_ p4_lds2(unsigned a)
{
  T_ v;
  asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(v.x), "=d"(v.y) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF> P4_FN void p4_sts(unsigned a, T_ v)
{
  asm volatile("st.shared.v2.f64 [%0+%1], {%2, %3};" ::"r"(a), "n"(OFF), "d"(v.x), "d"(v.y) : "memory");
}
P4_FN T_ p4_mul(T_ a, T_ b) { T_ r; r.x = a.x * b.x; r.y = a.y * b.y; return r; }
P4_FN T_ p4_add(T_ a, T_ b) { T_ r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
P4_FN T_ p4_fma(T_ a, T_ b, T_ c) { T_ r; r.x = fma(a.x, b.x, c.x); r.y = fma(a.y, b.y, c.y); return r; }
P4_FN T_ p4_mulu(double u, T_ a) { T_ r; r.x = u * a.x; r.y = u * a.y; return r; }
P4_FN T_ p4_fmau(double u, T_ a, T_ c) { T_ r; r.x = fma(u, a.x, c.x); r.y = fma(u, a.y, c.y); return r; }
P4_FN T_ p4_splat(double u) { T_ r; r.x = u; r.y = u; return r; }
#define LD(r) p4_lds2<(r) * P4_NA * 8>(sb)
#define ST(r, v) p4_sts<(r) * P4_NA * 8>(sb, (v))
#define MUL(a, b) p4_mul((a), (b))
#define ADD(a, b) p4_add((a), (b))
#define FMA(a, b, c) p4_fma((a), (b), (c))
#define MULK(k, a) p4_mulu((k), (a))
#define FMAK(k, a, c) p4_fmau((k), (a), (c))
#define MULU(u, a) p4_mulu((u), (a))
#define FMAU(u, a, c) p4_fmau((u), (a), (c))
#define SPLAT(u) p4_splat(u)
#define ZERO p4_splat(0.0)
#define GBST(slot, v) do { const T_ v_ = (v); double *q_ = gb + (long long) (slot) * ld; \
    if (flags & 2) __stcg(reinterpret_cast<double2 *>(q_), make_double2(v_.x, v_.y)); else if (flags & 1) __stcg(q_, v_.x); } while (0)
#define GBACC(slot, v) do { const T_ v_ = (v); double *q_ = gb + (long long) (slot) * ld; \
    if (flags & 1) p4_red(q_, v_.x); if (flags & 2) p4_red(q_ + 1, v_.y); } while (0)
#define ESC(s, v) do { const T_ v_ = (v); e = p4_fmau(LIN(s), v_, e); \
    if (flags & 4) { if (flags & 1) cand[s] = v_.x; if (flags & 2) cand[cand_ld + (s)] = v_.y; } } while (0)
#endif
#endif
)P4";

const char *kKernel = R"P4(
#ifndef P4_HOST
extern "C" __global__ void __launch_bounds__(P4_W * 32, P4_MINB) mtp_program_p4(const P4Args a)
{
  extern __shared__ __align__(16) double S[];
  double *s_lin = S + (size_t) P4_ROWS * P4_NA;
  double *epart = s_lin + ((P4_A + 1) & ~1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = tid; t < P4_A; t += P4_W * 32) s_lin[t] = a.lin[t];
  const int al = (P4_APL * lane) % P4_NA;        // first atom of this lane within the chunk
  const bool owner = P4_APL * lane < P4_NA;      // lanes beyond the chunk width repeat the work of another lane, never write
  double e_thread = 0.0;
  __syncthreads();
  for (int chunk0 = blockIdx.x * P4_NA; chunk0 < a.inum; chunk0 += gridDim.x * P4_NA) {
    const int na = min(P4_NA, a.inum - chunk0);
    // basic moments of the chunk -> rows 0 .. K-1 (16-byte cp.async, zero fill past the end of the list)
    for (int t = tid; t < P4_K * (P4_NA / 2); t += P4_W * 32) {
      const int k = t / (P4_NA / 2), c = (t % (P4_NA / 2)) * 2;
      const int nb = max(0, min(2, na - c)) * 8;
      const unsigned dst = (unsigned) __cvta_generic_to_shared(S + k * P4_NA + c);
      const double *src = a.mb + (long long) p4_slot_of_k[k] * a.ld + chunk0 + (nb ? c : 0);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(nb));
    }
    asm volatile("cp.async.commit_group;\n" ::);
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();
    P4Ctx x;
    x.sb = (unsigned) __cvta_generic_to_shared(S + al);
    x.lb = (unsigned) __cvta_generic_to_shared(s_lin);
    x.ld = a.ld;
    x.gb = a.gb + chunk0 + al;
    x.cand_ld = a.cand_ld;
    x.cand = a.grade ? a.cand_rows + (long long) (chunk0 + al) * a.cand_ld + a.cand_col0 : nullptr;
    x.flags = ((owner && al < na) ? 1 : 0) | ((P4_APL == 2 && owner && al + 1 < na) ? 2 : 0) | (a.grade ? 4 : 0);
#if P4_APL == 2
    x.e.x = x.e.y = 0.0;
#else
    x.e = 0.0;
#endif
#pragma unroll 1
    for (int st = 0; st < P4_NSTAGE; st++) {
      p4_run_stage(st, warp, x);
      __syncthreads();
    }
    for (int t = tid; t < P4_NZERO * P4_NA; t += P4_W * 32) {    // rows of gb that no basic moment owns
      const int c = t % P4_NA;
      if (c < na) a.gb[(long long) p4_zero_slot[t / P4_NA] * a.ld + chunk0 + c] = 0.0;
    }
    if (a.eflag_global || a.eflag_atom) {    // fixed-order sum over the warps, species term (pair_mtp.cpp:204-212)
      if (owner) {
#if P4_APL == 2
        epart[warp * P4_NA + al] = x.e.x;
        epart[warp * P4_NA + al + 1] = x.e.y;
#else
        epart[warp * P4_NA + al] = x.e;
#endif
      }
      __syncthreads();
      if (tid < na) {
        const int i = a.ilist ? a.ilist[a.first_ii + chunk0 + tid] : a.first_ii + chunk0 + tid;
        int itype = (int) *reinterpret_cast<const long long *>(reinterpret_cast<const char *>(a.xt) + 32 * (size_t) i + 24);
        if (itype < 0 || itype >= a.S) itype = 0;
        double es = 0.0;
        for (int w = 0; w < P4_W; w++) es += epart[w * P4_NA + tid];
        es += a.species[itype];
        if (a.eflag_atom) a.eatom[i] = es;
        if (a.eflag_global) e_thread += es;
      }
      __syncthreads();
    }
  }
  // per-CTA energy partial, fixed order
  if (tid < P4_NA) epart[tid] = e_thread;
  __syncthreads();
  if (tid < 8) {
    double s = 0.0;
    if (tid == 0)
      for (int t = 0; t < P4_NA && t < P4_W * 32; t++) s += epart[t];
    a.partials[(size_t) blockIdx.x * 8 + tid] = s;
  }
}
#endif
)P4";

// Sparse / grouped form: a CTA iteration takes P4_G groups of P4_NA atoms; the rows of group g start at
// S + g * P4_ROWS * P4_NA; at the first stage of a round the basic moments that round reads are staged for all groups.
const char *kKernelSparse = R"P4(
#ifndef P4_HOST
extern "C" __global__ void __launch_bounds__(P4_NT, P4_MINB) mtp_program_p4(const P4Args a)
{
  extern __shared__ __align__(16) double S[];
  constexpr int NAC = P4_NA * P4_G;    // atoms per CTA iteration
  double *s_lin = S + (size_t) P4_ROWS * NAC;
  double *epart = s_lin + ((P4_A + 1) & ~1);
  const int tid = threadIdx.x, lane = tid & 31;
#if P4_SPATIAL
  const int warp = (tid >> 5) % P4_W, grp = (tid >> 5) / P4_W;    // warp = which share of a stage, grp = whose atoms
#else
  const int warp = tid >> 5, grp = 0;
#endif
  for (int t = tid; t < P4_A; t += P4_NT) s_lin[t] = a.lin[t];
  const int al = (P4_APL * lane) % P4_NA;        // first atom of this lane within a group
  const bool owner = P4_APL * lane < P4_NA;      // lanes beyond the group width repeat the work of another lane, never write
  double e_thread = 0.0;
#if P4_RPAR
  const int r_only = blockIdx.y;    // rounds in parallel: this CTA evaluates one round (gb was zeroed, every share is a RED.ADD)
  const int st_begin = p4_round_stage0[r_only], st_end = r_only + 1 < P4_NROUND ? p4_round_stage0[r_only + 1] : P4_NSTAGE;
#else
  const int r_only = 0, st_begin = 0, st_end = P4_NSTAGE;
#endif
  __syncthreads();
  for (int chunk0 = blockIdx.x * NAC; chunk0 < a.inum; chunk0 += gridDim.x * NAC) {
    const int na = min(NAC, a.inum - chunk0);
    P4Ctx x;
    x.sb = (unsigned) __cvta_generic_to_shared(S + al);
    x.lb = (unsigned) __cvta_generic_to_shared(s_lin);
    x.ld = a.ld;
    x.gb = a.gb + chunk0 + al;
    x.cand_ld = a.cand_ld;
    x.cand = a.grade ? a.cand_rows + (long long) (chunk0 + al) * a.cand_ld + a.cand_col0 : nullptr;
    x.na = na;
    x.al = al;
    x.fbase = (owner ? 8 : 0) | (a.grade ? 4 : 0);
    x.grp = grp;
#pragma unroll
    for (int g = 0; g < P4_GE; g++) {
#if P4_APL == 2
      x.e[g].x = x.e[g].y = 0.0;
#else
      x.e[g] = 0.0;
#endif
    }
    int round = r_only;
#pragma unroll 1
    for (int st = st_begin; st < st_end; st++) {
      if (st == p4_round_stage0[round]) {    // basic moments this round reads -> their rows, every group (16-byte cp.async, zero fill past the end)
        const int i0 = p4_stage_off[round], nrow = p4_stage_off[round + 1] - i0;
        for (int t = tid; t < nrow * (NAC / 2); t += P4_NT) {
          const int i = i0 + t / (NAC / 2), c = (t % (NAC / 2)) * 2;
          const int nb = max(0, min(2, na - c)) * 8;
          const unsigned dst = (unsigned) __cvta_generic_to_shared(S + ((size_t) (c / P4_NA) * P4_ROWS + p4_stage_row[i]) * P4_NA + c % P4_NA);
          const double *src = a.mb + (long long) p4_stage_slot[i] * a.ld + chunk0 + (nb ? c : 0);
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(nb));
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        round++;
      }
      p4_run_stage(st, warp, x);
      __syncthreads();
    }
    for (int t = tid; t < (P4_RPAR ? 0 : P4_NZERO * NAC); t += P4_NT) {    // rows of gb that no basic moment owns
      const int c = t % NAC;
      if (c < na) a.gb[(long long) p4_zero_slot[t / NAC] * a.ld + chunk0 + c] = 0.0;
    }
    if (a.eflag_global || a.eflag_atom) {    // fixed-order sum over the warps, species term (pair_mtp.cpp:204-212)
      if (owner) {
#pragma unroll
        for (int ge = 0; ge < P4_GE; ge++) {
          const int g = P4_SPATIAL ? grp : ge;
#if P4_APL == 2
          epart[(warp * P4_G + g) * P4_NA + al] = x.e[ge].x;
          epart[(warp * P4_G + g) * P4_NA + al + 1] = x.e[ge].y;
#else
          epart[(warp * P4_G + g) * P4_NA + al] = x.e[ge];
#endif
        }
      }
      __syncthreads();
      for (int c = tid; c < na; c += P4_NT) {
        const int i = a.ilist ? a.ilist[a.first_ii + chunk0 + c] : a.first_ii + chunk0 + c;
        int itype = (int) *reinterpret_cast<const long long *>(reinterpret_cast<const char *>(a.xt) + 32 * (size_t) i + 24);
        if (itype < 0 || itype >= a.S) itype = 0;
        double es = 0.0;
        for (int w = 0; w < P4_W; w++) es += epart[(w * P4_G + c / P4_NA) * P4_NA + c % P4_NA];
#if P4_RPAR
        if (r_only == 0) es += a.species[itype];
        if (a.eflag_atom) a.esite[(long long) r_only * a.ld + chunk0 + c] = es;    // summed in round order afterwards
#else
        es += a.species[itype];
        if (a.eflag_atom) a.eatom[i] = es;
#endif
        if (a.eflag_global) e_thread += es;
      }
      __syncthreads();
    }
  }
  // per-CTA energy partial, fixed order
  epart[tid] = e_thread;
  __syncthreads();
  if (tid < 8) {
    double s = 0.0;
    if (tid == 0)
      for (int t = 0; t < P4_NT; t++) s += epart[t];
    a.partials[((size_t) blockIdx.y * gridDim.x + blockIdx.x) * 8 + tid] = s;
  }
}
#endif
)P4";

struct Plan {
  Analysis an;
  std::vector<std::vector<std::vector<Task>>> work;    // [stage][warp] -> tasks in order
};

bool make_plan(const Potential &p, const P4Params &prm, Plan &pl, std::string &why, const std::vector<char> *active = nullptr,
               bool first_round = true)
{
  if (!(prm.na == 8 || prm.na == 16 || prm.na == 32 || prm.na == 64) || prm.warps < 1 || prm.warps > 32 || prm.warps * 32 < prm.na) {
    why = "bad generator parameters";
    return false;
  }
  if (!analyse(p, pl.an, why, active, first_round, prm.sparse != 0)) return false;
  const Analysis &an = pl.an;
  pl.work.assign(an.nstages, std::vector<std::vector<Task>>(prm.warps));
  for (int st = 0; st < an.nstages; st++) {
    std::vector<Task> tasks;
    if (st == 0)
      for (int n = 0; n < an.K; n++)
        if (an.scalar[n] >= 0) tasks.push_back({2, n, 2});
    for (int n = an.K; n < an.M; n++)
      if (an.target[n] && an.fstage[n] == st && (an.operand[n] || an.scalar[n] >= 0))
        tasks.push_back({0, n, (int) an.in[n].size() + 2});
    for (int n = 0; n < an.M; n++)
      if ((n < an.K ? (an.first_round || !an.out[n].empty() || an.scalar[n] >= 0) : an.operand[n]) && an.rstage[n] == st) {
        int c = 2;
        for (const RevPair &rp : an.out[n]) c += (int) rp.terms.size() + 1;
        tasks.push_back({1, n, c});
      }
    // contiguous runs of the task order (locality: the components of one contraction stay in one warp) with the
    // smallest possible maximum load: binary search on the load, greedy feasibility check
    long long total = 0, biggest = 0;
    for (const Task &t : tasks) {
      total += t.cost;
      biggest = std::max<long long>(biggest, t.cost);
    }
    auto parts_needed = [&](long long cap) {
      int parts = 1;
      long long run = 0;
      for (const Task &t : tasks) {
        if (run + t.cost > cap && run > 0) {
          parts++;
          run = 0;
        }
        run += t.cost;
      }
      return parts;
    };
    long long lo = std::max(biggest, (total + prm.warps - 1) / prm.warps), hi = std::max(total, 1LL);
    while (lo < hi) {
      const long long mid = (lo + hi) / 2;
      if (parts_needed(mid) <= prm.warps) hi = mid;
      else
        lo = mid + 1;
    }
    int w = 0;
    long long run = 0;
    for (const Task &t : tasks) {
      if (run + t.cost > lo && run > 0 && w + 1 < prm.warps) {
        w++;
        run = 0;
      }
      pl.work[st][w].push_back(t);
      run += t.cost;
    }
  }
  return true;
}

size_t smem_of_rows(int rows, int A, const P4Params &prm)
{
  const size_t g = (size_t) std::max(1, prm.groups);
  // rows, linear coefficients, energy partials: [warp][atom] per iteration, and one entry per THREAD for the CTA's sum
  return ((size_t) rows * prm.na * g + (size_t) ((A + 1) & ~1) + (size_t) prm.warps * std::max(prm.na, 32) * g) * 8;
}

// Rounds.  When the rows of the whole program (moments of every operand node + adjoints of the non-basic ones) do not
// fit the shared memory of a CTA, the basis functions are dealt to several ROUNDS: a round evaluates the ancestors of
// its basis functions only -- forward, energy, reverse -- in rows that the next round reuses, and adds its share of the
// basic-moment adjoints to gb (the reverse pass is linear in the seeds, so the shares add up).  Intermediates that two
// rounds need are computed twice; the basis functions are taken in file order, which keeps related contractions together.
bool make_rounds(const Potential &p, const P4Params &prm, std::vector<std::vector<char>> &rounds, std::string &why)
{
  Analysis all;
  if (!analyse(p, all, why, nullptr, true, prm.sparse != 0)) return false;
  rounds.clear();
  if (prm.smem_budget == 0 || smem_of_rows(all.m_rows + all.g_rows, all.A, prm) <= prm.smem_budget) return true;    // one round
  const int K = all.K, M = all.M, T = all.T, A = all.A;
  const long long fixed = (long long) smem_of_rows(0, A, prm);
  const long long row_budget = ((long long) prm.smem_budget - fixed) / ((long long) prm.na * std::max(1, prm.groups) * 8);
  const bool sparse = prm.sparse != 0;
  const int *tm = p.alpha_index_times.data();
  // proper non-basic ancestors of every basis function (bit sets): exactly the nodes that need a moment row and an
  // adjoint row in the round that evaluates it
  const int W64 = (M + 63) / 64;
  std::vector<std::vector<int>> in_ops(M);
  for (int e = 0; e < T; e++) {
    in_ops[tm[4 * e + 3]].push_back(tm[4 * e]);
    in_ops[tm[4 * e + 3]].push_back(tm[4 * e + 1]);
  }
  std::vector<std::vector<unsigned long long>> anc(M);    // ancestors of node n, n itself excluded (memoised, file order)
  std::vector<int> order;
  {
    std::vector<char> seen(M, 0);
    for (int e = 0; e < T; e++)
      if (!seen[tm[4 * e + 3]]) {
        seen[tm[4 * e + 3]] = 1;
        order.push_back(tm[4 * e + 3]);
      }
  }
  for (int n = 0; n < M; n++) anc[n].assign(W64, 0ULL);
  for (int t : order)    // a target's factors were completed before its first product (topological file order)
    for (int o : in_ops[t]) {
      if (o >= K || sparse) anc[t][o >> 6] |= 1ULL << (o & 63);    // sparse: basic factors cost a row too
      for (int w = 0; w < W64; w++) anc[t][w] |= anc[o][w];
    }
  if (sparse)    // a basic moment that is a basis function itself is read by its round
    for (int sc = 0; sc < A; sc++) {
      const int n = p.alpha_moment_mapping[sc];
      if (n < K) anc[n][n >> 6] |= 1ULL << (n & 63);
    }
  // rows a set of nodes costs: one per basic moment (all K when not sparse), moment + adjoint per other node
  std::vector<unsigned long long> basic_mask(W64, 0ULL);
  for (int n = 0; n < K; n++) basic_mask[n >> 6] |= 1ULL << (n & 63);
  auto rows_of = [&](const std::vector<unsigned long long> &a, const std::vector<unsigned long long> *b) {
    long long nb = 0, no = 0;
    for (int w = 0; w < W64; w++) {
      const unsigned long long v = a[w] | (b ? (*b)[w] : 0ULL);
      nb += __builtin_popcountll(v & basic_mask[w]);
      no += __builtin_popcountll(v & ~basic_mask[w]);
    }
    return (sparse ? nb : (long long) K) + 2 * no;
  };
  std::vector<char> assigned(A, 0);
  int left = A;
  while (left > 0) {
    // seed: the unassigned basis function with the most ancestors; then always the one that adds the fewest rows
    std::vector<unsigned long long> cur(W64, 0ULL);
    std::vector<char> act(A, 0);
    int in_cur = 0;
    while (true) {
      int best = -1;
      long long best_add = 0, best_size = -1;
      for (int sc = 0; sc < A; sc++) {
        if (assigned[sc]) continue;
        const std::vector<unsigned long long> &as = anc[p.alpha_moment_mapping[sc]];
        long long size = 0;
        for (int w = 0; w < W64; w++) size += __builtin_popcountll(as[w]);
        const long long add = rows_of(cur, &as) - rows_of(cur, nullptr);
        const bool better = in_cur == 0 ? size > best_size : (best < 0 || add < best_add || (add == best_add && size > best_size));
        if (better) {
          best = sc;
          best_add = add;
          best_size = size;
        }
      }
      if (best < 0) break;
      if (rows_of(cur, &anc[p.alpha_moment_mapping[best]]) > row_budget) {
        if (in_cur == 0) {
          why = "one basis function alone needs more shared-memory rows than a CTA has";
          return false;
        }
        break;
      }
      const std::vector<unsigned long long> &as = anc[p.alpha_moment_mapping[best]];
      for (int w = 0; w < W64; w++) cur[w] |= as[w];
      act[best] = 1;
      assigned[best] = 1;
      in_cur++;
      left--;
    }
    rounds.push_back(act);
  }
  if ((int) rounds.size() > 64) {
    why = "the program would need more than 64 rounds";
    return false;
  }
  return true;
}

}    // namespace

namespace {
P4Params normalised(P4Params prm)
{
  prm.groups = std::max(1, prm.groups);
  if (prm.groups > 1) prm.sparse = 1;
  else
    prm.spatial = 0;
  return prm;
}
}    // namespace

size_t p4_smem_bytes(const Potential &p, const P4Params &prm_in, int *rounds_out)
{
  if (rounds_out) *rounds_out = 1;
  const P4Params prm = normalised(prm_in);
  std::vector<std::vector<char>> rounds;
  std::string why;
  if (!make_rounds(p, prm, rounds, why)) return 0;
  int rows = 0, A = 0;
  if (rounds.empty()) {
    Analysis an;
    if (!analyse(p, an, why, nullptr, true, prm.sparse != 0)) return 0;
    rows = an.m_rows + an.g_rows;
    A = an.A;
  } else
    for (size_t r = 0; r < rounds.size(); r++) {
      Analysis an;
      if (!analyse(p, an, why, &rounds[r], r == 0, prm.sparse != 0)) return 0;
      rows = std::max(rows, an.m_rows + an.g_rows);
      A = an.A;
    }
  if (rounds_out && !rounds.empty()) *rounds_out = (int) rounds.size();
  return smem_of_rows(rows, A, prm);
}

bool p4_generate(const Potential &p, const P4Params &prm_in, const short *slot_of_k, int nslots, std::string &src, P4Info &info,
                 std::string &why)
{
  const P4Params prm = normalised(prm_in);
  if (prm.groups > 8) {
    why = "bad generator parameters";
    return false;
  }
  std::vector<std::vector<char>> rounds;
  if (!make_rounds(p, prm, rounds, why)) return false;
  const int nrounds = std::max<int>(1, (int) rounds.size());
  std::vector<Plan> plans((size_t) nrounds);
  const bool rpar = prm.rpar && prm.sparse && nrounds > 1;    // rounds in parallel: no round is the first, all of them ADD
  for (int r = 0; r < nrounds; r++)
    if (!make_plan(p, prm, plans[r], why, rounds.empty() ? nullptr : &rounds[r], r == 0 && !rpar)) return false;
  const Analysis &an0 = plans[0].an;
  const int K = an0.K;
  int rows = 0, nstages = 0, m_rows = 0;
  long long terms = 0;
  for (const Plan &pl : plans) {
    rows = std::max(rows, pl.an.m_rows + pl.an.g_rows);
    m_rows = std::max(m_rows, pl.an.m_rows);
    nstages += pl.an.nstages;
    terms += pl.an.terms;
  }
  info = P4Info();
  info.rows = rows;
  info.m_rows = m_rows;
  info.g_rows = rows - m_rows;
  info.stages = nstages;
  info.rounds = nrounds;
  info.smem_bytes = smem_of_rows(rows, an0.A, prm);
  info.terms = terms;
  info.threads = prm.warps * 32 * (prm.spatial ? prm.groups : 1);
  info.groups = prm.groups;
  info.rpar_rounds = rpar ? nrounds : 0;

  unsigned long long h = 1469598103934665603ULL;
  h = fnv(h, kGeneratorVersion, strlen(kGeneratorVersion));
  const long long hdr[14] = {K, an0.M, an0.T, an0.A, prm.na, prm.warps, prm.cache, prm.acc_max, nslots, prm.fn_cost,
                             (long long) prm.smem_budget, prm.groups, prm.sparse, prm.spatial + 2 * (rpar ? 1 : 0)};
  h = fnv(h, hdr, sizeof(hdr));
  h = fnv(h, p.alpha_index_times.data(), p.alpha_index_times.size() * sizeof(int));
  h = fnv(h, p.alpha_moment_mapping.data(), p.alpha_moment_mapping.size() * sizeof(int));
  if (slot_of_k) h = fnv(h, slot_of_k, (size_t) K * sizeof(short));
  info.hash = h;

  const int apl = prm.na == 64 ? 2 : 1;
  src.clear();
  src.reserve((size_t) 64 * 1024 + (size_t) terms * 96);
  char buf[640];
  snprintf(buf, sizeof(buf),
           "// generated by mtp_codegen (%s): contraction program of one potential structure, K=%d M=%d T=%d A=%d, %d round(s)\n"
           "#define P4_NA %d\n#define P4_APL %d\n#define P4_W %d\n#define P4_ROWS %d\n#define P4_MROWS %d\n#define P4_K %d\n"
           "#define P4_A %d\n#define P4_M %d\n#define P4_NSTAGE %d\n#define P4_NSLOTS %d\n#ifndef P4_MINB\n#define P4_MINB 1\n#endif\n",
           kGeneratorVersion, K, an0.M, an0.T, an0.A, nrounds, prm.na, apl, prm.warps, rows, m_rows, K, an0.A, an0.M, nstages, nslots);
  src += buf;
  snprintf(buf, sizeof(buf), "#define P4_G %d\n#define P4_NROUND %d\n#define P4_SPATIAL %d\n#define P4_GE %d\n#define P4_NT %d\n#define P4_RPAR %d\n%s",
           prm.groups, nrounds, prm.spatial ? 1 : 0, prm.spatial ? 1 : prm.groups, info.threads, rpar ? 1 : 0,
           prm.sparse ? "#define P4_SPARSE 1\n" : "");
  src += buf;
  src += kDevicePrelude;
  // tables
  std::vector<char> slot_used((size_t) std::max(nslots, 1), 0);
  src += "P4_TABLE short p4_slot_of_k[P4_K] = {";
  for (int k = 0; k < K; k++) {
    const int s = slot_of_k ? (int) slot_of_k[k] : k;
    if (s < 0 || s >= nslots) {
      why = "slot map out of range";
      return false;
    }
    slot_used[s] = 1;
    src += std::to_string(s) + (k + 1 < K ? "," : "");
  }
  src += "};\n";
  std::vector<int> zero;
  for (int s = 0; s < nslots; s++)
    if (!slot_used[s]) zero.push_back(s);
  src += "#define P4_NZERO " + std::to_string(zero.size()) + "\nP4_TABLE short p4_zero_slot[P4_NZERO + 1] = {";
  for (int s : zero) src += std::to_string(s) + ",";
  src += "0};\n";
  if (prm.sparse) {    // per round: the basic moments it reads -> (row of mb, shared-memory row), and its first stage
    std::string off = "P4_TABLE short p4_stage_off[P4_NROUND + 1] = {0", slot = "P4_TABLE short p4_stage_slot[] = {",
                row = "P4_TABLE short p4_stage_row[] = {", st0 = "P4_TABLE short p4_round_stage0[P4_NROUND + 1] = {";
    int count = 0, stage = 0;
    for (const Plan &pl : plans) {
      for (int k = 0; k < K; k++)
        if (pl.an.mrow[k] >= 0) {
          slot += std::to_string(slot_of_k ? (int) slot_of_k[k] : k) + ",";
          row += std::to_string(pl.an.mrow[k]) + ",";
          count++;
        }
      off += "," + std::to_string(count);
      st0 += std::to_string(stage) + ",";
      stage += pl.an.nstages;
    }
    if (count > 32767) {
      why = "too many staged rows";
      return false;
    }
    src += off + "};\n" + slot + "0};\n" + row + "0};\n" + st0 + "-1};\n";
  }
  // m-row of every node in the LAST round (host harness: lets the checker read back the stored moments); -1 = not stored
  src += "#ifdef P4_HOST\nstatic const int p4_mrow[P4_M] = {";
  for (int n = 0; n < an0.M; n++) src += std::to_string(plans.back().an.mrow[n]) + (n + 1 < an0.M ? "," : "");
  src += "};\n#endif\n";

  // stage functions, round after round; the kernel walks the stages of all rounds in order
  std::vector<std::pair<int, std::string>> present;    // (global stage * W + warp, function name)
  int stage0 = 0;
  for (int r = 0; r < nrounds; r++) {
    const Plan &pl = plans[r];
    const Analysis &an = pl.an;
    const int rrows = an.m_rows + an.g_rows;
    for (int st = 0; st < an.nstages; st++) {
      long long crit = 0;
      for (int w = 0; w < prm.warps; w++) {
        const std::vector<Task> &tasks = pl.work[st][w];
        if (tasks.empty()) continue;
        long long wt = 0;
        for (const Task &t : tasks) wt += t.cost;
        crit = std::max(crit, wt);
        // a warp's share of a stage is emitted as a sequence of separately compiled (noinline) functions of bounded
        // size: ptxas allocates registers per function, and its compile time is superlinear in the function length
        size_t i = 0;
        int part = 0;
        while (i < tasks.size()) {
          size_t j = i;
          long long c = 0;
          while (j < tasks.size() && (j == i || c + tasks[j].cost <= prm.fn_cost)) c += tasks[j++].cost;
          Emitter E;
          E.capacity = std::max(4, prm.cache);
          E.uniform_base = rrows;
          std::string body;
          E.out = &body;
          for (int pass = 0; pass < 2; pass++) {
            int ntmp = 0;
            long long st_count = 0;
            if (pass == 1) E.begin_emit();
            size_t q = i;
            while (q < j) {
              const Task &t = tasks[q];
              if (t.kind == 2) {
                const std::string m = E.use(an.mrow[t.node]);
                E.line("  ESC(" + std::to_string(an.scalar[t.node]) + ", " + m + ");\n");
                q++;
              } else if (t.kind == 0) {
                emit_forward(an, t.node, E, ntmp, st_count);
                q++;
              } else {
                std::vector<int> blk;
                while (q < j && tasks[q].kind == 1 && (int) blk.size() < std::max(1, prm.acc_max)) blk.push_back(tasks[q++].node);
                emit_reverse_block(an, blk, slot_of_k, E, ntmp, st_count);
              }
            }
            if (pass == 1) info.stores += st_count;
          }
          info.loads += E.loads;
          snprintf(buf, sizeof(buf), "p4_r%d_s%d_w%d_%d", r, st, w, part);
          src += std::string("P4_STAGE_FN P4_RET ") + buf + "(P4_PARAMS)\n{\n";
          src += body;
          src += "  P4_RETURN;\n}\n";
          present.push_back({(stage0 + st) * prm.warps + w, buf});
          part++;
          i = j;
        }
      }
      info.crit_terms += crit;
      if (getenv("MTP_B200_P4_VERBOSE")) {
        long long tot = 0, big = 0;
        for (int w = 0; w < prm.warps; w++)
          for (const Task &t : pl.work[st][w]) {
            tot += t.cost;
            big = std::max<long long>(big, t.cost);
          }
        fprintf(stderr, "round %d stage %d: max warp load %lld, mean %lld, biggest task %lld\n", r, st, crit, tot / prm.warps, big);
      }
    }
    stage0 += an.nstages;
  }
  src += "P4_FN void p4_run_stage(int stage, int warp, P4Ctx &x)\n{\n  switch (stage * P4_W + warp) {\n";
  for (size_t q = 0; q < present.size();) {
    const int key = present[q].first;
    src += "    case " + std::to_string(key) + ":\n";
    for (; q < present.size() && present[q].first == key; q++)
      src += std::string(prm.sparse ? "      P4_GCALL(" : "      P4_CALL(") + present[q].second + ");\n";
    src += "      break;\n";
  }
  src += "    default: break;\n  }\n}\n";
  src += prm.sparse ? kKernelSparse : kKernel;
  return true;
}

}    // namespace mtpb200
