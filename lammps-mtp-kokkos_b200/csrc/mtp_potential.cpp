// MLIP-3 .almtp parser + contraction-program compiler (host, plain C++17).
// Grammar and messages follow the reference's loader so that a file the reference accepts loads to the
// same tables and a file it rejects is rejected with the same text:
//   PairMTP::read_file                  pair_mtp.cpp:345-570
//   RadialMTPBasis::ReadBasisProperties mtp_radial_basis.cpp:59-102
//   PairMTPExtrapolation::read_file     pair_mtp_extrapolation.cpp:545-612
#include "mtp_potential.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace mtpb200 {

namespace {

const char *const kBlank = " \t\r\n\f";    // LAMMPS TOKENIZER_DEFAULT_SEPARATORS

struct Tokens {
  std::vector<std::string> tok;
  size_t pos = 0;
  Tokens(const std::string &line, const std::string &seps)
  {
    size_t p = 0;
    while ((p = line.find_first_not_of(seps, p)) != std::string::npos) {
      size_t e = line.find_first_of(seps, p);
      if (e == std::string::npos) e = line.size();
      tok.emplace_back(line.substr(p, e - p));
      p = e;
    }
  }
  bool more() const { return pos < tok.size(); }
  std::string word()
  {
    if (!more()) throw std::out_of_range("No more tokens");
    return tok[pos++];
  }
  int integer()
  {
    std::string t = word();
    char *end = nullptr;
    long v = strtol(t.c_str(), &end, 10);
    if (end == t.c_str() || *end) throw std::runtime_error("Not a valid integer number: '" + t + "'");
    return (int) v;
  }
  double real()
  {
    std::string t = word();
    char *end = nullptr;
    double v = strtod(t.c_str(), &end);
    if (end == t.c_str() || *end) throw std::runtime_error("Not a valid floating-point number: '" + t + "'");
    return v;
  }
};

// Line cursor over the whole file image.  Like LAMMPS's TextFileReader it returns the next physical line
// that still holds at least one word after (optionally) cutting a '#' comment; blank lines are skipped.
struct LineCursor {
  const std::string &buf;
  size_t pos = 0;
  explicit LineCursor(const std::string &b) : buf(b) {}
  bool raw_line(std::string &out)
  {
    if (pos >= buf.size()) return false;
    size_t e = buf.find('\n', pos);
    e = (e == std::string::npos) ? buf.size() : e + 1;
    out.assign(buf, pos, e - pos);
    pos = e;
    return true;
  }
  bool next(std::string &out, bool strip_comments = true)
  {
    std::string ln;
    while (raw_line(ln)) {
      if (strip_comments) {
        size_t h = ln.find('#');
        if (h != std::string::npos) ln.erase(h);
      }
      if (ln.find_first_not_of(kBlank) != std::string::npos) {
        out.swap(ln);
        return true;
      }
    }
    return false;
  }
  std::string must(const char *what_if_eof)
  {
    std::string ln;
    if (!next(ln)) throw std::runtime_error(what_if_eof);
    return ln;
  }
};

std::string read_file_image(const std::string &path)
{
  FILE *fp = fopen(path.c_str(), "rb");
  if (!fp) {
    const char *dir = getenv("LAMMPS_POTENTIALS");    // utils::open_potential also searches this folder
    if (dir) fp = fopen((std::string(dir) + "/" + path).c_str(), "rb");
  }
  if (!fp) throw std::runtime_error("Cannot open potential file " + path);
  std::string img;
  char chunk[1 << 16];
  size_t n;
  while ((n = fread(chunk, 1, sizeof(chunk), fp)) > 0) img.append(chunk, n);
  fclose(fp);
  return img;
}

}    // namespace

void parse_almtp(const std::string &path, bool want_selection_state, Potential &p)
{
  const std::string img = read_file_image(path);
  LineCursor cur(img);
  const std::string seps = std::string(kBlank) + "=, ";
  const std::string brace = seps + "{},";
  const char *eof = "Error reading MTP file. Unexpected end of file.";

  {
    Tokens t(cur.must("Only MTP potential files are accepted."), seps);
    if (t.word() != "MTP") throw std::runtime_error("Only MTP potential files are accepted.");
  }
  if (cur.must("MTP file must have version \"1.1.0\"") != "version = 1.1.0\n")
    throw std::runtime_error("MTP file must have version \"1.1.0\"");

  std::string line = cur.must(eof);
  Tokens t(line, seps);
  std::string key = t.word();
  auto advance = [&]() {
    line = cur.must(eof);
    t = Tokens(line, seps);
    key = t.word();
  };

  if (key == "potential_name") {
    p.potential_name = t.more() ? t.word() : "";
    advance();
  }
  if (key == "scaling") {
    p.scaling = t.real();
    advance();
  } else
    p.scaling = 1;
  char fmtbuf[64];
  snprintf(fmtbuf, sizeof(fmtbuf), "The scaling is : %.2e.\n", p.scaling);
  p.log += fmtbuf;

  if (key != "species_count") throw std::runtime_error("Error reading MTP file. Species count not found.");
  p.species_count = t.integer();
  p.log += "There are " + std::to_string(p.species_count) + " species.\n";
  const int S = p.species_count;
  if (S < 1) throw std::runtime_error("Error reading MTP file. Species count not found.");
  p.setflag.assign((size_t) (S + 1) * (S + 1), 0);

  advance();
  if (key == "potential_tag") {
    p.potential_tag = t.more() ? t.word() : "";
    advance();
  }

  if (key != "radial_basis_type")
    throw std::runtime_error("Error reading MTP file. No radial basis set type is specified.");
  const std::string rb_type = t.word();
  if (rb_type != "RBChebyshev")
    throw std::runtime_error("Error reading MTP file. The specified radial basis set type, " + rb_type +
                             ", was not found..");
  // ---- basis properties block (mtp_radial_basis.cpp:59-102).  A "scaling" line placed here (MLIP-2
  // style) is read and then overridden by the top-level value, as in the reference (pair_mtp.cpp:416).
  advance();
  if (key == "scaling") {
    (void) t.real();
    advance();
  }
  if (key != "min_val" && key != "min_dist")
    throw std::runtime_error("Error in reading MTP file. Cannot read lower cutoff.");
  p.min_cutoff = t.real();
  advance();
  if (key != "max_val" && key != "max_dist")
    throw std::runtime_error("Error in reading MTP file. Cannot read upper cutoff.");
  p.max_cutoff = t.real();
  advance();
  if (key != "radial_basis_size")
    throw std::runtime_error("Error in reading MTP file. Cannot read radial basis set size.");
  p.radial_basis_size = t.integer();
  advance();
  if (key != "radial_funcs_count")
    throw std::runtime_error("Error in reading MTP file. Cannot read radial function count.");
  p.radial_func_count = t.integer();
  advance();
  if (key != "radial_coeffs") {
    if (key == "magnetic_basis_type") throw std::runtime_error("Magnetic basis is currently not supported.");
    throw std::runtime_error("Error in reading MTP file. Cannot read radial coeffs count.");
  }
  const int R = p.radial_func_count, B = p.radial_basis_size;
  if (R < 1 || B < 1) throw std::runtime_error("Error in reading MTP file. Cannot read radial function count.");
  p.radial_basis_coeffs.assign((size_t) S * S * R * B, 0.0);
  for (int n = 0; n < S * S; n++) {
    Tokens hd(cur.must(eof), seps + "-");
    const int t1 = hd.integer(), t2 = hd.integer();
    if (t1 < 0 || t1 >= S || t2 < 0 || t2 >= S)
      throw std::runtime_error("Error reading MTP file. Species pair out of range in radial_coeffs.");
    p.setflag[(size_t) (t1 + 1) * (S + 1) + (t2 + 1)] = 1;
    const size_t off = (size_t) (t1 * S + t2) * R * B;
    for (int mu = 0; mu < R; mu++) {
      Tokens row(cur.must(eof), brace);
      for (int k = 0; k < B; k++) p.radial_basis_coeffs[off + (size_t) mu * B + k] = row.real();
    }
  }

  advance();
  if (key != "alpha_moments_count") throw std::runtime_error("Error reading MTP file. Alpha moment count not found.");
  p.alpha_moment_count = t.integer();
  advance();
  if (key != "alpha_index_basic_count")
    throw std::runtime_error("Error reading MTP file. Alpha moment count not found.");
  p.alpha_index_basic_count = t.integer();
  {
    Tokens b(cur.must(eof), brace);
    if (b.word() != "alpha_index_basic") throw std::runtime_error("Error reading MTP file. Alpha index basic not found.");
    p.alpha_index_basic.resize((size_t) p.alpha_index_basic_count * 4);
    for (auto &v : p.alpha_index_basic) v = b.integer();
  }
  advance();
  if (key != "alpha_index_times_count")
    throw std::runtime_error("Error reading MTP file. Alpha index times count not found.");
  p.alpha_index_times_count = t.integer();
  {
    Tokens b(cur.must(eof), brace);
    if (b.word() != "alpha_index_times") throw std::runtime_error("Error reading MTP file. Alpha index times not found.");
    p.alpha_index_times.resize((size_t) p.alpha_index_times_count * 4);
    for (auto &v : p.alpha_index_times) v = b.integer();
  }
  advance();
  if (key != "alpha_scalar_moments")
    throw std::runtime_error("Error reading MTP file. Alpha scalar moment count not found.");
  p.alpha_scalar_count = t.integer();
  {
    Tokens b(cur.must(eof), brace);
    if (b.word() != "alpha_moment_mapping")
      throw std::runtime_error("Error reading MTP file. Alpha moment mappings not found.");
    p.alpha_moment_mapping.resize((size_t) p.alpha_scalar_count);
    for (auto &v : p.alpha_moment_mapping) v = b.integer();
  }
  {
    Tokens b(cur.must("Error reading MTP file. Species coefficients not found."), brace);
    if (b.word() != "species_coeffs") throw std::runtime_error("Error reading MTP file. Species coefficients not found.");
    p.species_coeffs.resize((size_t) S);
    for (auto &v : p.species_coeffs) v = b.real();
  }
  {
    Tokens b(cur.must("Error reading MTP file. Moment coefficients not found."), brace);
    if (b.word() != "moment_coeffs") throw std::runtime_error("Error reading MTP file. Moment coefficients not found.");
    p.linear_coeffs.resize((size_t) p.alpha_scalar_count);
    for (auto &v : p.linear_coeffs) v = b.real();
  }
  finalize_tables(p);

  if (!want_selection_state) return;

  // ---- MaxVol selection state (pair_mtp_extrapolation.cpp:545-612) ----
  std::string ln;
  if (!cur.next(ln, /*strip_comments=*/false))
    throw std::runtime_error(
        "No selection state found! Consider training/retraining or disabling extrapolation!\n");
  {
    Tokens v(ln, seps);
    if (v.word() != "#MVS_v1.1")
      throw std::runtime_error(
          "Error in reading MTP file selection state. Please verify MVS version is #MVS_v1.1!");
  }
  int energy_weight = 0, site_en_weight = 0;    // the reference truncates the weights to int (:570,576,592)
  const char *names[5] = {"energy_weight", "force_weight", "stress_weight", "site_en_weight", "weight_scaling"};
  for (int n = 0; n < 5; n++) {
    std::string l2;
    if (!cur.next(l2)) throw std::runtime_error(std::string("Error in reading MTP file, ") + names[n]);
    Tokens w(l2, seps);
    if (w.word() != names[n]) throw std::runtime_error(std::string("Error in reading MTP file, ") + names[n]);
    if (n == 0) energy_weight = (int) w.real();
    if (n == 3) site_en_weight = (int) w.real();
  }
  if (energy_weight + site_en_weight > 1)
    throw std::runtime_error(
        "Error, the MTP currently only supports configuration mode (energy_weight=1) or neighbourhood mode "
        "(site_en_weight=1). Please retrain the MTP with the correct modes!");
  p.configuration_mode = (energy_weight == 1);
  const size_t Q = (size_t) p.coeff_count, nd = Q * Q;
  size_t at = cur.pos + 1;    // one '#' byte precedes the binary block (:607)
  if (at + 2 * nd * sizeof(double) > img.size())
    throw std::runtime_error("Unexpected end of file or read error while reading binary data");
  p.active_set.resize(nd);
  p.inverse_active_set.resize(nd);
  memcpy(p.active_set.data(), img.data() + at, nd * sizeof(double));
  memcpy(p.inverse_active_set.data(), img.data() + at + nd * sizeof(double), nd * sizeof(double));
  p.has_selection_state = true;
}

void finalize_tables(Potential &p)
{
  const int S = p.species_count, R = p.radial_func_count, B = p.radial_basis_size;
  const int K = p.alpha_index_basic_count, T = p.alpha_index_times_count, A = p.alpha_scalar_count;
  const int M = p.alpha_moment_count;
  if (S < 1 || R < 1 || B < 1 || K < 1 || T < 0 || A < 1 || M < K)
    throw std::runtime_error("Error reading MTP file. Inconsistent table sizes.");
  if (M > 65535) throw std::runtime_error("alpha_moments_count above 65535 is not supported.");
  if ((int) p.radial_basis_coeffs.size() != S * S * R * B || (int) p.alpha_index_basic.size() != 4 * K ||
      (int) p.alpha_index_times.size() != 4 * T || (int) p.alpha_moment_mapping.size() != A ||
      (int) p.species_coeffs.size() != S || (int) p.linear_coeffs.size() != A)
    throw std::runtime_error("Error reading MTP file. Inconsistent table sizes.");
  int radial_func_max = 0, pmax = 0;
  for (int k = 0; k < K; k++) {
    const int *e = &p.alpha_index_basic[4 * (size_t) k];
    if (e[0] < 0 || e[1] < 0 || e[2] < 0 || e[3] < 0)
      throw std::runtime_error("Error reading MTP file. Negative alpha_index_basic entry.");
    radial_func_max = std::max(radial_func_max, e[0]);
    pmax = std::max(pmax, e[1] + e[2] + e[3]);
  }
  if (radial_func_max != R - 1) throw std::runtime_error("Wrong number of radial functions specified!");
  if (pmax > 63) throw std::runtime_error("alpha_index_basic rank above 63 is not supported.");
  p.max_alpha_index_basic = pmax + 1;
  for (int e = 0; e < T; e++) {
    const int *q = &p.alpha_index_times[4 * (size_t) e];
    if (q[0] < 0 || q[0] >= M || q[1] < 0 || q[1] >= M || q[3] < 0 || q[3] >= M)
      throw std::runtime_error("Error reading MTP file. alpha_index_times entry out of range.");
    if (std::abs((long) q[2]) >= (1L << 24))
      throw std::runtime_error("Error reading MTP file. alpha_index_times multiplicity too large.");
  }
  for (int s = 0; s < A; s++)
    if (p.alpha_moment_mapping[s] < 0 || p.alpha_moment_mapping[s] >= M)
      throw std::runtime_error("Error reading MTP file. alpha_moment_mapping entry out of range.");
  if (p.setflag.empty()) {
    p.setflag.assign((size_t) (S + 1) * (S + 1), 0);
    for (int i = 1; i <= S; i++)
      for (int j = 1; j <= S; j++) p.setflag[(size_t) i * (S + 1) + j] = 1;
  }
  p.coeff_count = S * S * R * B + S + A;
}

// --------------------------------------------------------------------------------------------------

namespace {

struct NodeList {
  int node;
  std::vector<ProgramTerm> terms;
};

void pack_pass(const std::vector<std::vector<NodeList>> &per_level, ProgramPass &out)
{
  out = ProgramPass();
  out.level_group_begin.push_back(0);
  int slot_rows = 0;
  for (const auto &lists_in : per_level) {
    std::vector<const NodeList *> lists;
    for (const auto &l : lists_in) lists.push_back(&l);
    // longest lists first so that the 32 lanes of a group have similar trip counts
    std::stable_sort(lists.begin(), lists.end(),
                     [](const NodeList *a, const NodeList *b) { return a->terms.size() > b->terms.size(); });
    for (size_t g0 = 0; g0 < lists.size(); g0 += 32) {
      const size_t g1 = std::min(lists.size(), g0 + 32);
      int mx = 0;
      for (size_t i = g0; i < g1; i++) mx = std::max(mx, (int) lists[i]->terms.size());
      out.group_term_base.push_back(slot_rows);
      out.group_max_terms.push_back(mx);
      out.terms.resize((size_t) (slot_rows + mx) * 32, ProgramTerm{0, 0, 0.0f});
      for (int lane = 0; lane < 32; lane++) {
        const size_t i = g0 + lane;
        if (i < g1) {
          out.node.push_back(lists[i]->node);
          out.nterms.push_back((int) lists[i]->terms.size());
          for (size_t t = 0; t < lists[i]->terms.size(); t++)
            out.terms[(size_t) (slot_rows + t) * 32 + lane] = lists[i]->terms[t];
        } else {
          out.node.push_back(-1);
          out.nterms.push_back(0);
        }
      }
      slot_rows += mx;
    }
    out.level_group_begin.push_back(out.ngroups());
  }
}

}    // namespace

int count_adjoint_rows(const Potential &p)
{
  const int M = p.alpha_moment_count, K = p.alpha_index_basic_count, T = p.alpha_index_times_count;
  std::vector<char> src(std::max(M, 1), 0);
  for (int e = 0; e < T; e++) src[p.alpha_index_times[4 * (size_t) e]] = src[p.alpha_index_times[4 * (size_t) e + 1]] = 1;
  int rows = K;
  for (int n = K; n < M; n++) rows += src[n];
  return std::max(rows, 1);
}

void compile_program(const Potential &p, Program &prog, int na_large, int na_small, int na_v3)
{
  const int M = p.alpha_moment_count, T = p.alpha_index_times_count, A = p.alpha_scalar_count;
  const int *times = p.alpha_index_times.data();
  prog = Program();

  // sequential-consistency check: every write to a node precedes every read of it
  std::vector<int> last_write(M, -1), first_read(M, T);
  for (int e = 0; e < T; e++) {
    const int a0 = times[4 * e], a1 = times[4 * e + 1], a3 = times[4 * e + 3];
    first_read[a0] = std::min(first_read[a0], e);
    first_read[a1] = std::min(first_read[a1], e);
    last_write[a3] = e;
  }
  for (int n = 0; n < M; n++)
    if (last_write[n] >= first_read[n])
      throw std::runtime_error(
          "Error in the alpha times indicies! alpha_index_times is not a topologically ordered program.");

  // dependency level of every node (0 = never a target)
  prog.level.assign(M, 0);
  for (int e = 0; e < T; e++) {
    const int a0 = times[4 * e], a1 = times[4 * e + 1], a3 = times[4 * e + 3];
    prog.level[a3] = std::max(prog.level[a3], 1 + std::max(prog.level[a0], prog.level[a1]));
  }
  // (levels of sources are final when an edge is visited because all their writes precede this read)
  prog.depth = 0;
  for (int n = 0; n < M; n++) prog.depth = std::max(prog.depth, prog.level[n]);

  // forward: per target, terms in file order
  std::vector<std::vector<NodeList>> fwd(prog.depth);
  {
    std::vector<int> slot(M, -1);
    for (int e = 0; e < T; e++) {
      const int a0 = times[4 * e], a1 = times[4 * e + 1], mult = times[4 * e + 2], a3 = times[4 * e + 3];
      auto &lv = fwd[prog.level[a3] - 1];
      if (slot[a3] < 0) {
        slot[a3] = (int) lv.size();
        lv.push_back(NodeList{a3, {}});
      }
      lv[slot[a3]].terms.push_back(ProgramTerm{(uint16_t) a0, (uint16_t) a1, (float) mult});
    }
  }
  pack_pass(fwd, prog.fwd);

  // reverse: per source, terms in reverse file order; levels visited from depth-1 down to 0
  std::vector<std::vector<NodeList>> rev(std::max(prog.depth, 1));
  {
    std::vector<int> slot(M, -1);
    auto add = [&](int src, int a3, int other, int mult) {
      auto &lv = rev[prog.depth - 1 - prog.level[src]];
      if (slot[src] < 0) {
        slot[src] = (int) lv.size();
        lv.push_back(NodeList{src, {}});
      }
      lv[slot[src]].terms.push_back(ProgramTerm{(uint16_t) a3, (uint16_t) other, (float) mult});
    };
    for (int e = T - 1; e >= 0; e--) {
      const int a0 = times[4 * e], a1 = times[4 * e + 1], mult = times[4 * e + 2], a3 = times[4 * e + 3];
      add(a1, a3, a0, mult);    // g[a1] += g[a3]*mult*m[a0]   (pair_mtp.cpp:231)
      add(a0, a3, a1, mult);    // g[a0] += g[a3]*mult*m[a1]   (pair_mtp.cpp:232)
    }
    // basic moments that feed no product still need their adjoint (= the seed): give them an empty list
    for (int k = 0; k < p.alpha_index_basic_count; k++)
      if (slot[k] < 0) rev.back().push_back(NodeList{k, {}});
  }
  pack_pass(rev, prog.rev);

  prog.ginit.assign(M, 0.0);
  for (int s = 0; s < A; s++) prog.ginit[p.alpha_moment_mapping[s]] = p.linear_coeffs[s];

  // ---- chunk form ----
  std::vector<char> is_source(M, 0);
  for (int e = 0; e < T; e++) is_source[times[4 * e]] = is_source[times[4 * e + 1]] = 1;
  prog.grow.assign(M, -1);
  prog.adjoint_rows = p.alpha_index_basic_count;
  for (int n = 0; n < p.alpha_index_basic_count; n++) prog.grow[n] = n;
  for (int n = p.alpha_index_basic_count; n < M; n++)
    if (is_source[n]) prog.grow[n] = prog.adjoint_rows++;
  prog.adjoint_rows = std::max(prog.adjoint_rows, 1);
  auto pack_chunk = [&](const std::vector<std::vector<NodeList>> &levels, bool reverse, ChunkPass &out) {
    out = ChunkPass();
    out.level_begin.push_back(0);
    out.term_begin.push_back(0);
    for (const auto &lv : levels) {
      std::vector<const NodeList *> lists;
      for (const auto &l : lv) lists.push_back(&l);
      std::stable_sort(lists.begin(), lists.end(),
                       [](const NodeList *a, const NodeList *b) { return a->terms.size() > b->terms.size(); });
      for (const NodeList *l : lists) {
        out.node.push_back(l->node);
        out.init.push_back(reverse ? prog.ginit[l->node] : 0.0);
        for (const ProgramTerm &t : l->terms) {
          if (reverse && !is_source[t.a]) {
            // g[a3] is never updated: it stays ginit[a3], fold it into the coefficient (drop exact zeros)
            const double c = (double) t.mult * prog.ginit[t.a];
            if (c == 0.0) continue;
            out.term_idx.push_back(0xFFFFu | ((uint32_t) t.b << 16));
            out.term_coef.push_back(c);
          } else {
            out.term_idx.push_back((uint32_t) t.a | ((uint32_t) t.b << 16));
            out.term_coef.push_back((double) t.mult);
          }
        }
        out.term_begin.push_back((int) out.term_idx.size());
      }
      out.level_begin.push_back((int) out.node.size());
    }
  };
  if (M >= 0xFFFF) throw std::runtime_error("alpha_moments_count above 65534 is not supported.");
  pack_chunk(fwd, false, prog.cfwd);
  pack_chunk(rev, true, prog.crev);

  // ---- flat predicated streams ----
  struct Raw {
    int a, b, node;
    bool store;
    double coef;
  };
  const int ONE = M;    // row of 1.0
  auto pack_flat = [&](const std::vector<std::vector<NodeList>> &levels, bool reverse, int na, FlatPass &out) {
    out = FlatPass();
    const int vw = 16 * 32 / na;
    const uint32_t row_bytes = (uint32_t) na * 8;    // rows are [node][atom], no padding
    out.vw = vw;
    out.na = na;
    out.nlevels = (int) levels.size();
    out.stream_begin.push_back(0);
    for (const auto &lv : levels) {
      // node -> uniform terms (base first, then the list in its original order)
      std::vector<std::vector<Raw>> per_node;
      for (const NodeList &l : lv) {
        std::vector<Raw> t;
        if (!reverse) {
          if (l.node < p.alpha_index_basic_count) t.push_back(Raw{l.node, ONE, 0, false, 1.0});
          for (const ProgramTerm &q : l.terms) t.push_back(Raw{q.a, q.b, 0, false, (double) q.mult});
        } else {
          t.push_back(Raw{ONE, ONE, 0, false, prog.ginit[l.node]});
          for (size_t qi = 0; qi < l.terms.size(); qi++) {
            const ProgramTerm &q = l.terms[qi];
            double mult = (double) q.mult;
            // a self product m[a3] += mult*m[s]*m[s] lists the same reverse term twice in a row: merge (x + x == 2x)
            if (qi + 1 < l.terms.size() && l.terms[qi + 1].a == q.a && l.terms[qi + 1].b == q.b &&
                l.terms[qi + 1].mult == q.mult && q.b == (uint16_t) l.node) {
              mult *= 2.0;
              qi++;
            }
            if (!is_source[q.a]) {
              const double c = mult * prog.ginit[q.a];    // g[a3] stays ginit[a3]
              if (c == 0.0) continue;
              t.push_back(Raw{ONE, q.b, 0, false, c});
            } else
              t.push_back(Raw{q.a, q.b, 0, false, mult});
          }
        }
        if (t.empty()) t.push_back(Raw{ONE, ONE, 0, false, 0.0});
        for (auto &x : t) x.node = l.node;
        t.back().store = true;
        per_node.push_back(std::move(t));
      }
      std::vector<int> order(per_node.size());
      for (size_t i = 0; i < order.size(); i++) order[i] = (int) i;
      std::stable_sort(order.begin(), order.end(),
                       [&](int x, int y) { return per_node[x].size() > per_node[y].size(); });
      std::vector<std::vector<Raw>> bins(vw);
      for (int idx : order) {
        int best = 0;
        for (int b = 1; b < vw; b++)
          if (bins[b].size() < bins[best].size()) best = b;
        bins[best].insert(bins[best].end(), per_node[idx].begin(), per_node[idx].end());
      }
      for (int b = 0; b < vw; b++) {
        while (bins[b].size() % FLAT_UNROLL) bins[b].push_back(Raw{ONE, ONE, ONE, false, 0.0});
        for (const Raw &r : bins[b]) {
          // operand flag (offsets are multiples of 8): bit 0 = the operand is the constant 1.0 (broadcast read)
          // reverse pass: operand a and the destination are rows of the (compact) adjoint table
          auto arow = [&](int node) { return reverse ? prog.grow[node] : node; };
          if (reverse && ((r.a != ONE && prog.grow[r.a] < 0) || (r.node != ONE && prog.grow[r.node] < 0)))
            throw std::runtime_error("internal: adjoint row missing for a node of the reverse pass");
          const uint32_t ao = r.a == ONE ? 1u : (uint32_t) arow(r.a) * row_bytes;
          const uint32_t bo = r.b == ONE ? 1u : (uint32_t) r.b * row_bytes;
          out.terms.push_back(FlatTerm{ao, bo, r.coef});
          // (padding no-ops never store; their destination field is unused)
          out.st.push_back((r.node == ONE ? 0u : (uint32_t) arow(r.node) * row_bytes) | (r.store ? 1u : 0u));
        }
        out.stream_begin.push_back((int) out.terms.size());
      }
    }
  };
  for (int v = 0; v < 2; v++) {
    pack_flat(fwd, false, v == 0 ? na_large : na_small, prog.ffwd[v]);
    pack_flat(rev, true, v == 0 ? na_large : na_small, prog.frev[v]);
  }

  // ---- grouped streams of the 4-atoms-per-lane kernel ----
  struct Raw3 {
    int a, b;
    double coef;
  };
  auto pack_flat3 = [&](const std::vector<std::vector<NodeList>> &levels, bool reverse, int na, Flat3Pass &out) -> bool {
    out = Flat3Pass();
    const int vpw = 128 / na;
    const uint32_t row_bytes = (uint32_t) na * 8;
    const int SCRATCH = M + 1;
    out.vpw = vpw;
    out.na = na;
    out.nlevels = (int) levels.size();
    out.row_begin.push_back(0);
    out.group_begin.push_back(0);
    const Raw3 noop{ONE, ONE, 0.0};
    auto emit_term = [&](const Raw3 &r) { out.terms.push_back(G3Term{(uint32_t) r.a * row_bytes, (uint32_t) r.b * row_bytes, r.coef}); };
    struct Group {
      std::vector<int> members;    // indices into per_node, -1 = dummy
      int rows;
      bool split;                  // one node, its terms dealt to all virtual warps
    };
    for (const auto &lv : levels) {
      std::vector<std::vector<Raw3>> per_node;
      for (const NodeList &l : lv) {
        std::vector<Raw3> t;
        if (!reverse) {
          if (l.node < p.alpha_index_basic_count) t.push_back(Raw3{l.node, ONE, 1.0});
          for (const ProgramTerm &q : l.terms) t.push_back(Raw3{q.a, q.b, (double) q.mult});
        } else {
          for (size_t qi = 0; qi < l.terms.size(); qi++) {
            const ProgramTerm &q = l.terms[qi];
            double mult = (double) q.mult;
            if (qi + 1 < l.terms.size() && l.terms[qi + 1].a == q.a && l.terms[qi + 1].b == q.b &&
                l.terms[qi + 1].mult == q.mult && q.b == (uint16_t) l.node) {
              mult *= 2.0;    // self product: the same reverse term twice in a row (x + x == 2x)
              qi++;
            }
            if (!is_source[q.a]) {
              const double c = mult * prog.ginit[q.a];    // g[a3] stays ginit[a3]
              if (c == 0.0) continue;
              t.push_back(Raw3{ONE, q.b, c});
            } else
              t.push_back(Raw3{q.a, q.b, mult});
          }
        }
        per_node.push_back(std::move(t));
      }
      // long lists are split over the virtual warps of a group (partial sums combined by shuffles), the others are
      // grouped vpw at a time in order of length so that the common row count wastes little
      std::vector<int> order;
      std::vector<Group> groups;
      for (size_t i = 0; i < per_node.size(); i++) {
        if ((int) per_node[i].size() > G3_SPLIT_ABOVE) {
          Group g;
          g.members.assign(vpw, (int) i);
          g.rows = ((int) per_node[i].size() + vpw - 1) / vpw;
          g.split = true;
          groups.push_back(std::move(g));
        } else
          order.push_back((int) i);
      }
      std::stable_sort(order.begin(), order.end(),
                       [&](int x, int y) { return per_node[x].size() > per_node[y].size(); });
      for (size_t i = 0; i < order.size(); i += vpw) {
        Group g;
        size_t mx = 0;
        for (int v = 0; v < vpw; v++) {
          const int idx = i + v < order.size() ? order[i + v] : -1;
          g.members.push_back(idx);
          if (idx >= 0) mx = std::max(mx, per_node[idx].size());
        }
        g.rows = std::max<int>(1, (int) mx);
        g.split = false;
        groups.push_back(std::move(g));
      }
      // longest group first onto the least loaded warp (cost: rows + the end-of-group work)
      std::vector<int> gorder(groups.size());
      for (size_t i = 0; i < gorder.size(); i++) gorder[i] = (int) i;
      std::stable_sort(gorder.begin(), gorder.end(), [&](int x, int y) { return groups[x].rows > groups[y].rows; });
      std::vector<std::vector<int>> bins(G3_WARPS);
      std::vector<int> load(G3_WARPS, 0);
      for (int gi : gorder) {
        int best = 0;
        for (int b = 1; b < G3_WARPS; b++)
          if (load[b] < load[best]) best = b;
        bins[best].push_back(gi);
        load[best] += groups[gi].rows + 1;
      }
      for (int b = 0; b < G3_WARPS; b++) {
        int rows = 0;
        auto emit_group = [&](const Group &g) {
          for (int v = 0; v < vpw; v++) {
            const int idx = g.members[v];
            const int node = idx >= 0 ? lv[idx].node : SCRATCH;
            const bool seeded = reverse && idx >= 0 && (!g.split || v == 0);
            out.heads.push_back(G3Head{(uint32_t) node * row_bytes, (uint32_t) g.rows | (g.split ? 0x80000000u : 0u),
                                       seeded ? prog.ginit[node] : 0.0});
          }
          for (int r = 0; r < g.rows; r++)
            for (int v = 0; v < vpw; v++) {
              const int idx = g.members[v];
              const int q = g.split ? r * vpw + v : r;    // split: terms dealt round-robin to the virtual warps
              emit_term(idx >= 0 && q < (int) per_node[idx].size() ? per_node[idx][q] : noop);
            }
          rows += g.rows;
        };
        for (int gi : bins[b]) emit_group(groups[gi]);
        if (rows % 4) {    // whole 4-row trips
          Group g;
          g.members.assign(vpw, -1);
          g.rows = 4 - rows % 4;
          g.split = false;
          emit_group(g);
        }
        out.row_begin.push_back((int) out.terms.size() / vpw);
        out.group_begin.push_back((int) out.heads.size() / vpw);
      }
    }
    for (int i = 0; i < G3_PAD_ROWS * vpw; i++) emit_term(noop);
    for (int i = 0; i < 2 * vpw; i++) out.heads.push_back(G3Head{(uint32_t) SCRATCH * row_bytes, 1u << 30, 0.0});
    return true;
  };
  prog.f3_na = 0;
  if (na_v3 == 32 || na_v3 == 16) {
    if (pack_flat3(fwd, false, na_v3, prog.f3fwd) && pack_flat3(rev, true, na_v3, prog.f3rev)) prog.f3_na = na_v3;
  }
}

double check_grouped_streams(const Potential &p, const Program &prog)
{
  if (!prog.f3_na) throw std::runtime_error("the grouped streams were not built for this potential");
  const int M = p.alpha_moment_count, K = p.alpha_index_basic_count, T = p.alpha_index_times_count;
  const int *times = p.alpha_index_times.data();
  // sequential reference
  std::vector<double> m(M, 0.0), g(M, 0.0);
  uint64_t lcg = 0x9E3779B97F4A7C15ull;
  for (int k = 0; k < K; k++) {
    lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
    m[k] = ((double) (lcg >> 11) / 9007199254740992.0 - 0.5) * 1.5;
  }
  for (int e = 0; e < T; e++) m[times[4 * e + 3]] += times[4 * e + 2] * m[times[4 * e]] * m[times[4 * e + 1]];
  for (int n = 0; n < M; n++) g[n] = prog.ginit[n];
  for (int e = T - 1; e >= 0; e--) {
    const int a0 = times[4 * e], a1 = times[4 * e + 1], a3 = times[4 * e + 3];
    const double mult = times[4 * e + 2];
    g[a1] += g[a3] * mult * m[a0];
    g[a0] += g[a3] * mult * m[a1];
  }
  // the streams
  const int na = prog.f3_na;
  const uint32_t row_bytes = (uint32_t) na * 8;
  std::vector<double> cm(M + 2, 0.0), cg(M + 2, 0.0);
  for (int k = 0; k < K; k++) cm[k] = m[k];
  cm[M] = cg[M] = 1.0;
  auto run = [&](const Flat3Pass &f, std::vector<double> &A, const std::vector<double> &B) {
    const int vpw = f.vpw;
    for (int lv = 0; lv < f.nlevels; lv++) {
      std::vector<std::pair<int, double>> stores;
      for (int w = 0; w < G3_WARPS; w++) {
        const int sidx = lv * G3_WARPS + w;
        int row = f.row_begin[sidx];
        if ((f.row_begin[sidx + 1] - row) % 4) throw std::runtime_error("stream is not a whole number of 4-row trips");
        for (int gi = f.group_begin[sidx]; gi < f.group_begin[sidx + 1]; gi++) {
          const G3Head *hd = &f.heads[(size_t) gi * vpw];
          const int rows = (int) (hd[0].rows & 0x7FFFFFFFu);
          const bool split = (hd[0].rows >> 31) != 0;
          std::vector<double> acc(vpw);
          for (int v = 0; v < vpw; v++) {
            if (hd[v].rows != hd[0].rows) throw std::runtime_error("group heads disagree");
            acc[v] = hd[v].init;
            for (int r = 0; r < rows; r++) {
              const G3Term &t = f.terms[(size_t) (row + r) * vpw + v];
              if (t.a_off % row_bytes || t.b_off % row_bytes) throw std::runtime_error("operand offset is not a row");
              acc[v] += t.coef * A[t.a_off / row_bytes] * B[t.b_off / row_bytes];
            }
          }
          if (split) {
            double sum = 0.0;
            for (int v = 0; v < vpw; v++) sum += acc[v];
            stores.push_back({(int) (hd[0].dst_off / row_bytes), sum});
          } else
            for (int v = 0; v < vpw; v++) stores.push_back({(int) (hd[v].dst_off / row_bytes), acc[v]});
          row += rows;
        }
        if (row != f.row_begin[sidx + 1]) throw std::runtime_error("group rows do not add up to the stream length");
      }
      for (auto &st : stores)
        if (st.first != M + 1) A[st.first] = st.second;    // row M + 1 is scratch
    }
  };
  run(prog.f3fwd, cm, cm);
  run(prog.f3rev, cg, cm);
  double err = 0.0, scale_m = 1e-300, scale_g = 1e-300;
  for (int n = 0; n < M; n++) scale_m = std::max(scale_m, std::fabs(m[n]));
  for (int k = 0; k < K; k++) scale_g = std::max(scale_g, std::fabs(g[k]));
  for (int n = 0; n < M; n++) err = std::max(err, std::fabs(cm[n] - m[n]) / scale_m);
  for (int k = 0; k < K; k++) err = std::max(err, std::fabs(cg[k] - g[k]) / scale_g);
  return err;
}

}    // namespace mtpb200
