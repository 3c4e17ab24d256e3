// TEST INFRASTRUCTURE -- not part of the shipped product.
//
// Minimal stand-in for the upstream-LAMMPS classes that the reference's
// LAMMPS/ML-MTP/*.cpp sources include (pair_mtp.cpp:18-31,
// pair_mtp_extrapolation.cpp:18-36).  Upstream LAMMPS is not vendored by the
// reference and is not installed in this image, so this shim reproduces exactly
// the API surface those four files call, single-rank, so that they compile
// UNMODIFIED from /root/reference into oracle/_ref/libmtp_ref.so (see
// oracle/Makefile).  Semantics follow upstream LAMMPS (Pair::ev_setup flag
// decoding, TextFileReader::next_line, ValueTokenizer, utils::*), restated from
// their documented behaviour; nothing here is copied from the reference.
//
// The same shim is used to compile-check and drive the product's own
// PairStyle sources (lammps-mtp-kokkos_b200/lammps/) in the test-suite.
#ifndef LMP_SHIM_H
#define LMP_SHIM_H

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <format>
#include <stdexcept>
#include <string>
#include <vector>

#include "fmt/format.h"    // upstream pointers.h -> lmptype.h/utils.h pull in {fmt}

// ---------------------------------------------------------------- MPI stub (1 rank)
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
struct MPI_Status { int count; };
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_DOUBLE 2
#define MPI_CHAR 3
#define MPI_LONG_LONG 4
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_IN_PLACE ((void *) 1)
static inline size_t shim_mpi_size(MPI_Datatype t)
{
  return t == MPI_INT ? 4 : t == MPI_DOUBLE ? 8 : t == MPI_CHAR ? 1 : 8;
}
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{
  if (s != MPI_IN_PLACE) memcpy(r, s, n * shim_mpi_size(t));
  return 0;
}
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm)
{
  if (s != MPI_IN_PLACE) memcpy(r, s, n * shim_mpi_size(t));
  return 0;
}
static inline int MPI_Scan(const void *s, void *r, int n, MPI_Datatype t, MPI_Op, MPI_Comm)
{
  if (s != MPI_IN_PLACE) memcpy(r, s, n * shim_mpi_size(t));
  return 0;
}
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return 0; }
static inline int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *c) { *c = 0; return 0; }
static inline int MPI_Gather(const void *s, int n, MPI_Datatype t, void *r, int, MPI_Datatype, int, MPI_Comm)
{
  memcpy(r, s, n * shim_mpi_size(t));
  return 0;
}
static inline int MPI_Gatherv(const void *s, int n, MPI_Datatype t, void *r, const int *, const int *displs, MPI_Datatype, int,
                              MPI_Comm)
{
  memcpy((char *) r + (displs ? displs[0] : 0) * shim_mpi_size(t), s, n * shim_mpi_size(t));
  return 0;
}

namespace LAMMPS_NS {

typedef int64_t bigint;
typedef int tagint;
#define MPI_LMP_BIGINT MPI_LONG_LONG
#define NEIGHMASK 0x1FFFFFFF
#define FLERR __FILE__, __LINE__
#ifndef MIN
#define MIN(a, b) ((a) < (b) ? (a) : (b))
#define MAX(a, b) ((a) > (b) ? (a) : (b))
#endif

// Fatal errors become C++ exceptions so the test driver can observe them.
class LAMMPSAbortException : public std::runtime_error {
 public:
  explicit LAMMPSAbortException(const std::string &m) : std::runtime_error(m) {}
};

class LAMMPS;

class Error {
 public:
  std::string last_warning;
  template <typename... Args>
  [[noreturn]] void all(const std::string &file, int line, const std::string &fmt, Args &&...args)
  {
    raise(file, line, std::vformat(fmt, std::make_format_args(args...)));
  }
  template <typename... Args>
  [[noreturn]] void one(const std::string &file, int line, const std::string &fmt, Args &&...args)
  {
    raise(file, line, std::vformat(fmt, std::make_format_args(args...)));
  }
  template <typename... Args>
  void warning(const std::string &, int, const std::string &fmt, Args &&...args)
  {
    last_warning = std::vformat(fmt, std::make_format_args(args...));
  }

 private:
  [[noreturn]] static void raise(const std::string &file, int line, const std::string &msg)
  {
    throw LAMMPSAbortException("ERROR: " + msg + " (" + file + ":" + std::to_string(line) + ")");
  }
};

class Memory {
 public:
  // 1-D
  template <typename T> T *create(T *&a, int n, const char *)
  {
    a = (T *) malloc(sizeof(T) * (size_t) (n > 0 ? n : 1));
    return a;
  }
  template <typename T> T *grow(T *&a, int n, const char *)
  {
    a = (T *) realloc(a, sizeof(T) * (size_t) (n > 0 ? n : 1));
    return a;
  }
  template <typename T> void destroy(T *&a)
  {
    free(a);
    a = nullptr;
  }
  // 2-D, contiguous data with row pointers (upstream layout: a[0] is the data block)
  template <typename T> T **create(T **&a, int n1, int n2, const char *)
  {
    size_t nn1 = n1 > 0 ? n1 : 1, nn2 = n2 > 0 ? n2 : 1;
    T *data = (T *) malloc(sizeof(T) * nn1 * nn2);
    a = (T **) malloc(sizeof(T *) * nn1);
    for (size_t i = 0; i < nn1; i++) a[i] = data + i * nn2;
    return a;
  }
  template <typename T> T **grow(T **&a, int n1, int n2, const char *name)
  {
    if (a == nullptr) return create(a, n1, n2, name);
    size_t nn1 = n1 > 0 ? n1 : 1, nn2 = n2 > 0 ? n2 : 1;
    T *data = (T *) realloc(a[0], sizeof(T) * nn1 * nn2);
    a = (T **) realloc(a, sizeof(T *) * nn1);
    for (size_t i = 0; i < nn1; i++) a[i] = data + i * nn2;
    return a;
  }
  template <typename T> void destroy(T **&a)
  {
    if (a == nullptr) return;
    free(a[0]);
    free(a);
    a = nullptr;
  }
  // 3-D
  template <typename T> T ***create(T ***&a, int n1, int n2, int n3, const char *)
  {
    size_t nn1 = n1 > 0 ? n1 : 1, nn2 = n2 > 0 ? n2 : 1, nn3 = n3 > 0 ? n3 : 1;
    T *data = (T *) malloc(sizeof(T) * nn1 * nn2 * nn3);
    T **plane = (T **) malloc(sizeof(T *) * nn1 * nn2);
    a = (T ***) malloc(sizeof(T **) * nn1);
    for (size_t i = 0; i < nn1; i++) {
      a[i] = plane + i * nn2;
      for (size_t j = 0; j < nn2; j++) a[i][j] = data + (i * nn2 + j) * nn3;
    }
    return a;
  }
  template <typename T> T ***grow(T ***&a, int n1, int n2, int n3, const char *name)
  {
    if (a == nullptr) return create(a, n1, n2, n3, name);
    size_t nn1 = n1 > 0 ? n1 : 1, nn2 = n2 > 0 ? n2 : 1, nn3 = n3 > 0 ? n3 : 1;
    T *data = (T *) realloc(a[0][0], sizeof(T) * nn1 * nn2 * nn3);
    T **plane = (T **) realloc(a[0], sizeof(T *) * nn1 * nn2);
    a = (T ***) realloc(a, sizeof(T **) * nn1);
    for (size_t i = 0; i < nn1; i++) {
      a[i] = plane + i * nn2;
      for (size_t j = 0; j < nn2; j++) a[i][j] = data + (i * nn2 + j) * nn3;
    }
    return a;
  }
  template <typename T> void destroy(T ***&a)
  {
    if (a == nullptr) return;
    free(a[0][0]);
    free(a[0]);
    free(a);
    a = nullptr;
  }
};

class Atom {
 public:
  double **x = nullptr;
  double **f = nullptr;
  int *type = nullptr;
  tagint *tag = nullptr;
  bigint natoms = 0;
  int nlocal = 0, nghost = 0, nmax = 0, ntypes = 0;
};

class Comm {
 public:
  int me = 0, nprocs = 1;
};

class Force {
 public:
  int newton_pair = 1;
  int newton = 1;
  class Pair *pair = nullptr;    // the top-level pair style (this style itself unless pair_style hybrid wraps it)
};

class Domain {
 public:
  double xprd = 0, yprd = 0, zprd = 0, xy = 0, xz = 0, yz = 0;
  double boxlo[3] = {0, 0, 0}, boxhi[3] = {0, 0, 0};
};

namespace NeighConst {
  enum { REQ_DEFAULT = 0, REQ_FULL = 1 << 0, REQ_GHOST = 1 << 1 };
}

class NeighRequest {
 public:
  int flags = 0;
  int kokkos_host = 0, kokkos_device = 0;
  void set_kokkos_host(int v) { kokkos_host = v; }
  void set_kokkos_device(int v) { kokkos_device = v; }
};

class Neighbor {
 public:
  std::vector<NeighRequest *> requests;
  int ago = 0;
  NeighRequest *add_request(class Pair *, int flags = 0)
  {
    auto *r = new NeighRequest;
    r->flags = flags;
    requests.push_back(r);
    return r;
  }
  ~Neighbor()
  {
    for (auto *r : requests) delete r;
  }
};

class NeighList {
 public:
  int inum = 0, gnum = 0;
  int *ilist = nullptr;
  int *numneigh = nullptr;
  int **firstneigh = nullptr;
};

class LAMMPS {
 public:
  Memory *memory;
  Error *error;
  Atom *atom;
  Comm *comm;
  Force *force;
  Domain *domain;
  Neighbor *neighbor;
  MPI_Comm world = MPI_COMM_WORLD;
  std::string log;    // utils::logmesg sink
  char *suffix = nullptr;
  int suffix_enable = 0;
  LAMMPS() :
      memory(new Memory), error(new Error), atom(new Atom), comm(new Comm), force(new Force),
      domain(new Domain), neighbor(new Neighbor)
  {
  }
  ~LAMMPS()
  {
    delete neighbor;
    delete domain;
    delete force;
    delete comm;
    delete atom;
    delete error;
    delete memory;
  }
};

class Pointers {
 public:
  explicit Pointers(LAMMPS *ptr) :
      lmp(ptr), memory(ptr->memory), error(ptr->error), atom(ptr->atom), comm(ptr->comm),
      force(ptr->force), domain(ptr->domain), neighbor(ptr->neighbor), world(ptr->world)
  {
  }
  virtual ~Pointers() = default;

 protected:
  LAMMPS *lmp;
  Memory *&memory;
  Error *&error;
  Atom *&atom;
  Comm *&comm;
  Force *&force;
  Domain *&domain;
  Neighbor *&neighbor;
  MPI_Comm &world;
};

// energy / virial flag bits (upstream pair.h / integrate.h conventions)
enum { ENERGY_NONE = 0, ENERGY_GLOBAL = 1, ENERGY_ATOM = 2 };
enum { VIRIAL_NONE = 0, VIRIAL_PAIR = 1, VIRIAL_FDOTR = 2, VIRIAL_ATOM = 4, VIRIAL_CENTROID = 8 };

class Pair : protected Pointers {
 public:
  double eng_vdwl = 0, eng_coul = 0;
  double virial[6] = {0, 0, 0, 0, 0, 0};
  double *eatom = nullptr, **vatom = nullptr, **cvatom = nullptr;
  double cutforce = 0;
  double **cutsq = nullptr;
  int **setflag = nullptr;
  int comm_forward = 0, comm_reverse = 0;
  int single_enable = 1, respa_enable = 0, one_coeff = 0, manybody_flag = 0, restartinfo = 1;
  int no_virial_fdotr_compute = 0;
  int nextra = 0;
  double *pvector = nullptr;
  int evflag = 0, eflag_either = 0, eflag_global = 0, eflag_atom = 0;
  int vflag_either = 0, vflag_global = 0, vflag_atom = 0, cvflag_atom = 0, vflag_fdotr = 0;
  int maxeatom = 0, maxvatom = 0;
  int copymode = 0, kokkosable = 0;
  int execution_space = 0;
  unsigned int datamask_read = 0, datamask_modify = 0;
  int allocated = 0;
  NeighList *list = nullptr;
  char *suffix = nullptr;

  explicit Pair(LAMMPS *l) : Pointers(l) {}
  ~Pair() override
  {
    if (copymode) return;
    memory->destroy(eatom);
    memory->destroy(vatom);
  }
  virtual void compute(int, int) = 0;
  virtual void settings(int, char **) = 0;
  virtual void coeff(int, char **) = 0;
  virtual void init_style() {}
  virtual double init_one(int, int) { return 0.0; }
  virtual void init_list(int, NeighList *ptr) { list = ptr; }
  virtual void *extract(const char *, int &) { return nullptr; }
  virtual void *extract_peratom(const char *, int &) { return nullptr; }

  void ev_init(int eflag, int vflag, int alloc = 1)
  {
    if (eflag || vflag) ev_setup(eflag, vflag, alloc);
    else
      ev_unset();
  }
  void ev_unset()
  {
    evflag = eflag_either = eflag_global = eflag_atom = 0;
    vflag_either = vflag_global = vflag_atom = cvflag_atom = vflag_fdotr = 0;
  }

  // upstream Pair::ev_setup: decode flags, (re)allocate and zero the accumulators
  void ev_setup(int eflag, int vflag, int alloc = 1)
  {
    evflag = 1;
    eflag_either = eflag;
    eflag_global = eflag & ENERGY_GLOBAL;
    eflag_atom = eflag & ENERGY_ATOM;
    vflag_global = vflag & (VIRIAL_PAIR | VIRIAL_FDOTR);
    vflag_atom = vflag & VIRIAL_ATOM;
    cvflag_atom = vflag & VIRIAL_CENTROID;
    vflag_either = vflag_global || vflag_atom || cvflag_atom;

    if (eflag_atom && atom->nmax > maxeatom) {
      maxeatom = atom->nmax;
      if (alloc) {
        memory->destroy(eatom);
        memory->create(eatom, maxeatom, "pair:eatom");
      }
    }
    if (vflag_atom && atom->nmax > maxvatom) {
      maxvatom = atom->nmax;
      if (alloc) {
        memory->destroy(vatom);
        memory->create(vatom, maxvatom, 6, "pair:vatom");
      }
    }
    if (eflag_global) eng_vdwl = eng_coul = 0.0;
    if (vflag_global)
      for (int i = 0; i < 6; i++) virial[i] = 0.0;
    int n = atom->nlocal;
    if (force->newton) n += atom->nghost;
    if (eflag_atom && alloc)
      for (int i = 0; i < n; i++) eatom[i] = 0.0;
    if (vflag_atom && alloc)
      for (int i = 0; i < n; i++)
        for (int k = 0; k < 6; k++) vatom[i][k] = 0.0;

    if (vflag_global == VIRIAL_FDOTR && no_virial_fdotr_compute == 0) {
      vflag_fdotr = 1;
      vflag_global = 0;
      if (vflag_atom == 0 && cvflag_atom == 0) vflag_either = 0;
      if (vflag_either == 0 && eflag_either == 0) evflag = 0;
    } else
      vflag_fdotr = 0;
  }

  // upstream Pair::virial_fdotr_compute (single rank: all ghosts included)
  void virial_fdotr_compute()
  {
    double **x = atom->x, **f = atom->f;
    int nall = atom->nlocal + atom->nghost;
    for (int i = 0; i < nall; i++) {
      virial[0] += x[i][0] * f[i][0];
      virial[1] += x[i][1] * f[i][1];
      virial[2] += x[i][2] * f[i][2];
      virial[3] += x[i][1] * f[i][0];
      virial[4] += x[i][2] * f[i][0];
      virial[5] += x[i][2] * f[i][1];
    }
  }
};

// ---------------------------------------------------------------- tokenizer / file reader
#define TOKENIZER_DEFAULT_SEPARATORS " \t\r\n\f"

class TokenizerException : public std::exception {
  std::string message;

 public:
  TokenizerException(const std::string &msg, const std::string &token) :
      message(token.empty() ? msg : msg + ": '" + token + "'")
  {
  }
  const char *what() const noexcept override { return message.c_str(); }
};
class InvalidIntegerException : public TokenizerException {
 public:
  explicit InvalidIntegerException(const std::string &t) : TokenizerException("Not a valid integer number", t) {}
};
class InvalidFloatException : public TokenizerException {
 public:
  explicit InvalidFloatException(const std::string &t) : TokenizerException("Not a valid floating-point number", t) {}
};

class ValueTokenizer {
  std::string text, seps;
  size_t pos = 0;

 public:
  ValueTokenizer(const std::string &str, const std::string &separators = TOKENIZER_DEFAULT_SEPARATORS) :
      text(str), seps(separators)
  {
  }
  bool has_next()
  {
    size_t p = text.find_first_not_of(seps, pos);
    return p != std::string::npos;
  }
  std::string next_string()
  {
    size_t b = text.find_first_not_of(seps, pos);
    if (b == std::string::npos) throw TokenizerException("No more tokens", "");
    size_t e = text.find_first_of(seps, b);
    if (e == std::string::npos) e = text.size();
    pos = e;
    return text.substr(b, e - b);
  }
  int next_int()
  {
    std::string t = next_string();
    char *end = nullptr;
    long v = strtol(t.c_str(), &end, 10);
    if (end == t.c_str() || *end != '\0') throw InvalidIntegerException(t);
    return (int) v;
  }
  bigint next_bigint()
  {
    std::string t = next_string();
    char *end = nullptr;
    long long v = strtoll(t.c_str(), &end, 10);
    if (end == t.c_str() || *end != '\0') throw InvalidIntegerException(t);
    return (bigint) v;
  }
  double next_double()
  {
    std::string t = next_string();
    char *end = nullptr;
    double v = strtod(t.c_str(), &end);
    if (end == t.c_str() || *end != '\0') throw InvalidFloatException(t);
    return v;
  }
  int count()
  {
    int n = 0;
    size_t p = 0;
    while ((p = text.find_first_not_of(seps, p)) != std::string::npos) {
      n++;
      p = text.find_first_of(seps, p);
      if (p == std::string::npos) break;
    }
    return n;
  }
};

class TextFileReader {
  std::string filetype;
  bool closefp = false;
  int bufsize = 1024;
  char *line;
  FILE *fp;

  static int count_words(const char *s)
  {
    int n = 0;
    const char *seps = TOKENIZER_DEFAULT_SEPARATORS;
    while (*s) {
      s += strspn(s, seps);
      if (!*s) break;
      n++;
      s += strcspn(s, seps);
    }
    return n;
  }

 public:
  bool ignore_comments = true;
  TextFileReader(FILE *f, std::string ftype) : filetype(std::move(ftype)), line(new char[1024]), fp(f) {}
  TextFileReader(const TextFileReader &) = delete;
  ~TextFileReader()
  {
    if (closefp) fclose(fp);
    delete[] line;
  }
  void set_bufsize(int newsize)
  {
    if (newsize < 100) throw std::runtime_error("line buffer size must be >= 100 bytes");
    delete[] line;
    bufsize = newsize;
    line = new char[bufsize];
  }
  void rewind() { ::rewind(fp); }
  void skip_line()
  {
    if (!fgets(line, bufsize, fp)) throw std::runtime_error("Missing line");
  }
  // upstream semantics: read physical lines until at least one word (or nparams
  // words) has been seen; strip '#' comments when ignore_comments; nullptr on EOF
  char *next_line(int nparams = 0)
  {
    int n = 0, nwords = 0;
    char *ptr = fgets(line, bufsize, fp);
    if (ptr == nullptr) return nullptr;
    if (ignore_comments && (ptr = strchr(line, '#'))) *ptr = '\0';
    nwords = count_words(line);
    if (nwords > 0) n = (int) strlen(line);
    while (nwords == 0 || nwords < nparams) {
      ptr = fgets(&line[n], bufsize - n, fp);
      if (ptr == nullptr) {
        if (nwords > 0 && nwords < nparams) throw std::runtime_error("Unexpected end of file");
        return nullptr;
      }
      if (ignore_comments && (ptr = strchr(line, '#'))) *ptr = '\0';
      nwords += count_words(&line[n]);
      if (nwords > 0) n = (int) strlen(line);
    }
    return line;
  }
};

class PotentialFileReader;    // declared only; the reference never instantiates it

namespace utils {
  template <typename... Args> void logmesg(LAMMPS *lmp, const std::string &fmt, Args &&...args)
  {
    lmp->log += std::vformat(fmt, std::make_format_args(args...));
  }
  inline std::string lowercase(const std::string &s)
  {
    std::string r(s);
    for (auto &c : r) c = (char) ::tolower((unsigned char) c);
    return r;
  }
  inline FILE *open_potential(const std::string &name, LAMMPS *lmp, int *)
  {
    FILE *fp = fopen(name.c_str(), "r");
    if (!fp) {
      const char *dir = getenv("LAMMPS_POTENTIALS");
      if (dir) {
        std::string alt = std::string(dir) + "/" + name;
        fp = fopen(alt.c_str(), "r");
      }
    }
    if (!fp) lmp->error->one(FLERR, "Cannot open potential file {}", name);
    return fp;
  }
  inline double numeric(const char *file, int line, const std::string &str, bool, LAMMPS *lmp)
  {
    char *end = nullptr;
    double v = strtod(str.c_str(), &end);
    if (str.empty() || end == str.c_str() || *end != '\0')
      lmp->error->all(file, line, "Expected floating point parameter instead of '{}' in input script or data file", str);
    return v;
  }
  inline int inumeric(const char *file, int line, const std::string &str, bool, LAMMPS *lmp)
  {
    char *end = nullptr;
    long v = strtol(str.c_str(), &end, 10);
    if (str.empty() || end == str.c_str() || *end != '\0')
      lmp->error->all(file, line, "Expected integer parameter instead of '{}' in input script or data file", str);
    return (int) v;
  }
  inline void sfread(const char *srcname, int srcline, void *s, size_t size, size_t num, FILE *fp,
                     const char *, Error *error)
  {
    size_t rv = fread(s, size, num, fp);
    if (rv != num) error->one(srcname, srcline, "Unexpected end of file or read error while reading binary data");
  }
}    // namespace utils

}    // namespace LAMMPS_NS

#endif
