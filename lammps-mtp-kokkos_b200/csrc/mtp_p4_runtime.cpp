// NVRTC compilation, cubin cache and module loading for the generated contraction-program kernel.
#include "mtp_p4_runtime.hpp"

#include <nvrtc.h>

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <stdexcept>
#include <sys/stat.h>
#include <unistd.h>

namespace mtpb200 {

P4Choice p4_choose(const Potential &p, size_t smem_optin, bool latency_shape)
{
  P4Choice ch;
  if (const char *e = getenv(latency_shape ? "MTP_B200_P4_SMALL" : "MTP_B200_P4")) {
    int na = 0, w = 0, c = 0, acc = 0, mb = 0;
    long budget = 0;
    int groups = 1, fn_cost = 0, sparse = 0, spatial = 0, rpar = 0;
    const int nf = sscanf(e, "%d,%d,%d,%d,%d,%ld,%d,%d,%d,%d,%d", &na, &w, &c, &acc, &mb, &budget, &groups, &fn_cost, &sparse, &spatial,
                          &rpar);
    if (nf >= 5) {
      if (nf >= 7) ch.prm.groups = groups;
      if (nf >= 8 && fn_cost > 0) ch.prm.fn_cost = fn_cost;
      if (nf >= 9) ch.prm.sparse = sparse;
      if (nf >= 10) ch.prm.spatial = spatial;
      if (nf >= 11) ch.prm.rpar = rpar;
      ch.prm.na = na;
      ch.prm.warps = w;
      ch.prm.cache = c;
      ch.prm.acc_max = acc;
      ch.min_blocks = mb;
      ch.prm.smem_budget = (nf >= 6 && budget > 0) ? (size_t) budget : smem_optin;
      const size_t b = p4_smem_bytes(p, ch.prm);
      ch.ok = b > 0 && b <= smem_optin;
      return ch;
    }
  }
  // one SM has 228 KB of shared memory, 1 KB of which is reserved per resident CTA
  const size_t two_ctas = (smem_optin + 1024) / 2 - 1024;
  // The emitted code is straight-line and executed once per chunk: it streams through the instruction caches, and what
  // bounds it is the request rate of the GPC-level instruction cache the SMs of a GPC share.  The throughput
  // shape therefore always takes a full warp of atoms per row (32 per CTA: atoms per fetched instruction).
  // Latency shape (mtp/small/kk: few atoms, every SM must get a chunk): a chunk streams the whole program through one SM
  // whatever its width, so the widest chunk that fits (16 atoms, else 8) halves the number of chunks per SM; 8 warps
  // shorten the chunk's critical path.
  // Throughput shape (measured on B200, profiles/r2_p4_shapes.txt): what the kernel waits for is instruction fetch
  // (straight-line code, every CTA streams it), the barriers between stages and the latency of its own global traffic,
  // so MORE RESIDENT CTAs beat bigger CTAs: each keeps only the basic moments its current round reads (sparse rounds),
  // which lets four 4-warp CTAs (level <= 16 or so) or two 8-warp CTAs (levels 20-24) share an SM.
  if (!latency_shape && !getenv("MTP_B200_P4_DENSE")) {
    const size_t four_ctas = (smem_optin + 1024) / 4 - 1024;
    P4Params prm;
    prm.na = 32;
    prm.cache = 40;    // 128 registers per thread at these occupancies
    prm.acc_max = 12;
    prm.sparse = 1;
    int rounds = 0;
    prm.warps = 4;
    prm.smem_budget = four_ctas;
    size_t b = p4_smem_bytes(p, prm, &rounds);
    if (b > 0 && b <= four_ctas && rounds <= 4) {
      ch.prm = prm;
      ch.min_blocks = 4;
      ch.ok = true;
      return ch;
    }
    prm.warps = 8;
    prm.smem_budget = two_ctas;
    b = p4_smem_bytes(p, prm, &rounds);
    if (b > 0 && b <= two_ctas && rounds <= 24) {
      ch.prm = prm;
      ch.min_blocks = 2;
      ch.ok = true;
      return ch;
    }
  }
  // Latency shape, large programs: ROUNDS IN PARALLEL.  A chunk's critical path is the whole program streamed through one
  // SM; cutting the basis functions into 2-8 sparse rounds that run as separate CTAs (gridDim.y = rounds, shares added to a
  // zeroed gb by RED.ADD) divides that path by the number of rounds and fills the SMs a small system leaves idle.
  // MEASURED (config 3, 2,000 atoms, level 20): 66 us against 54 us for the single-CTA-per-chunk shape -- the rounds of all
  // chunks stream more distinct code through the GPC instruction caches, which is what bounds the kernel, and the zeroing
  // of gb is on the path.  Not the default; MTP_B200_P4_RPAR=1 selects it (levels 18-20 parity-tested on the GPU).
  if (latency_shape && getenv("MTP_B200_P4_RPAR")) {
    const size_t four_ctas = (smem_optin + 1024) / 4 - 1024;
    P4Params prm;
    prm.na = 16;
    prm.warps = 4;
    prm.cache = 40;
    prm.acc_max = 12;
    prm.sparse = 1;
    prm.rpar = 1;
    prm.smem_budget = four_ctas;
    int rounds = 0;
    const size_t b = p4_smem_bytes(p, prm, &rounds);
    if (b > 0 && b <= four_ctas && rounds >= 2 && rounds <= 8) {
      ch.prm = prm;
      ch.min_blocks = 4;
      ch.ok = true;
      return ch;
    }
  }
  const int nas[2] = {latency_shape ? 16 : 32, latency_shape ? 8 : 32};
  for (int t = 0; t < (latency_shape ? 2 : 1); t++) {
    P4Params prm;
    prm.na = nas[t];
    prm.warps = latency_shape ? 8 : 4;
    size_t b = p4_smem_bytes(p, prm);    // smem_budget = 0: the whole program in one round
    if (b == 0) return ch;               // structure not supported
    if (b <= two_ctas) {                 // two CTAs per SM: their instruction streams overlap
      ch.prm = prm;
      ch.min_blocks = 2;
      ch.ok = true;
      return ch;
    }
    prm.warps = 8;
    b = p4_smem_bytes(p, prm);
    if (b <= smem_optin) {
      ch.prm = prm;
      ch.min_blocks = 1;
      ch.ok = true;
      return ch;
    }
  }
  // Too many rows for one CTA (levels >= 20): the basis functions are dealt to rounds that reuse the rows (mtp_codegen.cpp,
  // make_rounds); a full warp of atoms per row for the throughput shape, 16 for the latency shape.
  P4Params prm;
  prm.na = latency_shape ? 16 : 32;
  prm.warps = 8;
  prm.smem_budget = smem_optin;
  const size_t b = p4_smem_bytes(p, prm);
  if (b > 0 && b <= smem_optin) {
    ch.prm = prm;
    ch.min_blocks = 1;
    ch.ok = true;
  }
  return ch;
}

std::string p4_cache_dir()
{
  if (const char *e = getenv("MTP_B200_KCACHE")) return e;
  Dl_info di;
  if (dladdr((const void *) &p4_cache_dir, &di) && di.dli_fname) {
    std::string path = di.dli_fname;
    const size_t slash = path.rfind('/');
    if (slash != std::string::npos) return path.substr(0, slash) + "/kcache";
  }
  return "/tmp/mtp_b200_kcache";
}

namespace {

std::string cache_file(const P4Info &info, const P4Choice &ch)
{
  char name[96];
  snprintf(name, sizeof(name), "/p4_%016llx_b%d%s.cubin", info.hash, ch.min_blocks, getenv("MTP_B200_P4_LINEINFO") ? "_li" : "");
  return p4_cache_dir() + name;
}

bool read_file(const std::string &path, std::vector<char> &out)
{
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) return false;
  fseek(f, 0, SEEK_END);
  const long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  out.resize(n > 0 ? (size_t) n : 0);
  const bool ok = n > 0 && fread(out.data(), 1, (size_t) n, f) == (size_t) n;
  fclose(f);
  return ok;
}

void write_file_atomic(const std::string &path, const std::vector<char> &data)
{
  const std::string dir = path.substr(0, path.rfind('/'));
  mkdir(dir.c_str(), 0755);    // best effort: an unwritable cache only costs a recompilation next time
  const std::string tmp = path + ".tmp." + std::to_string((long) getpid());
  FILE *f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const bool ok = fwrite(data.data(), 1, data.size(), f) == data.size();
  fclose(f);
  if (!ok || rename(tmp.c_str(), path.c_str()) != 0) unlink(tmp.c_str());
}

std::vector<char> nvrtc_compile(const std::string &src, int min_blocks)
{
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, src.c_str(), "mtp_program_p4.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS)
    throw std::runtime_error("nvrtcCreateProgram failed");
  const std::string minb = "-DP4_MINB=" + std::to_string(min_blocks);
  // (-lineinfo triples the size of these cubins; MTP_B200_P4_LINEINFO=1 adds it for an ncu source view)
  const char *opts[] = {"-arch=sm_100a", "-std=c++17", minb.c_str(), "-lineinfo"};
  const nvrtcResult rc = nvrtcCompileProgram(prog, getenv("MTP_B200_P4_LINEINFO") ? 4 : 3, opts);
  if (rc != NVRTC_SUCCESS) {
    size_t n = 0;
    nvrtcGetProgramLogSize(prog, &n);
    std::string log(n, '\0');
    if (n) nvrtcGetProgramLog(prog, &log[0]);
    nvrtcDestroyProgram(&prog);
    throw std::runtime_error(std::string("NVRTC failed to compile the contraction-program kernel: ") + nvrtcGetErrorString(rc) + "\n" +
                             log.substr(0, 2000));
  }
  size_t n = 0;
  nvrtcGetCUBINSize(prog, &n);
  std::vector<char> cubin(n);
  if (n) nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  if (cubin.empty()) throw std::runtime_error("NVRTC produced no cubin for sm_100a");
  return cubin;
}

}    // namespace

std::vector<char> p4_cubin(const Potential &p, const P4Choice &ch, const short *slot_of_k, int nslots, P4Info &info, bool *compiled)
{
  std::string src, why;
  if (compiled) *compiled = false;
  if (!p4_generate(p, ch.prm, slot_of_k, nslots, src, info, why))
    throw std::runtime_error("contraction-program generator: " + why);
  const std::string path = cache_file(info, ch);
  std::vector<char> cubin;
  if (!getenv("MTP_B200_KCACHE_OFF") && read_file(path, cubin)) return cubin;
  cubin = nvrtc_compile(src, ch.min_blocks);
  if (compiled) *compiled = true;
  if (!getenv("MTP_B200_KCACHE_OFF")) write_file_atomic(path, cubin);
  return cubin;
}

#define P4_CUDA(expr)                                                                                              \
  do {                                                                                                             \
    cudaError_t e__ = (expr);                                                                                      \
    if (e__ != cudaSuccess) throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(e__));          \
  } while (0)

void P4Module::load(const std::vector<char> &cubin, int device, int sm_count)
{
  unload();
  P4_CUDA(cudaSetDevice(device));
  P4_CUDA(cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  P4_CUDA(cudaLibraryGetKernel(&kernel, lib, "mtp_program_p4"));
  P4_CUDA(cudaFuncSetAttribute((const void *) kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) info.smem_bytes));    // module-private kernel
  P4_CUDA(cudaFuncSetAttribute((const void *) kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  int per_sm = 0;
  P4_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *) kernel, info.threads, info.smem_bytes));
  grid_cap = (per_sm > 0 ? per_sm : 1) * sm_count;
}

void P4Module::unload()
{
  if (lib) cudaLibraryUnload(lib);
  lib = nullptr;
  kernel = nullptr;
}

}    // namespace mtpb200
