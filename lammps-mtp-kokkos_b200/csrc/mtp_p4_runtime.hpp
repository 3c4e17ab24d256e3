// Run-time side of the generated contraction-program kernel: NVRTC compilation of the emitted source to an sm_100a
// cubin, an on-disk cubin cache keyed by the structure hash, and the module handle the launch path uses.
// NVRTC is a pure compiler: it runs without a GPU, so cubins can be produced ahead of time (mtp_codegen_prebuild).
#pragma once

#include "mtp_codegen.hpp"

#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace mtpb200 {

constexpr size_t kSm100SmemOptin = 232448;    // sharedMemPerBlockOptin of sm_100 (227 KB)

struct P4Choice {
  P4Params prm;
  int min_blocks = 1;    // CTAs per SM the kernel is compiled for (__launch_bounds__)
  bool ok = false;
  int atoms_per_cta() const { return prm.na * (prm.groups > 1 ? prm.groups : 1); }    // per CTA iteration
};

// atoms per CTA / warps for potential p given the shared-memory limit;
// override: MTP_B200_P4="na,warps,cache,acc,minb[,smem_budget[,groups[,fn_cost[,sparse[,spatial[,rpar]]]]]]"
P4Choice p4_choose(const Potential &p, size_t smem_optin, bool latency_shape);

// directory of the cubin cache: $MTP_B200_KCACHE, else <directory of this shared library>/kcache
std::string p4_cache_dir();

// cubin for (p, choice): from the cache, else generated + compiled (and stored).  Throws std::runtime_error.
// `compiled` (optional) is set when NVRTC actually ran.
std::vector<char> p4_cubin(const Potential &p, const P4Choice &ch, const short *slot_of_k, int nslots, P4Info &info,
                           bool *compiled = nullptr);

struct P4Module {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t kernel = nullptr;
  P4Info info;
  P4Choice choice;
  int grid_cap = 0;
  bool loaded() const { return kernel != nullptr; }
  void load(const std::vector<char> &cubin, int device, int sm_count);    // throws on CUDA errors
  void unload();
};

}    // namespace mtpb200
