// TEST INFRASTRUCTURE -- the handful of LAMMPS-KOKKOS types the LMP_KOKKOS flavour of the product's PairStyle
// (lammps-mtp-kokkos_b200/lammps/pair_mtp_b200.cpp) touches, reduced to plain CUDA-runtime allocations so that the
// flavour compiles and RUNS here without Kokkos or upstream LAMMPS: device views with run-time strides (LayoutLeft for
// the 2-D neighbor view, as in LAMMPS-KOKKOS on CUDA), DualView sync / modify, AtomKokkos, NeighListKokkos, MemoryKokkos.
// What the reference itself uses of them: LAMMPS/KOKKOS/pair_mtp_kokkos.cpp:37-45,215-240,379-390.
#pragma once

#include "lammps_shim.h"

#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

struct LMPDeviceType {};
struct LMPHostType {};

namespace LAMMPS_NS {

enum ExecutionSpace { Host, Device };
#define EMPTY_MASK 0x00000000
#define X_MASK 0x00000001
#define F_MASK 0x00000004
#define TYPE_MASK 0x00000040

inline void kk_check(cudaError_t e, const char *what)
{
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}

// rank-1 / rank-2 view over device memory with explicit strides (in elements)
template <class T> struct KKView {
  T *ptr = nullptr;
  long ext[2] = {0, 0}, str[2] = {1, 1};
  T *data() const { return ptr; }
  long extent(int d) const { return ext[d]; }
  long stride(int d) const { return str[d]; }
  long span() const { return ext[1] > 0 ? ext[0] * ext[1] : ext[0]; }
};

// host array + device mirror; sync / modify as in Kokkos::DualView (whole-array copies)
template <class T> struct KKDualView {
  KKView<T> d_view;
  T *h_ptr = nullptr;
  bool host_dirty = false, device_dirty = false;
  template <class Space> KKView<T> view() const { return d_view; }
  template <class Space> void modify() { (std::is_same<Space, LMPDeviceType>::value ? device_dirty : host_dirty) = true; }
  template <class Space> void sync()
  {
    const size_t bytes = sizeof(T) * (size_t) d_view.span();
    if (std::is_same<Space, LMPHostType>::value) {
      if (device_dirty && bytes) kk_check(cudaMemcpy(h_ptr, d_view.ptr, bytes, cudaMemcpyDeviceToHost), "DualView sync to host");
      device_dirty = false;
    } else {
      if (host_dirty && bytes) kk_check(cudaMemcpy(d_view.ptr, h_ptr, bytes, cudaMemcpyHostToDevice), "DualView sync to device");
      host_dirty = false;
    }
  }
  void allocate(T *host, long n0, long n1 = 0)    // row-major (LayoutRight) like t_x_array / t_f_array
  {
    release();
    h_ptr = host;
    d_view.ext[0] = n0;
    d_view.ext[1] = n1;
    d_view.str[0] = n1 > 0 ? n1 : 1;
    d_view.str[1] = 1;
    if (d_view.span()) {    // Kokkos views are zero-initialised
      kk_check(cudaMalloc((void **) &d_view.ptr, sizeof(T) * (size_t) d_view.span()), "cudaMalloc");
      kk_check(cudaMemset(d_view.ptr, 0, sizeof(T) * (size_t) d_view.span()), "cudaMemset");
    }
  }
  void release()
  {
    if (d_view.ptr) cudaFree(d_view.ptr);
    d_view = KKView<T>();
  }
};

struct DAT {
  typedef KKDualView<double> tdual_x_array, tdual_f_array, tdual_efloat_1d, tdual_virial_array;
  typedef KKDualView<int> tdual_int_1d;
};

class AtomKokkos : public Atom {
 public:
  DAT::tdual_x_array k_x;
  DAT::tdual_f_array k_f;
  DAT::tdual_int_1d k_type;
  int nsync = 0, nmodified = 0;
  void sync(ExecutionSpace space, unsigned mask)
  {
    nsync++;
    if (space != Device) return;
    if (mask & X_MASK) k_x.sync<LMPDeviceType>();
    if (mask & F_MASK) k_f.sync<LMPDeviceType>();
    if (mask & TYPE_MASK) k_type.sync<LMPDeviceType>();
  }
  void modified(ExecutionSpace space, unsigned mask)
  {
    nmodified++;
    if (space == Device && (mask & F_MASK)) k_f.modify<LMPDeviceType>();
  }
};

template <class DeviceType> class NeighListKokkos : public NeighList {
 public:
  KKView<int> d_ilist, d_numneigh, d_neighbors;    // d_neighbors(i, jj): LayoutLeft on CUDA -> stride(0) = 1
  int maxneighs = 0;
};

class MemoryKokkos {
 public:
  template <class T> void create_kokkos(KKDualView<T> &k, T *&host, long n, const char *)
  {
    host = n > 0 ? new T[n]() : nullptr;
    k.allocate(host, n);
  }
  template <class T> void create_kokkos(KKDualView<T> &k, T **&host, long n0, long n1, const char *)
  {
    T *flat = n0 * n1 > 0 ? new T[n0 * n1]() : nullptr;
    host = n0 > 0 ? new T *[n0] : nullptr;
    for (long i = 0; i < n0; i++) host[i] = flat + i * n1;
    k.allocate(flat, n0, n1);
  }
  template <class T> void destroy_kokkos(KKDualView<T> &k, T *&host)
  {
    k.release();
    delete[] host;
    host = nullptr;
  }
  template <class T> void destroy_kokkos(KKDualView<T> &k, T **&host)
  {
    k.release();
    if (host) {
      delete[] host[0];
      delete[] host;
    }
    host = nullptr;
  }
};

}    // namespace LAMMPS_NS
