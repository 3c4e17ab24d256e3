"""ctypes front-end of tests/shim/pair_b200_driver.cpp: the product's LAMMPS PairStyle classes driven the way
LAMMPS drives them (pair_style string + arguments, coeff, init, compute), against oracle/lammps_shim/."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "lammps-mtp-kokkos_b200", "lammps", "libpair_mtp_b200_shim.so")
SO_KK = os.path.join(ROOT, "lammps-mtp-kokkos_b200", "lammps", "libpair_mtp_b200_kk_shim.so")    # LMP_KOKKOS flavour

_dp, _ip, _lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long)


def build():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build_cuda()
    g.build_lammps_plugin(kokkos=True)
    return g.build_lammps_plugin()


class LammpsError(RuntimeError):
    pass


class PairB200:
    def __init__(self, style, *args, species=1, newton=1, kokkos=False, lone=False):
        """kokkos: the LMP_KOKKOS flavour of the style (device views through tests/shim/kokkos_stub);
        lone: the style is force->pair itself (no pair_style hybrid above it)."""
        if not os.path.exists(SO) or not os.path.exists(SO_KK):
            build()
        self.lib = C.CDLL(SO_KK if kokkos else SO)
        assert bool(self.lib.b200drv_is_kokkos()) == kokkos
        self.lib.b200drv_create.restype = C.c_void_p
        self.lib.b200drv_log.restype = C.c_char_p
        err = C.create_string_buffer(2048)
        argv = (C.c_char_p * max(len(args), 1))(*[str(a).encode() for a in args])
        self.h = self.lib.b200drv_create(style.encode(), C.c_int(len(args)), argv, C.c_int(species), err, C.c_int(2048))
        if not self.h:
            raise LammpsError(err.value.decode())
        self.lib.b200drv_set_lone_pair(C.c_void_p(self.h), C.c_int(1 if lone else 0))

    @property
    def log(self):
        return self.lib.b200drv_log(C.c_void_p(self.h)).decode()

    def set_domain(self, prd, natoms):
        a = np.zeros(6)
        a[: len(prd)] = prd
        self.lib.b200drv_set_domain(C.c_void_p(self.h), a.ctypes.data_as(_dp), C.c_long(natoms))

    def compute(self, x, type_, nlocal, ilist, numneigh, neigh, offsets, eflag=3, vflag=5, ago=0, grade=False, f_init=None):
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        ilist = np.ascontiguousarray(ilist, dtype=np.int32)
        numneigh = np.ascontiguousarray(numneigh, dtype=np.int32)
        neigh = np.ascontiguousarray(neigh, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        nall = x.shape[0]

        class R:
            pass
        r = R()
        r.f = np.zeros((nall, 3)) if f_init is None else np.array(f_init, dtype=np.float64)
        r.eatom, r.vatom, r.ev, r.grades = np.zeros(nall), np.zeros((nall, 6)), np.zeros(8), np.zeros(nall)
        err = C.create_string_buffer(2048)
        rc = self.lib.b200drv_compute(
            C.c_void_p(self.h), C.c_int(nlocal), C.c_int(nall - nlocal), x.ctypes.data_as(_dp), type_.ctypes.data_as(_ip),
            C.c_int(len(ilist)), ilist.ctypes.data_as(_ip), numneigh.ctypes.data_as(_ip), neigh.ctypes.data_as(_ip),
            offsets.ctypes.data_as(_lp), C.c_int(eflag), C.c_int(vflag), C.c_int(ago), C.c_int(1 if grade else 0),
            r.f.ctypes.data_as(_dp), r.eatom.ctypes.data_as(_dp), r.vatom.ctypes.data_as(_dp), r.ev.ctypes.data_as(_dp),
            r.grades.ctypes.data_as(_dp), err, C.c_int(2048))
        if rc != 0:
            raise LammpsError(err.value.decode())
        r.energy, r.virial, r.max_grade = r.ev[0], r.ev[1:7], r.ev[7]
        return r

    def close(self):
        if self.h:
            self.lib.b200drv_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
