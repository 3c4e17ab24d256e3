"""TEST INFRASTRUCTURE -- ctypes front-ends for the two parity checkers.

* ``OracleMTP``   : oracle/libmtp_oracle.so, the plain-C restatement (mtp_oracle.c)
* ``ReferenceMTP``: oracle/_ref/libmtp_ref.so, the reference's own unmodified CPU sources
                    (pair_mtp.cpp / pair_mtp_extrapolation.cpp) behind oracle/ref_driver.cpp

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  Neither library is ever used by the product path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libmtp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libmtp_ref.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)
_bp = C.POINTER(C.c_ubyte)


def build(ref: bool | None = None) -> None:
    """Compile the checkers (gcc / g++ only).  The reference build needs /root/reference."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref is None:
        ref = os.path.isdir("/root/reference/LAMMPS/ML-MTP")
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _ptr(a, t):
    return None if a is None else a.ctypes.data_as(t)


class _Params(C.Structure):
    _fields_ = [("species_count", C.c_int), ("radial_func_count", C.c_int), ("radial_basis_size", C.c_int),
                ("alpha_moment_count", C.c_int), ("alpha_index_basic_count", C.c_int),
                ("alpha_index_times_count", C.c_int), ("alpha_scalar_count", C.c_int),
                ("max_alpha_index_basic", C.c_int), ("min_cutoff", C.c_double), ("max_cutoff", C.c_double),
                ("scaling", C.c_double), ("radial_basis_coeffs", _dp), ("alpha_index_basic", _ip),
                ("alpha_index_times", _ip), ("alpha_moment_mapping", _ip), ("species_coeffs", _dp),
                ("linear_coeffs", _dp), ("coeff_count", C.c_int), ("configuration_mode", C.c_int),
                ("inverse_active_set", _dp)]


class Result:
    def __init__(self, nall, q=0):
        self.f = np.zeros((nall, 3))
        self.eatom = np.zeros(nall)
        self.vatom = np.zeros((nall, 6))
        self.ev = np.zeros(8)
        self.grades = np.zeros(nall)
        self.candidate = np.zeros(max(q, 1))
        self.mask = None

    @property
    def energy(self):
        return self.ev[0]

    @property
    def virial(self):
        return self.ev[1:7]

    @property
    def max_grade(self):
        return self.ev[7]


def _prep(x, type_, ilist, numneigh, neigh_flat, offsets):
    x = np.ascontiguousarray(x, dtype=np.float64)
    type_ = np.ascontiguousarray(type_, dtype=np.int32)
    ilist = np.ascontiguousarray(ilist, dtype=np.int32)
    numneigh = np.ascontiguousarray(numneigh, dtype=np.int32)
    neigh_flat = np.ascontiguousarray(neigh_flat, dtype=np.int32)
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    return x, type_, ilist, numneigh, neigh_flat, offsets


class OracleMTP:
    """The C restatement, driven from an ``MTPPotential`` (mtp_b200.almtp)."""

    def __init__(self, pot):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.mtp_oracle_compute.restype = C.c_int
        self.lib.mtp_oracle_grade.restype = C.c_double
        self.pot = pot
        self._keep = dict(
            rc=np.ascontiguousarray(pot.radial_coeffs, dtype=np.float64),
            basic=np.ascontiguousarray(pot.alpha_index_basic, dtype=np.int32),
            times=np.ascontiguousarray(pot.alpha_index_times, dtype=np.int32),
            mapping=np.ascontiguousarray(pot.alpha_moment_mapping, dtype=np.int32),
            sc=np.ascontiguousarray(pot.species_coeffs, dtype=np.float64),
            lc=np.ascontiguousarray(pot.moment_coeffs, dtype=np.float64))
        k = self._keep
        inv = None
        if pot.inverse_active_set is not None:
            k["inv"] = np.ascontiguousarray(pot.inverse_active_set, dtype=np.float64)
            inv = k["inv"]
        cfg = int(pot.energy_weight) == 1 if pot.energy_weight is not None else 0
        self.params = _Params(
            pot.species_count, pot.radial_funcs_count, pot.radial_basis_size, pot.alpha_moments_count,
            pot.K, pot.T, pot.A, pot.max_alpha_index_basic, pot.min_dist, pot.max_dist, pot.scaling,
            _ptr(k["rc"], _dp), _ptr(k["basic"], _ip), _ptr(k["times"], _ip), _ptr(k["mapping"], _ip),
            _ptr(k["sc"], _dp), _ptr(k["lc"], _dp), pot.coeff_count, int(cfg), _ptr(inv, _dp))

    def chebyshev(self, dist):
        B = self.pot.radial_basis_size
        v = np.zeros(B)
        d = np.zeros(B)
        self.lib.mtp_oracle_chebyshev(C.c_double(dist), C.c_double(self.pot.min_dist),
                                      C.c_double(self.pot.max_dist), C.c_double(self.pot.scaling),
                                      C.c_int(B), _ptr(v, _dp), _ptr(d, _dp))
        return v, d

    def compute(self, x, type_, ilist, numneigh, neigh_flat, offsets, eflag=3, vflag=5, grade=False,
                natoms_total=None, want_mask=False, f_init=None) -> Result:
        x, type_, ilist, numneigh, neigh_flat, offsets = _prep(x, type_, ilist, numneigh, neigh_flat, offsets)
        nall = x.shape[0]
        r = Result(nall, self.pot.coeff_count)
        if f_init is not None:
            r.f[:] = f_init
        if want_mask:
            r.mask = np.zeros(neigh_flat.shape[0], dtype=np.uint8)
        nat = len(ilist) if natoms_total is None else natoms_total
        rc = self.lib.mtp_oracle_compute(
            C.byref(self.params), C.c_int(nall), _ptr(x, _dp), _ptr(type_, _ip), C.c_int(len(ilist)),
            _ptr(ilist, _ip), _ptr(numneigh, _ip), _ptr(neigh_flat, _ip), _ptr(offsets, _lp), C.c_int(eflag),
            C.c_int(vflag), C.c_int(1 if grade else 0), C.c_long(nat), _ptr(r.f, _dp), _ptr(r.eatom, _dp),
            _ptr(r.vatom, _dp), _ptr(r.ev, _dp), _ptr(r.grades, _dp), _ptr(r.candidate, _dp),
            _ptr(r.mask, _bp))
        if rc != 0:
            raise RuntimeError("Too few species count in the MTP potential!")
        return r


class ReferenceMTP:
    """The reference's own CPU pair style (``mtp`` or ``mtp/extrapolation``), from a potential FILE."""

    def __init__(self, style: str, *args: str):
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(REF_SO + " (build it here with `make -C oracle ref`)")
        self.lib = C.CDLL(REF_SO)
        self.lib.mtpref_create.restype = C.c_void_p
        self.lib.mtpref_log.restype = C.c_char_p
        err = C.create_string_buffer(1024)
        argv = (C.c_char_p * len(args))(*[a.encode() for a in args])
        self.h = self.lib.mtpref_create(style.encode(), C.c_int(len(args)), argv, err, C.c_int(1024))
        if not self.h:
            raise RuntimeError(err.value.decode())
        iv = (C.c_int * 10)()
        dv = (C.c_double * 3)()
        self.lib.mtpref_info(C.c_void_p(self.h), iv, dv)
        names = ["species_count", "K", "T", "M", "A", "R", "B", "P", "Q", "configuration_mode"]
        self.info = dict(zip(names, list(iv)))
        self.info.update(min_cutoff=dv[0], max_cutoff=dv[1], scaling=dv[2])

    @property
    def log(self):
        return self.lib.mtpref_log(C.c_void_p(self.h)).decode()

    def set_domain(self, prd, natoms):
        a = np.zeros(6)
        a[: len(prd)] = prd            # xprd, yprd, zprd [, xy, xz, yz]
        self.lib.mtpref_set_domain(C.c_void_p(self.h), _ptr(a, _dp), C.c_long(natoms))

    def compute(self, x, type_, nlocal, ilist, numneigh, neigh_flat, offsets, eflag=3, vflag=5, grade=False,
                want_mask=False, f_init=None) -> Result:
        x, type_, ilist, numneigh, neigh_flat, offsets = _prep(x, type_, ilist, numneigh, neigh_flat, offsets)
        nall = x.shape[0]
        r = Result(nall, self.info["Q"])
        if f_init is not None:
            r.f[:] = f_init
        if want_mask:
            assert len(ilist) == 1
            r.mask = np.zeros(int(numneigh[ilist[0]]), dtype=np.uint8)
        err = C.create_string_buffer(1024)
        rc = self.lib.mtpref_compute(
            C.c_void_p(self.h), C.c_int(nlocal), C.c_int(nall - nlocal), _ptr(x, _dp), _ptr(type_, _ip),
            C.c_int(len(ilist)), _ptr(ilist, _ip), _ptr(numneigh, _ip), _ptr(neigh_flat, _ip),
            _ptr(offsets, _lp), C.c_int(eflag), C.c_int(vflag), C.c_int(1 if grade else 0), _ptr(r.f, _dp),
            _ptr(r.eatom, _dp), _ptr(r.vatom, _dp), _ptr(r.ev, _dp), _ptr(r.grades, _dp), _ptr(r.mask, _bp),
            err, C.c_int(1024))
        if rc != 0:
            raise RuntimeError(err.value.decode())
        if grade and self.info["Q"]:
            self.lib.mtpref_candidate(C.c_void_p(self.h), _ptr(r.candidate, _dp), C.c_int(self.info["Q"]))
        return r

    def close(self):
        if self.h:
            self.lib.mtpref_destroy(C.c_void_p(self.h))
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
