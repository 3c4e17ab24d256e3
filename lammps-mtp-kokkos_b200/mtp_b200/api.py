"""ctypes binding of include/mtp_b200.h (the C ABI of the CUDA library) for tests, bench and harness.

The product's host side for LAMMPS is the C++ PairStyle in ``lammps-mtp-kokkos_b200/lammps/``; this
module is the same boundary seen from Python.  There is no fallback: if ``libmtp_b200.so`` is missing
or no sm_100 device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG), "libmtp_b200.so")

VARIANT_LARGE = 0
VARIANT_SMALL = 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_llp = C.POINTER(C.c_longlong)
_bp = C.POINTER(C.c_ubyte)


class MTPInfo(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "abi_version", "species_count", "radial_func_count", "radial_basis_size", "alpha_moment_count",
        "alpha_index_basic_count", "alpha_index_times_count", "alpha_scalar_count", "max_alpha_index_basic",
        "coeff_count", "configuration_mode", "has_selection_state", "wave_count", "chunksize", "device")] + [
        (n, C.c_double) for n in ("min_cutoff", "max_cutoff", "scaling")]


class MTPParamsHost(C.Structure):
    _fields_ = [("species_count", C.c_int), ("radial_func_count", C.c_int), ("radial_basis_size", C.c_int),
                ("alpha_moment_count", C.c_int), ("alpha_index_basic_count", C.c_int),
                ("alpha_index_times_count", C.c_int), ("alpha_scalar_count", C.c_int),
                ("min_cutoff", C.c_double), ("max_cutoff", C.c_double), ("scaling", C.c_double),
                ("radial_basis_coeffs", _dp), ("alpha_index_basic", _ip), ("alpha_index_times", _ip),
                ("alpha_moment_mapping", _ip), ("species_coeffs", _dp), ("linear_coeffs", _dp),
                ("configuration_mode", C.c_int), ("inverse_active_set", _dp)]


class MTPComputeArgs(C.Structure):
    _fields_ = [("variant", C.c_int), ("inum", C.c_int), ("nall", C.c_int),
                ("x", C.c_void_p), ("type", C.c_void_p), ("ilist", C.c_void_p), ("numneigh", C.c_void_p),
                ("neighbors", C.c_void_p), ("neigh_offsets", C.c_void_p),
                ("stride_i", C.c_longlong), ("stride_jj", C.c_longlong), ("neighmask", C.c_int),
                ("eflag", C.c_int), ("vflag", C.c_int), ("want_grade", C.c_int), ("natoms_total", C.c_longlong),
                ("f", C.c_void_p), ("eatom", C.c_void_p), ("vatom", C.c_void_p), ("ev_out", C.c_void_p),
                ("grades", C.c_void_p), ("cfg_candidate", C.c_void_p), ("within_cutoff", C.c_void_p),
                ("stream", C.c_void_p), ("max_numneigh", C.c_int), ("f_overwrite", C.c_int)]


EXPORTS = ["mtp_create_from_file", "mtp_create", "mtp_destroy", "mtp_last_error", "mtp_get_info",
           "mtp_get_tables", "mtp_set_chunksize", "mtp_compute", "mtp_synchronize", "mtp_compute_host",
           "mtp_halo_pack_x", "mtp_halo_unpack_add_f", "mtp_fp64_peak", "mtp_kernel_launch_count", "mtp_last_kernel_path", "mtp_neigh_build", "mtp_program_check",
           "mtp_nve_initial_integrate", "mtp_nve_final_integrate", "mtp_select_grades"]

_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                           "the MTP B200 path has no fallback")
    lib = C.CDLL(LIB_PATH)
    lib.mtp_create_from_file.restype = C.c_void_p
    lib.mtp_create_from_file.argtypes = [C.c_char_p, C.c_int, C.c_int]
    lib.mtp_create.restype = C.c_void_p
    lib.mtp_create.argtypes = [C.POINTER(MTPParamsHost), C.c_int]
    lib.mtp_destroy.argtypes = [C.c_void_p]
    lib.mtp_destroy.restype = None
    lib.mtp_last_error.restype = C.c_char_p
    lib.mtp_potential_check.argtypes = [C.c_char_p, C.c_int, C.POINTER(MTPInfo)]
    lib.mtp_get_info.argtypes = [C.c_void_p, C.POINTER(MTPInfo)]
    lib.mtp_get_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    lib.mtp_set_chunksize.argtypes = [C.c_void_p, C.c_int]
    lib.mtp_set_lanes.argtypes = [C.c_void_p, C.c_int]
    lib.mtp_compute.argtypes = [C.c_void_p, C.POINTER(MTPComputeArgs)]
    lib.mtp_synchronize.argtypes = [C.c_void_p]
    lib.mtp_compute_host.argtypes = [C.c_void_p, C.POINTER(MTPComputeArgs), C.c_int]
    lib.mtp_halo_pack_x.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _dp, C.c_void_p, C.c_void_p]
    lib.mtp_halo_pack_x_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.mtp_halo_unpack_add_f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.mtp_profile_enable.argtypes = [C.c_void_p, C.c_int]
    lib.mtp_profile_read.argtypes = [C.c_void_p, _dp, _llp]
    lib.mtp_fp64_peak.argtypes = [C.c_int, _dp, _dp]
    lib.mtp_kernel_launch_count.restype = C.c_longlong
    lib.mtp_last_kernel_path.argtypes = [C.c_void_p]
    lib.mtp_program_kernel_note.argtypes = [C.c_void_p]
    lib.mtp_program_kernel_note.restype = C.c_char_p
    lib.mtp_program_kernel_note_small.argtypes = [C.c_void_p]
    lib.mtp_program_kernel_note_small.restype = C.c_char_p
    lib.mtp_codegen_source.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_longlong, _llp, _llp]
    lib.mtp_codegen_prebuild.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    lib.mtp_nve_initial_integrate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                              C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    lib.mtp_select_grades.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
    lib.mtp_nve_final_integrate.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    lib.mtp_neigh_build.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int,
                                    C.POINTER(C.c_int), C.c_void_p]
    _lib = lib
    return lib


class MTPError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _check(lib, rc):
    if rc != 0:
        raise MTPError(rc, lib.mtp_last_error().decode())


class HostResult:
    def __init__(self, nall, q):
        self.f = np.zeros((nall, 3))
        self.eatom = np.zeros(nall)
        self.vatom = np.zeros((nall, 6))
        self.ev = np.zeros(8)
        self.grades = np.zeros(nall)
        self.candidate = np.zeros(max(q, 1))
        self.mask = None

    energy = property(lambda s: s.ev[0])
    virial = property(lambda s: s.ev[1:7])
    max_grade = property(lambda s: s.ev[7])


class MTPB200:
    """One loaded potential on one GPU (``mtp_handle``)."""

    def __init__(self, path: str | None = None, *, selection_state: bool = False, device: int = -1, params=None):
        self.lib = load_library()
        if path is not None:
            self.h = self.lib.mtp_create_from_file(os.fsencode(path), int(selection_state), device)
        else:
            self.h = self.lib.mtp_create(C.byref(params), device)
        if not self.h:
            raise MTPError(-2, self.lib.mtp_last_error().decode())
        self.info = MTPInfo()
        _check(self.lib, self.lib.mtp_get_info(self.h, C.byref(self.info)))

    @classmethod
    def from_potential(cls, pot, device: int = -1):
        """Through mtp_create(): tables handed over as arrays (an ``almtp.MTPPotential``)."""
        keep = dict(rc=np.ascontiguousarray(pot.radial_coeffs, dtype=np.float64),
                    basic=np.ascontiguousarray(pot.alpha_index_basic, dtype=np.int32),
                    times=np.ascontiguousarray(pot.alpha_index_times, dtype=np.int32),
                    mapping=np.ascontiguousarray(pot.alpha_moment_mapping, dtype=np.int32),
                    sc=np.ascontiguousarray(pot.species_coeffs, dtype=np.float64),
                    lc=np.ascontiguousarray(pot.moment_coeffs, dtype=np.float64))
        inv = None
        if pot.inverse_active_set is not None:
            keep["inv"] = np.ascontiguousarray(pot.inverse_active_set, dtype=np.float64)
            inv = keep["inv"].ctypes.data_as(_dp)
        cfg = int(pot.energy_weight) == 1 if pot.energy_weight is not None else 0
        p = MTPParamsHost(pot.species_count, pot.radial_funcs_count, pot.radial_basis_size, pot.alpha_moments_count,
                          pot.K, pot.T, pot.A, pot.min_dist, pot.max_dist, pot.scaling,
                          keep["rc"].ctypes.data_as(_dp), keep["basic"].ctypes.data_as(_ip),
                          keep["times"].ctypes.data_as(_ip), keep["mapping"].ctypes.data_as(_ip),
                          keep["sc"].ctypes.data_as(_dp), keep["lc"].ctypes.data_as(_dp), int(cfg), inv)
        return cls(None, device=device, params=p)

    def tables(self):
        i = self.info
        S, R, B = i.species_count, i.radial_func_count, i.radial_basis_size
        out = dict(radial=np.zeros((S, S, R, B)), basic=np.zeros((i.alpha_index_basic_count, 4), dtype=np.int32),
                   times=np.zeros((i.alpha_index_times_count, 4), dtype=np.int32),
                   mapping=np.zeros(i.alpha_scalar_count, dtype=np.int32), species=np.zeros(S),
                   linear=np.zeros(i.alpha_scalar_count))
        inv = None
        if i.has_selection_state:
            out["inverse_active_set"] = np.zeros((i.coeff_count, i.coeff_count))
            inv = out["inverse_active_set"].ctypes.data
        _check(self.lib, self.lib.mtp_get_tables(self.h, out["radial"].ctypes.data, out["basic"].ctypes.data,
                                                 out["times"].ctypes.data, out["mapping"].ctypes.data,
                                                 out["species"].ctypes.data, out["linear"].ctypes.data, inv))
        return out

    def set_chunksize(self, n: int):
        _check(self.lib, self.lib.mtp_set_chunksize(self.h, int(n)))
        self.info.chunksize = int(n)

    def set_lanes(self, n: int):
        _check(self.lib, self.lib.mtp_set_lanes(self.h, int(n)))

    # ---- host buffers ---------------------------------------------------------------------------
    def compute_host(self, x, type_, ilist, numneigh, neigh, offsets=None, *, stride_i=0, stride_jj=1, eflag=3,
                     vflag=5, grade=False, natoms_total=0, want_mask=False, variant=VARIANT_LARGE, f_init=None,
                     list_changed=True, f_overwrite=False, out: HostResult | None = None) -> HostResult:
        x = np.ascontiguousarray(x, dtype=np.float64)
        type_ = np.ascontiguousarray(type_, dtype=np.int32)
        numneigh = np.ascontiguousarray(numneigh, dtype=np.int32)
        neigh = np.ascontiguousarray(neigh, dtype=np.int32)
        il = None if ilist is None else np.ascontiguousarray(ilist, dtype=np.int32)
        off = None if offsets is None else np.ascontiguousarray(offsets, dtype=np.int64)
        nall = x.shape[0]
        inum = nall if il is None else len(il)
        r = out or HostResult(nall, self.info.coeff_count)
        if f_init is not None:
            r.f[:] = f_init
        if want_mask:
            r.mask = np.zeros(max(neigh.size, 1), dtype=np.uint8)
        a = MTPComputeArgs()
        a.variant, a.inum, a.nall = variant, inum, nall
        a.x, a.type = x.ctypes.data, type_.ctypes.data
        a.ilist = None if il is None else il.ctypes.data
        a.numneigh, a.neighbors = numneigh.ctypes.data, neigh.ctypes.data
        a.neigh_offsets = None if off is None else off.ctypes.data
        a.stride_i, a.stride_jj, a.neighmask = stride_i, stride_jj, 0x1FFFFFFF
        a.eflag, a.vflag, a.want_grade, a.natoms_total = eflag, vflag, int(grade), natoms_total
        a.f, a.eatom, a.vatom, a.ev_out = r.f.ctypes.data, r.eatom.ctypes.data, r.vatom.ctypes.data, r.ev.ctypes.data
        a.grades, a.cfg_candidate = r.grades.ctypes.data, r.candidate.ctypes.data
        a.within_cutoff = r.mask.ctypes.data if want_mask else None
        a.max_numneigh = 0          # mtp_compute_host derives it from the host numneigh array
        a.f_overwrite = int(f_overwrite)
        self._keep = (x, type_, numneigh, neigh, il, off)
        _check(self.lib, self.lib.mtp_compute_host(self.h, C.byref(a), int(list_changed)))
        return r

    def compute_system(self, sysm, **kw) -> HostResult:
        return self.compute_host(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets, **kw)

    # ---- device buffers (torch tensors as plain device memory) ----------------------------------------
    def compute_device(self, x, type_, ilist, numneigh, neigh, offsets, f, ev_out, *, eatom=None, vatom=None,
                       grades=None, cfg_candidate=None, within=None, stride_i=0, stride_jj=1, eflag=1, vflag=1,
                       grade=False, natoms_total=0, variant=VARIANT_LARGE, stream=None, max_numneigh=0):
        def ptr(t):
            return None if t is None else t.data_ptr()

        a = MTPComputeArgs()
        a.variant, a.nall = variant, x.shape[0]
        a.inum = x.shape[0] if ilist is None else ilist.shape[0]
        a.x, a.type, a.ilist, a.numneigh, a.neighbors = ptr(x), ptr(type_), ptr(ilist), ptr(numneigh), ptr(neigh)
        a.neigh_offsets = ptr(offsets)
        a.stride_i, a.stride_jj, a.neighmask = stride_i, stride_jj, 0x1FFFFFFF
        a.eflag, a.vflag, a.want_grade, a.natoms_total = eflag, vflag, int(grade), natoms_total
        a.f, a.eatom, a.vatom, a.ev_out = ptr(f), ptr(eatom), ptr(vatom), ptr(ev_out)
        a.grades, a.cfg_candidate, a.within_cutoff = ptr(grades), ptr(cfg_candidate), ptr(within)
        a.stream = stream
        a.max_numneigh = int(max_numneigh)
        _check(self.lib, self.lib.mtp_compute(self.h, C.byref(a)))

    def compute_device_phased(self, phase_inum, wait_events, done_events, x, type_, ilist, numneigh, neigh, offsets, f, ev_out, *,
                              stride_i=0, stride_jj=1, eflag=1, vflag=1, variant=VARIANT_LARGE, stream=None, max_numneigh=0):
        """mtp_compute_phased: ilist is cut into consecutive phases of phase_inum[p] centres; phase p starts after the
        torch.cuda.Event wait_events[p] (or None) and done_events[p] (or None) is recorded when it is complete."""
        a = MTPComputeArgs()
        a.variant, a.nall, a.inum = variant, x.shape[0], ilist.shape[0]
        a.x, a.type, a.ilist = x.data_ptr(), type_.data_ptr(), ilist.data_ptr()
        a.numneigh, a.neighbors = numneigh.data_ptr(), neigh.data_ptr()
        a.neigh_offsets = None if offsets is None else offsets.data_ptr()
        a.stride_i, a.stride_jj, a.neighmask = stride_i, stride_jj, 0x1FFFFFFF
        a.eflag, a.vflag = eflag, vflag
        a.f, a.ev_out = f.data_ptr(), ev_out.data_ptr()
        a.stream = stream
        a.max_numneigh = int(max_numneigh)
        n = len(phase_inum)
        cnt = (C.c_int * n)(*[int(v) for v in phase_inum])
        wv = (C.c_void_p * n)(*[None if e is None else e.cuda_event for e in wait_events])
        dv = (C.c_void_p * n)(*[None if e is None else e.cuda_event for e in done_events])
        self.lib.mtp_compute_phased.argtypes = [C.c_void_p, C.POINTER(MTPComputeArgs), C.c_int, C.POINTER(C.c_int),
                                                C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
        _check(self.lib, self.lib.mtp_compute_phased(self.h, C.byref(a), n, cnt, wv, dv))

    ERR_CAPACITY = -7

    def neigh_build(self, x, nlocal: int, cutneigh: float, width: int | None = None, stream=None):
        """Device build of the full neighbor list (mtp_neigh_build).  x: CUDA float64 tensor [nall, 3].
        Returns (numneigh int32 [nlocal], table int32 [nlocal, width], longest row); pass the table to
        compute_device(neigh=table, offsets=None, stride_i=width, stride_jj=1).  The table is rebuilt wider when a
        row does not fit (LAMMPS-KOKKOS's resize-and-retry)."""
        import torch
        nall = int(x.shape[0])
        if width is None:
            width = getattr(self, "_neigh_width", 128)
        numneigh = torch.empty(max(nlocal, 1), dtype=torch.int32, device=x.device)
        while True:
            table = torch.empty((max(nlocal, 1), width), dtype=torch.int32, device=x.device)
            mx = C.c_int(0)
            rc = self.lib.mtp_neigh_build(self.h, int(nlocal), nall, x.data_ptr(), float(cutneigh), numneigh.data_ptr(),
                                          table.data_ptr(), int(width), C.byref(mx), stream)
            if rc == self.ERR_CAPACITY:
                width = (mx.value + 7) // 8 * 8
                continue
            _check(self.lib, rc)
            self._neigh_width = width
            return numneigh[:nlocal], table[:nlocal], mx.value

    def select_grades(self, grades, threshold: float, stream=None):
        """Ascending ids of the atoms whose grade is >= threshold, selected on the device (mtp_select_grades).
        grades: CUDA float64 tensor [n]; returns a CUDA int32 tensor."""
        import torch
        n = int(grades.shape[0])
        idx = torch.empty(max(n, 1), dtype=torch.int32, device=grades.device)
        cnt = C.c_int(0)
        _check(self.lib, self.lib.mtp_select_grades(self.h, grades.data_ptr(), n, float(threshold), idx.data_ptr(), C.byref(cnt),
                                                    stream))
        return idx[: cnt.value]

    PROF_CLASSES = ("pack", "gather", "moments", "program", "forces", "grade", "finalize", "site")

    def profile_enable(self, on: bool = True):
        _check(self.lib, self.lib.mtp_profile_enable(self.h, int(on)))

    def profile_read(self):
        """{class: (ms, spans)} of device time per kernel class since the last read."""
        ms = (C.c_double * 8)()
        cnt = (C.c_longlong * 8)()
        _check(self.lib, self.lib.mtp_profile_read(self.h, ms, cnt))
        return {n: (ms[i], cnt[i]) for i, n in enumerate(self.PROF_CLASSES)}

    def last_kernel_path(self) -> dict:
        """Which kernels the last compute launched (mtp_last_kernel_path)."""
        v = int(self.lib.mtp_last_kernel_path(self.h))
        return {"family": v & 15, "program_v3": bool(v & 16), "program_generated": bool(v & 32),
                "program_atoms_per_cta": (v >> 8) & 4095}

    def program_kernel_note(self, latency_shape: bool = False) -> str:
        """Empty when the generated contraction-program kernel serves this handle, else why it does not."""
        fn = self.lib.mtp_program_kernel_note_small if latency_shape else self.lib.mtp_program_kernel_note
        return fn(self.h).decode()

    def synchronize(self):
        _check(self.lib, self.lib.mtp_synchronize(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mtp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def potential_check(path: str, selection_state: bool = False) -> MTPInfo:
    """Host-only parse + program compile of a potential file (no CUDA device needed)."""
    lib = load_library()
    info = MTPInfo()
    _check(lib, lib.mtp_potential_check(os.fsencode(path), int(selection_state), C.byref(info)))
    return info


def program_check(path: str, atoms_per_cta: int = 32) -> float:
    """Host-only check of the grouped contraction-program streams (mtp_program_check)."""
    lib = load_library()
    err = C.c_double(0.0)
    lib.mtp_program_check.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_double)]
    _check(lib, lib.mtp_program_check(os.fsencode(path), int(atoms_per_cta), C.byref(err)))
    return float(err.value)


CODEGEN_INFO = ("atoms_per_cta", "warps", "ctas_per_sm", "rows", "stages", "smem_bytes", "terms", "loads", "stores",
                "critical_terms", "nslots", "hash", "rounds")


def codegen_source(path: str, latency_shape: bool = False):
    """Source of the generated contraction-program kernel for a potential file + generator statistics (no device)."""
    lib = load_library()
    need = C.c_longlong(0)
    info = (C.c_longlong * 13)()
    _check(lib, lib.mtp_codegen_source(os.fsencode(path), int(latency_shape), None, 0, C.byref(need), info))
    buf = C.create_string_buffer(need.value)
    _check(lib, lib.mtp_codegen_source(os.fsencode(path), int(latency_shape), buf, need.value, C.byref(need), info))
    return buf.value.decode(), dict(zip(CODEGEN_INFO, [int(v) for v in info]))


def codegen_prebuild(path: str, latency_shape: bool = False) -> bool:
    """Compile the generated kernel of a potential with NVRTC into the cubin cache (no device). True if NVRTC ran."""
    lib = load_library()
    ran = C.c_int(0)
    _check(lib, lib.mtp_codegen_prebuild(os.fsencode(path), int(latency_shape), C.byref(ran)))
    return bool(ran.value)


def fp64_peaks(device: int = -1):
    lib = load_library()
    a, b = C.c_double(0), C.c_double(0)
    _check(lib, lib.mtp_fp64_peak(device, C.byref(a), C.byref(b)))
    return a.value, b.value


def kernel_launch_count() -> int:
    return int(load_library().mtp_kernel_launch_count())
