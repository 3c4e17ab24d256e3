"""Mini MD harness standing in for the parts of upstream LAMMPS around the pair style
(lattice creation, ghost atoms, full neighbor list, reverse communication of ghost
forces).  Upstream LAMMPS is not in this image; the pair style
itself only sees what LAMMPS would hand it: ``x[nall][3]``, ``type[nall]`` (1-based),
``ilist``, ``numneigh``, the neighbor rows, and it accumulates into ``f[nall][3]``
including ghost rows (pair_mtp.cpp:77-88,248-254).

Configs of BASELINE.json are reproduced by ``make_config``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_HARNESS_DIR = os.path.join(os.path.dirname(_HERE), "csrc", "harness")
_HARNESS_SO = os.path.join(_HARNESS_DIR, "libmtp_harness.so")


def build_harness() -> str:
    src = os.path.join(_HARNESS_DIR, "neigh_host.c")
    if (not os.path.exists(_HARNESS_SO)) or os.path.getmtime(_HARNESS_SO) < os.path.getmtime(src):
        subprocess.run(["gcc", "-O2", "-fopenmp", "-fPIC", "-shared", "-o", _HARNESS_SO, src, "-lm"], check=True)
    return _HARNESS_SO


_lib = None


def _harness():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_harness())
        _lib.mtp_harness_neigh_count.restype = C.c_long
    return _lib


_BASES = {
    "sc": np.array([[0.0, 0.0, 0.0]]),
    "bcc": np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]]),
    "fcc": np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]]),
    "diamond": np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5],
                         [0.25, 0.25, 0.25], [0.75, 0.75, 0.25], [0.75, 0.25, 0.75], [0.25, 0.75, 0.75]]),
}


@dataclass
class System:
    """What LAMMPS hands the pair style for one rank."""
    box: np.ndarray          # [3] orthorhombic periodic box lengths (lo = 0)
    nlocal: int
    x: np.ndarray            # [nall, 3] owned atoms first, then ghosts
    type: np.ndarray         # [nall] int32, 1-based
    owner: np.ndarray        # [nall] int32: owned-atom index each row is an image of
    ilist: np.ndarray        # [inum] int32
    numneigh: np.ndarray     # [nall] int32
    offsets: np.ndarray      # [nall + 1] int64 CSR row starts
    neigh: np.ndarray        # [npairs] int32
    rlist: float

    @property
    def nall(self):
        return int(self.x.shape[0])

    @property
    def inum(self):
        return int(self.ilist.shape[0])

    def reverse_comm(self, f: np.ndarray) -> np.ndarray:
        """LAMMPS Comm::reverse_comm for newton on: add ghost rows of f onto their owners."""
        out = f[: self.nlocal].copy()
        np.add.at(out, self.owner[self.nlocal:], f[self.nlocal:])
        return out

    def padded_neighbors(self, width: int | None = None):
        """Row-major [nall_rows, width] 2-D neighbor table (LAMMPS-KOKKOS d_neighbors shape)."""
        w = int(self.numneigh.max()) if width is None else width
        tab = np.zeros((self.nlocal, w), dtype=np.int32)
        for i in range(self.nlocal):
            n = self.numneigh[i]
            tab[i, :n] = self.neigh[self.offsets[i]: self.offsets[i] + n]
        return tab


def lattice(kind: str, a: float, cells, jitter: float = 0.0, seed: int = 2024):
    cells = np.asarray(cells, dtype=np.int64)
    base = _BASES[kind]
    gx, gy, gz = np.meshgrid(np.arange(cells[0]), np.arange(cells[1]), np.arange(cells[2]), indexing="ij")
    origin = np.stack([gx, gy, gz], axis=-1).reshape(-1, 1, 3).astype(np.float64)
    x = ((origin + base[None, :, :]) * a).reshape(-1, 3)
    box = cells.astype(np.float64) * a
    if jitter:
        rng = np.random.default_rng(seed)
        x = x + rng.uniform(-jitter, jitter, size=x.shape)
    x = np.mod(x, box)    # wrap into [0, box)
    return np.ascontiguousarray(x), box


def add_ghosts(x: np.ndarray, box: np.ndarray, rghost: float, periodic=(True, True, True)):
    """Periodic images within rghost of the box faces, dimension by dimension (so edge and
    corner images come for free), like LAMMPS's six-swap forward communication."""
    pos = x
    owner = np.arange(x.shape[0], dtype=np.int32)
    for d in range(3):
        if not periodic[d]:
            continue
        if rghost >= box[d]:
            raise ValueError("ghost cutoff exceeds the box length; use a larger box")
        lo = pos[:, d] < rghost
        hi = pos[:, d] >= box[d] - rghost
        shift = np.zeros(3)
        shift[d] = box[d]
        pos_new = [pos, pos[lo] + shift, pos[hi] - shift]
        own_new = [owner, owner[lo], owner[hi]]
        pos = np.concatenate(pos_new)
        owner = np.concatenate(own_new)
    return np.ascontiguousarray(pos), owner


def neighbor_list(x_all: np.ndarray, nlocal: int, rlist: float):
    lib = _harness()
    nall = x_all.shape[0]
    x_all = np.ascontiguousarray(x_all, dtype=np.float64)
    numneigh = np.zeros(nall, dtype=np.int32)
    dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_long)
    tot = lib.mtp_harness_neigh_count(C.c_int(nlocal), C.c_int(nall), x_all.ctypes.data_as(dp), C.c_double(rlist),
                                      numneigh.ctypes.data_as(ip))
    if tot < 0:
        raise MemoryError("neighbor list build failed")
    offsets = np.zeros(nall + 1, dtype=np.int64)
    np.cumsum(numneigh, out=offsets[1:])
    flat = np.empty(max(int(tot), 1), dtype=np.int32)
    lib.mtp_harness_neigh_fill(C.c_int(nlocal), C.c_int(nall), x_all.ctypes.data_as(dp), C.c_double(rlist),
                               offsets.ctypes.data_as(lp), flat.ctypes.data_as(ip))
    return numneigh, offsets, flat[: int(tot)]


def make_system(x: np.ndarray, types: np.ndarray, box, cutoff: float, skin: float = 2.0,
                periodic=(True, True, True)) -> System:
    box = np.asarray(box, dtype=np.float64)
    nlocal = x.shape[0]
    rlist = cutoff + skin
    x_all, owner = add_ghosts(x, box, rlist, periodic)
    t_all = np.ascontiguousarray(np.asarray(types, dtype=np.int32)[owner])
    numneigh, offsets, flat = neighbor_list(x_all, nlocal, rlist)
    return System(box=box, nlocal=nlocal, x=x_all, type=t_all, owner=owner,
                  ilist=np.arange(nlocal, dtype=np.int32), numneigh=numneigh, offsets=offsets, neigh=flat,
                  rlist=rlist)


def random_types(n: int, fractions, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    fr = np.asarray(fractions, dtype=np.float64)
    return (rng.choice(len(fr), size=n, p=fr / fr.sum()) + 1).astype(np.int32)


# BASELINE.json configs.  ``scale`` shrinks the cell counts for tests.
CONFIGS = {
    1: dict(name="fcc Al 4000 atoms, level 10, S=1", kind="fcc", a=4.05, cells=(10, 10, 10), level=10,
            species=1, fractions=(1.0,), type_seed=0, masses=(26.9815,)),
    2: dict(name="bcc W/Mo 262144 atoms, level 16, S=2", kind="bcc", a=3.165, cells=(64, 64, 32), level=16,
            species=2, fractions=(0.5, 0.5), type_seed=7, masses=(183.84, 95.95)),
    3: dict(name="diamond Si 2000 atoms, level 20, S=1", kind="diamond", a=5.431, cells=(5, 5, 10), level=20,
            species=1, fractions=(1.0,), type_seed=0, masses=(28.0855,)),
    4: dict(name="fcc Al-Cu 256000 atoms, level 16, S=2, active set", kind="fcc", a=4.05, cells=(40, 40, 40),
            level=16, species=2, fractions=(0.95, 0.05), type_seed=11, active_set=True, masses=(26.9815, 63.546)),
    5: dict(name="fcc CoCrNi 4M atoms/GPU, level 22, S=3", kind="fcc", a=3.56, cells=(100, 100, 100), level=22,
            species=3, fractions=(1, 1, 1), type_seed=5, masses=(58.9332, 51.9961, 58.6934)),
}


def make_config(idx: int, cells=None, cutoff: float = 5.0, skin: float = 2.0, jitter: float = 0.05) -> System:
    cfg = CONFIGS[idx]
    x, box = lattice(cfg["kind"], cfg["a"], cells or cfg["cells"], jitter=jitter, seed=2024)
    types = random_types(x.shape[0], cfg["fractions"], cfg["type_seed"])
    return make_system(x, types, box, cutoff, skin)
