#!/usr/bin/env python
"""Generates tests/golden/*.npz from the REFERENCE ITSELF.

The reference ships no tests or golden vectors, so the pins are outputs of
its own unmodified CPU sources (LAMMPS/ML-MTP/pair_mtp.cpp, pair_mtp_extrapolation.cpp,
mtp_radial_basis.cpp, mtp_rb_chevbyshev_basis.cpp) compiled against oracle/lammps_shim/ into
oracle/_ref/libmtp_ref.so (`make -C oracle ref`; needs /root/reference, i.e. the build container).

    python tests/golden/make_golden.py          # rewrites every fixture

Each fixture holds the exact potential FILE bytes, the exact inputs LAMMPS would hand the pair style
(x, type, ilist, numneigh, CSR neighbor rows) and what the reference returned (energy, virial, f,
eatom, vatom, within_cutoff mask, grades, candidate vector).  The fixtures travel to the GPU box;
/root/reference does not.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "lammps-mtp-kokkos_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import oracle_py  # noqa: E402
from mtp_b200 import almtp, harness  # noqa: E402

# name, level, species, kind, a, cells, grade mode (None | "nbh" | "cfg"), centres kept in ilist
CASES = [
    ("L08_S1_fcc", 8, 1, "fcc", 4.05, (4, 4, 4), None, 64),
    ("L10_S1_fcc", 10, 1, "fcc", 4.05, (4, 4, 4), None, 64),          # config 1, shrunk
    ("L16_S2_bcc", 16, 2, "bcc", 3.165, (5, 5, 5), None, 48),         # config 2, shrunk
    ("L20_S1_dia", 20, 1, "diamond", 5.431, (3, 3, 3), None, 24),     # config 3, shrunk
    ("L22_S3_fcc", 22, 3, "fcc", 3.56, (4, 4, 4), None, 16),          # config 5, shrunk
    ("L12_S3_cluster", 12, 3, "cluster", 0.0, (0, 0, 0), None, 63),   # ragged / empty neighborhoods
    # grade cases keep the identity ilist: the reference sizes its grade array by inum but indexes it by atom
    # id (pair_mtp_extrapolation.cpp:91-94,335), which is only in bounds for LAMMPS's own ilist = 0..inum-1
    ("L10_S2_nbh", 10, 2, "fcc", 4.05, (4, 4, 4), "nbh", 0),          # config 4 semantics, shrunk
    ("L10_S2_cfg", 10, 2, "bcc", 3.165, (5, 5, 5), "cfg", 0),
]


def make_case(name, level, species, kind, a, cells, mode, ncentres, outdir):
    import util
    pot = almtp.random_potential(level, species, with_active_set=mode is not None, configuration_mode=mode == "cfg")
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, name + ".almtp")
        almtp.write_almtp(path, pot)
        pot_bytes = np.frombuffer(open(path, "rb").read(), dtype=np.uint8)
        if kind == "cluster":
            sysm = util.random_cluster(60, species)
        else:
            sysm = util.small_system(kind, a, cells, species, seed=11)
        rng = np.random.default_rng(5)
        if ncentres:
            ilist = np.sort(rng.choice(sysm.nlocal, size=min(ncentres, sysm.nlocal), replace=False)).astype(np.int32)
        else:
            ilist = np.arange(sysm.nlocal, dtype=np.int32)
        # keep only the rows of the chosen centres (CSR re-packed) so that the fixture stays small
        nn = np.zeros_like(sysm.numneigh)
        nn[ilist] = sysm.numneigh[ilist]
        off = np.zeros(sysm.nall + 1, dtype=np.int64)
        np.cumsum(nn, out=off[1:])
        neigh = np.concatenate([sysm.neigh[sysm.offsets[i]: sysm.offsets[i] + sysm.numneigh[i]] for i in ilist]
                               + [np.zeros(0, np.int32)]).astype(np.int32)
        neigh[::5] |= (1 << 30)      # special-bond bits, cleared by NEIGHMASK (pair_mtp.cpp:114)

        if mode is None:
            ref = oracle_py.ReferenceMTP("mtp", path)
        else:
            ref = oracle_py.ReferenceMTP("mtp/extrapolation", path)
        ref.set_domain(sysm.box, len(ilist))
        r = ref.compute(sysm.x, sysm.type, sysm.nlocal, ilist, nn, neigh, off, eflag=3, vflag=5, grade=mode is not None)
        # within_cutoff mask: the reference keeps one row at a time (pair_mtp.h:75) -> one centre per call
        mask = np.zeros(neigh.size, dtype=np.uint8)
        for i in ilist:
            one = ref.compute(sysm.x, sysm.type, sysm.nlocal, np.array([i], np.int32), nn, neigh, off, eflag=0, vflag=0,
                              want_mask=True)
            mask[off[i]: off[i] + nn[i]] = one.mask
        ref.close()
        # MLIP-3 style run (output file + thresholds): the text block the reference's write_config() emits
        cfg_text = np.zeros(0, dtype=np.uint8)
        if mode is not None:
            out = os.path.join(td, "preselected.cfg")
            ref2 = oracle_py.ReferenceMTP("mtp/extrapolation", path, out, "0.0", "1e300")
            ref2.set_domain(sysm.box, len(ilist))
            ref2.compute(sysm.x, sysm.type, sysm.nlocal, ilist, nn, neigh, off, eflag=1, vflag=0, grade=False)
            ref2.close()
            cfg_text = np.frombuffer(open(out, "rb").read(), dtype=np.uint8)
        np.savez_compressed(
            os.path.join(outdir, name + ".npz"), potential=pot_bytes, mode=np.array(mode or ""),
            x=sysm.x, type=sysm.type, nlocal=np.array(sysm.nlocal), box=sysm.box, ilist=ilist, numneigh=nn,
            offsets=off, neigh=neigh, energy=np.array(r.energy), virial=r.virial.copy(), f=r.f, eatom=r.eatom,
            vatom=r.vatom, mask=mask, grades=r.grades, max_grade=np.array(r.max_grade), candidate=r.candidate,
            cfg_text=cfg_text)
        print(f"{name}: nall={sysm.nall} centres={len(ilist)} pairs={neigh.size} in-cutoff={int(mask.sum())} "
              f"E={r.energy:.12g} max|F|={np.abs(r.f).max():.6g} max_grade={r.max_grade:.6g}")


def main():
    oracle_py.build(ref=True)
    for case in CASES:
        make_case(*case, outdir=HERE)


if __name__ == "__main__":
    main()
