#!/usr/bin/env python
"""Diagnostic: per-phase clock shares of the site kernel (needs the -DMTP_PHASE_CLOCKS build).

    python profiles/phase_clocks.py build          # nvcc ... -DMTP_PHASE_CLOCKS -> libmtp_b200_prof.so
    python profiles/phase_clocks.py [config] [cx cy cz]
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lammps-mtp-kokkos_b200"))
sys.path.insert(0, ROOT)


def build_prof():
    import __graft_entry__ as g
    csrc = os.path.join(g.PKG, "csrc")
    out = os.path.join(g.PKG, "libmtp_b200_prof.so")
    subprocess.run(["nvcc"] + g.NVCC_FLAGS + ["-DMTP_PHASE_CLOCKS", "-o", out, os.path.join(csrc, "mtp_api.cu"),
                    os.path.join(csrc, "mtp_potential.cpp")], check=True)
    return out


def main():
    from mtp_b200 import almtp, api, harness
    api.LIB_PATH = os.path.join(ROOT, "lammps-mtp-kokkos_b200", "libmtp_b200_prof.so")
    cfg_idx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    cells = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (32, 32, 32)
    cfg = harness.CONFIGS[cfg_idx]
    pot = almtp.random_potential(cfg["level"], cfg["species"])
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "p.almtp")
        almtp.write_almtp(path, pot)
        mtp = api.MTPB200(path)
    sysm = harness.make_config(cfg_idx, cells=cells)
    lib = api.load_library()
    buf = (C.c_ulonglong * 8)()
    mtp.compute_system(sysm, eflag=1, vflag=1)
    lib.mtp_debug_phase_clocks(buf)
    mtp.compute_system(sysm, eflag=1, vflag=1)
    lib.mtp_debug_phase_clocks(buf)
    names = ["sweep0 gather+radial+DMMA", "moment write-out", "program forward", "energy + adjoint seed",
             "program reverse", "canonical adjoints", "sweep1 gather+radial+Horner+scatter", "atom epilogue"]
    tot = float(sum(buf))
    print(f"config {cfg_idx} cells {cells}: {sysm.nlocal} atoms; warp-clocks per atom and share")
    for n, v in zip(names, buf):
        print(f"  {n:38s} {v / sysm.nlocal:10.0f} {100 * v / tot:6.1f}%")
    print(f"  {'total':38s} {tot / sysm.nlocal:10.0f}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build_prof()
    else:
        main()
