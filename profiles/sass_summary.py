#!/usr/bin/env python
"""SASS evidence for the kernels of the step: opcode histogram per kernel from `cuobjdump -sass` of libmtp_b200.so and of a
generated program cubin.  usage: python profiles/sass_summary.py > profiles/r2_sass_summary.txt"""
import glob
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["mtp_gather_radial_kernelILi4ELi2ELi3E", "mtp_moments_v2ILi6E", "mtp_forces_v2ILi6ELi32ELb0E", "grade_dmma_reg_kernel",
        "neigh_build_kernel", "mtp_program_v3ILi32ELb0ELb1E"]
NOTE = {"LDG.E.ENL2.256": "256-bit global load (sm_100): one request per 32-byte position record",
        "LDGSTS": "cp.async (asynchronous global -> shared copies)", "DMMA": "FP64 tensor-core MMA (mma.sync.m8n8k4.f64)",
        "RED": "fire-and-forget global atomics (force scatter)", "DFMA": "FP64 fused multiply-add"}


def histogram(path, want):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, hist = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1) if any(w in m.group(1) for w in want) else None
            if cur:
                hist[cur] = Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if cur and m:
            hist[cur][m.group(1)] += 1
    return hist


def show(name, h):
    tot = sum(h.values())
    base = Counter()
    for k, v in h.items():
        base[k.split(".")[0]] += v
    print(f"## {name}\n   {tot} SASS instructions; " + ", ".join(f"{k} {v}" for k, v in base.most_common(14)))
    for key, why in NOTE.items():
        n = sum(v for k, v in h.items() if k.startswith(key))
        if n:
            print(f"   {key:16s} x{n:<6d} {why}")


so = os.path.join(ROOT, "lammps-mtp-kokkos_b200", "libmtp_b200.so")
print("# cuobjdump -sass of", os.path.relpath(so, ROOT), "(sm_100a)")
for name, h in histogram(so, WANT).items():
    show(name, h)
sys.path.insert(0, os.path.join(ROOT, "lammps-mtp-kokkos_b200"))
import tempfile                                    # noqa: E402
from mtp_b200 import almtp, api                    # noqa: E402
os.environ["MTP_B200_KCACHE"] = tempfile.mkdtemp()
with tempfile.TemporaryDirectory() as td:
    p = os.path.join(td, "L16.almtp")
    almtp.write_almtp(p, almtp.random_potential(16, 2))
    api.codegen_prebuild(p)
    cubin = glob.glob(os.path.join(os.environ["MTP_B200_KCACHE"], "*.cubin"))[0]
    print("\n# generated contraction-program kernel, level 16 (NVRTC cubin):", os.path.basename(cubin))
    for name, h in histogram(cubin, ["mtp_program_p4", "p4_r0_s"]).items():
        if name == "mtp_program_p4" or name.endswith("w0_0"):
            show(name, h)
    allh = Counter()
    for name, h in histogram(cubin, ["p4_r0_s"]).items():
        allh.update(h)
    show("all stage functions of one 32-atom chunk", allh)
