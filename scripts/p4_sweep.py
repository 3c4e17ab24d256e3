"""Sweep generator parameters of the contraction-program kernel on one GPU (development tool).
usage: python scripts/p4_sweep.py CONFIG "na,warps,cache,acc,minb" ...     ("default" = the library's own choice)
Prints per-kernel-class ms per force evaluation (kernels serialised, lanes = 1, L2 flushed between evaluations)."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lammps-mtp-kokkos_b200"))

import numpy as np
import torch
from mtp_b200 import almtp, api, harness
from mtp_b200.api import MTPB200

config = int(sys.argv[1])
cells = None
specs = sys.argv[2:]
if specs and specs[0].startswith("cells="):
    cells = tuple(int(v) for v in specs[0][6:].split("x"))
    specs = specs[1:]
cfg = harness.CONFIGS[config]
td = tempfile.mkdtemp()
path = os.path.join(td, "p.almtp")
almtp.write_almtp(path, almtp.random_potential(cfg["level"], cfg["species"]))
sysm = harness.make_config(config, cells=cells)
dev = torch.device("cuda", 0)
t = {k: torch.from_numpy(getattr(sysm, k)).to(dev) for k in ("x", "type", "ilist", "numneigh", "neigh", "offsets")}
f = torch.zeros((sysm.nall, 3), dtype=torch.float64, device=dev)
ev = torch.zeros(8, dtype=torch.float64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
mx = int(sysm.numneigh[: sysm.nlocal].max())
e_ref = None
for spec in specs:
    envs = {}
    for kv in spec.split(";"):
        if "=" in kv:
            k, v = kv.split("=", 1)
            envs[k] = v
        elif kv != "default":
            envs["MTP_B200_P4"] = kv
    for k in ("MTP_B200_P4", "MTP_B200_NO_P4"):
        os.environ.pop(k, None)
    os.environ.update(envs)
    mtp = MTPB200(path)
    mtp.set_chunksize(int(os.environ.get("SWEEP_CHUNK", "131072")))
    mtp.set_lanes(1)

    def step():
        f.zero_()
        mtp.compute_device(t["x"], t["type"], t["ilist"], t["numneigh"], t["neigh"], t["offsets"], f, ev, eflag=1, vflag=1,
                           max_numneigh=mx)
    for _ in range(3):
        step()
    mtp.profile_enable(True)
    n = 5
    for s in range(n):
        flush.fill_(s)
        step()
    prof = mtp.profile_read()
    mtp.profile_enable(False)
    torch.cuda.synchronize()
    e = float(ev[0])
    e_ref = e if e_ref is None else e_ref
    print(f"{spec:40s} note='{mtp.program_kernel_note()}' path={mtp.last_kernel_path()} dE={abs(e - e_ref) / abs(e_ref):.1e} " +
          " ".join(f"{nm}={v[0] / n:.4f}" for nm, v in prof.items() if v[0] > 0), flush=True)
    mtp.close()
    for k in envs:
        os.environ.pop(k, None)
