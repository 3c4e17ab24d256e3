"""Worker of tests/test_gpu_multi.py::test_decomposed_nve_conserves_energy (torch.distributed.run, NCCL, one rank per
GPU): the device-resident NVE loop on a brick decomposition -- integrate, NCCL ghost halo, device list rebuild
(collective decision), pair style, reverse halo -- conserves the all-reduced total energy."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lammps-mtp-kokkos_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from mtp_b200 import almtp, decomp  # noqa: E402
from mtp_b200.api import MTPB200  # noqa: E402
from mtp_b200.md import NVE, scale_to_rms_force  # noqa: E402


def main():
    out, tmp = sys.argv[1], sys.argv[2]
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    grid = decomp.brick_grid(world)
    pot0 = almtp.random_potential(10, 2)
    path0 = os.path.join(tmp, f"raw{rank}.almtp")
    almtp.write_almtp(path0, pot0)
    mtp0 = MTPB200(path0, device=local)
    sysm, halo = decomp.make_rank_system(2, (6, 6, 6), grid, rank, dev, mtp0.lib, direct=True)
    # RMS force of the raw potential over all ranks -> one global scale factor
    probe = NVE(mtp0, sysm, halo, masses=[183.84, 95.95], temperature=0.0)
    s = torch.stack([(probe.f[: sysm.nlocal] ** 2).sum(), torch.tensor(float(sysm.nlocal), device=dev, dtype=torch.float64)])
    dist.all_reduce(s)
    rms = float((s[0] / s[1]).sqrt())
    mtp0.close()
    pot = scale_to_rms_force(pot0, rms, 0.05)
    path = os.path.join(tmp, f"scaled{rank}.almtp")
    almtp.write_almtp(path, pot)
    mtp = MTPB200(path, device=local)
    md = NVE(mtp, sysm, halo, masses=[183.84, 95.95], dt=0.001, temperature=50.0, seed=12345 + rank, rebuild_trigger=0.05)

    def total():
        e = torch.tensor([md.potential_energy() + md.kinetic_energy(), float(sysm.nlocal)], dtype=torch.float64, device=dev)
        dist.all_reduce(e)
        return float(e[0]), int(e[1])

    e0, natoms = total()
    es = []
    for _ in range(4):
        md.run(20)
        es.append(total()[0])
    ke = torch.tensor([md.kinetic_energy()], dtype=torch.float64, device=dev)
    dist.all_reduce(ke)
    if rank == 0:
        np.savez(out, e0=e0, es=np.array(es), natoms=natoms, ke=float(ke[0]), rebuilds=md.rebuilds, launches=halo.launches)
    dist.barrier()
    mtp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
