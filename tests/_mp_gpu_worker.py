"""Worker of tests/test_gpu_multi.py (torch.distributed.run, NCCL, one rank per GPU): brick + device halo
pack -> NCCL send/recv -> CUDA kernels -> reverse halo -> all-reduce, compared on rank 0 with the same global
system evaluated by the CPU oracle in one piece."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lammps-mtp-kokkos_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from mtp_b200 import almtp, decomp, harness  # noqa: E402
from mtp_b200.api import MTPB200  # noqa: E402


def main():
    out, tmp = sys.argv[1], sys.argv[2]
    kind = sys.argv[3] if len(sys.argv) > 3 else "staged"
    direct = kind in ("direct", "overlap")
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    grid = decomp.brick_grid(world)
    cells = (7, 7, 7) if kind == "overlap" else (5, 5, 5)
    pot = almtp.random_potential(10, 2)
    path = os.path.join(tmp, f"p{rank}.almtp")
    almtp.write_almtp(path, pot)
    mtp = MTPB200(path, device=local)
    sysm, halo = decomp.make_rank_system(2, cells, grid, rank, dev, mtp.lib, direct=direct)
    nlocal, nall = sysm.nlocal, sysm.nall
    disp = np.random.default_rng(99).uniform(-0.05, 0.05, size=(nlocal, 3))
    x = torch.from_numpy(sysm.x.copy()).to(dev)
    x[:nlocal] += torch.from_numpy(disp).to(dev)
    t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    t_type, t_il, t_nn, t_ne, t_off = t(sysm.type), t(sysm.ilist), t(sysm.numneigh), t(sysm.neigh), t(sysm.offsets)
    f = torch.zeros((nall, 3), dtype=torch.float64, device=dev)
    ev = torch.zeros(8, dtype=torch.float64, device=dev)
    eatom = torch.zeros(nall, dtype=torch.float64, device=dev)
    if kind == "overlap":
        # halo exchange hidden behind the interior atoms: three partial evaluations, one reduction
        ov = decomp.OverlappedStep(halo, sysm.x[:nlocal], halo.sublo, halo.subhi, halo.rghost, dev, min_part=1)
        assert ov.enabled == (world > 1) and (world == 1 or min(ov.counts) > 0), ov.counts

        def part(il, evbuf):
            mtp.compute_device(x, t_type, il, t_nn, t_ne, t_off, f, evbuf, eatom=eatom, eflag=3, vflag=1,
                               stream=torch.cuda.current_stream().cuda_stream)
        def phased(il, counts, waits, dones, evbuf):
            mtp.compute_device_phased(counts, waits, dones, x, t_type, il, t_nn, t_ne, t_off, f, evbuf, eflag=1, vflag=1,
                                      stream=torch.cuda.current_stream().cuda_stream)
        f.fill_(7.0)      # run() zeroes f itself
        ov.run(x, f, ev, part, t_il, compute_phased=phased if world > 1 else None)
        if world > 1:      # per-atom energies are not part of the phased call: one plain evaluation for them
            f2, ev2 = torch.zeros_like(f), torch.zeros_like(ev)
            mtp.compute_device(x, t_type, t_il, t_nn, t_ne, t_off, f2, ev2, eatom=eatom, eflag=3, vflag=1,
                               stream=torch.cuda.current_stream().cuda_stream)
    else:
        halo.forward(x)
        mtp.compute_device(x, t_type, t_il, t_nn, t_ne, t_off, f, ev, eatom=eatom, eflag=3, vflag=1,
                           stream=torch.cuda.current_stream().cuda_stream)
        halo.reverse(f)
        halo.allreduce_ev(ev)
    mtp.synchronize()
    gathered = [None] * world
    dist.all_gather_object(gathered, (x[:nlocal].cpu().numpy(), sysm.type[:nlocal], f[:nlocal].cpu().numpy(),
                                      eatom[:nlocal].cpu().numpy()))
    if rank == 0:
        from oracle_py import OracleMTP
        cfg = harness.CONFIGS[2]
        _, box = harness.lattice(cfg["kind"], cfg["a"], cells)
        gbox = box * np.array(grid)
        gx = np.concatenate([g[0] for g in gathered])
        gt = np.concatenate([g[1] for g in gathered])
        gs = harness.make_system(np.mod(gx, gbox), gt, gbox, 5.0, 2.0)
        ref = OracleMTP(pot).compute(gs.x, gs.type, gs.ilist, gs.numneigh, gs.neigh, gs.offsets, eflag=3, vflag=5)
        np.savez(out, f=np.concatenate([g[2] for g in gathered]), fref=gs.reverse_comm(ref.f), ev=ev.cpu().numpy(),
                 evref=ref.ev, eatom=np.concatenate([g[3] for g in gathered]), eatomref=ref.eatom[: gs.nlocal],
                 launches=np.array([halo.launches]))
    dist.barrier()
    mtp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
