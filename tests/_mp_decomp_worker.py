"""Worker of tests/test_decomp_gloo.py (run under torch.distributed.run, gloo, CPU).

Every rank builds its brick + halo, displaces its owned atoms, runs forward halo -> pair compute (the CPU
oracle stands in for the CUDA kernels: tests may use it as the checker) -> reverse halo -> all-reduce, and
rank 0 compares against the same global system evaluated in one piece.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lammps-mtp-kokkos_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from mtp_b200 import almtp, decomp, harness  # noqa: E402
from oracle_py import OracleMTP  # noqa: E402


def main():
    out = sys.argv[1]
    kind = sys.argv[2] if len(sys.argv) > 2 else "staged"
    direct = kind in ("direct", "overlap")
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    grid = decomp.brick_grid(world)
    dev = torch.device("cpu")
    cells = (7, 7, 7) if kind == "overlap" else (5, 5, 5)
    pot = almtp.random_potential(10, 2, with_active_set=(kind == "cfg"), configuration_mode=(kind == "cfg"))
    sysm, halo = decomp.make_rank_system(2, cells, grid, rank, dev, direct=direct)
    nlocal = sysm.nlocal
    # move the owned atoms after setup (same displacement field in every brick keeps the global reference simple)
    disp = np.random.default_rng(99).uniform(-0.05, 0.05, size=(nlocal, 3))
    x = torch.from_numpy(sysm.x.copy())
    x[:nlocal] += torch.from_numpy(disp)
    stale_ghosts = x[nlocal:].clone()
    halo.forward(x)
    assert not torch.equal(stale_ghosts, x[nlocal:])
    orc = OracleMTP(pot)
    cfg_grade = None
    if kind == "cfg":
        # configuration mode: per-rank candidate vectors summed over ranks, grade from the sum
        natoms = nlocal * world
        rr = orc.compute(x.numpy(), sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets, eflag=1, vflag=0,
                         grade=True, natoms_total=natoms)
        cand = torch.from_numpy(rr.candidate.copy())
        ainv = torch.from_numpy(np.ascontiguousarray(pot.inverse_active_set))
        cfg_grade = float(decomp.allreduce_cfg_grade(cand, ainv, natoms))
    if kind == "overlap":
        # the split-phase exchange with the interior / boundary partition (the oracle stands in for the kernels)
        x[nlocal:] = stale_ghosts
        ov = decomp.OverlappedStep(halo, sysm.x[:nlocal], halo.sublo, halo.subhi, halo.rghost, dev, min_part=1)
        assert ov.enabled and min(ov.counts) > 0 and sum(ov.counts) == nlocal, ov.counts
        f = torch.zeros((sysm.nall, 3), dtype=torch.float64)
        ev = torch.zeros(8, dtype=torch.float64)
        eatom = np.zeros(sysm.nall)

        def part(il, evbuf):
            ids = il.numpy()
            if il is ov.parts[0]:      # interior atoms must not need the halo: evaluate them on the STALE ghosts
                xs = x.numpy().copy()
                xs[nlocal:] = stale_ghosts.numpy()
            else:
                xs = x.numpy()
            r = orc.compute(xs, sysm.type, ids, sysm.numneigh, sysm.neigh, sysm.offsets, eflag=3, vflag=5)
            f.add_(torch.from_numpy(r.f))
            evbuf.copy_(torch.from_numpy(r.ev))
            eatom[ids] = r.eatom[ids]
        ov.run(x, f, ev, part, torch.from_numpy(sysm.ilist))

        class R:      # what the gather below expects
            pass
        r = R()
        r.eatom = eatom
    else:
        r = orc.compute(x.numpy(), sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets, eflag=3, vflag=5)
        f = torch.from_numpy(r.f.copy())
        halo.reverse(f)
        ev = torch.from_numpy(r.ev.copy())
        halo.allreduce_ev(ev)
    gathered = [None] * world
    dist.all_gather_object(gathered, (x[:nlocal].numpy(), sysm.type[:nlocal], f[:nlocal].numpy(), r.eatom[:nlocal]))
    if rank == 0:
        cfg = harness.CONFIGS[2]
        _, box = harness.lattice(cfg["kind"], cfg["a"], cells)
        gbox = box * np.array(grid)
        gx = np.concatenate([g[0] for g in gathered])
        gt = np.concatenate([g[1] for g in gathered])
        gs = harness.make_system(np.mod(gx, gbox), gt, gbox, 5.0, 2.0)
        ref = orc.compute(gs.x, gs.type, gs.ilist, gs.numneigh, gs.neigh, gs.offsets, eflag=3, vflag=5, grade=(kind == "cfg"))
        fref = gs.reverse_comm(ref.f)
        np.savez(out, f=np.concatenate([g[2] for g in gathered]), fref=fref, ev=ev.numpy(), evref=ref.ev,
                 eatom=np.concatenate([g[3] for g in gathered]), eatomref=ref.eatom[: gs.nlocal],
                 ghosts=np.array([sysm.nall - nlocal]), halo_bytes=np.array([halo.bytes_per_step]),
                 cfg=np.array([cfg_grade if cfg_grade is not None else 0.0, ref.ev[7]]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
