"""Device-resident NVE loop around the pair style: the reference's own example deck (`velocity all create ...`,
`fix 1 all nve`, README.md:148-149 of the reference) with every per-step stage on the GPU --

    FixNVE::initial_integrate  ->  mtp_nve_initial_integrate   (v += dt/2 f/m, x += dt v, displacement check)
    Comm::forward_comm         ->  halo.forward(x)              (ghost images, device pack [+ NCCL])
    Neighbor::build (if moved) ->  mtp_neigh_build              (every `check_every` steps when an atom moved > skin/2)
    Pair::compute              ->  mtp_compute
    Comm::reverse_comm         ->  halo.reverse(f)
    FixNVE::final_integrate    ->  mtp_nve_final_integrate

so positions, velocities, forces and the neighbor list never cross PCIe.  The
harness keeps its ghost set fixed (no atom migration): a run ends with an error if an atom travels farther than half
the skin from where the ghosts were created.  Metal units, as in the reference's decks.
"""
from __future__ import annotations

import numpy as np
import torch

from . import api

FTM2V = 1.0 / 1.0364269e-4      # LAMMPS metal units: force->ftm2v
MVV2E = 1.0364269e-4            # force->mvv2e
BOLTZ = 8.617343e-5             # force->boltz, eV/K


def maxwell_velocities(types: np.ndarray, masses: np.ndarray, temperature: float, seed: int) -> np.ndarray:
    """`velocity all create T seed mom yes`: Gaussian velocities, zero total momentum, rescaled to exactly T."""
    rng = np.random.default_rng(seed)
    m = masses[types]
    v = rng.normal(size=(types.shape[0], 3)) / np.sqrt(m)[:, None]
    v -= (m[:, None] * v).sum(axis=0) / m.sum()
    n = types.shape[0]
    dof = max(3 * n - 3, 1)
    t_now = MVV2E * (m[:, None] * v * v).sum() / (dof * BOLTZ)
    if t_now > 0.0:
        v *= np.sqrt(temperature / t_now)
    return v


class NVE:
    def __init__(self, mtp: api.MTPB200, sysm, halo, masses, *, dt: float = 0.001, temperature: float = 300.0,
                 seed: int = 12345, check_every: int = 10, skin: float = 2.0, rebuild_trigger: float | None = None,
                 variant=api.VARIANT_LARGE):
        dev = torch.device("cuda", torch.cuda.current_device())
        self.mtp, self.halo, self.lib = mtp, halo, mtp.lib
        self.nlocal, self.nall = sysm.nlocal, sysm.nall
        self.dt, self.check_every, self.skin, self.variant = float(dt), int(check_every), float(skin), variant
        self.rlist = float(sysm.rlist)
        # distance from the positions of the last list build that triggers a rebuild (LAMMPS: half the skin)
        self.trigger = 0.5 * self.skin if rebuild_trigger is None else float(rebuild_trigger)
        masses = np.asarray(masses, dtype=np.float64)
        mass1 = np.concatenate([[0.0], masses])                 # indexed by the 1-based type
        self.x = torch.from_numpy(np.ascontiguousarray(sysm.x)).to(dev)
        self.type = torch.from_numpy(np.ascontiguousarray(sysm.type)).to(dev)
        self.mass = torch.from_numpy(mass1).to(dev)
        v0 = maxwell_velocities(sysm.type[: self.nlocal], mass1, temperature, seed)
        self.v = torch.from_numpy(np.ascontiguousarray(v0)).to(dev)
        self.f = torch.zeros((self.nall, 3), dtype=torch.float64, device=dev)
        self.ev = torch.zeros(8, dtype=torch.float64, device=dev)
        self.ilist = torch.arange(self.nlocal, dtype=torch.int32, device=dev)
        self.nn = torch.zeros(self.nall, dtype=torch.int32, device=dev)
        self.moved = torch.zeros(1, dtype=torch.int32, device=dev)
        self.x_ghost_origin = self.x[: self.nlocal].clone()
        self.x_at_build = self.x[: self.nlocal].clone()
        self.m_local = self.mass[self.type[: self.nlocal].long()]
        self.rebuilds = 0
        self.steps_done = 0
        self._rebuild()
        self._force()

    # ---- stages ------------------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream().cuda_stream

    def _rebuild(self):
        far = float(((self.x[: self.nlocal] - self.x_ghost_origin) ** 2).sum(dim=1).max().sqrt())
        if far > 0.5 * self.skin:
            raise RuntimeError(f"an atom moved {far:.3f} A from where the ghost atoms were created (> skin/2): this harness "
                               "does not migrate atoms / rebuild ghosts")
        nn, self.table, self.max_nn = self.mtp.neigh_build(self.x, self.nlocal, self.rlist, stream=self._stream())
        self.nn[: self.nlocal] = nn
        self.x_at_build.copy_(self.x[: self.nlocal])
        self.moved.zero_()
        self.rebuilds += 1

    def _moved_anywhere(self) -> bool:
        """Re-neighboring is a collective decision (LAMMPS: MPI_Allreduce in Neighbor::decide)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.moved, op=dist.ReduceOp.MAX)
        return bool(int(self.moved.item()))

    def _force(self, forward: bool = True):
        if forward:
            self.halo.forward(self.x)
        self.f.zero_()
        self.mtp.compute_device(self.x, self.type, self.ilist, self.nn, self.table, None, self.f, self.ev,
                                stride_i=self.table.shape[1], stride_jj=1, eflag=1, vflag=1, variant=self.variant,
                                stream=self._stream(), max_numneigh=self.max_nn)
        self.halo.reverse(self.f)

    def run(self, nsteps: int):
        dtf = 0.5 * self.dt * FTM2V
        for _ in range(nsteps):
            api._check(self.lib, self.lib.mtp_nve_initial_integrate(
                self.nlocal, self.x.data_ptr(), self.v.data_ptr(), self.f.data_ptr(), self.type.data_ptr(),
                self.mass.data_ptr(), dtf, self.dt, self.x_at_build.data_ptr(), self.trigger, self.moved.data_ptr(),
                self._stream()))
            self.steps_done += 1
            self.halo.forward(self.x)
            if self.steps_done % self.check_every == 0 and self._moved_anywhere():
                self._rebuild()                 # from the current owned AND ghost positions
            self._force(forward=False)
            api._check(self.lib, self.lib.mtp_nve_final_integrate(
                self.nlocal, self.v.data_ptr(), self.f.data_ptr(), self.type.data_ptr(), self.mass.data_ptr(), dtf,
                self._stream()))

    # ---- thermo ------------------------------------------------------------------------------------------------
    def potential_energy(self) -> float:
        return float(self.ev[0].item())

    def kinetic_energy(self) -> float:
        return float(0.5 * MVV2E * (self.m_local[:, None] * self.v * self.v).sum().item())

    def temperature(self) -> float:
        return 2.0 * self.kinetic_energy() / (max(3 * self.nlocal - 3, 1) * BOLTZ)


def scale_to_rms_force(pot, rms_now: float, rms_target: float = 0.05):
    """Random-init MTPs are unphysically stiff; forces are linear in the moment coefficients, so one global factor
    brings the RMS force on the jittered lattice to rms_target eV/A."""
    import copy
    out = copy.deepcopy(pot)
    s = rms_target / max(rms_now, 1e-300)
    out.moment_coeffs = np.asarray(pot.moment_coeffs) * s
    out.species_coeffs = np.asarray(pot.species_coeffs) * s
    return out
