"""Device-resident NVE loop (mtp_nve_*_integrate + mtp_neigh_build + mtp_compute):
the integrator kernels replay FixNVE bit for bit, and the whole loop conserves energy."""
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def test_integrator_kernels_replay_fixnve_bit_for_bit(built):
    import torch
    from mtp_b200 import api
    from mtp_b200.md import FTM2V
    lib = api.load_library()
    rng = np.random.default_rng(3)
    n, nall = 1000, 1300
    x = rng.uniform(0, 30, size=(nall, 3))
    v = rng.normal(size=(n, 3))
    f = rng.normal(size=(nall, 3))
    typ = rng.integers(1, 4, size=nall).astype(np.int32)
    mass = np.array([0.0, 26.98, 63.55, 183.84])
    dt = 0.001
    dtf = 0.5 * dt * FTM2V
    x0 = x[:n] + rng.normal(scale=0.3, size=(n, 3))
    tx, tv, tf = (torch.from_numpy(a.copy()).cuda() for a in (x, v, f))
    tt, tm, t0 = torch.from_numpy(typ).cuda(), torch.from_numpy(mass).cuda(), torch.from_numpy(x0).cuda()
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")
    api._check(lib, lib.mtp_nve_initial_integrate(n, tx.data_ptr(), tv.data_ptr(), tf.data_ptr(), tt.data_ptr(), tm.data_ptr(),
                                                  dtf, dt, t0.data_ptr(), 1.0, flag.data_ptr(), None))
    dtfm = (dtf / mass[typ[:n]])[:, None]
    v1 = v + dtfm * f[:n]
    x1 = x.copy()
    x1[:n] = x[:n] + dt * v1
    assert np.array_equal(tv.cpu().numpy(), v1)
    assert np.array_equal(tx.cpu().numpy(), x1)          # ghost rows untouched
    moved = bool((((x1[:n] - x0) ** 2).sum(axis=1) > 1.0).any())
    assert bool(flag.item()) == moved and moved
    api._check(lib, lib.mtp_nve_final_integrate(n, tv.data_ptr(), tf.data_ptr(), tt.data_ptr(), tm.data_ptr(), dtf, None))
    assert np.array_equal(tv.cpu().numpy(), v1 + dtfm * f[:n])
    # both-or-neither contract of the displacement check
    rc = lib.mtp_nve_initial_integrate(n, tx.data_ptr(), tv.data_ptr(), tf.data_ptr(), tt.data_ptr(), tm.data_ptr(), dtf, dt,
                                       t0.data_ptr(), 1.0, None, None)
    assert rc == -1


def test_nve_conserves_energy_with_device_built_list(tmp_path, built):
    import torch
    from mtp_b200 import almtp, decomp, harness
    from mtp_b200.api import MTPB200
    from mtp_b200.md import NVE, scale_to_rms_force
    dev = torch.device("cuda", 0)
    pot0 = almtp.random_potential(10, 2)
    x, box = harness.lattice("bcc", 3.165, (8, 8, 8), jitter=0.05)
    types = harness.random_types(len(x), [0.5, 0.5], 7)
    path0 = os.path.join(str(tmp_path), "raw.almtp")
    almtp.write_almtp(path0, pot0)
    mtp0 = MTPB200(path0)
    sysm, halo = decomp.build_rank_system_direct(x, types, np.zeros(3), box, (1, 1, 1), 0, box, 5.0, 2.0, dev, mtp0.lib)
    r0 = mtp0.compute_system(sysm)
    f_own = sysm.reverse_comm(r0.f) if (sysm.owner[sysm.nlocal:] >= 0).all() else None
    if f_own is None:      # DirectHalo systems do not carry owner ids for ghosts: fold ghost forces on the device instead
        tf = torch.from_numpy(r0.f).to(dev)
        halo.reverse(tf)
        f_own = tf[: sysm.nlocal].cpu().numpy()
    rms = float(np.sqrt((f_own ** 2).sum(axis=1).mean()))
    mtp0.close()
    pot = scale_to_rms_force(pot0, rms, 0.05)
    path = os.path.join(str(tmp_path), "scaled.almtp")
    almtp.write_almtp(path, pot)
    mtp = MTPB200(path)
    # (a random-init potential does not confine the atoms, so the run is kept cold and short enough that nobody leaves
    # the ghost shell; the rebuild trigger is lowered to 0.05 A so that the device list IS rebuilt several times)
    md = NVE(mtp, sysm, halo, masses=[183.84, 95.95], dt=0.001, temperature=50.0, seed=12345, rebuild_trigger=0.05)
    rms_scaled = float((md.f[: sysm.nlocal] ** 2).sum(dim=1).mean().sqrt())
    assert abs(rms_scaled - 0.05) < 1e-6
    assert abs(md.temperature() - 50.0) < 1e-6
    e0 = md.potential_energy() + md.kinetic_energy()
    es = []
    for _ in range(6):
        md.run(20)
        es.append(md.potential_energy() + md.kinetic_energy())
    drift = max(abs(e - e0) for e in es) / sysm.nlocal
    ke_atom = md.kinetic_energy() / sysm.nlocal
    print(f"NVE 120 steps: max |dE|/atom = {drift:.3e} eV, KE/atom = {ke_atom:.3e} eV, T = {md.temperature():.1f} K, "
          f"rebuilds = {md.rebuilds}")
    assert md.rebuilds >= 3
    assert drift < 1e-4 * ke_atom        # four orders of magnitude below the thermal energy per atom
    # total momentum stays zero (Newton's third law through ghosts and the reverse halo)
    p = (md.m_local[:, None] * md.v).sum(dim=0).abs().max().item()
    assert p < 1e-8 * float((md.m_local[:, None] * md.v).abs().sum())
    mtp.close()
