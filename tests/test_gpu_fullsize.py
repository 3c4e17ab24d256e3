"""BASELINE.json config 2 at its FULL size (262,144-atom bcc W/Mo, level 16): the oracle is too slow here, so parity is
checked through size-independent properties -- Newton's third law, energy = sum of per-atom energies, the virial
identity, independence of the result from how the work is cut (super-chunks, lanes, program-kernel form, host- or
device-built neighbor list) -- plus an oracle comparison on a slab of centre atoms of the same system."""
import os

import numpy as np
import pytest

import util
from util import TOL_AUX, TOL_E_REL, TOL_F_MAXABSREL, maxabsrel

pytestmark = pytest.mark.gpu


def test_config2_full_size_properties(tmp_path, built):
    import torch
    from mtp_b200 import harness
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    cfg = harness.CONFIGS[2]
    path, pot = util.write_potential(tmp_path, cfg["level"], cfg["species"])
    sysm = harness.make_config(2)
    assert sysm.nlocal == 262144
    mtp = MTPB200(path)
    res = mtp.compute_system(sysm, eflag=3, vflag=1)
    assert mtp.last_kernel_path()["program_generated"]
    f = sysm.reverse_comm(res.f)
    fmax = np.abs(f).max()
    # Newton's third law through the ghosts: the total force vanishes
    assert np.abs(f.sum(axis=0)).max() <= 1e-9 * fmax * np.sqrt(sysm.nlocal)
    # total energy is the sum of the per-atom energies (fixed-order device reduction vs numpy)
    assert abs(res.eatom[: sysm.nlocal].sum() - res.energy) <= 1e-12 * abs(res.energy)
    # virial identity (pair_mtp.cpp:255-266): W_ab = sym(sum_k x_ka f_kb) over owned + ghost rows, before reverse comm
    a, b = [0, 1, 2, 0, 0, 1], [0, 1, 2, 1, 2, 2]
    w = 0.5 * ((sysm.x[:, a] * res.f[:, b]).sum(axis=0) + (sysm.x[:, b] * res.f[:, a]).sum(axis=0))
    assert maxabsrel(w, res.virial) <= 1e-9

    # the same evaluation cut differently: one lane / one super-chunk per launch limit, small chunks, three lanes
    for lanes, chunk in ((1, 1 << 30), (3, 20000), (2, 65536)):
        mtp.set_lanes(lanes)
        mtp.set_chunksize(chunk)
        r2 = mtp.compute_system(sysm, eflag=3, vflag=1)
        assert abs(r2.energy - res.energy) <= TOL_E_REL * abs(res.energy)
        assert maxabsrel(r2.f, res.f) <= TOL_F_MAXABSREL
        assert maxabsrel(r2.virial, res.virial) <= TOL_AUX
        assert maxabsrel(r2.eatom, res.eatom) <= TOL_AUX
    mtp.set_lanes(2)
    mtp.set_chunksize(32768)

    # the interpreting forms of the contraction program (separate handles: the choice is made at load)
    for env in ({"MTP_B200_NO_P4": "1"}, {"MTP_B200_NO_P4": "1", "MTP_B200_NO_PROG_V3": "1"}):
        os.environ.update(env)
        try:
            old = MTPB200(path)
        finally:
            for k in env:
                del os.environ[k]
        r3 = old.compute_system(sysm, eflag=3, vflag=1)
        used = old.last_kernel_path()
        assert not used["program_generated"] and used["program_v3"] == ("MTP_B200_NO_PROG_V3" not in env)
        assert abs(r3.energy - res.energy) <= TOL_E_REL * abs(res.energy)
        assert maxabsrel(r3.f, res.f) <= TOL_F_MAXABSREL
        old.close()

    # neighbor list built on the device: same counts as the host list, same energies and forces
    x = torch.from_numpy(sysm.x).cuda()
    numneigh, table, mx = mtp.neigh_build(x, sysm.nlocal, sysm.rlist)
    assert np.array_equal(numneigh.cpu().numpy(), sysm.numneigh[: sysm.nlocal])
    # (row contents: checksum of checksums -- per-row sum and sum of squares of the neighbor ids)
    t = table.cpu().numpy().astype(np.int64)
    t[np.arange(t.shape[1])[None, :] >= sysm.numneigh[: sysm.nlocal, None]] = 0
    ii = np.repeat(np.arange(sysm.nlocal), sysm.numneigh[: sysm.nlocal])
    assert np.array_equal(t.sum(axis=1), np.bincount(ii, weights=sysm.neigh, minlength=sysm.nlocal).astype(np.int64))
    assert np.array_equal((t * t).sum(axis=1),
                          np.bincount(ii, weights=sysm.neigh.astype(np.float64) ** 2, minlength=sysm.nlocal).astype(np.int64))
    typ = torch.from_numpy(sysm.type).cuda()
    nn_all = torch.zeros(sysm.nall, dtype=torch.int32, device="cuda")
    nn_all[: sysm.nlocal] = numneigh
    fd = torch.zeros((sysm.nall, 3), dtype=torch.float64, device="cuda")
    ev = torch.zeros(8, dtype=torch.float64, device="cuda")
    ilist = torch.arange(sysm.nlocal, dtype=torch.int32, device="cuda")
    mtp.compute_device(x, typ, ilist, nn_all, table, None, fd, ev, stride_i=table.shape[1], stride_jj=1, eflag=1, vflag=1,
                       max_numneigh=mx)
    mtp.synchronize()
    assert abs(float(ev[0]) - res.energy) <= TOL_E_REL * abs(res.energy)
    assert maxabsrel(fd.cpu().numpy(), res.f) <= TOL_F_MAXABSREL

    # the oracle on a slab of 2,000 centre atoms of the full system (eatom is per centre; forces need all centres)
    slab = np.arange(100000, 102000, dtype=np.int32)
    ref = OracleMTP(pot).compute(sysm.x, sysm.type, slab, sysm.numneigh, sysm.neigh, sysm.offsets, eflag=3, vflag=0)
    gpu = mtp.compute_host(sysm.x, sysm.type, slab, sysm.numneigh, sysm.neigh, sysm.offsets, eflag=3, vflag=0)
    assert abs(gpu.energy - ref.energy) <= TOL_E_REL * abs(ref.energy)
    assert maxabsrel(gpu.f, ref.f) <= TOL_F_MAXABSREL
    assert maxabsrel(gpu.eatom[slab], ref.eatom[slab]) <= TOL_AUX
    mtp.close()
