"""Device neighbor-list build (mtp_neigh_build) against the host list of the harness:
neighbor SETS are bit-exact (same rsq rounding), and the device-built table drives the pair style to the oracle's
energies and forces."""
import ctypes as C

import numpy as np
import pytest

import util
from util import TOL_AUX, TOL_E_REL, TOL_F_MAXABSREL, maxabsrel

pytestmark = pytest.mark.gpu


def _host_rows(sysm):
    """Host CSR list -> rows sorted ascending, padded with -1."""
    nl = sysm.nlocal
    w = int(sysm.numneigh[:nl].max()) if nl else 0
    tab = np.full((nl, max(w, 1)), -1, dtype=np.int64)
    for i in range(nl):
        n = sysm.numneigh[i]
        tab[i, :n] = np.sort(sysm.neigh[sysm.offsets[i]: sysm.offsets[i] + n])
    return tab


def _device_rows(numneigh, table):
    nn = numneigh.cpu().numpy()
    t = table.cpu().numpy().astype(np.int64)
    w = t.shape[1]
    t[np.arange(w)[None, :] >= nn[:, None]] = np.iinfo(np.int64).max
    t.sort(axis=1)
    t[t == np.iinfo(np.int64).max] = -1
    return nn, t


SYSTEMS = [
    ("fcc", 4.05, (6, 6, 6), 1),
    ("bcc", 3.165, (9, 8, 7), 2),        # config 2 shape, unequal box edges
    ("diamond", 5.431, (4, 4, 5), 1),
]


@pytest.mark.parametrize("kind,a,cells,species", SYSTEMS)
def test_neighbor_sets_are_bit_exact(tmp_path, built, kind, a, cells, species):
    import torch
    from mtp_b200.api import MTPB200
    path, _ = util.write_potential(tmp_path, 8, species)
    sysm = util.small_system(kind, a, cells, species)
    mtp = MTPB200(path)
    x = torch.from_numpy(sysm.x).cuda()
    numneigh, table, mx = mtp.neigh_build(x, sysm.nlocal, sysm.rlist)
    nn, rows = _device_rows(numneigh, table)
    ref = _host_rows(sysm)
    assert np.array_equal(nn, sysm.numneigh[: sysm.nlocal])
    assert mx == int(sysm.numneigh[: sysm.nlocal].max())
    w = min(rows.shape[1], ref.shape[1])
    assert np.array_equal(rows[:, :w], ref[:, :w])
    assert (rows[:, w:] == -1).all() and (ref[:, w:] == -1).all()
    mtp.close()


def test_ragged_cluster_empty_rows_and_capacity(tmp_path, built):
    import torch
    from mtp_b200.api import MTPB200, load_library
    path, _ = util.write_potential(tmp_path, 8, 3)
    sysm = util.random_cluster(60, 3)
    assert sysm.numneigh[: sysm.nlocal].min() == 0
    mtp = MTPB200(path)
    x = torch.from_numpy(sysm.x).cuda()
    numneigh, table, mx = mtp.neigh_build(x, sysm.nlocal, sysm.rlist, width=8)    # too narrow: rebuilt wider
    assert table.shape[1] >= mx > 8
    nn, rows = _device_rows(numneigh, table)
    ref = _host_rows(sysm)
    assert np.array_equal(nn, sysm.numneigh[: sysm.nlocal])
    assert np.array_equal(rows[:, : ref.shape[1]], ref)
    # the raw call reports the capacity error and the width it needs
    lib = load_library()
    nn_d = torch.empty(sysm.nlocal, dtype=torch.int32, device="cuda")
    tab_d = torch.empty((sysm.nlocal, 4), dtype=torch.int32, device="cuda")
    need = C.c_int(0)
    rc = lib.mtp_neigh_build(mtp.h, sysm.nlocal, sysm.nall, x.data_ptr(), float(sysm.rlist), nn_d.data_ptr(), tab_d.data_ptr(), 4,
                             C.byref(need), None)
    assert rc == MTPB200.ERR_CAPACITY and need.value == mx
    assert b"too narrow" in lib.mtp_last_error()
    assert np.array_equal(nn_d.cpu().numpy(), sysm.numneigh[: sysm.nlocal])    # counts stay valid
    mtp.close()


def test_device_built_list_drives_the_pair_style(tmp_path, built):
    import torch
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, 12, 2)
    sysm = util.small_system("bcc", 3.165, (8, 8, 8), 2)
    ref = OracleMTP(pot).compute(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets)
    mtp = MTPB200(path)
    x = torch.from_numpy(sysm.x).cuda()
    typ = torch.from_numpy(sysm.type).cuda()
    numneigh, table, mx = mtp.neigh_build(x, sysm.nlocal, sysm.rlist)
    nn_all = torch.zeros(sysm.nall, dtype=torch.int32, device="cuda")
    nn_all[: sysm.nlocal] = numneigh
    f = torch.zeros((sysm.nall, 3), dtype=torch.float64, device="cuda")
    ev = torch.zeros(8, dtype=torch.float64, device="cuda")
    eatom = torch.zeros(sysm.nall, dtype=torch.float64, device="cuda")
    ilist = torch.arange(sysm.nlocal, dtype=torch.int32, device="cuda")
    mtp.compute_device(x, typ, ilist, nn_all, table, None, f, ev, eatom=eatom, stride_i=table.shape[1], stride_jj=1,
                       eflag=3, vflag=1, max_numneigh=mx)
    mtp.synchronize()
    evh = ev.cpu().numpy()
    assert abs(evh[0] - ref.energy) <= TOL_E_REL * abs(ref.energy)
    assert maxabsrel(f.cpu().numpy(), ref.f) <= TOL_F_MAXABSREL
    assert maxabsrel(evh[1:7], ref.virial) <= TOL_AUX
    assert maxabsrel(eatom.cpu().numpy()[: sysm.nlocal], ref.eatom[: sysm.nlocal]) <= TOL_AUX
    mtp.close()
