"""GPU parity: the CUDA path (through the C ABI) against the oracle on the same seeded inputs."""
import numpy as np
import pytest

import util
from util import TOL_AUX, TOL_E_REL, TOL_F_MAXABSREL, maxabsrel

pytestmark = pytest.mark.gpu


def _check(gpu, ref, sysm, what=""):
    assert abs(gpu.energy - ref.energy) <= TOL_E_REL * max(abs(ref.energy), 1e-300), what
    assert maxabsrel(gpu.f, ref.f) <= TOL_F_MAXABSREL, what
    assert maxabsrel(gpu.virial, ref.virial) <= TOL_AUX, what
    assert maxabsrel(gpu.eatom[: sysm.nlocal], ref.eatom[: sysm.nlocal]) <= TOL_AUX, what
    assert maxabsrel(gpu.vatom, ref.vatom) <= TOL_AUX, what


CASES = [
    (8, 1, "fcc", 4.05, (4, 4, 4)),
    (10, 1, "fcc", 4.05, (5, 5, 5)),       # config 1 (shrunk)
    (16, 2, "bcc", 3.165, (6, 6, 6)),      # config 2 (shrunk)
    (20, 1, "diamond", 5.431, (3, 3, 3)),  # config 3 (shrunk)
    (22, 3, "fcc", 3.56, (4, 4, 4)),       # config 5 (shrunk)
    (2, 1, "fcc", 4.05, (4, 4, 4)),        # level-2 closed form: K=1, T=0
    (6, 2, "bcc", 3.165, (5, 5, 5)),
]


@pytest.mark.parametrize("level,species,kind,a,cells", CASES)
def test_energy_force_virial_mask(tmp_path, built, level, species, kind, a, cells):
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, level, species)
    sysm = util.small_system(kind, a, cells, species)
    ref = OracleMTP(pot).compute(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets, want_mask=True)
    mtp = MTPB200(path)
    gpu = mtp.compute_system(sysm, want_mask=True)
    _check(gpu, ref, sysm, f"L{level}")
    # neighbor indexing / cutoff mask is bit-exact
    assert np.array_equal(gpu.mask[: ref.mask.size], ref.mask)
    # Newton's third law including ghosts
    assert np.abs(gpu.f.sum(axis=0)).max() <= 1e-9 * np.abs(gpu.f).max()
    mtp.close()


def test_ragged_cluster_and_padded_2d_list(tmp_path, built):
    """Empty / ragged neighborhoods, listed-but-outside-cutoff pairs, and the 2-D strided list form."""
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, 12, 3)
    sysm = util.random_cluster(60, 3)
    assert sysm.numneigh[: sysm.nlocal].min() == 0
    ref = OracleMTP(pot).compute(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets)
    mtp = MTPB200(path)
    gpu = mtp.compute_system(sysm)
    _check(gpu, ref, sysm)
    tab = sysm.padded_neighbors()
    gpu2 = mtp.compute_host(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, tab.ravel(), None, stride_i=tab.shape[1], stride_jj=1)
    _check(gpu2, ref, sysm)
    tabT = np.ascontiguousarray(tab.T)     # LayoutLeft: jj slow, i fast
    gpu3 = mtp.compute_host(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, tabT.ravel(), None, stride_i=1, stride_jj=tab.shape[0])
    _check(gpu3, ref, sysm)
    mtp.close()


def test_flags_ilist_subset_accumulate_and_neighmask(tmp_path, built):
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, 10, 2)
    sysm = util.small_system("bcc", 3.165, (5, 5, 5), 2)
    orc = OracleMTP(pot)
    rng = np.random.default_rng(0)
    ilist = np.sort(rng.choice(sysm.nlocal, size=sysm.nlocal // 3, replace=False)).astype(np.int32)
    f0 = rng.normal(size=(sysm.nall, 3))
    neigh = sysm.neigh.copy()
    neigh[::3] |= (1 << 30)        # special-bond bits must be masked off (pair_mtp.cpp:114)
    ref = orc.compute(sysm.x, sysm.type, ilist, sysm.numneigh, neigh, sysm.offsets, eflag=1, vflag=1, f_init=f0)
    mtp = MTPB200(path)
    gpu = mtp.compute_host(sysm.x, sysm.type, ilist, sysm.numneigh, neigh, sysm.offsets, eflag=1, vflag=1, f_init=f0)
    assert abs(gpu.energy - ref.energy) <= TOL_E_REL * abs(ref.energy)
    assert maxabsrel(gpu.f, ref.f) <= TOL_F_MAXABSREL
    assert maxabsrel(gpu.virial, ref.virial) <= TOL_AUX
    # no energy / virial requested -> ev stays zero, forces unchanged
    gpu0 = mtp.compute_host(sysm.x, sysm.type, ilist, sysm.numneigh, neigh, sysm.offsets, eflag=0, vflag=0, f_init=f0)
    assert np.all(gpu0.ev[:7] == 0.0)
    assert maxabsrel(gpu0.f, ref.f) <= TOL_F_MAXABSREL
    mtp.close()


@pytest.mark.parametrize("level,species,kind,a,cells", [(8, 1, "fcc", 4.05, (4, 4, 4)), (16, 2, "fcc", 4.05, (4, 4, 4))])
def test_neighborhood_grades(tmp_path, built, level, species, kind, a, cells):
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, level, species, active_set=True)
    sysm = util.small_system(kind, a, cells, species, seed=11)
    ref = OracleMTP(pot).compute(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets, grade=True)
    mtp = MTPB200(path, selection_state=True)
    for chunk in (1 << 30, 37):
        mtp.set_chunksize(chunk)
        gpu = mtp.compute_system(sysm, grade=True)
        _check(gpu, ref, sysm)
        assert maxabsrel(gpu.grades[: sysm.nlocal], ref.grades[: sysm.nlocal]) <= TOL_AUX
        assert abs(gpu.max_grade - ref.max_grade) <= TOL_AUX * ref.max_grade
    mtp.close()


def test_configuration_grade(tmp_path, built):
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, 10, 2, active_set=True, cfg_mode=True)
    sysm = util.small_system("bcc", 3.165, (5, 5, 5), 2)
    ref = OracleMTP(pot).compute(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets, grade=True,
                                 natoms_total=sysm.nlocal)
    mtp = MTPB200(path, selection_state=True)
    assert mtp.info.configuration_mode == 1
    mtp.set_chunksize(64)
    gpu = mtp.compute_system(sysm, grade=True, natoms_total=sysm.nlocal)
    _check(gpu, ref, sysm)
    assert maxabsrel(gpu.candidate[: pot.coeff_count], ref.candidate[: pot.coeff_count]) <= TOL_AUX
    assert abs(gpu.max_grade - ref.max_grade) <= TOL_AUX * ref.max_grade
    mtp.close()


def test_species_bound_is_reported(tmp_path, built):
    from mtp_b200.api import MTPB200, MTPError
    path, _ = util.write_potential(tmp_path, 8, 1)
    sysm = util.small_system("fcc", 4.05, (4, 4, 4), 1)
    t = sysm.type.copy()
    t[5] = 2
    mtp = MTPB200(path)
    with pytest.raises(MTPError, match="Too few species"):
        mtp.compute_host(sysm.x, t, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets)
    mtp.close()


def test_fp64_roofs_are_measurable(built):
    from mtp_b200.api import fp64_peaks
    dfma, dmma = fp64_peaks()
    print("FP64 roofs TFLOP/s: DFMA", dfma, "DMMA", dmma)
    assert dfma > 5.0 and dmma > 0.5


# systems large enough that every SM gets a chunk, so the throughput shape of the contraction program is the one that
# runs: the kernel generated for the potential (default), or -- MTP_B200_NO_P4 -- the interpreting 4-atoms-per-lane
# kernel (32 or 16 atoms per CTA); ilist subsets give ragged tail chunks
@pytest.mark.parametrize("generated", [True, False])
@pytest.mark.parametrize("level,species,kind,a,cells,atoms_per_cta", [
    (16, 2, "bcc", 3.165, (14, 14, 14), 32),     # config 2 shape
    (12, 3, "fcc", 3.56, (11, 11, 11), 32),
    (18, 1, "fcc", 4.05, (9, 9, 9), 16),
])
def test_program_kernel_throughput_shape(tmp_path, built, monkeypatch, level, species, kind, a, cells, atoms_per_cta, generated):
    from mtp_b200 import api
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, level, species)
    sysm = util.small_system(kind, a, cells, species, seed=5)
    orc = OracleMTP(pot)
    if generated:
        import torch
        sms = torch.cuda.get_device_properties(0).multi_processor_count
        # fewer atoms than one 32-atom chunk per SM: the library takes the latency shape of the generated kernel
        atoms_per_cta = api.codegen_source(path, sysm.nlocal < 32 * sms)[1]["atoms_per_cta"]
    else:
        monkeypatch.setenv("MTP_B200_NO_P4", "1")
    mtp = MTPB200(path)
    assert (mtp.program_kernel_note() == "") == generated, mtp.program_kernel_note()
    for lanes in (1, 2):
        mtp.set_lanes(lanes)
        mtp.set_chunksize(1 << 30 if lanes == 1 else 2500)
        ilist = sysm.ilist if lanes == 1 else sysm.ilist[: sysm.nlocal - 13]
        ref = orc.compute(sysm.x, sysm.type, ilist, sysm.numneigh, sysm.neigh, sysm.offsets)
        gpu = mtp.compute_host(sysm.x, sysm.type, ilist, sysm.numneigh, sysm.neigh, sysm.offsets)
        if lanes == 1:
            path_used = mtp.last_kernel_path()
            assert path_used["program_generated"] == generated and path_used["program_v3"] == (not generated), path_used
            assert path_used["program_atoms_per_cta"] == atoms_per_cta, path_used
        assert abs(gpu.energy - ref.energy) <= TOL_E_REL * abs(ref.energy)
        assert maxabsrel(gpu.f, ref.f) <= TOL_F_MAXABSREL
        assert maxabsrel(gpu.virial, ref.virial) <= TOL_AUX
        assert maxabsrel(gpu.eatom[ilist], ref.eatom[ilist]) <= TOL_AUX
    mtp.close()


def test_program_kernel_throughput_shape_grades(tmp_path, built):
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, 16, 2, active_set=True)
    sysm = util.small_system("fcc", 4.05, (11, 11, 11), 2, seed=11)
    ref = OracleMTP(pot).compute(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets, grade=True)
    mtp = MTPB200(path, selection_state=True)
    gpu = mtp.compute_system(sysm, grade=True)
    assert mtp.last_kernel_path()["program_generated"]
    _check(gpu, ref, sysm)
    assert maxabsrel(gpu.grades[: sysm.nlocal], ref.grades[: sysm.nlocal]) <= TOL_AUX
    assert abs(gpu.max_grade - ref.max_grade) <= TOL_AUX * ref.max_grade
    mtp.close()


def test_device_side_grade_selection(tmp_path, built):
    """Atoms with grade >= threshold compacted on the device == numpy on the oracle's grades."""
    import torch
    from mtp_b200.api import MTPB200
    from oracle_py import OracleMTP
    path, pot = util.write_potential(tmp_path, 8, 2, active_set=True)
    sysm = util.small_system("bcc", 3.165, (6, 6, 6), 2, seed=11)
    ref = OracleMTP(pot).compute(sysm.x, sysm.type, sysm.ilist, sysm.numneigh, sysm.neigh, sysm.offsets, grade=True)
    mtp = MTPB200(path, selection_state=True)
    gpu = mtp.compute_system(sysm, grade=True)
    g = torch.from_numpy(np.ascontiguousarray(gpu.grades[: sysm.nlocal])).cuda()
    for q in (0.0, 0.5, 0.9, 1.0, 2.0):
        thr = float(np.quantile(ref.grades[: sysm.nlocal], min(q, 1.0))) * (2.0 if q > 1.0 else 1.0)
        sel = mtp.select_grades(g, thr).cpu().numpy()
        assert np.array_equal(sel, np.nonzero(gpu.grades[: sysm.nlocal] >= thr)[0])
    assert mtp.select_grades(g[:0], 1.0).numel() == 0
    mtp.close()
