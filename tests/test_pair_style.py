"""The LAMMPS PairStyle host side (lammps-mtp-kokkos_b200/lammps/pair_mtp_b200.cpp), driven like LAMMPS drives
a pair style.  CPU part: argument grammar and error behaviour of the reference's KOKKOS styles
(pair_mtp_kokkos.cpp:104-117, pair_mtp_extrapolation_kokkos.cpp:116-138, pair_mtp.cpp:303-329).
GPU part: the styles against the golden outputs of the reference's CPU styles."""
import os

import numpy as np
import pytest

import golden_util
import util
from pair_driver import LammpsError, PairB200
from util import TOL_AUX, TOL_E_REL, TOL_F_MAXABSREL, maxabsrel


@pytest.fixture(scope="module")
def plugin():
    import pair_driver
    return pair_driver.build()


def test_unknown_style(plugin):
    with pytest.raises(LammpsError, match="Unrecognized pair style"):
        PairB200("mtp/kk/host", "x.almtp", "chunksize", "10")


@pytest.mark.parametrize("args", [("pot.almtp",), ("pot.almtp", "chunk", "10"), ("pot.almtp", "chunksize", "10", "x")])
def test_inference_style_needs_chunksize_keyword(plugin, args):
    with pytest.raises(LammpsError, match="requires 3 arguments"):
        PairB200("mtp/kk", *args)
    with pytest.raises(LammpsError, match="requires 3 arguments"):
        PairB200("mtp/small/kk", *args)


def test_extrapolation_style_grammar(plugin):
    with pytest.raises(LammpsError, match="requires 3 :"):
        PairB200("mtp/extrapolation/kk", "pot.almtp", "out.cfg", "2", "10")
    with pytest.raises(LammpsError, match="Chunksize not found"):
        PairB200("mtp/extrapolation/kk", "pot.almtp", "chunk", "10")
    with pytest.raises(LammpsError, match="Chunksize not found"):
        PairB200("mtp/extrapolation/small/kk", "pot.almtp", "out.cfg", "2", "10", "size", "10")
    with pytest.raises(LammpsError, match="Expected integer"):
        PairB200("mtp/extrapolation/kk", "pot.almtp", "chunksize", "many")


def test_missing_file_and_no_gpu_are_fatal(plugin, tmp_path):
    with pytest.raises(LammpsError, match="Cannot open potential file"):
        PairB200("mtp/kk", str(tmp_path / "nope.almtp"), "chunksize", "32768")
    import torch
    if not torch.cuda.is_available():
        path, _ = util.write_potential(tmp_path, 8, 1)
        with pytest.raises(LammpsError, match="no CPU fallback|CUDA"):
            PairB200("mtp/kk", path, "chunksize", "32768")


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("kokkos", [False, True], ids=["host-buffers", "kokkos-views"])
@pytest.mark.parametrize("style", ["mtp/kk", "mtp/small/kk"])
@pytest.mark.parametrize("name", ["L10_S1_fcc", "L16_S2_bcc", "L12_S3_cluster"])
def test_inference_styles_match_reference_cpu_style(plugin, tmp_path, style, name, kokkos):
    """Both build flavours of the style: host buffers (plain LAMMPS, mtp_compute_host) and device views (LMP_KOKKOS,
    mtp_compute on AtomKokkos / NeighListKokkos views, the 2-D neighbor view in LayoutLeft)."""
    g = golden_util.Golden(name, tmp_path)
    pair = PairB200(style, g.path, "chunksize", "32768", species=g.pot.species_count, kokkos=kokkos)
    assert "species" in pair.log
    r = pair.compute(g.x, g.type, g.nlocal, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=3, vflag=5)
    assert abs(r.energy - g.energy) <= TOL_E_REL * abs(g.energy)
    assert maxabsrel(r.f, g.f) <= TOL_F_MAXABSREL
    assert maxabsrel(r.virial, g.virial) <= TOL_AUX
    assert maxabsrel(r.eatom, g.eatom) <= TOL_AUX
    assert maxabsrel(r.vatom, g.vatom) <= TOL_AUX
    # second step without re-neighboring (ago > 0): the list is not re-sent, forces accumulate into f
    r2 = pair.compute(g.x, g.type, g.nlocal, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=1, vflag=1, ago=1, f_init=r.f)
    assert maxabsrel(r2.f, 2 * g.f) <= TOL_F_MAXABSREL
    pair.close()
    if not kokkos:
        # the style alone (force->pair): LAMMPS cleared f, so the result is stored, whatever the array held
        lone = PairB200(style, g.path, "chunksize", "32768", species=g.pot.species_count, lone=True)
        r3 = lone.compute(g.x, g.type, g.nlocal, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=1, vflag=1, f_init=r.f)
        assert maxabsrel(r3.f, g.f) <= TOL_F_MAXABSREL
        lone.close()


@pytest.mark.gpu
def test_pair_requires_newton_on_and_pair_coeff_star_star(plugin, tmp_path):
    path, _ = util.write_potential(tmp_path, 8, 1)
    pair = PairB200("mtp/kk", path, "CHUNKSIZE", "64")     # keyword is case-insensitive (utils::lowercase)
    pair.close()


def _parse_cfg(text):
    lines = text.strip().split("\n")
    i = lines.index("Size")
    n = int(lines[i + 1])
    cell = [[float(v) for v in lines[i + 3 + k].split()] for k in range(3)]
    hdr = lines[i + 6]
    rows = [ln.split("\t") for ln in lines[i + 7: i + 7 + n]]
    feat = lines[i + 7 + n]
    return n, cell, hdr, rows, feat, lines[0], lines[-1]


@pytest.mark.gpu
@pytest.mark.parametrize("kokkos", [False, True], ids=["host-buffers", "kokkos-views"])
@pytest.mark.parametrize("style", ["mtp/extrapolation/kk", "mtp/extrapolation/small/kk"])
@pytest.mark.parametrize("name", ["L10_S2_nbh", "L10_S2_cfg"])
def test_extrapolation_styles(plugin, tmp_path, style, name, kokkos):
    g = golden_util.Golden(name, tmp_path)
    S = g.pot.species_count
    from functools import partial
    import pair_driver
    PairB200 = partial(pair_driver.PairB200, kokkos=kokkos)      # noqa: N806  (every style of this test in one flavour)
    # LAMMPS-style: grades only when fix pair raises extrapolation_flag
    pair = PairB200(style, g.path, "chunksize", "100", species=S)
    assert ("Configuration" if g.mode == "cfg" else "Neighborhood") in pair.log
    pair.set_domain(g.box, len(g.ilist))
    r0 = pair.compute(g.x, g.type, g.nlocal, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=1, vflag=1, grade=False)
    assert r0.max_grade == 0.0 and abs(r0.energy - g.energy) <= TOL_E_REL * abs(g.energy)
    if g.mode == "nbh":
        r = pair.compute(g.x, g.type, g.nlocal, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=1, vflag=1, grade=True)
        assert maxabsrel(r.grades[: g.nlocal], g.grades[: g.nlocal]) <= TOL_AUX
        assert abs(r.max_grade - g.max_grade) <= TOL_AUX * g.max_grade      # pvector[0]
    else:
        with pytest.raises(LammpsError, match="MLIP-3 style extrapolation"):
            pair.compute(g.x, g.type, g.nlocal, g.ilist, g.numneigh, g.neigh, g.offsets, grade=True)
    pair.close()
    # MLIP-3 style: thresholds + preselected .cfg file, compared with the block the reference wrote
    out = str(tmp_path / "preselected.cfg")
    pair = PairB200(style, g.path, out, "0.0", "1e300", "chunksize", "32768", species=S)
    pair.set_domain(g.box, len(g.ilist))
    pair.compute(g.x, g.type, g.nlocal, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=1, vflag=0)
    pair.close()
    got = _parse_cfg(open(out).read())
    want = _parse_cfg(g.cfg_text.tobytes().decode())
    assert got[0] == want[0] and got[1] == want[1] and got[2] == want[2] and got[5:] == want[5:]
    for a, b in zip(got[3], want[3]):
        assert a[:5] == b[:5]                                   # id, type, coordinates: identical text
        if g.mode == "nbh":
            assert abs(float(a[5]) - float(b[5])) <= 1e-5 + 1e-9 * abs(float(b[5]))
    assert abs(float(got[4].split()[-1]) - float(want[4].split()[-1])) <= 1e-6 + 1e-9 * g.max_grade
    # break threshold: the run is aborted after the block is flushed
    pair = PairB200(style, g.path, out, "0.0", "1.0", "chunksize", "32768", species=S)
    pair.set_domain(g.box, len(g.ilist))
    with pytest.raises(LammpsError, match="Exceeded Break Threshold"):
        pair.compute(g.x, g.type, g.nlocal, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=1, vflag=0)
    assert open(out).read().rstrip().endswith("END_CFG")
