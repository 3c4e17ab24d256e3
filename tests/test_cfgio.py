"""MLIP-3 .cfg reader against the text the REFERENCE's own writer produced (tests/golden/*.npz: cfg_text comes from
PairMTPExtrapolation::write_config of oracle/_ref, pair_mtp_extrapolation.cpp:401-479)."""
import os

import numpy as np
import pytest

from mtp_b200 import cfgio

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden(name):
    z = np.load(os.path.join(HERE, "golden", name), allow_pickle=True)
    t = z["cfg_text"]
    return z, (t.item() if t.shape == () else bytes(t.astype(np.uint8)).decode())


@pytest.mark.parametrize("name,has_grades", [("L10_S2_nbh.npz", True), ("L10_S2_cfg.npz", False)])
def test_reader_inverts_the_reference_writer(name, has_grades):
    z, text = _golden(name)
    cfgs = cfgio.parse_cfg(text)
    assert len(cfgs) >= 1
    c = cfgs[0]
    nlocal = int(z["nlocal"])
    assert c.size == nlocal and c.positions.shape == (nlocal, 3)
    assert np.array_equal(c.ids, np.arange(1, nlocal + 1))                      # global ids start at 1 (:421)
    assert np.array_equal(c.types, z["type"][:nlocal] - 1)                      # 0-based species (:420)
    assert np.abs(c.positions - z["x"][:nlocal]).max() <= 0.5e-6 + 1e-12        # "{:.6f}"
    assert np.allclose(np.diag(c.supercell), z["box"], atol=0.5e-6) and c.supercell.shape == (3, 3)
    assert abs(float(c.features["MV_grade"]) - float(z["max_grade"])) <= 0.5e-6 * max(1.0, abs(float(z["max_grade"])))
    if has_grades:
        g = z["grades"][:nlocal]
        assert np.abs(c.nbh_grades - g).max() <= 0.5e-5 + 1e-9 * np.abs(g).max()   # "{:.5f}"
    else:
        assert c.nbh_grades is None
    assert c.energy is None and c.forces is None


def test_replay_of_a_selected_configuration_rebuilds_the_same_neighborhoods():
    """A configuration read back becomes a harness system with the same atoms and the same neighbor counts."""
    z, text = _golden("L10_S2_nbh.npz")
    c = cfgio.parse_cfg(text)[0]
    sysm = c.to_system(cutoff=5.0, skin=2.0)
    nlocal = int(z["nlocal"])
    assert sysm.nlocal == nlocal and np.array_equal(sysm.type[:nlocal], z["type"][:nlocal])
    # positions were rounded to 1e-6 by the writer: the neighbor COUNTS within cutoff + skin survive for this lattice
    assert np.array_equal(sysm.numneigh[:nlocal], z["numneigh"][:nlocal])


def test_general_mlip_blocks_and_error_reporting(tmp_path):
    text = ("BEGIN_CFG\n Size\n    2\n Supercell\n    4.0 0.0 0.0\n    0.0 4.0 0.0\n    0.0 0.0 4.0\n"
            " AtomData:  id type cartes_x cartes_y cartes_z fx fy fz\n"
            "    1 0 0.0 0.0 0.0 0.1 0.2 0.3\n    2 1 2.0 2.0 2.0 -0.1 -0.2 -0.3\n"
            " Energy\n    -7.5\n PlusStress:  xx yy zz yz xz xy\n    1 2 3 4 5 6\n Feature   EFS_by  VASP\nEND_CFG\n\n"
            "BEGIN_CFG\nSize\n1\nSupercell\n3 0 0\nAtomData: id type cartes_x cartes_y cartes_z\n1 0 0 0 0\nEND_CFG\n")
    p = tmp_path / "two.cfg"
    p.write_text(text)
    a, b = cfgio.read_cfg(str(p))
    assert a.energy == -7.5 and np.array_equal(a.stress, [1, 2, 3, 4, 5, 6]) and a.features["EFS_by"] == "VASP"
    assert np.allclose(a.forces, [[0.1, 0.2, 0.3], [-0.1, -0.2, -0.3]])
    assert b.size == 1 and b.supercell.shape == (1, 3)
    with pytest.raises(ValueError):
        b.to_system(5.0)                       # not a 3-D orthorhombic cell
    with pytest.raises(cfgio.CfgError, match="END_CFG"):
        cfgio.parse_cfg("BEGIN_CFG\nSize\n0\n")
    with pytest.raises(cfgio.CfgError, match="fields"):
        cfgio.parse_cfg("BEGIN_CFG\nSize\n1\nAtomData: id type cartes_x cartes_y cartes_z\n1 0 0 0\nEND_CFG\n")
    with pytest.raises(cfgio.CfgError, match="unknown keyword"):
        cfgio.parse_cfg("BEGIN_CFG\nSize\n0\nBogus\nEND_CFG\n")
