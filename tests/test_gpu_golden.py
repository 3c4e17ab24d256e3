"""GPU parity against tests/golden/: the CUDA path (through the C ABI) vs what the reference's own unmodified
CPU sources returned on the same inputs (fixtures made by golden/make_golden.py in the build container)."""
import numpy as np
import pytest

import golden_util
from util import TOL_AUX, TOL_E_REL, TOL_F_MAXABSREL, maxabsrel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 1], ids=["large", "small"])
@pytest.mark.parametrize("name", golden_util.NAMES)
def test_cuda_matches_reference_golden(tmp_path, name, variant):
    from mtp_b200.api import MTPB200
    g = golden_util.Golden(name, tmp_path)
    grade = g.mode in ("nbh", "cfg")
    mtp = MTPB200(g.path, selection_state=grade)
    r = mtp.compute_host(g.x, g.type, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=3, vflag=5, grade=grade,
                         natoms_total=len(g.ilist), want_mask=True, variant=variant)
    mtp.close()
    assert abs(r.energy - g.energy) <= TOL_E_REL * abs(g.energy)          # <= 1e-10 relative (north_star)
    assert maxabsrel(r.f, g.f) <= TOL_F_MAXABSREL                          # <= 1e-9 max-abs-relative (north_star)
    assert maxabsrel(r.virial, g.virial) <= TOL_AUX
    assert maxabsrel(r.eatom, g.eatom) <= TOL_AUX
    assert maxabsrel(r.vatom, g.vatom) <= TOL_AUX
    assert np.array_equal(r.mask[: g.mask.size], g.mask)                   # neighbor indexing: bit-exact
    if g.mode == "nbh":
        assert maxabsrel(r.grades[: g.nlocal], g.grades[: g.nlocal]) <= TOL_AUX
    if grade:
        assert abs(r.max_grade - g.max_grade) <= TOL_AUX * abs(g.max_grade)
    if g.mode == "cfg":
        q = g.pot.coeff_count
        assert maxabsrel(r.candidate[:q], g.candidate[:q]) <= TOL_AUX
