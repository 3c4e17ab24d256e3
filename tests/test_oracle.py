"""CPU tests of the parity checkers themselves (no GPU).

* the C restatement (oracle/mtp_oracle.c) against tests/golden/ -- outputs of the reference's own
  unmodified CPU sources (pair_mtp.cpp / pair_mtp_extrapolation.cpp), see golden/make_golden.py;
* when oracle/_ref/libmtp_ref.so is present (build container), the restatement against the reference
  directly on fresh seeded inputs;
* the physical invariants of SURVEY.md section 4 / App. A.5 (rotation invariance, finite differences,
  Newton's third law, virial identity).
"""
import os

import numpy as np
import pytest

import golden_util
import util
from util import maxabsrel

# The restatement follows the reference's expression order and both are compiled -O2 -ffp-contract=off,
# so agreement is expected to the last bit; 1e-14 leaves room for libm differences between boxes only.
TIGHT = 1e-14


@pytest.mark.parametrize("name", golden_util.NAMES)
def test_restatement_matches_reference_golden(tmp_path, built, name):
    from oracle_py import OracleMTP
    g = golden_util.Golden(name, tmp_path)
    grade = g.mode in ("nbh", "cfg")
    r = OracleMTP(g.pot).compute(g.x, g.type, g.ilist, g.numneigh, g.neigh, g.offsets, eflag=3, vflag=5, grade=grade,
                                 natoms_total=len(g.ilist), want_mask=True)
    assert abs(r.energy - g.energy) <= TIGHT * abs(g.energy)
    assert maxabsrel(r.f, g.f) <= TIGHT
    assert maxabsrel(r.virial, g.virial) <= TIGHT
    assert maxabsrel(r.eatom, g.eatom) <= TIGHT
    assert maxabsrel(r.vatom, g.vatom) <= TIGHT
    # neighbor indexing / cutoff mask: bit-exact
    assert np.array_equal(r.mask[: g.mask.size], g.mask)
    if g.mode == "nbh":
        assert maxabsrel(r.grades[: g.nlocal], g.grades[: g.nlocal]) <= 1e-13
    if grade:
        assert abs(r.max_grade - g.max_grade) <= 1e-13 * abs(g.max_grade)
    if g.mode == "cfg":
        q = g.pot.coeff_count
        assert maxabsrel(r.candidate[:q], g.candidate[:q]) <= 1e-13


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref", "libmtp_ref.so")),
                    reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("level,species", [(6, 1), (14, 2), (18, 1)])
def test_restatement_matches_reference_live(tmp_path, built, level, species):
    from oracle_py import OracleMTP, ReferenceMTP
    path, pot = util.write_potential(tmp_path, level, species, seed=100 + level)
    sysm = util.small_system("fcc", 3.9, (4, 4, 4), species, seed=3)
    il = sysm.ilist[:40]
    a = OracleMTP(pot).compute(sysm.x, sysm.type, il, sysm.numneigh, sysm.neigh, sysm.offsets)
    ref = ReferenceMTP("mtp", path)
    b = ref.compute(sysm.x, sysm.type, sysm.nlocal, il, sysm.numneigh, sysm.neigh, sysm.offsets)
    ref.close()
    assert a.energy == b.energy
    assert np.array_equal(a.f, b.f)
    assert np.array_equal(a.virial, b.virial)


def _cluster(seed=0, n=15):
    rng = np.random.default_rng(seed)
    pts = [np.zeros(3)]
    while len(pts) < n:
        p = rng.normal(size=3)
        p *= rng.uniform(2.2, 5.5) / np.linalg.norm(p)
        if all(np.linalg.norm(p - q) > 1.8 for q in pts):
            pts.append(p)
    return np.array(pts)


def _eval_full(orc, x, types):
    """Every atom a centre, everyone listed as everyone's neighbor (the cutoff mask does the rest)."""
    n = len(x)
    il = np.arange(n, dtype=np.int32)
    nn = np.full(n, n - 1, dtype=np.int32)
    off = np.arange(n + 1, dtype=np.int64) * (n - 1)
    neigh = np.concatenate([np.delete(il, i) for i in range(n)]).astype(np.int32)
    return orc.compute(x, types, il, nn, neigh, off)


@pytest.mark.parametrize("level,species", [(8, 1), (12, 2)])
def test_invariants(tmp_path, built, level, species):
    from oracle_py import OracleMTP
    _, pot = util.write_potential(tmp_path, level, species)
    orc = OracleMTP(pot)
    x = _cluster(level) + 20.0
    types = ((np.arange(len(x)) % species) + 1).astype(np.int32)
    r = _eval_full(orc, x, types)
    fmax = np.abs(r.f).max()
    # Newton's third law
    assert np.abs(r.f.sum(axis=0)).max() <= 1e-12 * fmax
    # rotation invariance of the energy, covariance of the forces
    q, _ = np.linalg.qr(np.random.default_rng(1).normal(size=(3, 3)))
    r2 = _eval_full(orc, (x - 20.0) @ q.T + 20.0, types)
    assert abs(r2.energy - r.energy) <= 1e-12 * abs(r.energy)
    assert maxabsrel(r2.f, r.f @ q.T) <= 1e-11
    # forces = -dE/dx by central differences
    h = 1e-5
    for (i, c) in [(0, 0), (3, 1), (7, 2)]:
        xp, xm = x.copy(), x.copy()
        xp[i, c] += h
        xm[i, c] -= h
        fd = -(_eval_full(orc, xp, types).energy - _eval_full(orc, xm, types).energy) / (2 * h)
        assert abs(fd - r.f[i, c]) <= 2e-7 * fmax
    # pairwise virial -sym(F (x) r) summed over pairs == sym(sum_i x_i (x) f_i)
    w = x.T @ r.f
    v = np.array([w[0, 0], w[1, 1], w[2, 2], (w[0, 1] + w[1, 0]) / 2, (w[0, 2] + w[2, 0]) / 2, (w[1, 2] + w[2, 1]) / 2])
    assert maxabsrel(r.virial, v) <= 1e-11
    # per-atom energies / virials sum to the totals
    assert abs(r.eatom.sum() - r.energy) <= 1e-12 * abs(r.energy)
    assert maxabsrel(r.vatom.sum(axis=0), r.virial) <= 1e-12


def test_chebyshev_matches_closed_form(built, tmp_path):
    """phi_n(d) = s T_n(xi) (d - r_max)^2 (mtp_rb_chevbyshev_basis.cpp:29-54)."""
    from oracle_py import OracleMTP
    _, pot = util.write_potential(tmp_path, 8, 1)
    orc = OracleMTP(pot)
    for d in (2.0, 2.7, 3.9, 4.999, 5.0):
        v, dv = orc.chebyshev(d)
        xi = (2 * d - (pot.min_dist + pot.max_dist)) / (pot.max_dist - pot.min_dist)
        n = np.arange(pot.radial_basis_size)
        tn = np.cos(n * np.arccos(np.clip(xi, -1, 1)))
        assert np.allclose(v, pot.scaling * tn * (d - pot.max_dist) ** 2, rtol=1e-11, atol=1e-13)
        h = 1e-6
        vp, _ = orc.chebyshev(d + h)
        vm, _ = orc.chebyshev(d - h)
        assert np.allclose(dv, (vp - vm) / (2 * h), rtol=1e-6, atol=1e-7)


def test_grade_is_maxabs_matvec(built):
    import ctypes as C

    from oracle_py import ORACLE_SO
    lib = C.CDLL(ORACLE_SO)
    lib.mtp_oracle_grade.restype = C.c_double
    rng = np.random.default_rng(0)
    q = 37
    a = np.ascontiguousarray(rng.normal(size=(q, q)))
    b = np.ascontiguousarray(rng.normal(size=q))
    dp = C.POINTER(C.c_double)
    g = lib.mtp_oracle_grade(a.ctypes.data_as(dp), b.ctypes.data_as(dp), C.c_int(q))
    assert abs(g - np.abs(a @ b).max()) <= 1e-13 * g
