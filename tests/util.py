"""Shared helpers for the parity tests: seeded potentials, small systems, error metrics."""
import os

import numpy as np

from mtp_b200 import almtp, harness

# tolerances stated by BASELINE.json north_star
TOL_E_REL = 1e-10       # relative, total energy
TOL_F_MAXABSREL = 1e-9  # max |dF| / max |F|
TOL_AUX = 1e-9          # virial / eatom / vatom / grades, max-abs relative to the largest entry


def maxabsrel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = max(np.abs(b).max(), 1e-300) if b.size else 1.0
    return float(np.abs(a - b).max() / den) if b.size else 0.0


def write_potential(tmpdir, level, species, *, seed=None, active_set=False, cfg_mode=False, name=None, **kw):
    pot = almtp.random_potential(level, species, seed, with_active_set=active_set, configuration_mode=cfg_mode, **kw)
    path = os.path.join(str(tmpdir), name or f"L{level}_S{species}{'_as' if active_set else ''}{'_cfg' if cfg_mode else ''}.almtp")
    almtp.write_almtp(path, pot)
    return path, almtp.read_almtp(path)


def small_system(kind, a, cells, species, *, seed=7, jitter=0.05, cutoff=5.0, skin=2.0):
    x, box = harness.lattice(kind, a, cells, jitter=jitter)
    types = harness.random_types(len(x), [1.0] * species, seed)
    return harness.make_system(x, types, box, cutoff, skin)


def random_cluster(n, species, *, seed=1, extent=12.0, dmin=1.9, cutoff=5.0, skin=2.0):
    """Non-periodic blob: ragged neighbor counts, some atoms with no neighbor inside the cutoff."""
    rng = np.random.default_rng(seed)
    pts = []
    while len(pts) < n:
        p = rng.uniform(0, extent, size=3)
        if all(np.linalg.norm(p - q) > dmin for q in pts):
            pts.append(p)
    x = np.array(pts)
    # two far-away atoms: one isolated, a pair only inside the skin shell (listed but outside the cutoff)
    x = np.vstack([x, [extent + 30, 0, 0], [extent + 60, 0, 0], [extent + 60 + cutoff + 0.5 * skin, 0, 0]])
    types = harness.random_types(len(x), [1.0] * species, seed)
    box = np.array([1e6, 1e6, 1e6])
    return harness.make_system(x, types, box, cutoff, skin, periodic=(False, False, False))
