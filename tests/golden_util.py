"""Loader for tests/golden/*.npz -- outputs of the reference's own CPU sources (see golden/make_golden.py)."""
import glob
import os

import numpy as np

from mtp_b200 import almtp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    def __init__(self, name, tmpdir):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.mode = str(z["mode"])
        self.path = os.path.join(str(tmpdir), name + ".almtp")
        with open(self.path, "wb") as fh:
            fh.write(z["potential"].tobytes())
        self.pot = almtp.read_almtp(self.path)
        for k in ("x", "type", "box", "ilist", "numneigh", "offsets", "neigh", "virial", "f", "eatom", "vatom", "mask",
                  "grades", "candidate", "cfg_text"):
            setattr(self, k, z[k])
        self.nlocal = int(z["nlocal"])
        self.energy = float(z["energy"])
        self.max_grade = float(z["max_grade"])
        self.nall = self.x.shape[0]
