import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lammps-mtp-kokkos_b200"), os.path.join(ROOT, "oracle"), ROOT, os.path.dirname(__file__)):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests must fail loudly on a GPU box, and are simply deselected by -m "not gpu" elsewhere.
    pass


@pytest.fixture(scope="session")
def built():
    """Build the checkers (gcc) once per session; the CUDA library is built by __graft_entry__.build()."""
    import oracle_py
    oracle_py.build()
    from mtp_b200 import harness
    harness.build_harness()
    return True
