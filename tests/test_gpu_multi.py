"""Multi-GPU parity (NCCL, one rank per GPU): decomposed CUDA evaluation == undecomposed oracle evaluation."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from util import TOL_AUX, TOL_E_REL, TOL_F_MAXABSREL

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, tmp_path, kind="direct", worker="_mp_gpu_worker.py"):
    out = str(tmp_path / "res.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, worker), out, str(tmp_path), kind]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    return np.load(out)


def _check(z):
    fmax = np.abs(z["fref"]).max()
    assert np.abs(z["f"] - z["fref"]).max() <= TOL_F_MAXABSREL * fmax
    assert abs(z["ev"][0] - z["evref"][0]) <= TOL_E_REL * abs(z["evref"][0])
    assert np.abs(z["ev"][1:7] - z["evref"][1:7]).max() <= TOL_AUX * np.abs(z["evref"][1:7]).max()
    assert np.abs(z["eatom"] - z["eatomref"]).max() <= TOL_AUX * np.abs(z["eatomref"]).max()
    assert z["launches"][0] > 0          # the halo went through the CUDA pack / unpack kernels


@pytest.mark.parametrize("kind", ["staged", "direct"])
def test_one_rank_device_halo(tmp_path, built, kind):
    """world = 1: periodic self-images through the device pack / unpack kernels."""
    _check(_run(1, tmp_path, kind))


@pytest.mark.parametrize("kind", ["staged", "direct", "overlap"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_decomposed_cuda_equals_global(tmp_path, built, world, kind):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _check(_run(world, tmp_path, kind))


@pytest.mark.parametrize("world", [1, 2])
def test_decomposed_nve_conserves_energy(tmp_path, built, world):
    """Device-resident NVE loop on the brick decomposition (NCCL ghost halo, collective re-neighboring decision,
    device list rebuilds): the all-reduced total energy is conserved."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    z = _run(world, tmp_path, worker="_mp_gpu_md_worker.py")
    drift = np.abs(z["es"] - z["e0"]).max() / int(z["natoms"])
    ke_atom = float(z["ke"]) / int(z["natoms"])
    assert int(z["rebuilds"]) >= 2
    assert drift < 1e-4 * ke_atom, (drift, ke_atom)
