"""Multi-rank host logic on CPU: brick decomposition + forward/reverse ghost halo + EV all-reduce over gloo,
checked against the undecomposed global system."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("kind", ["staged", "direct", "overlap", "cfg"])
@pytest.mark.parametrize("world", [2, 4])
def test_decomposed_equals_global(tmp_path, built, world, kind):
    out = str(tmp_path / "res.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "_mp_decomp_worker.py"), out, kind]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    z = np.load(out)
    fmax = np.abs(z["fref"]).max()
    assert np.abs(z["f"] - z["fref"]).max() <= 1e-12 * fmax
    assert abs(z["ev"][0] - z["evref"][0]) <= 1e-12 * abs(z["evref"][0])
    assert np.abs(z["ev"][1:7] - z["evref"][1:7]).max() <= 1e-11 * np.abs(z["evref"][1:7]).max()
    assert np.abs(z["eatom"] - z["eatomref"]).max() <= 1e-12 * np.abs(z["eatomref"]).max()
    assert z["ghosts"][0] > 0 and z["halo_bytes"][0] > 0
    if kind == "cfg":      # grade of the summed candidate == grade of the undecomposed configuration
        assert z["cfg"][1] > 0 and abs(z["cfg"][0] - z["cfg"][1]) <= 1e-11 * z["cfg"][1]


def test_single_rank_direct_halo(built):
    """world = 1, single-stage halo: same ghost set as the staged scheme, forward/reverse are self-images."""
    import torch

    from mtp_b200 import decomp, harness
    sysm, halo = decomp.make_rank_system(1, (3, 3, 3), (1, 1, 1), 0, torch.device("cpu"), direct=True)
    ref = harness.make_config(1, cells=(3, 3, 3))
    assert sysm.nall == ref.nall
    key = lambda a: a[np.lexsort(np.round(a, 9).T)]  # noqa: E731
    assert np.allclose(key(sysm.x), key(ref.x))
    assert np.array_equal(np.sort(sysm.numneigh[: sysm.nlocal]), np.sort(ref.numneigh[: ref.nlocal]))
    x = torch.from_numpy(sysm.x.copy())
    x[sysm.nlocal:] = 0.0
    halo.forward(x)
    assert np.array_equal(x.numpy(), sysm.x)
    f = torch.from_numpy(np.random.default_rng(0).normal(size=(sysm.nall, 3)))
    tot = f.sum(dim=0).clone()
    halo.reverse(f)
    assert np.allclose(f[: sysm.nlocal].sum(dim=0).numpy(), tot.numpy(), atol=1e-10)


def test_single_rank_halo_is_periodic_self_image(built):
    """world = 1: all six swaps are self-copies and reproduce harness.add_ghosts / System.reverse_comm."""
    import torch

    from mtp_b200 import decomp, harness
    sysm, halo = decomp.make_rank_system(1, (3, 3, 3), (1, 1, 1), 0, torch.device("cpu"))
    ref = harness.make_config(1, cells=(3, 3, 3))
    assert sysm.nall == ref.nall and np.array_equal(sysm.x, ref.x) and np.array_equal(sysm.neigh, ref.neigh)
    x = torch.from_numpy(sysm.x.copy())
    x[sysm.nlocal:] = 0.0
    halo.forward(x)
    assert np.array_equal(x.numpy(), ref.x)
    f = torch.from_numpy(np.random.default_rng(0).normal(size=(sysm.nall, 3)))
    want = ref.reverse_comm(f.numpy().copy())
    halo.reverse(f)
    assert np.allclose(f[: sysm.nlocal].numpy(), want, rtol=0, atol=1e-14)
