"""CPU tests of the contraction-program code generator (mtp_codegen.cpp): the emitted kernel source is compiled for
the HOST and executed stage by stage / warp by warp / lane by lane against the reference's sequential program
(pair_mtp.cpp:196-233, restated in tests/shim/p4_host_check.cpp).  No GPU."""
import os
import subprocess

import numpy as np
import pytest

import util
from mtp_b200 import almtp, api, mtp_basis

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "tests", "shim", "p4_host_check.cpp")


def _write_tables(path, pot):
    times = np.asarray(pot.alpha_index_times, dtype=np.int64).reshape(-1, 4)
    with open(path, "w") as f:
        f.write(f"{pot.K} {pot.alpha_moments_count} {len(times)} {pot.A}\n")
        f.write(" ".join(str(int(v)) for v in times.ravel()) + "\n")
        f.write(" ".join(str(int(v)) for v in pot.alpha_moment_mapping) + "\n")
        f.write(" ".join(repr(float(v)) for v in pot.moment_coeffs) + "\n")


def _host_check(tmp_path, pot_path, pot, latency, env=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        src, info = api.codegen_source(pot_path, latency)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    gen = os.path.join(str(tmp_path), "gen_p4.cu")
    open(gen, "w").write(src)
    tables = os.path.join(str(tmp_path), "tables.txt")
    _write_tables(tables, pot)
    exe = os.path.join(str(tmp_path), "p4_host_check")
    two = ["-DP4_TWO"] if "#define P4_APL 2" in src else []    # two atoms per lane
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", f'-DP4_SOURCE="{gen}"'] + two + ["-o", exe, HARNESS], check=True)
    r = subprocess.run([exe, tables], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return info


@pytest.mark.parametrize("level,species,latency", [(8, 1, False), (10, 2, False), (12, 3, False), (16, 2, False),
                                                   (20, 1, True), (22, 1, True), (22, 1, False)])
def test_generated_program_matches_the_sequential_program(tmp_path, level, species, latency):
    path, pot = util.write_potential(tmp_path, level, species)
    info = _host_check(tmp_path, path, pot, latency)
    tb = mtp_basis.build_mtp_tables(level)
    slack = 1.0 if info["rounds"] == 1 else 1.2                    # rounds recompute shared intermediates
    assert info["terms"] <= slack * 3 * len(tb.alpha_index_times)   # forward T + reverse 2T, squares merged
    assert info["loads"] < info["terms"] or level <= 10            # the register cache removes most operand loads


def test_program_too_large_for_one_cta_is_evaluated_in_rounds(tmp_path):
    """Level 20: the rows of the whole program exceed a CTA's shared memory at 32 atoms per row, so the basis functions
    are dealt to rounds that reuse the rows; the emulation checks the sum of the rounds against the sequential program."""
    path, pot = util.write_potential(tmp_path, 20, 1)
    info = _host_check(tmp_path, path, pot, False)
    assert info["atoms_per_cta"] == 32 and info["rounds"] > 1 and info["smem_bytes"] <= 232448
    tb = mtp_basis.build_mtp_tables(20)
    assert info["terms"] <= 1.2 * 3 * len(tb.alpha_index_times)      # shared intermediates recomputed: < 20 % extra work
    assert api.codegen_source(path, True)[1]["atoms_per_cta"] == 16


def test_rounds_on_a_small_program(tmp_path):
    path, pot = util.write_potential(tmp_path, 12, 2)
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "32,4,24,8,2,18000"})
    assert info["rounds"] >= 3
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "64,8,24,8,1,36000"})
    assert info["rounds"] >= 3 and info["atoms_per_cta"] == 64


def test_sparse_rounds_and_atom_groups(tmp_path):
    """Sparse rounds (a round stages only the basic moments it reads; later rounds ADD their adjoint shares) and atom groups
    (every emitted function runs once per group of 32 atoms); the emulation plays one group, round by round, with every row
    stale at a round's start."""
    path, pot = util.write_potential(tmp_path, 12, 2)
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "32,4,40,12,4,18000,1,4000,1"})    # sparse, one group
    assert info["rounds"] >= 2 and info["atoms_per_cta"] == 32 and info["rows"] * 32 * 8 <= 18000
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "32,4,24,8,2,30000,2"})            # two groups
    assert info["rounds"] >= 2 and info["atoms_per_cta"] == 64
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "32,4,24,8,1,60000,4,60"})         # four groups, small functions
    assert info["atoms_per_cta"] == 128
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "64,8,24,8,1,60000,2"})            # two atoms per lane x two groups
    assert info["atoms_per_cta"] == 128 and info["warps"] == 8


def test_rounds_in_parallel_for_the_latency_shape(tmp_path):
    """mtp/small/kk: the rounds of a large program run as separate CTAs (gridDim.y = rounds); every round ADDS its adjoint
    shares to a zeroed gb, none of them is "the first" (the emulation plays the rounds in order on a zeroed gb)."""
    path, pot = util.write_potential(tmp_path, 12, 2)
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "16,4,24,8,4,9000,1,4000,1,0,1"})
    assert info["rounds"] >= 2 and info["atoms_per_cta"] == 16
    path, pot = util.write_potential(tmp_path, 20, 1)
    os.environ["MTP_B200_P4_RPAR"] = "1"                            # opt-in: measured slower than one CTA per chunk
    try:
        src, info = api.codegen_source(path, True)
    finally:
        del os.environ["MTP_B200_P4_RPAR"]
    assert "#define P4_RPAR 1" in src and 2 <= info["rounds"] <= 8 and info["atoms_per_cta"] == 16
    assert "\n  GBST(" not in src and "\n  GBACC(" in src           # no plain adjoint store in any stage function
    src, _ = api.codegen_source(path, False)
    assert "#define P4_RPAR 0" in src                              # the throughput shape keeps its rounds in sequence
    src, info = api.codegen_source(path, True)
    assert "P4_RPAR 1" not in src and info["rounds"] == 1          # the default latency shape: one CTA per chunk, one round


def test_throughput_shape_prefers_resident_ctas(tmp_path):
    """Level 16 (config 2): four 4-warp CTAs per SM in sparse rounds; level 22 (config 5): two 8-warp CTAs per SM."""
    path, pot = util.write_potential(tmp_path, 16, 2)
    info = _host_check(tmp_path, path, pot, False)
    assert (info["atoms_per_cta"], info["warps"], info["ctas_per_sm"]) == (32, 4, 4) and info["smem_bytes"] <= 57344
    path, pot = util.write_potential(tmp_path, 22, 1)
    src, info = api.codegen_source(path, False)
    assert (info["atoms_per_cta"], info["warps"], info["ctas_per_sm"]) == (32, 8, 2) and info["smem_bytes"] <= 115712
    assert "red.global.add.f64" in src and "#define P4_SPARSE 1" in src
    tb = mtp_basis.build_mtp_tables(22)
    assert info["terms"] <= 1.2 * 3 * len(tb.alpha_index_times)


def test_latency_shape_and_two_atoms_per_lane(tmp_path):
    path, pot = util.write_potential(tmp_path, 12, 2)
    info = _host_check(tmp_path, path, pot, True)
    assert info["atoms_per_cta"] == 16
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "64,8,24,8,1"})
    assert info["atoms_per_cta"] == 64 and info["warps"] == 8
    info = _host_check(tmp_path, path, pot, False, env={"MTP_B200_P4": "16,3,6,2,1"})    # tiny cache, odd warp count
    assert info["atoms_per_cta"] == 16 and info["warps"] == 3


def test_permuted_tables_generate_a_correct_program(tmp_path):
    """A file whose basic moments and products come in another order (as real MLIP-3 files may) is still handled."""
    path, pot = util.write_potential(tmp_path, 12, 1)
    rng = np.random.default_rng(5)
    K = pot.K
    perm = rng.permutation(K)                    # new index of old basic k
    basic = np.asarray(pot.alpha_index_basic).reshape(K, 4)
    new_basic = np.zeros_like(basic)
    new_basic[perm] = basic
    ren = np.arange(pot.alpha_moments_count)
    ren[:K] = perm
    times = np.asarray(pot.alpha_index_times).reshape(-1, 4).copy()
    times[:, 0], times[:, 1], times[:, 3] = ren[times[:, 0]], ren[times[:, 1]], ren[times[:, 3]]
    # shuffle the products within each dependency wave (any order that writes before it reads is a valid file)
    waves = mtp_basis.prepare_waves(pot.alpha_index_times, K)
    out, at = [], 0
    for n in waves:
        blk = times[at:at + n]
        out.append(blk[rng.permutation(n)])
        at += n
    pot.alpha_index_basic = new_basic.astype(np.int32)
    pot.alpha_index_times = np.concatenate(out).astype(np.int32)
    pot.alpha_moment_mapping = np.array([int(ren[m]) for m in pot.alpha_moment_mapping], dtype=np.int32)
    path2 = os.path.join(str(tmp_path), "permuted.almtp")
    almtp.write_almtp(path2, pot)
    pot2 = almtp.read_almtp(path2)
    _host_check(tmp_path, path2, pot2, False)


def test_prebuild_compiles_with_nvrtc_and_caches(tmp_path, monkeypatch):
    """NVRTC needs no GPU: the cubin of a small potential is built here and found in the cache the second time."""
    monkeypatch.setenv("MTP_B200_KCACHE", str(tmp_path / "kcache"))
    path, _ = util.write_potential(tmp_path, 8, 1)
    assert api.codegen_prebuild(path) is True
    assert api.codegen_prebuild(path) is False
    files = os.listdir(tmp_path / "kcache")
    assert len(files) == 1 and files[0].endswith(".cubin")
