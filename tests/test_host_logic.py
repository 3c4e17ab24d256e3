"""CPU tests of the host side (no GPU): C-ABI surface, .almtp parser, basis-table generator, wave scheduler."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from mtp_b200 import almtp, api, mtp_basis

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libmtp_ref.so")


def test_c_abi_exports_every_declared_symbol():
    """Every function declared in include/*.h is exported by the shared library (loads without a GPU)."""
    lib = api.load_library()
    declared = set()
    for hdr in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(mtp_[a-z0-9_]+)\s*\(", text))
    assert {"mtp_create_from_file", "mtp_compute", "mtp_compute_host", "mtp_destroy", "mtp_last_error",
            "mtp_halo_pack_x", "mtp_halo_unpack_add_f", "mtp_potential_check"} <= declared
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"


def test_no_gpu_means_loud_failure_not_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    path, _ = util.write_potential(tmp_path, 8, 1)
    with pytest.raises(api.MTPError, match="no CPU fallback|CUDA"):
        api.MTPB200(path)


@pytest.mark.parametrize("level,species,sel", [(8, 1, False), (16, 2, True), (22, 3, False)])
def test_parser_reads_what_the_writer_wrote(tmp_path, level, species, sel):
    path, pot = util.write_potential(tmp_path, level, species, active_set=sel)
    info = api.potential_check(path, sel)
    assert info.species_count == species
    assert info.radial_func_count == pot.radial_funcs_count
    assert info.radial_basis_size == pot.radial_basis_size
    assert info.alpha_moment_count == pot.alpha_moments_count
    assert info.alpha_index_basic_count == pot.K
    assert info.alpha_index_times_count == pot.T
    assert info.alpha_scalar_count == pot.A
    assert info.max_alpha_index_basic == pot.max_alpha_index_basic
    assert info.coeff_count == (pot.coeff_count if sel else 0)
    assert info.has_selection_state == int(sel)
    assert (info.min_cutoff, info.max_cutoff, info.scaling) == (pot.min_dist, pot.max_dist, pot.scaling)
    assert info.wave_count == len(mtp_basis.build_mtp_tables(level).wave_sizes)


def _mutate(path, out, fn):
    raw = open(path, "rb").read()
    open(out, "wb").write(fn(raw))
    return out


BAD_FILES = [
    ("not_mtp", lambda r: r.replace(b"MTP\n", b"XTP\n", 1), "Only MTP potential files are accepted."),
    ("version", lambda r: r.replace(b"version = 1.1.0", b"version = 1.0.0", 1), 'MTP file must have version "1.1.0"'),
    ("version_spacing", lambda r: r.replace(b"version = 1.1.0", b"version=1.1.0", 1), 'MTP file must have version "1.1.0"'),
    ("no_species", lambda r: r.replace(b"species_count", b"species_kount", 1), "Species count not found"),
    ("rb_type", lambda r: r.replace(b"RBChebyshev", b"RBShapeev", 1), "RBShapeev"),
    ("magnetic", lambda r: r.replace(b"\tradial_coeffs\n", b"\tmagnetic_basis_type = x\n\tradial_coeffs\n", 1),
     "Magnetic basis is currently not supported."),
    ("radial_max", lambda r: r.replace(b"radial_funcs_count = 2", b"radial_funcs_count = 3", 1), None),
    ("truncated", lambda r: r[: r.index(b"alpha_index_times_count")], "Unexpected end of file"),
]


@pytest.mark.parametrize("name,fn,msg", BAD_FILES, ids=[b[0] for b in BAD_FILES])
def test_parser_rejects_like_the_reference(tmp_path, name, fn, msg):
    """Files the reference rejects (pair_mtp.cpp:352-569) are rejected, with the reference's wording."""
    path, _ = util.write_potential(tmp_path, 8, 1)
    bad = _mutate(path, os.path.join(str(tmp_path), name + ".almtp"), fn)
    with pytest.raises(api.MTPError) as ei:
        api.potential_check(bad)
    if msg:
        assert msg in str(ei.value)
    if os.path.exists(REF_SO) and name != "radial_max":
        from oracle_py import ReferenceMTP
        with pytest.raises(RuntimeError) as er:
            ReferenceMTP("mtp", bad)
        # (the reference dereferences a null line on a truncated file -- any failure counts there)
        if msg and name != "truncated":
            assert msg in str(er.value), str(er.value)


def test_selection_state_grammar(tmp_path):
    """pair_mtp_extrapolation.cpp:545-612: missing state, bad MVS tag, both weights set, weights truncated to int."""
    path, pot = util.write_potential(tmp_path, 8, 2, active_set=True)
    assert api.potential_check(path, True).configuration_mode == 0
    plain, _ = util.write_potential(tmp_path, 8, 2, name="plain.almtp")
    with pytest.raises(api.MTPError, match="No selection state found"):
        api.potential_check(plain, True)
    bad = _mutate(path, str(tmp_path / "mvs.almtp"), lambda r: r.replace(b"#MVS_v1.1", b"#MVS_v1.0", 1))
    with pytest.raises(api.MTPError, match="MVS version"):
        api.potential_check(bad, True)
    both = _mutate(path, str(tmp_path / "both.almtp"), lambda r: r.replace(b"energy_weight = 0.0", b"energy_weight = 1.0", 1))
    with pytest.raises(api.MTPError, match="configuration mode"):
        api.potential_check(both, True)
    # 0.9 truncates to 0 (B8): still neighbourhood mode, and 0.9 + 1 does not trip the "> 1" check
    frac = _mutate(path, str(tmp_path / "frac.almtp"), lambda r: r.replace(b"energy_weight = 0.0", b"energy_weight = 0.9", 1))
    assert api.potential_check(frac, True).configuration_mode == 0
    short = _mutate(path, str(tmp_path / "short.almtp"), lambda r: r[:-64])
    with pytest.raises(api.MTPError, match="binary data"):
        api.potential_check(short, True)
    cfgp, _ = util.write_potential(tmp_path, 8, 2, active_set=True, cfg_mode=True, name="cfg.almtp")
    assert api.potential_check(cfgp, True).configuration_mode == 1


def test_comments_blank_lines_and_mlip2_scaling_placement(tmp_path):
    """'#' comments are ignored (pair_mtp.cpp:347); a scaling line inside the basis block is parsed and then
    overridden by the top-level default 1.0."""
    path, pot = util.write_potential(tmp_path, 8, 1)
    raw = open(path, "rb").read()
    raw = raw.replace(b"species_count = 1\n", b"# a comment line\n\nspecies_count = 1   # trailing comment\n", 1)
    raw = raw.replace(b"scaling = 1.0\n", b"", 1)
    raw = raw.replace(b"radial_basis_type = RBChebyshev\n", b"radial_basis_type = RBChebyshev\n\tscaling = 0.5\n", 1)
    p2 = str(tmp_path / "c.almtp")
    open(p2, "wb").write(raw)
    info = api.potential_check(p2)
    assert info.species_count == 1 and info.scaling == 1.0


def test_level8_known_answer():
    """the generator reproduces the public MLIP level-8 table."""
    t = mtp_basis.build_mtp_tables(8)
    kat = mtp_basis.LEVEL8_KAT
    assert t.radial_funcs_count == kat["radial_funcs_count"]
    assert t.alpha_moments_count == kat["alpha_moments_count"]
    assert sorted(map(tuple, t.alpha_index_basic)) == sorted(map(tuple, kat["alpha_index_basic"]))
    assert len(t.alpha_index_times) == len(kat["alpha_index_times"])
    assert len(t.alpha_moment_mapping) == len(kat["alpha_moment_mapping"])
    assert mtp_basis.prepare_waves(kat["alpha_index_times"], 11) == [11, 2, 1]


def test_scalar_counts_match_published_mlip_levels():
    known = {2: 1, 4: 2, 6: 5, 8: 9, 10: 16, 12: 29, 14: 52, 16: 92, 18: 163, 20: 288, 22: 500}
    for level, count in known.items():
        assert len(mtp_basis.build_mtp_tables(level).alpha_moment_mapping) == count


@pytest.mark.parametrize("level", [8, 12, 16, 20, 22])
def test_generated_programs_are_ordered_and_at_most_three_waves(level):
    t = mtp_basis.build_mtp_tables(level)
    times = np.array(t.alpha_index_times)
    k = len(t.alpha_index_basic)
    written = set(range(k))
    targets_done = set()
    for a0, a1, mult, a3 in times:
        assert a0 in written and a1 in written and mult >= 1
        assert a3 >= k
        written.add(a3)
    waves = mtp_basis.prepare_waves(t.alpha_index_times, k)
    assert len(waves) <= 3 and sum(waves) == len(times)          # pair_mtps_kokkos.cpp:189-194 would abort on a 4th
    assert waves == [w for w in t.wave_sizes if w]


def test_kat_potential_evaluates_like_generated_one(tmp_path, built):
    """The level-8 KAT table and the generated level-8 table span the same basis: with matched coefficients
    they give the same energy (the tables may order scalars differently, so compare the basis VALUES as sets)."""
    from oracle_py import OracleMTP
    gen = almtp.random_potential(8, 1)
    kat = mtp_basis.LEVEL8_KAT
    katp = almtp.MTPPotential(
        species_count=1, min_dist=gen.min_dist, max_dist=gen.max_dist, radial_basis_size=gen.radial_basis_size,
        radial_funcs_count=2, radial_coeffs=gen.radial_coeffs, alpha_moments_count=18,
        alpha_index_basic=np.array(kat["alpha_index_basic"], dtype=np.int32),
        alpha_index_times=np.array(kat["alpha_index_times"], dtype=np.int32),
        alpha_moment_mapping=np.array(kat["alpha_moment_mapping"], dtype=np.int32),
        species_coeffs=np.zeros(1), moment_coeffs=np.zeros(9))
    sysm = util.small_system("fcc", 4.05, (3, 3, 3), 1)
    il = sysm.ilist[:1]

    def basis_values(p):
        vals = []
        for s in range(p.A):
            p.moment_coeffs = np.zeros(p.A)
            p.moment_coeffs[s] = 1.0
            p.species_coeffs = np.zeros(1)
            vals.append(OracleMTP(p).compute(sysm.x, sysm.type, il, sysm.numneigh, sysm.neigh, sysm.offsets).energy)
        return np.sort(np.array(vals))

    a, b = basis_values(gen), basis_values(katp)
    assert np.allclose(a, b, rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("level,atoms_per_cta", [(2, 32), (6, 32), (8, 32), (10, 32), (12, 16), (14, 32), (16, 32), (16, 16),
                                                  (18, 16), (20, 16), (22, 32)])
def test_grouped_program_streams_reproduce_the_sequential_program(tmp_path, level, atoms_per_cta):
    """The stream packer of the 4-atoms-per-lane program kernel (node groups, split long lists, padding, scratch row)
    interpreted on the host against pair_mtp.cpp:196-233 -- forward moments and reverse-mode adjoints."""
    path, _ = util.write_potential(tmp_path, level, 2)
    err = api.program_check(path, atoms_per_cta)
    assert err <= 1e-13, err


def test_md_helpers_velocity_create_and_force_scaling():
    """`velocity all create T seed mom yes` (README.md:148 of the reference) and the one-factor force rescaling."""
    pytest.importorskip("torch")
    from mtp_b200 import almtp
    from mtp_b200.md import BOLTZ, MVV2E, maxwell_velocities, scale_to_rms_force
    types = np.random.default_rng(0).integers(1, 3, size=5000)
    mass1 = np.array([0.0, 183.84, 95.95])
    v = maxwell_velocities(types, mass1, 300.0, 12345)
    m = mass1[types]
    assert np.abs((m[:, None] * v).sum(axis=0)).max() <= 1e-9 * np.abs(m[:, None] * v).sum()
    t = MVV2E * (m[:, None] * v * v).sum() / ((3 * len(types) - 3) * BOLTZ)
    assert abs(t - 300.0) < 1e-9
    assert np.array_equal(v, maxwell_velocities(types, mass1, 300.0, 12345))          # seeded
    assert np.abs(maxwell_velocities(types, mass1, 0.0, 1)).max() == 0.0
    pot = almtp.random_potential(8, 2)
    half = scale_to_rms_force(pot, rms_now=0.1, rms_target=0.05)
    assert np.allclose(half.moment_coeffs, 0.5 * np.asarray(pot.moment_coeffs))
    assert np.allclose(half.species_coeffs, 0.5 * np.asarray(pot.species_coeffs))
    assert np.array_equal(half.radial_coeffs, pot.radial_coeffs)
