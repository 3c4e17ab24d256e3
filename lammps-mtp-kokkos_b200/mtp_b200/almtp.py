"""MLIP-3 ``.almtp`` potential files: writer, reader and a seeded random-potential factory.

Grammar = what the reference parser accepts (pair_mtp.cpp:345-570,
mtp_radial_basis.cpp:59-102) followed, for the extrapolation styles, by the
MaxVol selection state (pair_mtp_extrapolation.cpp:545-612): a ``#MVS_v1.1``
text header, five weight lines, one ``#`` byte, then the raw little-endian
``double[Q*Q]`` active set and its inverse, ``Q = S*S*R*B + S + A``.

The reader here is a convenience for tests and tools; the product's parser is the
C++ one behind ``mtp_create_from_file`` (csrc/mtp_potential.cpp).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field

import numpy as np

from .mtp_basis import MTPTables, build_mtp_tables


@dataclass
class MTPPotential:
    species_count: int
    min_dist: float
    max_dist: float
    radial_basis_size: int
    radial_funcs_count: int
    radial_coeffs: np.ndarray          # [S, S, R, B]
    alpha_moments_count: int
    alpha_index_basic: np.ndarray      # [K, 4] int32
    alpha_index_times: np.ndarray      # [T, 4] int32
    alpha_moment_mapping: np.ndarray   # [A] int32
    species_coeffs: np.ndarray         # [S]
    moment_coeffs: np.ndarray          # [A]
    scaling: float = 1.0
    potential_name: str = "MTP1m"
    potential_tag: str = ""
    # selection state (optional)
    energy_weight: float | None = None
    site_en_weight: float | None = None
    active_set: np.ndarray | None = None           # [Q, Q]
    inverse_active_set: np.ndarray | None = None   # [Q, Q]
    extra: dict = field(default_factory=dict)

    @property
    def K(self):
        return int(self.alpha_index_basic.shape[0])

    @property
    def T(self):
        return int(self.alpha_index_times.shape[0])

    @property
    def A(self):
        return int(self.alpha_moment_mapping.shape[0])

    @property
    def coeff_count(self):
        S, R, B = self.species_count, self.radial_funcs_count, self.radial_basis_size
        return S * S * R * B + S + self.A

    @property
    def max_alpha_index_basic(self):
        return 1 + int(self.alpha_index_basic[:, 1:].sum(axis=1).max())


def _fmt(v: float) -> str:
    return repr(float(v))    # shortest round-trip representation: text -> strtod is exact


def _brace(rows) -> str:
    return "{" + ", ".join("{" + ", ".join(str(int(v)) for v in r) + "}" for r in rows) + "}"


def write_almtp(path: str, p: MTPPotential) -> None:
    S, R, B = p.species_count, p.radial_funcs_count, p.radial_basis_size
    L = ["MTP", "version = 1.1.0", f"potential_name = {p.potential_name}"]
    if p.scaling != 1.0 or p.extra.get("write_scaling", True):
        L.append(f"scaling = {_fmt(p.scaling)}")
    L.append(f"species_count = {S}")
    L.append(f"potential_tag = {p.potential_tag}")
    L.append("radial_basis_type = RBChebyshev")
    L.append(f"\tmin_dist = {_fmt(p.min_dist)}")
    L.append(f"\tmax_dist = {_fmt(p.max_dist)}")
    L.append(f"\tradial_basis_size = {B}")
    L.append(f"\tradial_funcs_count = {R}")
    L.append("\tradial_coeffs")
    for i in range(S):
        for j in range(S):
            L.append(f"\t\t{i}-{j}")
            for mu in range(R):
                L.append("\t\t\t{" + ", ".join(_fmt(v) for v in p.radial_coeffs[i, j, mu]) + "}")
    L.append(f"alpha_moments_count = {p.alpha_moments_count}")
    L.append(f"alpha_index_basic_count = {p.K}")
    L.append("alpha_index_basic = " + _brace(p.alpha_index_basic))
    L.append(f"alpha_index_times_count = {p.T}")
    L.append("alpha_index_times = " + _brace(p.alpha_index_times))
    L.append(f"alpha_scalar_moments = {p.A}")
    L.append("alpha_moment_mapping = {" + ", ".join(str(int(v)) for v in p.alpha_moment_mapping) + "}")
    L.append("species_coeffs = {" + ", ".join(_fmt(v) for v in p.species_coeffs) + "}")
    L.append("moment_coeffs = {" + ", ".join(_fmt(v) for v in p.moment_coeffs) + "}")
    with open(path, "wb") as f:
        f.write(("\n".join(L) + "\n").encode())
        if p.inverse_active_set is not None:
            Q = p.coeff_count
            A = np.ascontiguousarray(p.active_set, dtype="<f8")
            Ai = np.ascontiguousarray(p.inverse_active_set, dtype="<f8")
            assert A.shape == (Q, Q) and Ai.shape == (Q, Q)
            ew = p.energy_weight if p.energy_weight is not None else 0.0
            sw = p.site_en_weight if p.site_en_weight is not None else 1.0
            hdr = ["#MVS_v1.1", f"energy_weight = {_fmt(ew)}", "force_weight = 0.0", "stress_weight = 0.0",
                   f"site_en_weight = {_fmt(sw)}", "weight_scaling = 1"]
            f.write(("\n".join(hdr) + "\n#").encode())
            f.write(A.tobytes())
            f.write(Ai.tobytes())


def read_almtp(path: str) -> MTPPotential:
    raw = open(path, "rb").read()
    cut = raw.find(b"#MVS_v1.1")
    text = (raw if cut < 0 else raw[:cut]).decode()
    lines = [ln.split("#")[0].strip() for ln in text.splitlines()]
    lines = [ln for ln in lines if ln]
    kv = {}
    it = iter(lines)
    assert next(it) == "MTP"
    radial_rows = []
    pair_order = []
    for ln in it:
        if "=" in ln:
            k, v = ln.split("=", 1)
            kv[k.strip()] = v.strip()
        elif re.fullmatch(r"\d+-\d+", ln):
            pair_order.append(tuple(int(t) for t in ln.split("-")))
        elif ln.startswith("{"):
            radial_rows.append([float(t) for t in re.split(r"[{},\s]+", ln) if t])
    S = int(kv["species_count"])
    B = int(kv["radial_basis_size"])
    R = int(kv["radial_funcs_count"])
    rc = np.zeros((S, S, R, B))
    for n, (i, j) in enumerate(pair_order):
        rc[i, j] = np.array(radial_rows[n * R:(n + 1) * R])

    def ints(s):
        return np.array([int(t) for t in re.split(r"[{},\s]+", s) if t], dtype=np.int32)

    def flts(s):
        return np.array([float(t) for t in re.split(r"[{},\s]+", s) if t])

    p = MTPPotential(
        species_count=S, min_dist=float(kv.get("min_dist", kv.get("min_val"))),
        max_dist=float(kv.get("max_dist", kv.get("max_val"))), radial_basis_size=B, radial_funcs_count=R,
        radial_coeffs=rc, alpha_moments_count=int(kv["alpha_moments_count"]),
        alpha_index_basic=ints(kv["alpha_index_basic"]).reshape(-1, 4),
        alpha_index_times=ints(kv["alpha_index_times"]).reshape(-1, 4),
        alpha_moment_mapping=ints(kv["alpha_moment_mapping"]), species_coeffs=flts(kv["species_coeffs"]),
        moment_coeffs=flts(kv["moment_coeffs"]), scaling=float(kv.get("scaling", 1.0)),
        potential_name=kv.get("potential_name", ""), potential_tag=kv.get("potential_tag", ""))
    if cut >= 0:
        tail = raw[cut:]
        # five weight lines after the version line, then '#', then binary
        pos = 0
        hdr = {}
        for _ in range(6):
            e = tail.index(b"\n", pos)
            ln = tail[pos:e].decode()
            if "=" in ln:
                k, v = ln.split("=", 1)
                hdr[k.strip()] = float(v)
            pos = e + 1
        assert tail[pos:pos + 1] == b"#"
        pos += 1
        Q = p.coeff_count
        buf = np.frombuffer(tail, dtype="<f8", count=2 * Q * Q, offset=pos)
        p.active_set = buf[:Q * Q].reshape(Q, Q).copy()
        p.inverse_active_set = buf[Q * Q:].reshape(Q, Q).copy()
        p.energy_weight = hdr.get("energy_weight")
        p.site_en_weight = hdr.get("site_en_weight")
    return p


def random_potential(level: int, species: int, seed: int | None = None, *, min_dist=2.0, max_dist=5.0,
                     radial_basis_size=8, scaling=1.0, with_active_set=False, configuration_mode=False,
                     active_set_seed=3, tables: MTPTables | None = None) -> MTPPotential:
    """Seeded random-init MTP of a given level (SURVEY.md section 8d recipe): radial coeffs ~ U(-0.1,0.1),
    moment coeffs ~ N(0,1), species coeffs ~ U(-1,0); optional synthetic well-conditioned active set
    A = I + 0.1 N(0,1) with its inverse (LU)."""
    t = tables or build_mtp_tables(level)
    rng = np.random.default_rng(level if seed is None else seed)
    S, R, B = species, t.radial_funcs_count, radial_basis_size
    p = MTPPotential(
        species_count=S, min_dist=min_dist, max_dist=max_dist, radial_basis_size=B, radial_funcs_count=R,
        radial_coeffs=rng.uniform(-0.1, 0.1, size=(S, S, R, B)),
        alpha_moments_count=t.alpha_moments_count,
        alpha_index_basic=np.array(t.alpha_index_basic, dtype=np.int32).reshape(-1, 4),
        alpha_index_times=np.array(t.alpha_index_times, dtype=np.int32).reshape(-1, 4),
        alpha_moment_mapping=np.array(t.alpha_moment_mapping, dtype=np.int32),
        species_coeffs=rng.uniform(-1.0, 0.0, size=S),
        moment_coeffs=rng.normal(size=len(t.alpha_moment_mapping)), scaling=scaling,
        potential_name=f"random_L{level}_S{species}")
    if with_active_set:
        Q = p.coeff_count
        r2 = np.random.default_rng(active_set_seed)
        A = np.eye(Q) + 0.1 * r2.normal(size=(Q, Q))
        p.active_set = A
        p.inverse_active_set = np.linalg.inv(A)
        p.energy_weight = 1.0 if configuration_mode else 0.0
        p.site_en_weight = 0.0 if configuration_mode else 1.0
    return p
