"""MLIP-3 configuration (.cfg) reader: the counterpart of the pair style's writer
(PairMTPExtrapolation::write_config, pair_mtp_extrapolation.cpp:401-479 of the reference),
so that configurations selected during active learning can be replayed as inputs.

Grammar (one or more blocks per file):

    BEGIN_CFG
    Size
    <natoms>
    Supercell
    <ax> <ay> <az>          (1 to 3 rows)
    AtomData:  id type cartes_x cartes_y cartes_z [fx fy fz] [nbh_grades] ...
    <one row per atom, columns as named>
    [Energy / <value> on the next line]  [PlusStress: xx yy zz yz xz xy / six values on the next line]
    [Feature <name> <value>]...
    END_CFG

The writer of the reference emits ids starting at 1 and 0-based types (`itype = type[i] - 1`, :420), positions with
six decimals and grades with five; `to_system` turns a configuration back into what the pair style consumes
(1-based LAMMPS types, ghost atoms, full neighbor list).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class Configuration:
    size: int
    supercell: np.ndarray                      # [rows <= 3, 3]
    columns: list                              # AtomData column names
    data: dict                                 # column name -> array [size]
    energy: float | None = None
    stress: np.ndarray | None = None           # PlusStress, MLIP order xx yy zz yz xz xy
    features: dict = field(default_factory=dict)

    @property
    def ids(self):
        return self.data["id"].astype(np.int64)

    @property
    def types(self):
        """0-based species as written in the file."""
        return self.data["type"].astype(np.int32)

    @property
    def positions(self):
        return np.stack([self.data["cartes_x"], self.data["cartes_y"], self.data["cartes_z"]], axis=1)

    @property
    def nbh_grades(self):
        return self.data.get("nbh_grades")

    @property
    def forces(self):
        if "fx" not in self.data:
            return None
        return np.stack([self.data["fx"], self.data["fy"], self.data["fz"]], axis=1)

    def to_system(self, cutoff: float, skin: float = 2.0):
        """Orthorhombic supercells only (what the harness can hold): a harness System with LAMMPS's 1-based types."""
        from . import harness
        cell = np.zeros((3, 3))
        cell[: self.supercell.shape[0]] = self.supercell
        if self.supercell.shape[0] != 3 or np.abs(cell - np.diag(np.diag(cell))).max() > 0.0:
            raise ValueError("only orthorhombic 3-D supercells can be turned into a harness system")
        box = np.diag(cell).copy()
        x = np.mod(self.positions, box)
        return harness.make_system(x, self.types + 1, box, cutoff, skin)


class CfgError(ValueError):
    pass


def parse_cfg(text: str) -> list:
    """All configurations of a .cfg text."""
    lines = text.splitlines()
    out, i, n = [], 0, len(lines)

    def need(cond, msg):
        if not cond:
            raise CfgError(f"line {i + 1}: {msg}")

    while i < n:
        if not lines[i].strip():
            i += 1
            continue
        need(lines[i].strip() == "BEGIN_CFG", f"expected BEGIN_CFG, found {lines[i].strip()!r}")
        i += 1
        size, cell, cols, data, energy, stress, feats = None, [], None, None, None, None, {}
        closed = False
        while i < n:
            s = lines[i].strip()
            if not s:
                i += 1
                continue
            key = s.split()[0]
            if s == "END_CFG":
                closed = True
                i += 1
                break
            if key == "Size":
                need(i + 1 < n, "Size without a value")
                size = int(lines[i + 1].split()[0])
                need(size >= 0, "negative Size")
                i += 2
            elif key in ("Supercell", "SuperCell"):
                i += 1
                while i < n and len(cell) < 3:
                    tok = lines[i].split()
                    try:
                        row = [float(t) for t in tok]
                    except ValueError:
                        break
                    if len(row) != 3:
                        break
                    cell.append(row)
                    i += 1
            elif key == "AtomData:":
                need(size is not None, "AtomData before Size")
                cols = s.split()[1:]
                need(len(cols) >= 5 and len(set(cols)) == len(cols), "bad AtomData column list")
                rows = []
                for k in range(size):
                    need(i + 1 + k < n, "file ends inside AtomData")
                    tok = lines[i + 1 + k].split()
                    need(len(tok) == len(cols), f"atom row {k + 1} has {len(tok)} fields, {len(cols)} columns named")
                    rows.append([float(t) for t in tok])
                arr = np.array(rows, dtype=np.float64).reshape(size, len(cols))
                data = {c: arr[:, j].copy() for j, c in enumerate(cols)}
                i += 1 + size
            elif key == "Energy":
                need(i + 1 < n, "Energy without a value")
                energy = float(lines[i + 1].split()[0])
                i += 2
            elif key == "PlusStress:":
                need(i + 1 < n, "PlusStress without values")
                stress = np.array([float(t) for t in lines[i + 1].split()], dtype=np.float64)
                need(stress.size == 6, "PlusStress needs six values")
                i += 2
            elif key == "Feature":
                tok = s.split(None, 2)
                need(len(tok) == 3, "Feature needs a name and a value")
                feats[tok[1]] = tok[2].strip()
                i += 1
            else:
                need(False, f"unknown keyword {key!r}")
        need(closed, "BEGIN_CFG without END_CFG")
        need(size is not None and data is not None, "configuration without Size / AtomData")
        for c in ("id", "type", "cartes_x", "cartes_y", "cartes_z"):
            need(c in data, f"AtomData lacks column {c}")
        out.append(Configuration(size=size, supercell=np.array(cell, dtype=np.float64).reshape(-1, 3), columns=cols,
                                 data=data, energy=energy, stress=stress, features=feats))
    return out


def read_cfg(path: str) -> list:
    with open(path) as fh:
        return parse_cfg(fh.read())
