"""Flat "system description": the marshalled form of what the reference keeps as CDK
object properties after ``MMFF94.assignParameters`` (``MMFF94.java:111-219``).

One :class:`MolecularSystem` is exactly the argument of ``jme_create``
(``include/jme_b200.h``): per-atom type / charge / linear flag, the vdW rows of the
types that occur, and the four bonded term lists with their ``float`` parameters.
The nonbonded pair map of the reference (``MMFF94.java:153,162-170``) is *not* part of
it - it is O(N^2) - the library rebuilds the same pair set from the bond list.

Term enumeration below restates ``MMFF94.java:158-212`` on index arrays:

* angle / stretch-bend keys ``(i, j, k)`` with ``type(i) < type(k)`` or, on a tie,
  ``index(i) < index(k)`` (``:180-183``);
* out-of-plane keys ``(i, j, k, l)`` for centres with >= 3 bonds, same ordering of
  ``i, k`` and every other neighbour as ``l`` (``:184-193``);
* torsion keys ``(i, j, k, l)`` with neither ``j`` nor ``k`` ideally linear,
  ``l != j, l != i`` and the canonical orientation of ``:205`` .
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

i32 = np.int32
f32 = np.float32


def _a(x, dt):
    return np.ascontiguousarray(np.asarray(x, dtype=dt))


@dataclass
class MolecularSystem:
    xyz: np.ndarray                      # (N,3) float64, Angstrom (Point3d)
    atom_type: np.ndarray                # (N,) int32 MMFF numeric type 1..82, 87..99
    charge: np.ndarray                   # (N,) float64 partial charge (atom.getCharge())
    linear: np.ndarray                   # (N,) uint8 ideallyLinear of the atom's type
    vdw_type: np.ndarray                 # (T,) int32 types that have a row below
    vdw_alpha: np.ndarray                # (T,) float32
    vdw_N: np.ndarray
    vdw_A: np.ndarray
    vdw_G: np.ndarray
    vdw_DA: np.ndarray                   # (T,) uint8 0/1(donor)/2(acceptor)
    bond_i: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    bond_j: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    bond_kb: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    bond_r0: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    angle_i: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    angle_j: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    angle_k: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    angle_ka: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    angle_theta0: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    stbn_kba_ijk: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    stbn_kba_kji: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    oop_i: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    oop_j: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    oop_k: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    oop_l: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    oop_koop: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    tor_i: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    tor_j: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    tor_k: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    tor_l: np.ndarray = field(default_factory=lambda: np.zeros(0, i32))
    tor_v1: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    tor_v2: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    tor_v3: np.ndarray = field(default_factory=lambda: np.zeros(0, f32))
    name: str = ""
    expected_energy: Optional[float] = None     # published OPTIMOL total, kcal/mol
    source: str = ""                            # where geometry / energy came from

    _F64 = ("xyz", "charge")
    _I32 = ("atom_type", "vdw_type", "bond_i", "bond_j", "angle_i", "angle_j", "angle_k",
            "oop_i", "oop_j", "oop_k", "oop_l", "tor_i", "tor_j", "tor_k", "tor_l")
    _U8 = ("linear", "vdw_DA")
    _F32 = ("vdw_alpha", "vdw_N", "vdw_A", "vdw_G", "bond_kb", "bond_r0", "angle_ka",
            "angle_theta0", "stbn_kba_ijk", "stbn_kba_kji", "oop_koop", "tor_v1", "tor_v2", "tor_v3")

    def __post_init__(self):
        for k in self._F64:
            setattr(self, k, _a(getattr(self, k), np.float64))
        for k in self._I32:
            setattr(self, k, _a(getattr(self, k), i32))
        for k in self._U8:
            setattr(self, k, _a(getattr(self, k), np.uint8))
        for k in self._F32:
            setattr(self, k, _a(getattr(self, k), f32))
        self.xyz = self.xyz.reshape(-1, 3)
        self.validate()

    # ------------------------------------------------------------------
    @property
    def n_atoms(self) -> int:
        return int(self.xyz.shape[0])

    def validate(self) -> None:
        n = self.n_atoms
        for k in ("atom_type", "charge", "linear"):
            if getattr(self, k).shape != (n,):
                raise ValueError(f"{k} must have one entry per atom")
        groups = (("bond_i", "bond_j", "bond_kb", "bond_r0"),
                  ("angle_i", "angle_j", "angle_k", "angle_ka", "angle_theta0", "stbn_kba_ijk", "stbn_kba_kji"),
                  ("oop_i", "oop_j", "oop_k", "oop_l", "oop_koop"),
                  ("tor_i", "tor_j", "tor_k", "tor_l", "tor_v1", "tor_v2", "tor_v3"),
                  ("vdw_type", "vdw_alpha", "vdw_N", "vdw_A", "vdw_G", "vdw_DA"))
        for g in groups:
            m = getattr(self, g[0]).shape[0]
            for k in g:
                if getattr(self, k).shape != (m,):
                    raise ValueError(f"{k} length differs from {g[0]}")
        for k in self._I32:
            if k in ("atom_type", "vdw_type"):
                continue
            a = getattr(self, k)
            if a.size and (a.min() < 0 or a.max() >= n):
                raise ValueError(f"{k} holds an atom index outside [0, {n})")
        missing = set(np.unique(self.atom_type).tolist()) - set(self.vdw_type.tolist())
        if missing:
            raise ValueError(f"no vdW row for atom types {sorted(missing)}")

    def with_coordinates(self, xyz) -> "MolecularSystem":
        import copy
        s = copy.copy(self)
        s.xyz = _a(xyz, np.float64).reshape(-1, 3)
        if s.xyz.shape[0] != self.n_atoms:
            raise ValueError("coordinate count differs")
        return s

    # ------------------------------------------------------------------
    def to_dict(self) -> dict:
        d = {"name": self.name, "expected_energy": self.expected_energy, "source": self.source}
        for k in self._F64 + self._F32:
            # repr of the widened value: json round-trips doubles exactly
            d[k] = np.asarray(getattr(self, k), dtype=np.float64).reshape(-1).tolist()
        for k in self._I32 + self._U8:
            d[k] = getattr(self, k).astype(np.int64).tolist()
        return d

    @classmethod
    def from_dict(cls, d: dict) -> "MolecularSystem":
        kw = {k: d[k] for k in cls._F64 + cls._F32 + cls._I32 + cls._U8}
        return cls(name=d.get("name", ""), expected_energy=d.get("expected_energy"),
                   source=d.get("source", ""), **kw)

    def save(self, path: str) -> None:
        with open(path, "w") as fh:
            json.dump(self.to_dict(), fh, indent=0, separators=(",", ":"))
            fh.write("\n")

    @classmethod
    def load(cls, path: str) -> "MolecularSystem":
        with open(path) as fh:
            return cls.from_dict(json.load(fh))


# ----------------------------------------------------------------------------
# term enumeration (MMFF94.java:158-212)
# ----------------------------------------------------------------------------
def adjacency(n_atoms: int, bonds: Sequence[Tuple[int, int]]) -> List[List[int]]:
    """Neighbour lists in bond-list order - the order ``atom.bonds()`` iterates."""
    adj: List[List[int]] = [[] for _ in range(n_atoms)]
    for a, b in bonds:
        adj[a].append(b)
        adj[b].append(a)
    return adj


def enumerate_terms(types: Sequence[int], linear: Sequence[int], bonds: Sequence[Tuple[int, int]]):
    """Return ``(angles, oops, torsions)`` as lists of index tuples, de-duplicated the way
    the reference's ``HashMap.put`` does, in first-insertion order."""
    n = len(types)
    adj = adjacency(n, bonds)
    angles: Dict[Tuple[int, int, int], None] = {}
    oops: Dict[Tuple[int, int, int, int], None] = {}
    torsions: Dict[Tuple[int, int, int, int], None] = {}
    for a in range(n):
        if len(adj[a]) < 2:
            continue
        tj = types[a]
        for a2 in adj[a]:
            for a3 in adj[a]:
                if a3 == a2:
                    continue
                ti, tk = types[a2], types[a3]
                first = ti < tk or (ti == tk and a2 < a3)
                if first:
                    angles[(a2, a, a3)] = None
                if len(adj[a]) >= 3 and first:
                    for a4 in adj[a]:
                        if a4 == a2 or a4 == a3:
                            continue
                        oops[(a2, a, a3, a4)] = None
                if not linear[a] and not linear[a3] and len(adj[a3]) >= 2:
                    for a4 in adj[a3]:
                        if a4 == a or a4 == a2:
                            continue
                        tl = types[a4]
                        if tj < tk or (tj == tk and ti < tl) or (tj == tk and ti == tl and a < a3):
                            torsions[(a2, a, a3, a4)] = None
    return list(angles), list(oops), list(torsions)


def pair_separation(n_atoms: int, bonds: Sequence[Tuple[int, int]]):
    """Topological distance classes used by the pair predicate
    (``MMFF94.java:162-170,1106-1140``): returns ``(excluded, is14)`` as sets of
    ``(i, j)`` with ``i < j`` - shortest path 1 or 2 bonds, and exactly 3 bonds."""
    adj = adjacency(n_atoms, bonds)
    excluded, is14 = set(), set()
    for a in range(n_atoms):
        dist = {a: 0}
        frontier = [a]
        for depth in (1, 2, 3):
            nxt = []
            for u in frontier:
                for v in adj[u]:
                    if v not in dist:
                        dist[v] = depth
                        nxt.append(v)
            frontier = nxt
        for v, d in dist.items():
            if v > a:
                if d in (1, 2):
                    excluded.add((a, v))
                elif d == 3:
                    is14.add((a, v))
    return excluded, is14
