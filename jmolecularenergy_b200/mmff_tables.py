"""Reader for the MMFF94 ``.par`` tables plus the table-hit half of the
reference's parameter lookup.

The tables themselves are NOT shipped with this package (Merck copyright
headers): the caller points :class:`MMFFTables`
at a directory that holds them, e.g. the reference's
``src/main/resources/org/jme/forcefield/mmff``.  The package only needs this
module to build synthetic systems and golden fixtures; the GPU path itself
consumes already-assigned parameters (``jme_create`` in ``include/jme_b200.h``).

What is restated here (reference file:line, all under
``src/main/java/org/jme/forcefield/mmff``):

* parsing with ``Float.parseFloat`` semantics - every tabulated number is a
  binary32 value that is widened at use (``MMFF94Parameters.java:189-356``);
* dense per-type row index ``type - (type >= 87 ? 5 : 1)`` (``MMFF94.java:119-120,160``);
* bond lookup by ``(bondType, ti<=tj)`` (``MMFF94.java:222-256``);
* angle lookup with the 4-level equivalence fallback (``MMFF94.java:868-900``);
* stretch-bend lookup incl. the ``swap`` of kbaIJK/kbaKJI (``MMFF94.java:983-1045``);
* out-of-plane lookup with sorted (i,k,l) and equivalences (``MMFF94.java:258-288``);
* torsion lookup with the 1-3 / 3-1 / 3-3 equivalence ladder (``MMFF94.java:629-706``);
* bond-charge-increment partial charges (``MMFF94.java:566-618``).

Empirical-rule fallbacks (``MMFF94.java:290-369,902-981,1046-1062,707-811``) are
deliberately not restated: a lookup that would need them raises
:class:`ParameterNotTabulated`, so fixtures only ever contain table hits.
Ring-dependent angle/torsion types (3-/4-/5-membered rings, ``MMFF94.java:838-856,
645-663``) are not handled either; callers must stick to acyclic systems or
pass the angle / torsion type explicitly.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32


class ParameterNotTabulated(KeyError):
    """Raised where the reference would fall back to an empirical rule."""


def dense_type_index(mmff_type: int) -> int:
    """Row index into the 95-row per-type tables (``MMFF94.java:119-120``)."""
    if not (1 <= mmff_type <= 82 or 87 <= mmff_type <= 99):
        raise ValueError(f"not an MMFF94 numeric atom type: {mmff_type}")
    return mmff_type - (5 if mmff_type >= 87 else 1)


def _rows(path: str) -> List[List[str]]:
    out = []
    with open(path, "r", encoding="utf-8") as fh:
        for line in fh:
            line = line.rstrip("\n").rstrip("\r")
            if not line or line[0] == "*":
                continue
            out.append(line.split("\t"))
    return out


@dataclass(frozen=True)
class VdwRow:
    type: int
    alpha: np.float32
    N: np.float32
    A: np.float32
    G: np.float32
    DA: int  # 0 none, 1 donor, 2 acceptor (MMFF94Parameters.java:363-365)


@dataclass(frozen=True)
class GeomRow:
    type: int
    atomic_number: int
    crd: int
    valence: int
    pilp: bool
    mltb: int
    arom: bool
    linear: bool
    sbmb: bool


class MMFFTables:
    """In-memory copy of the tables the hot path's inputs are drawn from."""

    def __init__(self, directory: str, mmff94s: bool = False):
        if not os.path.isdir(directory):
            raise FileNotFoundError(
                f"MMFF94 parameter directory not found: {directory!r}")
        self.directory = directory
        self.mmff94s = mmff94s
        p = lambda name: os.path.join(directory, name)

        self.vdw: List[VdwRow] = []
        for r in _rows(p("MMFF94-vdw.par")):
            da = {"D": 1, "A": 2}.get(r[5], 0)
            self.vdw.append(VdwRow(int(r[0]), f32(r[1]), f32(r[2]), f32(r[3]), f32(r[4]), da))
        self.geom: List[GeomRow] = []
        for r in _rows(p("MMFF94-geometric-properties.par")):
            self.geom.append(GeomRow(int(r[0]), int(r[1]), int(r[2]), int(r[3]), r[4] == "1",
                                     int(r[5]), r[6] == "1", r[7] == "1", r[8] == "1"))
        # (bondType, i, j) -> (kb, r0); first row wins like the binary search would
        self.bond: Dict[Tuple[int, int, int], Tuple[np.float32, np.float32]] = {}
        for r in _rows(p("MMFF94-bond-stretch.par")):
            self.bond.setdefault((int(r[0]), int(r[1]), int(r[2])), (f32(r[3]), f32(r[4])))
        self.angle: List[Tuple[int, int, int, int, np.float32, np.float32]] = [
            (int(float(r[0])), int(float(r[1])), int(float(r[2])), int(float(r[3])), f32(r[4]), f32(r[5]))
            for r in _rows(p("MMFF94-angle-bending.par"))]
        self.stbn: List[Tuple[int, int, int, int, np.float32, np.float32]] = [
            (int(float(r[0])), int(float(r[1])), int(float(r[2])), int(float(r[3])), f32(r[4]), f32(r[5]))
            for r in _rows(p("MMFF94-stretch-bend.par"))]
        # default stretch-bend constants by periodic-table rows (IR, JR, KR) -> (F(I_J,K), F(K_J,I)), used when
        # the typed table has no row (MMFF94.java:1047-1060)
        self.stbn_default: Dict[Tuple[int, int, int], Tuple[np.float32, np.float32]] = {}
        emp = p("MMFF94-stretch-bend-empirical.par")
        if os.path.exists(emp):
            for r in _rows(emp):
                self.stbn_default.setdefault((int(r[0]), int(r[1]), int(r[2])), (f32(r[3]), f32(r[4])))
        oop_file = "MMFF94s-out-of-plane.par" if mmff94s else "MMFF94-out-of-plane.par"
        self.oop: List[Tuple[int, int, int, int, np.float32]] = [
            (int(float(r[0])), int(float(r[1])), int(float(r[2])), int(float(r[3])), f32(r[4]))
            for r in _rows(p(oop_file))]
        tor_file = "MMFF94s-torsion.par" if mmff94s else "MMFF94-torsion.par"
        self.torsion: List[Tuple[int, int, int, int, int, np.float32, np.float32, np.float32]] = [
            (int(float(r[0])), int(float(r[1])), int(float(r[2])), int(float(r[3])), int(float(r[4])),
             f32(r[5]), f32(r[6]), f32(r[7])) for r in _rows(p(tor_file))]
        self.equiv: List[Tuple[int, int, int, int]] = [
            (int(r[2]), int(r[3]), int(r[4]), int(r[5]))
            for r in _rows(p("MMFF94-atom-type-equivalence.par"))]
        # stored negated, exactly like MMFF94Parameters.java:213
        self.bci: List[Tuple[int, int, int, np.float32]] = [
            (int(float(r[0])), int(float(r[1])), int(float(r[2])), f32(-f32(r[3])))
            for r in _rows(p("MMFF94-bond-charge-increment.par"))]
        self.pbci: List[Tuple[np.float32, np.float32]] = [
            (f32(r[1]), f32(r[2])) for r in _rows(p("MMFF94-partial-charge-increment.par"))]
        # indices for the full assignment port (assign.py): first row wins, like the reference's linear scans
        self.bci_by_key: Dict[Tuple[int, int, int], np.float32] = {}
        for row in self.bci:
            self.bci_by_key.setdefault((row[0], row[1], row[2]), row[3])
        self.angle_by_centre: Dict[int, List[tuple]] = {}
        for row in self.angle:
            self.angle_by_centre.setdefault(row[2], []).append(row)
        self.torsion_by_centre: Dict[Tuple[int, int], List[tuple]] = {}
        for row in self.torsion:
            self.torsion_by_centre.setdefault((row[2], row[3]), []).append(row)
        # (atomic number 1 <= atomic number 2) -> (kb, r0) of the default-rule bond table (MMFF94Parameters.java:236-248)
        self.bond_empirical: Dict[Tuple[int, int], Tuple[np.float32, np.float32]] = {}
        emp_b = p("MMFF94-bond-stretch-empirical.par")
        if os.path.exists(emp_b):
            for r in _rows(emp_b):
                self.bond_empirical.setdefault((int(r[0]), int(r[1])), (f32(r[3]), f32(r[2])))

    # ---- per-type rows -------------------------------------------------
    def vdw_row(self, t: int) -> VdwRow:
        row = self.vdw[dense_type_index(t)]
        assert row.type == t
        return row

    def geom_row(self, t: int) -> GeomRow:
        row = self.geom[dense_type_index(t)]
        assert row.type == t
        return row

    def equivalence(self, t: int, level: int) -> int:
        """``findAtomTypeEquivalence`` (``MMFF94.java:858-866``)."""
        return self.equiv[dense_type_index(t)][level]

    # ---- bond type (acyclic, non-aromatic subset) ----------------------
    def bond_type(self, ti: int, tj: int, order: int = 1) -> int:
        """``findBondType`` (``MMFF94.java:824-836``) without the aromatic-bond flag
        (callers here never build aromatic rings)."""
        if order == 1:
            gi, gj = self.geom_row(ti), self.geom_row(tj)
            if gi.sbmb and gj.sbmb:
                return 1
            if gi.arom and gj.arom:
                return 1
        return 0

    # ---- term lookups --------------------------------------------------
    def find_bond(self, ti: int, tj: int, bond_type: int) -> Tuple[np.float32, np.float32]:
        i, j = (ti, tj) if ti <= tj else (tj, ti)
        try:
            return self.bond[(bond_type, i, j)]
        except KeyError:
            raise ParameterNotTabulated(f"bond {bond_type}:{i}-{j}") from None

    def find_angle(self, ti: int, tj: int, tk: int, angle_type: int) -> Tuple[np.float32, np.float32]:
        i, k = ti, tk
        level = 0
        while True:
            i, k = min(i, k), max(i, k)
            for row in self.angle:
                if row[0] == angle_type and row[1] == i and row[2] == tj and row[3] == k:
                    if row[4] == f32(0.0):
                        raise ParameterNotTabulated(f"angle {angle_type}:{ti}-{tj}-{tk} needs the empirical ka")
                    return row[4], row[5]
            if level > 3:
                raise ParameterNotTabulated(f"angle {angle_type}:{ti}-{tj}-{tk}")
            i = self.equivalence(ti, level)
            k = self.equivalence(tk, level)
            level += 1

    def find_stretch_bend(self, ti: int, tj: int, tk: int, angle_type: int,
                          ij_bond_type: int, jk_bond_type: int) -> Tuple[np.float32, np.float32]:
        swap = False
        if ti == tk and ij_bond_type < jk_bond_type:
            ij_bond_type, jk_bond_type = jk_bond_type, ij_bond_type
            swap = True
        same_or_first = (ij_bond_type != 0) or (ij_bond_type == jk_bond_type)
        sbt = {0: 0, 1: 1 if same_or_first else 2, 2: 3, 4: 4, 3: 5,
               5: 6 if same_or_first else 7, 6: 8, 7: 9 if same_or_first else 10, 8: 11}[angle_type]
        for row in self.stbn:
            if row[2] == tj and row[0] == sbt and row[1] == ti and row[3] == tk:
                return (row[5], row[4]) if swap else (row[4], row[5])
        # no typed row: the default rule by periodic-table rows (MMFF94.java:1047-1060); the table is stored with
        # IR <= KR, the two constants change places when the atoms had to be swapped to look it up
        ri, rj, rk = (self.periodic_row(t) for t in (ti, tj, tk))
        hit = self.stbn_default.get((min(ri, rk), rj, max(ri, rk)))
        if hit is None:
            raise ParameterNotTabulated(f"stretch-bend {sbt}:{ti}-{tj}-{tk}")
        return (hit[1], hit[0]) if ri > rk else hit

    def periodic_row(self, t: int) -> int:
        """Row of the periodic table as the stretch-bend default rule counts it (MMFF94.java:1064-1080)."""
        z = self.geom_row(t).atomic_number
        for row, last in enumerate((2, 10, 18, 36, 54, 86)):
            if z <= last:
                return row
        return 6

    def find_oop(self, ti: int, tj: int, tk: int, tl: int) -> Optional[np.float32]:
        """Returns koop, or ``None`` where the reference stores a null row."""
        ikl = [ti, tk, tl]
        level = 0
        while True:
            ikl.sort()
            for row in self.oop:
                if row[1] == tj and row[0] == ikl[0] and row[2] == ikl[1] and row[3] == ikl[2]:
                    return row[4]
            if level > 3:
                return None
            ikl = [self.equivalence(ti, level), self.equivalence(tk, level), self.equivalence(tl, level)]
            level += 1

    def find_torsion(self, ti: int, tj: int, tk: int, tl: int,
                     torsion_type: int) -> Tuple[np.float32, np.float32, np.float32]:
        """Table part of ``findTorsionParameters`` for acyclic torsions: the requested
        type first, then (for a non-zero type) type 0, each over the equivalence ladder
        full / 1-3 / 3-1 / 3-3 (``MMFF94.java:664-703``)."""
        assert tj <= tk, "torsion key must be canonical (MMFF94.java:205)"
        ladder = [(None, None), (1, 3), (3, 1), (3, 3)]
        for tt in ([torsion_type] if torsion_type == 0 else [torsion_type, 0]):
            for li, ll in ladder:
                i = ti if li is None else self.equivalence(ti, li)
                l = tl if ll is None else self.equivalence(tl, ll)
                for row in self.torsion:
                    if row[0] == tt and row[2] == tj and row[3] == tk and row[1] == i and row[4] == l:
                        return row[5], row[6], row[7]
        raise ParameterNotTabulated(f"torsion {torsion_type}:{ti}-{tj}-{tk}-{tl}")

    # ---- charges ---------------------------------------------------------
    def partial_charges(self, types: Sequence[int], bonds: Sequence[Tuple[int, int]],
                        formal: Sequence[float], bond_orders: Optional[Sequence[int]] = None) -> np.ndarray:
        """``calculatePartialCharge`` (``MMFF94.java:566-618``) for atoms whose formal
        charge is not shared (the negative-neighbour sharing at ``:570-577`` and the
        type-62 rule at ``:578-585`` are applied as written).  float32 products and the
        double accumulation follow the Java expression types."""
        n = len(types)
        nbrs: List[List[int]] = [[] for _ in range(n)]
        order = {}
        for b, (a, c) in enumerate(bonds):
            nbrs[a].append(c)
            nbrs[c].append(a)
            o = 1 if bond_orders is None else bond_orders[b]
            order[(a, c)] = order[(c, a)] = o
        q = np.zeros(n, dtype=np.float64)
        for a in range(n):
            ta = types[a]
            q0 = float(formal[a])
            fcadj = self.pbci[ta - 1][1]
            if fcadj == f32(0.0):
                for nb in nbrs[a]:
                    k0 = float(formal[nb])
                    if k0 < 0.0:
                        q0 += k0 / (2.0 * len(nbrs[nb]))
            if ta == 62:
                for nb in nbrs[a]:
                    k0 = float(formal[nb])
                    if k0 > 0.0:
                        q0 -= k0 / 2.0
            qi = 0.0
            qk_sum = 0.0
            for nb in nbrs[a]:
                tn = types[nb]
                qk_sum += float(formal[nb])
                bt = self.bond_type(ta, tn, order[(a, nb)])
                i, j = min(ta, tn), max(ta, tn)
                reversed_ = ta > tn
                found = False
                for row in self.bci:
                    if row[1] == i and row[2] == j and row[0] == bt:
                        qi += float(f32(row[3] * f32(-1 if reversed_ else 1)))
                        found = True
                        break
                if not found:
                    qi += float(f32(self.pbci[ta - 1][0] - self.pbci[tn - 1][0]))
            crd = self.geom_row(ta).crd
            term = f32(f32(1) - f32(f32(crd) * fcadj))      # (1 - crd * fcadj): int*Float, int-Float -> float
            qi += float(term) * q0 + qk_sum * float(fcadj)
            q[a] = qi
        return q
