"""``MMFF94.assignParameters`` restated for molecules whose MMFF atom types are GIVEN.

The reference types atoms with CDK's ``Mmff.assignAtomTypes`` (third-party, needs a JVM); everything after that -
bond / angle / stretch-bend / torsion types with their ring rules, table lookups with the equivalence ladders, the
empirical fall-back rules, bond-charge-increment charges - is the reference's own code (``MMFF94.java:111-369,
566-1062``) and is restated here function by function.  Host-side set-up: it produces the arrays ``jme_create``
consumes; nothing of it runs on the hot path.

The ``.par`` tables are read from a caller-supplied directory (``MMFFTables``); the small constant tables of the
empirical rules (``MMFF94Parameters.java:60-133``: Halgren's published Z / C, U / V, covalent radii,
electronegativities and Herschbach-Laurie constants) are restated below as data.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Set, Tuple

import numpy as np

from .mmff_tables import MMFFTables, f32
from .system import MolecularSystem, enumerate_terms

# MMFF94Parameters.java:60-75: {atomic number: (Z, C)} of the empirical angle-bending rule (-1: not defined)
ANGLE_EMPIRICAL = {1: (1.395, -1), 5: (-1, 0.704), 6: (2.494, 1.016), 7: (2.711, 1.113), 8: (3.045, 1.337),
                   9: (2.847, -1), 14: (2.350, 0.811), 15: (2.350, 1.068), 16: (2.980, 1.249), 17: (2.909, 1.078),
                   33: (-1, 0.825), 35: (3.017, -1), 53: (3.086, -1)}
# MMFF94Parameters.java:77-84: {atomic number: (U, V)} of the empirical torsion rule
TORSION_EMPIRICAL = {6: (2.0, 2.12), 7: (2.0, 1.5), 8: (2.0, 0.2), 14: (1.25, 1.22), 15: (1.25, 2.4), 16: (1.25, 0.49)}
# MMFF94Parameters.java:86-105: {atomic number: (covalent radius, Pauling electronegativity)}
RADII_ELECTRONEGATIVITY = {1: (0.33, 2.20), 3: (1.34, 0.97), 6: (0.77, 2.50), 7: (0.73, 3.07), 8: (0.72, 3.50),
                           9: (0.74, 4.12), 11: (1.54, 1.01), 12: (1.30, 1.23), 14: (1.15, 1.74), 15: (1.09, 2.06),
                           16: (1.03, 2.44), 17: (1.01, 2.83), 19: (1.96, 0.91), 20: (1.74, 1.04), 29: (1.38, 1.75),
                           30: (1.31, 1.66), 35: (1.15, 2.74), 53: (1.33, 2.21)}
# MMFF94Parameters.java:107-133: Badger's rule {(row i, row j): (a_ij, d_ij)}
HERSCHBACH_LAURIE = {(0, 0): (1.26, 0.025), (0, 1): (1.66, 0.30), (0, 2): (1.84, 0.38), (0, 3): (1.98, 0.49),
                     (0, 4): (2.03, 0.51), (0, 5): (2.03, 0.25), (0, 30): (1.85, 0.15), (0, 40): (1.84, 0.61),
                     (0, 50): (1.78, 0.97), (1, 1): (1.91, 0.68), (1, 2): (2.28, 0.74), (1, 3): (2.35, 0.85),
                     (1, 4): (2.33, 0.68), (1, 5): (2.50, 0.97), (1, 30): (2.08, 1.14), (1, 40): (2.34, 1.17),
                     (2, 2): (2.41, 1.18), (2, 3): (2.52, 1.02), (2, 4): (2.61, 1.28), (2, 5): (2.60, 0.84),
                     (3, 3): (2.58, 1.41), (3, 4): (2.66, 0.86), (3, 5): (2.75, 1.14), (4, 4): (2.85, 1.62),
                     (4, 5): (2.76, 1.25)}


class AssignmentError(RuntimeError):
    """Where the reference throws (no empirical constants for an element, no stretch-bend default ...)."""


def find_rings(n: int, bonds: Sequence[Tuple[int, int]], max_size: int = 9) -> List[Tuple[int, ...]]:
    """All simple cycles of up to ``max_size`` atoms (what ``AllRingsFinder`` gives the reference for the ring sizes its
    rules look at: 3, 4, 5 and the aromatic six-rings), each once, as atom tuples starting at their smallest index."""
    nb: List[List[int]] = [[] for _ in range(n)]
    for a, b in bonds:
        nb[a].append(b)
        nb[b].append(a)
    rings: Set[Tuple[int, ...]] = set()

    def walk(start: int, path: List[int], seen: Set[int]) -> None:
        last = path[-1]
        for x in nb[last]:
            if x == start and len(path) >= 3:
                if path[1] < path[-1]:                    # one of the two directions
                    rings.add(tuple(path))
            elif x > start and x not in seen and len(path) < max_size:
                seen.add(x)
                path.append(x)
                walk(start, path, seen)
                path.pop()
                seen.discard(x)

    for s in range(n):
        walk(s, [s], {s})
    return sorted(rings, key=lambda r: (len(r), r))


class Assigner:
    """One molecule: types, bonds (order 1 / 2 / 3, aromatic flag), rings -> every parameter the hot path takes."""

    def __init__(self, tables: MMFFTables, types: Sequence[int], bonds: Sequence[Tuple[int, int]],
                 orders: Sequence[int], aromatic: Optional[Sequence[bool]] = None):
        self.t = tables
        self.types = [int(x) for x in types]
        self.n = len(self.types)
        self.bonds = [(int(a), int(b)) for a, b in bonds]
        self.order: Dict[Tuple[int, int], int] = {}
        self.arom: Dict[Tuple[int, int], bool] = {}
        aromatic = [False] * len(self.bonds) if aromatic is None else list(aromatic)
        self.nbrs: List[List[int]] = [[] for _ in range(self.n)]
        for (a, b), o, ar in zip(self.bonds, orders, aromatic):
            self.order[(a, b)] = self.order[(b, a)] = int(o)
            self.arom[(a, b)] = self.arom[(b, a)] = bool(ar)
            self.nbrs[a].append(b)
            self.nbrs[b].append(a)
        self.rings = find_rings(self.n, self.bonds)
        self.ring_sets = [frozenset(r) for r in self.rings]
        self.rings_of: List[List[int]] = [[] for _ in range(self.n)]
        for k, r in enumerate(self.rings):
            for a in r:
                self.rings_of[a].append(k)
        self.geom = [tables.geom_row(x) for x in self.types]
        self.r0: Dict[Tuple[int, int], np.float32] = {}

    def in_ring(self, a: int) -> bool:
        return bool(self.rings_of[a])

    # ---- MMFF94.java:824-836
    def bond_type(self, i: int, j: int) -> int:
        if self.order[(i, j)] == 1 and not self.arom[(i, j)]:
            gi, gj = self.geom[i], self.geom[j]
            if gi.sbmb and gj.sbmb:
                return 1
            if gi.arom and gj.arom:
                return 1
        return 0

    # ---- MMFF94.java:840-856
    def angle_type(self, i: int, j: int, k: int) -> int:
        bt = self.bond_type(i, j) + self.bond_type(j, k)
        if self.in_ring(i) and self.in_ring(j) and self.in_ring(k):
            for r in self.rings_of[i]:
                ring = self.ring_sets[r]
                if j in ring and k in ring:
                    if len(ring) == 3:
                        return 3 if bt == 0 else 3 + bt + 1
                    if len(ring) == 4:
                        return 4 if bt == 0 else 4 + bt + 2
        return bt

    def _row(self, a: int) -> int:
        """getPeriodicTableRow, MMFF94.java:1064-1080."""
        return self.t.periodic_row(self.types[a])

    def _hl_row(self, a: int) -> int:
        """getHerschbachLauriePeriodicTableRow, MMFF94.java:1082-1095, restated as written (hydrogen is not handled
        there: the reference throws)."""
        z = self.geom[a].atomic_number
        if z == 2:
            return 1
        if 3 <= z <= 10:
            return 2
        if 11 <= z <= 18:
            return 3
        if 19 <= z <= 36:
            return 40 if 21 <= z <= 30 else 4
        if 37 <= z <= 54:
            return 50 if 39 <= z <= 48 else 5
        raise AssignmentError(f"heavy atoms (type {self.types[a]}) are not supported")

    # ---- bonds: MMFF94.java:222-256, 290-345
    def bond_parameters(self, i: int, j: int) -> Tuple[np.float32, np.float32]:
        ti, tj = self.types[i], self.types[j]
        bt = self.bond_type(i, j)
        key = (bt, min(ti, tj), max(ti, tj))
        hit = self.t.bond.get(key)
        if hit is not None:
            return hit
        z1, z2 = sorted((self.geom[i].atomic_number, self.geom[j].atomic_number))
        if z1 not in RADII_ELECTRONEGATIVITY or z2 not in RADII_ELECTRONEGATIVITY:
            raise AssignmentError(f"could not find empirical parameters for the {ti}-{tj} bond")
        (r1, e1), (r2, e2) = RADII_ELECTRONEGATIVITY[z1], RADII_ELECTRONEGATIVITY[z2]
        c = 0.050 if (z1 == 1 or z2 == 1) else 0.085
        # Float covalent radii / electronegativities, double arithmetic, rounded to float at the end (:331)
        r0 = f32(float(f32(f32(r1) + f32(r2))) - c * math.pow(abs(float(f32(f32(e1) - f32(e2)))), 1.4))
        emp = self.t.bond_empirical.get((z1, z2))
        if emp is not None:
            kb = f32(float(emp[0]) * math.pow(float(f32(emp[1] / r0)), 6))
        else:
            ra, rb = sorted((self._hl_row(i), self._hl_row(j)))
            hl = HERSCHBACH_LAURIE.get((ra, rb))
            if hl is None:
                raise AssignmentError(f"could not find Herschbach-Laurie parameters for the {ti}-{tj} bond")
            kb = f32(math.pow(10.0, -float(f32(f32(r0 - f32(hl[0])) / f32(hl[1])))))
        return kb, r0

    # ---- angles: MMFF94.java:868-981
    def angle_parameters(self, i: int, j: int, k: int) -> Tuple[np.float32, np.float32]:
        ti, tj, tk = self.types[i], self.types[j], self.types[k]
        at = self.angle_type(i, j, k)
        a, c = ti, tk
        level = 0
        while True:
            a, c = min(a, c), max(a, c)
            for row in self.t.angle_by_centre.get(tj, ()):
                if row[0] == at and row[1] == a and row[3] == c:
                    if row[4] == f32(0.0):
                        return self._empirical_angle(i, j, k, at, row[5])
                    return row[4], row[5]
            if level > 3:
                return self._empirical_angle(i, j, k, at, None)
            a = self.t.equivalence(ti, level)
            c = self.t.equivalence(tk, level)
            level += 1

    def _empirical_angle(self, i: int, j: int, k: int, at: int, ideal: Optional[np.float32]):
        gj = self.geom[j]
        deg = len(self.nbrs[j])
        theta = f32(120.0)
        if deg == 4:
            theta = f32(109.45)
        elif deg == 3:
            if gj.valence == 3 and gj.mltb == 0:
                theta = f32(107.0) if gj.atomic_number == 7 else f32(92.0)
        elif deg == 2:
            if gj.atomic_number == 8:
                theta = f32(105.0)
            elif gj.atomic_number > 10:
                theta = f32(95.0)
            elif gj.linear:
                theta = f32(180.0)
        beta = f32(1.75)
        if self.in_ring(i) and self.in_ring(j) and self.in_ring(k):
            for r in self.rings_of[i]:
                ring = self.ring_sets[r]
                if j in ring and k in ring:
                    if len(ring) == 3:
                        theta = f32(60.0)
                        beta = f32(beta * f32(0.05))
                    elif len(ring) == 4:
                        theta = f32(90.0)
                        beta = f32(beta * f32(0.85))
                    break
        if ideal is not None:
            theta = ideal
        r_ij, r_jk = self.r0[(i, j)], self.r0[(j, k)]
        d = f32(math.pow(float(f32(r_ij - r_jk)), 2) / math.pow(float(f32(r_ij + r_jk)), 2))
        try:
            zi = ANGLE_EMPIRICAL[self.geom[i].atomic_number][0]
            cj = ANGLE_EMPIRICAL[gj.atomic_number][1]
            zk = ANGLE_EMPIRICAL[self.geom[k].atomic_number][0]
        except KeyError:
            raise AssignmentError(f"could not generate empirical angle bending parameters for "
                                  f"{self.types[i]}-{self.types[j]}-{self.types[k]}") from None
        num = f32(f32(f32(beta * f32(zi)) * f32(cj)) * f32(zk))           # float products, left to right
        ka = f32(float(num) / (float(f32(r_ij + r_jk)) * math.pow(math.radians(float(theta)), 2) * math.exp(2 * float(d))))
        return ka, theta

    # ---- stretch-bend: MMFF94.java:983-1062
    def stretch_bend_parameters(self, i: int, j: int, k: int) -> Tuple[np.float32, np.float32]:
        return self.t.find_stretch_bend(self.types[i], self.types[j], self.types[k], self.angle_type(i, j, k),
                                        self.bond_type(i, j), self.bond_type(j, k))

    # ---- torsions: MMFF94.java:629-813
    def torsion_parameters(self, i: int, j: int, k: int, l: int) -> Tuple[np.float32, np.float32, np.float32]:
        ti, tj, tk, tl = (self.types[x] for x in (i, j, k, l))
        ij, jk, kl = self.bond_type(i, j), self.bond_type(j, k), self.bond_type(k, l)
        tt, old = jk, 0
        if jk == 0 and self.order[(j, k)] == 1 and (ij == 1 or kl == 1):
            tt = 2
        if all(self.in_ring(x) for x in (i, j, k, l)):
            for r in self.rings_of[i]:
                ring = self.ring_sets[r]
                if j in ring and k in ring and l in ring:
                    if len(ring) == 4:
                        old, tt = tt, 4
                        break
                    if len(ring) == 5 and 1 in (ti, tj, tk, tl):
                        old, tt = tt, 5
                        break
        level, found, used_old = 0, None, False
        a, d = ti, tl
        rows = self.t.torsion_by_centre.get((tj, tk), ())
        while ((level < 5 and found is None and tt != old) or (level < 4 and found is None)
               or (level == 4 and tt == 5 and old != 0)):
            if level == 4 and not used_old:
                used_old, tt, level, found = True, old, 0, None
                a, d = ti, tl
            if level != 0:
                li, ll = {1: (1, 3), 2: (3, 1)}.get(level, (level, level))
                a, d = self.t.equivalence(ti, li), self.t.equivalence(tl, ll)
            for row in rows:
                if row[0] == tt and row[1] == a and row[4] == d:
                    found = row
                    break
            level += 1
        if found is not None:
            return found[5], found[6], found[7]
        return self._empirical_torsion(j, k)

    def _empirical_torsion(self, j: int, k: int):
        gj, gk = self.geom[j], self.geom[k]
        zero = (f32(0), f32(0), f32(0))

        def uv(a):
            try:
                return TORSION_EMPIRICAL[self.geom[a].atomic_number]
            except KeyError:
                raise AssignmentError(f"no empirical torsion parameters were found for type {self.types[a]}") from None

        (uj, vj), (uk, vk) = uv(j), uv(k)
        same_ring = any(k in self.ring_sets[r] for r in self.rings_of[j])
        v1 = v2 = v3 = f32(0)
        if gj.linear or gk.linear:                                           # rule (a)
            return zero
        if gj.arom and gk.arom and same_ring:                                # rule (b)
            pi = f32(0.5) if (not gj.pilp and not gk.pilp) else f32(0.3)
            beta = f32(3) if {gj.valence, gk.valence} == {3, 4} else f32(6)
            v2 = f32(float(f32(beta * pi)) * math.sqrt(float(f32(f32(uj) * f32(uk)))))
        elif self.order[(j, k)] == 2:                                        # rule (c)
            pi = f32(1) if (gj.mltb == 2 and gk.mltb == 2) else f32(0.4)
            v2 = f32(float(f32(f32(6) * pi)) * math.sqrt(float(f32(f32(uj) * f32(uk)))))
        elif gj.crd == 4 and gk.crd == 4:                                    # rule (d)
            v3 = f32(math.sqrt(float(f32(f32(vj) * f32(vk)))) / ((gj.crd - 1) * (gk.crd - 1)))
        elif (gj.crd == 4) != (gk.crd == 4):                                 # rules (e), (f)
            g = gk if gj.crd == 4 else gj
            if g.crd == 3 and (g.valence in (4, 34) or g.mltb != 0):
                return zero
            if g.crd == 2 and (g.valence == 3 or g.mltb != 0):
                return zero
            v3 = f32(math.sqrt(float(f32(f32(vj) * f32(vk)))) / ((gj.crd - 1) * (gk.crd - 1)))
        elif ((self.order[(j, k)] == 1 and gj.mltb != 0 and gk.mltb != 0) or (gj.mltb != 0 and gk.pilp)
              or (gj.pilp and gk.mltb != 0)):                                # rule (g)
            if gj.pilp and gk.pilp:
                return zero
            root = math.sqrt(float(f32(f32(uj) * f32(uk))))
            if gj.pilp and gk.mltb != 0:
                pi = f32(0.15)
                if gj.mltb == 1:
                    pi = f32(0.5)
                elif self._row(j) == 1 or self._row(k) == 1:
                    pi = f32(0.3)
                v2 = f32(float(f32(f32(6) * pi)) * root)
            elif gk.pilp and gj.mltb != 0:
                pi = f32(0.15)
                if gk.mltb == 1:
                    pi = f32(0.5)
                elif self._row(j) == 1 or self._row(k) == 1:
                    pi = f32(0.3)
                v2 = f32(float(f32(f32(6) * pi)) * root)
            elif (gj.mltb == 1 or gk.mltb == 1) and (gj.atomic_number != 6 or gk.atomic_number != 6):
                v2 = f32(6 * .4 * root)
            else:
                v2 = f32(6 * .15 * root)
        elif gj.atomic_number in (8, 16) and gk.atomic_number in (8, 16):    # rule (h)
            ww = (2.0 if gj.atomic_number == 8 else 8.0) * (2.0 if gk.atomic_number == 8 else 8.0)
            v2 = f32(-math.sqrt(float(f32(ww))))
        else:
            v3 = f32(f32(math.sqrt(float(f32(f32(vj) * f32(vk))))) / f32((gj.crd - 1) * (gk.crd - 1)))
        return v1, v2, v3

    # ---- formal charges: MMFF94.java:404-564
    _FIXED_FORMAL = {34: 1.0, 49: 1.0, 51: 1.0, 54: 1.0, 58: 1.0, 92: 1.0, 93: 1.0, 94: 1.0, 97: 1.0,
                     87: 2.0, 95: 2.0, 96: 2.0, 98: 2.0, 99: 2.0, 88: 3.0,
                     35: -1.0, 62: -1.0, 89: -1.0, 90: -1.0, 91: -1.0}

    def formal_charges(self, input_formal: Optional[Sequence[int]] = None) -> List[float]:
        """``calculateFormalCharge``: the MMFF formal charge of every atom from its type and surroundings - fixed
        values for the ionic types, a negative charge shared among the terminal O / S of carboxylates, phosphates,
        sulfonates ... (types 32, 72: ``:411-468``), among the nitrogens of a tetrazole-like anion ring (76,
        ``:470-482``), and the integer charges the INPUT carries (``IAtom.getFormalCharge``) averaged over the
        conjugated nitrogens of amidinium / guanidinium / imidazolium groups (55, 56, 81: ``:483-516``).  An atom counts
        as aromatic when one of its bonds does."""
        n = self.n
        input_formal = [0] * n if input_formal is None else [int(round(float(x))) for x in input_formal]
        z = [g.atomic_number for g in self.geom]
        aromatic_atom = [any(self.arom[(a, b)] for b in self.nbrs[a]) for a in range(n)]
        out = [0.0] * n
        for a in range(n):
            ta = self.types[a]
            if ta in (32, 72):
                amine2, terminal = 0, 0                       # (running counts over the neighbours, as in the reference)
                for nb in self.nbrs[a]:
                    for n2 in self.nbrs[nb]:
                        if z[n2] == 7 and len(self.nbrs[n2]) == 2 and not aromatic_atom[n2]:
                            amine2 += 1
                        elif z[n2] in (8, 16) and len(self.nbrs[n2]) == 1:
                            terminal += 1
                    if z[nb] == 16 and terminal == 2 and amine2 == 1:
                        amine2 = 0                            # sulfonamide
                    tn = self.types[nb]
                    if z[nb] == 6 and terminal != 0:
                        out[a] = -1.0 if terminal == 1 else -(terminal - 1) / float(terminal)
                        break
                    if tn == 45 and terminal == 3:
                        out[a] = -1.0 / 3.0
                        break
                    if tn in (25, 73) and terminal != 0:
                        out[a] = 0.0 if terminal == 1 else -(terminal - 1) / float(terminal)
                        break
                    if tn == 18 and terminal != 0:
                        out[a] = 0.0 if amine2 + terminal == 2 else -(amine2 + terminal - 2) / float(terminal)
                        break
                    if tn == 77 and terminal != 0:
                        out[a] = -1.0 / float(terminal)
                        break
            elif ta == 76:
                if self.rings_of[a]:
                    ring = self.rings[self.rings_of[a][0]]
                    out[a] = -1.0 / float(sum(1 for x in ring if self.types[x] == 76))
            elif ta in (55, 56, 81):
                group = {a}
                charge = input_formal[a]
                grown = True
                while grown:
                    grown = False
                    for x in sorted(group):
                        for nb in self.nbrs[x]:
                            if self.types[nb] not in (57, 80):
                                continue
                            for n2 in self.nbrs[nb]:
                                if self.types[n2] in (55, 56, 81) and n2 not in group:
                                    group.add(n2)
                                    charge += input_formal[n2]
                                    grown = True
                out[a] = charge / float(len(group))
            elif ta == 61:
                if any(self.types[nb] == 42 for nb in self.nbrs[a]):
                    out[a] = 1.0
            elif ta in self._FIXED_FORMAL:
                out[a] = self._FIXED_FORMAL[ta]
        return out

    # ---- charges: MMFF94.java:566-618 with ring-aware bond types
    def partial_charges(self, formal: Sequence[float]) -> np.ndarray:
        t = self.t
        q = np.zeros(self.n, dtype=np.float64)
        for a in range(self.n):
            ta = self.types[a]
            q0 = float(formal[a])
            fcadj = t.pbci[ta - 1][1]
            if fcadj == f32(0.0):
                for nb in self.nbrs[a]:
                    k0 = float(formal[nb])
                    if k0 < 0.0:
                        q0 += k0 / (2.0 * len(self.nbrs[nb]))
            if ta == 62:
                for nb in self.nbrs[a]:
                    k0 = float(formal[nb])
                    if k0 > 0.0:
                        q0 -= k0 / 2.0
            qi, qk_sum = 0.0, 0.0
            for nb in self.nbrs[a]:
                tn = self.types[nb]
                qk_sum += float(formal[nb])
                bt = self.bond_type(a, nb)
                i, j = min(ta, tn), max(ta, tn)
                hit = t.bci_by_key.get((bt, i, j))
                if hit is not None:
                    qi += float(f32(hit * f32(-1 if ta > tn else 1)))
                else:
                    qi += float(f32(t.pbci[ta - 1][0] - t.pbci[tn - 1][0]))
            term = f32(f32(1) - f32(f32(self.geom[a].crd) * fcadj))
            qi += float(term) * q0 + qk_sum * float(fcadj)
            q[a] = qi
        return q


def assign_parameters(tables: MMFFTables, xyz, types: Sequence[int], bonds: Sequence[Tuple[int, int]],
                      orders: Sequence[int], aromatic: Optional[Sequence[bool]] = None,
                      formal: Optional[Sequence[float]] = None, name: str = "", expected_energy: Optional[float] = None,
                      source: str = "", input_formal: Optional[Sequence[int]] = None) -> MolecularSystem:
    """The flattened result of ``MMFF94.assignParameters`` (``MMFF94.java:111-219``) for given atom types.  ``formal``:
    MMFF formal charges when the caller has them; otherwise they are derived as the reference does (``formal_charges``)
    from the types and the integer formal charges of the input structure (``input_formal``, default all zero)."""
    A = Assigner(tables, types, bonds, orders, aromatic)
    n = A.n
    formal = A.formal_charges(input_formal) if formal is None else [float(f) for f in formal]
    linear = [1 if g.linear else 0 for g in A.geom]
    charges = A.partial_charges(formal)
    bond_kb, bond_r0 = [], []
    for a, b in A.bonds:
        kb, r0 = A.bond_parameters(a, b)
        bond_kb.append(kb)
        bond_r0.append(r0)
        A.r0[(a, b)] = A.r0[(b, a)] = r0
    angles, oops, torsions = enumerate_terms(A.types, linear, A.bonds)
    ka, th0, kijk, kkji = [], [], [], []
    for (i, j, k) in angles:
        a_ka, a_th0 = A.angle_parameters(i, j, k)
        s_ijk, s_kji = A.stretch_bend_parameters(i, j, k)
        ka.append(a_ka); th0.append(a_th0); kijk.append(s_ijk); kkji.append(s_kji)
    oop_keep, koop = [], []
    for (i, j, k, l) in oops:
        v = tables.find_oop(A.types[i], A.types[j], A.types[k], A.types[l])
        if v is not None:
            oop_keep.append((i, j, k, l))
            koop.append(v)
    v1, v2, v3 = [], [], []
    for (i, j, k, l) in torsions:
        a, b, c = A.torsion_parameters(i, j, k, l)
        v1.append(a); v2.append(b); v3.append(c)
    used = sorted(set(A.types))
    rows = [tables.vdw_row(t) for t in used]

    def col(seq, idx):
        return [t[idx] for t in seq]

    return MolecularSystem(
        xyz=np.asarray(xyz, dtype=np.float64).reshape(-1, 3), atom_type=A.types, charge=charges, linear=linear,
        vdw_type=used, vdw_alpha=[r.alpha for r in rows], vdw_N=[r.N for r in rows],
        vdw_A=[r.A for r in rows], vdw_G=[r.G for r in rows], vdw_DA=[r.DA for r in rows],
        bond_i=col(A.bonds, 0), bond_j=col(A.bonds, 1), bond_kb=bond_kb, bond_r0=bond_r0,
        angle_i=col(angles, 0), angle_j=col(angles, 1), angle_k=col(angles, 2),
        angle_ka=ka, angle_theta0=th0, stbn_kba_ijk=kijk, stbn_kba_kji=kkji,
        oop_i=col(oop_keep, 0), oop_j=col(oop_keep, 1), oop_k=col(oop_keep, 2), oop_l=col(oop_keep, 3),
        oop_koop=koop,
        tor_i=col(torsions, 0), tor_j=col(torsions, 1), tor_k=col(torsions, 2), tor_l=col(torsions, 3),
        tor_v1=v1, tor_v2=v2, tor_v3=v3,
        name=name, expected_energy=expected_energy, source=source)
