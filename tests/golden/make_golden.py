#!/usr/bin/env python
"""Generate the golden fixtures under ``tests/golden/`` from the reference's own test data.

Run in the build container only (it reads ``/root/reference``, which does not exist on
the GPU box); the JSON it writes is committed and is what the tests read.

For each hand-typed molecule of the reference's MMFF94 validation suite
(``src/test/resources/org/jme/forcefield/mmff/MMFF94_dative.mol2`` geometry,
``MMFF94.energies`` OPTIMOL column - the same files ``MMFF94Test.java:70-90`` reads) it

1. takes coordinates and bonds from the mol2 record,
2. assigns MMFF numeric atom types BY HAND (the CDK typer is not available here) and
   formal charges for the bare ions,
3. looks every parameter up in the reference's ``.par`` tables (table hits only, see
   ``jmolecularenergy_b200/builder.py``),
4. evaluates the pure-Python restatement (``oracle/py_oracle.py``) and keeps the
   molecule only if it reproduces the published total AND its bond-charge-increment charges
   round to the USER_CHARGES column of the mol2 record.

A molecule that fails step 4 is reported and skipped - that means the typing (or a
needed empirical rule) is wrong, and it must not become a fixture.

Besides the hand-typed molecules, every ACYCLIC, neutral record of the suite is typed by a small rule
set (``auto_types``: neighbour counts and bond orders only - no ring or aromaticity perception, so ring
systems are left alone) and goes through the same two checks.  The checks are what makes this safe: over
the suite a correct typing lands within 6e-5 kcal/mol of the published value (the published numbers carry
5 decimals and OPTIMOL's own rounding), a wrong one is off by more than 1 kcal/mol, with nothing in
between.  The tolerance a fixture met (1e-5 or 1e-4) is recorded in MANIFEST.json and is what the tests use.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from jmolecularenergy_b200.assign import AssignmentError, assign_parameters, find_rings   # noqa: E402
from jmolecularenergy_b200.builder import parameterize            # noqa: E402
from jmolecularenergy_b200.mmff_tables import MMFFTables, ParameterNotTabulated  # noqa: E402
from oracle import py_oracle                                     # noqa: E402

REF = "/root/reference"
PAR_DIR = f"{REF}/src/main/resources/org/jme/forcefield/mmff"
MOL2 = f"{REF}/src/test/resources/org/jme/forcefield/mmff/MMFF94_dative.mol2"
ENERGIES = f"{REF}/src/test/resources/org/jme/forcefield/mmff/MMFF94.energies"

LAST_AROMATIC_BONDS = set()
LAST_INPUT_FORMAL = []
# Where the rules cannot tell two MMFF types apart from connectivity alone, pin_suite tries the alternatives in turn and
# the published energy and charges decide: "" = the first choice of every rule.
ALTERNATIVES = ("", "nn8", "enamine8", "amide40")
ALTERNATIVE = ""
WATER = {"O": 70, "H": 31}
# name -> (type per atom by mol2 atom order | callable(symbol)->type, {atom index: formal charge})
HAND_TYPED = {
    # ion - water complexes: OH2=70, HOH=31, ions 87..99 (MMFF94-atom-type-equivalence.par)
    "NAPW":   (lambda el: {"NA": 93, **WATER}[el], {"NA": 1.0}),
    "KPW1":   (lambda el: {"K": 94, **WATER}[el], {"K": 1.0}),
    "LIPW1":  (lambda el: {"LI": 92, **WATER}[el], {"LI": 1.0}),
    "FMW1":   (lambda el: {"F": 89, **WATER}[el], {"F": -1.0}),
    "CLMW1":  (lambda el: {"CL": 90, **WATER}[el], {"CL": -1.0}),
    "BRMW1":  (lambda el: {"BR": 91, **WATER}[el], {"BR": -1.0}),
    "CA2PW3": (lambda el: {"CA": 96, **WATER}[el], {"CA": 2.0}),
    "MG2PW3": (lambda el: {"MG": 99, **WATER}[el], {"MG": 2.0}),
    "ZN2PW3": (lambda el: {"ZN": 95, **WATER}[el], {"ZN": 2.0}),
    "CU1PW1": (lambda el: {"CU": 97, **WATER}[el], {"CU": 1.0}),
    "CU2PW3": (lambda el: {"CU": 98, **WATER}[el], {"CU": 2.0}),
    "FE2PW3": (lambda el: {"FE": 87, **WATER}[el], {"FE": 2.0}),
    "FE3PW3": (lambda el: {"FE": 88, **WATER}[el], {"FE": 3.0}),
    # formaldehyde: C=O 3, O=C 7, HC 5
    "CO01A":  (lambda el: {"C": 3, "O": 7, "H": 5}[el], {}),
    # formic acid (mol2 order O.3, O.2, C.2, H(C), H(O)): OC=O 6, O=C 7, C=O 3, HC 5, HOCO 24
    "KHDFRM11": ([6, 7, 3, 5, 24], {}),
    # ammonia: NR 8, HNR 23 - pins a non-zero Wilson angle
    "NH10A":  (lambda el: {"N": 8, "H": 23}[el], {}),
    # hydrogen sulfide: S 15, HS 71
    "SR01A":  (lambda el: {"S": 15, "H": 71}[el], {}),
    # phosphine: P 26, HP 71
    "PR01A":  (lambda el: {"P": 26, "H": 71}[el], {}),
    # small neutral organics, typed per atom in mol2 order (torsions, sp2 out-of-plane, 1-4 pairs)
    "HL13A":  ([11, 2, 2, 5, 5, 5], {}),                       # vinyl fluoride
    "OH10A":  ([2, 2, 6, 5, 5, 5, 29], {}),                    # vinyl alcohol (HOCC = 29)
    "CA04A":  ([7, 3, 3, 7, 6, 24, 5], {}),                    # glyoxylic acid
    "CE05A":  ([2, 2, 6, 3, 7, 5, 5, 5, 5], {}),               # vinyl formate
    "CO08A":  ([2, 1, 3, 7, 5, 2, 5, 5, 5, 5, 5], {}),         # but-3-enal
    "IM02A":  ([3, 9, 1, 5, 5, 5, 5, 5], {}),                  # N-methylformaldehydeimine
    "SR05A":  ([15, 15, 1, 71, 5, 5, 5], {}),                  # methyl hydrogen disulfide
    "PR02A":  ([1, 26, 71, 71, 5, 5, 5], {}),                  # methylphosphine
    "SI02A":  ([19, 1, 5, 5, 5, 5, 5, 5], {}),                 # methylsilane
    "SI03A":  ([19, 1, 6, 5, 5, 5, 5, 5, 21], {}),             # methyl hydroxyl silane
}


def auto_types(atoms, bonds, allow_rings=False):
    """MMFF numeric types for an acyclic, neutral molecule from connectivity and bond orders alone, or None when
    the molecule is outside what these rules cover.  Types (MMFFSYMB): C 1 alkyl / 2 vinylic / 3 C=O, C=N, C=S /
    4 sp; H 5 on C, Si / 21 alcohol / 24 acid, on P-O / 29 enol / 33 on S-O / 23 amine / 28 amide, enamine,
    sulfonamide / 27 imine / 71 on S, P / 31 water; O 6 divalent / 7 carbonyl, nitroso, sulfoxide / 32 terminal O
    on N (nitro), S (sulfone...), P / 70 water; N 8 amine / 9 imine, azo / 10 amide, hydrazone N / 40 enamine /
    42 nitrile / 43 sulfonamide / 45 nitro / 46 nitroso; S 15 / 16 thiocarbonyl / 17 sulfoxide / 18 sulfone;
    P 26 / 25; Si 19; halogens 11-14."""
    n = len(atoms)
    el = [a[0] for a in atoms]
    nb = [[] for _ in range(n)]
    aromatic_atom = [False] * n
    for a, b, o in bonds:
        o = "1" if o == "am" else o
        if o == "ar":                          # an aromatic bond counts as a multiple bond for the rules below
            aromatic_atom[a] = aromatic_atom[b] = True
            o = "2"
        if o not in ("1", "2", "3"):
            return None
        nb[a].append((b, int(o)))
        nb[b].append((a, int(o)))
    rings = find_rings(n, [(a, b) for a, b, _ in bonds])
    if rings and not allow_rings:
        return None
    if len(bonds) < n - 1:                    # not connected
        return None
    ring_sizes = [set() for _ in range(n)]
    for r in rings:
        for a in r:
            ring_sizes[a].add(len(r))
    arom_bonds = {frozenset((a, b)) for a, b, o in bonds if o == "ar"}
    for r in rings:                           # mol2 "ar" systems other than six-membered rings are left alone
        ar_in_ring = sum(1 for k in range(len(r)) if frozenset((r[k], r[(k + 1) % len(r)])) in arom_bonds)
        if ar_in_ring == len(r) and len(r) != 6:
            return None
    # Aromaticity from the Kekule structure (the suites write rings with alternating single / double bonds):
    #  * six-membered rings of C / N in which every atom has a double bond inside the ring or already belongs to an
    #    aromatic ring (fused benzenoids), iterated to a fixed point;
    #  * five-membered rings with two ring double bonds and exactly one atom that donates a lone pair (O, S, or N with
    #    three neighbours): furans, thiophenes, pyrroles, azoles.
    order_of = {frozenset((a, b)): ("2" if o == "ar" else ("1" if o == "am" else o)) for a, b, o in bonds}
    ring_bonds = lambda r: [frozenset((r[k], r[(k + 1) % len(r)])) for k in range(len(r))]
    arom_ring = []
    changed = True
    while changed:
        changed = False
        for r in rings:
            if r in arom_ring or len(r) != 6 or any(el[a] not in ("C", "N") for a in r):
                continue
            rb = ring_bonds(r)
            ok = all(any(order_of[b] == "2" and a in b for b in rb) or aromatic_atom[a] for a in r)
            if ok and all(len(nb[a]) <= 3 for a in r):
                arom_ring.append(r)
                for a in r:
                    aromatic_atom[a] = True
                arom_bonds.update(rb)
                changed = True
    five_role = {}                            # atom -> "X" (lone-pair donor), "A" (alpha to it), "B" (beta)
    for r in rings:
        if len(r) != 5:
            continue
        rb = ring_bonds(r)
        donors = [a for a in r if (el[a] in ("O", "S") and len(nb[a]) == 2) or (el[a] == "N" and len(nb[a]) == 3
                                                                                and all(o == 1 for _, o in nb[a]))]
        n_dbl = sum(1 for b in rb if order_of[b] == "2")
        others_sp2 = all(any(o == 2 for _, o in nb[a]) or aromatic_atom[a] for a in r if a not in donors)
        if len(donors) == 1 and others_sp2 and (n_dbl == 2 or (n_dbl == 1 and any(aromatic_atom[a] for a in r))) \
                and all(el[a] in ("C", "N", "O", "S") for a in r):
            x = r.index(donors[0])
            five_role[r[x]] = "X"
            for d, role in ((1, "A"), (4, "A"), (2, "B"), (3, "B")):
                five_role.setdefault(r[(x + d) % 5], role)
            arom_bonds.update(rb)
            for a in r:
                aromatic_atom[a] = True
    global LAST_AROMATIC_BONDS
    LAST_AROMATIC_BONDS = arom_bonds

    def double_to(i, els):
        return any(o == 2 and el[j] in els for j, o in nb[i])

    def top(i):
        return max((o for _, o in nb[i]), default=0)

    def terminal_chalcogens(i):
        return [j for j, _ in nb[i] if el[j] in ("O", "S") and len(nb[j]) == 1]

    def carboxylate_c(i):                       # CO2M / CS2M: a carbon with two terminal O / S and no acid hydrogen
        return el[i] == "C" and len(nb[i]) == 3 and len(terminal_chalcogens(i)) == 2

    def cation_n(i):                            # a nitrogen with three neighbours and a double bond that is not a nitro group
        return el[i] == "N" and len(nb[i]) == 3 and top(i) == 2 and not aromatic_atom[i] and len(terminal_chalcogens(i)) == 0

    def cation_c_nitrogens(i):                  # the nitrogens on a carbon that is double-bonded to a cationic nitrogen, else 0
        if el[i] != "C" or len(nb[i]) != 3 or not any(o == 2 and cation_n(j) for j, o in nb[i]):
            return 0
        return sum(1 for j, _ in nb[i] if el[j] == "N" and len(nb[j]) == 3 and not aromatic_atom[j])

    # integer formal charges as the input structure carries them (the reference reads IAtom.getFormalCharge for the
    # amidinium / guanidinium / imidazolium sharing rule): +1 on the nitrogen drawn with the double bond
    input_formal = [0] * n
    t = [0] * n
    for i in range(n):
        e, deg, mo = el[i], len(nb[i]), top(i)
        if e == "C":
            if deg == 4:
                t[i] = 22 if 3 in ring_sizes[i] else (20 if 4 in ring_sizes[i] else 1)     # CR3R, CR4R, CR
            elif carboxylate_c(i):
                t[i] = 41                                                                  # CO2M, CS2M
            elif cation_c_nitrogens(i) >= 2 and not aromatic_atom[i]:
                t[i] = 57                                                                  # CNN+, CGD+
            elif deg == 3 and five_role.get(i) in ("A", "B") and not double_to(i, ("O", "S")) \
                    and not any(o == 2 and not aromatic_atom[j] for j, o in nb[i]):
                t[i] = 63 if five_role[i] == "A" else 64                                   # C5A, C5B
            elif deg == 3 and aromatic_atom[i] and not any(o == 2 and not aromatic_atom[j] for j, o in nb[i]):
                t[i] = 37                                                                  # CB
            elif deg == 3 and mo == 2:
                if double_to(i, ("O", "N", "S")):
                    t[i] = 3
                elif 4 in ring_sizes[i] and any(o == 2 and 4 in ring_sizes[j] for j, o in nb[i]):
                    t[i] = 30                                                              # CE4R
                else:
                    t[i] = 2
            elif deg == 2 and (mo == 3 or sum(o == 2 for _, o in nb[i]) == 2):
                t[i] = 4
            else:
                return None
        elif e == "O":
            if deg == 2 and five_role.get(i) == "X":
                t[i] = 59                                                                  # OFUR
            elif deg == 2:
                t[i] = 70 if all(el[j] == "H" for j, _ in nb[i]) else 6
            elif deg == 1:
                j, o = nb[i][0]
                if el[j] == "C" and carboxylate_c(j):
                    t[i] = 32                                                              # O2CM
                elif el[j] == "C" and o == 2:
                    t[i] = 7
                elif el[j] == "C" and o == 1:
                    t[i] = 35                                                              # OM: alkoxide, phenoxide, enolate
                elif el[j] == "N":
                    t[i] = 7 if len(nb[j]) == 2 else 32
                elif el[j] == "S":
                    terminal = sum(1 for k, _ in nb[j] if el[k] == "O" and len(nb[k]) == 1)
                    t[i] = 7 if ((len(nb[j]) == 3 and terminal == 1) or len(nb[j]) == 2) else 32      # sulfoxide, sulfine
                elif el[j] == "P":
                    t[i] = 32
                else:
                    return None
            else:
                return None
        elif e == "S":
            t[i] = {(2, 1): 15, (1, 2): 16}.get((deg, mo), {3: 17, 4: 18}.get(deg, 0))
            if deg == 1:
                j, o = nb[i][0]
                if el[j] == "C" and o == 2 and not carboxylate_c(j):
                    t[i] = 16                                                              # S=C
                else:
                    t[i] = 72                                                              # S2CM, SM, terminal S on P / S
            if deg == 2 and sum(o == 2 for _, o in nb[i]) == 2:
                t[i] = 74                                                                  # =S=O
            if deg == 2 and five_role.get(i) == "X":
                t[i] = 44                                                                  # STHI
            if not t[i]:
                return None
        elif e == "P":
            t[i] = {3: 26, 4: 25}.get(deg, 0)
            if not t[i]:
                return None
        elif e == "SI":
            t[i] = 19
        elif e in ("F", "CL", "BR", "I"):
            if deg != 1:
                return None
            t[i] = {"F": 11, "CL": 12, "BR": 13, "I": 14}[e]
        elif e not in ("N", "H"):
            return None
    for i in range(n):                           # nitrogen: needs its neighbours' environment
        if el[i] != "N":
            continue
        deg, mo = len(nb[i]), top(i)
        if aromatic_atom[i]:
            if five_role.get(i) == "X" and deg == 3:
                if t[i] != 81:
                    t[i] = 39                                                              # NPYL (unless an imidazolium partner already claimed it)
                continue
            if five_role.get(i) in ("A", "B") and deg == 2:
                t[i] = 65 if five_role[i] == "A" else 66                                   # N5A, N5B
                continue
            if deg == 2:
                t[i] = 38                                                                  # NPYD
                continue
            if deg == 3 and 6 in ring_sizes[i] and i not in five_role:
                t[i] = 69 if terminal_chalcogens(i) else 58                                # NPOX, NPD+
                continue
            if deg == 3 and mo == 2 and five_role.get(i) == "B" and not terminal_chalcogens(i):
                # imidazolium-type cation: this nitrogen (drawn with the double bond) and the ring's lone-pair nitrogen two
                # places away become NIM+, the carbon between them CIM+, the ring's other carbons the general C5
                ring = [r for r in rings if len(r) == 5 and i in r and any(five_role.get(a) == "X" and el[a] == "N" for a in r)]
                if len(ring) != 1:
                    return None
                ring = ring[0]
                x = [a for a in ring if five_role.get(a) == "X"][0]
                between = [a for a in ring if el[a] == "C" and any(j == i for j, _ in nb[a]) and any(j == x for j, _ in nb[a])]
                if len(between) != 1 or len(nb[x]) != 3:
                    return None
                t[i] = t[x] = 81                                                           # NIM+
                t[between[0]] = 80                                                         # CIM+
                for a in ring:
                    if el[a] == "C" and a != between[0] and t[a] in (63, 64):
                        t[a] = 78                                                          # C5
                input_formal[i] = 1
                continue
            return None
        if deg == 1 and mo == 3:
            t[i] = 42
        elif deg == 1 and mo == 2 and el[nb[i][0][0]] == "N":
            t[i] = 47                                                                      # NAZT
        elif deg == 2 and sum(o == 2 for _, o in nb[i]) == 2:
            t[i] = 53                                                                      # =N=
        elif deg == 2 and mo == 2:
            t[i] = 46 if double_to(i, ("O",)) else 9
        elif deg == 2 and mo == 1 and any(el[j] == "S" and len(nb[j]) == 4 for j, _ in nb[i]):
            t[i] = 62                                                                      # NM: deprotonated sulfonamide
        elif deg == 4:
            t[i] = 68 if terminal_chalcogens(i) else 34                                    # N3OX, NR+
        elif deg == 3 and mo == 2 and len(terminal_chalcogens(i)) == 1 and double_to(i, ("C", "N")):
            t[i] = 67                                                                      # N2OX
        elif cation_n(i):
            c = [j for j, o in nb[i] if o == 2][0]
            k = cation_c_nitrogens(c)
            t[i] = 56 if k == 3 else (55 if k == 2 else 54)                                # NGD+, NCN+, N+=C
            input_formal[i] = 1
        elif deg == 3 and mo == 1 and any(cation_c_nitrogens(j) >= 2 for j, _ in nb[i]):
            c = [j for j, _ in nb[i] if cation_c_nitrogens(j) >= 2][0]
            t[i] = 56 if cation_c_nitrogens(c) == 3 else 55
        elif deg == 3 and mo == 2:
            t[i] = 45
        elif deg == 3 and mo == 1:
            kind = 8
            if any(el[j] == "S" and len(nb[j]) == 4 for j, _ in nb[i]):
                kind = 43                                                                  # NSO2 (also when acylated)
            elif any(el[j] == "C" and len(nb[j]) == 2 and any(o == 3 and el[k] == "N" for k, o in nb[j]) for j, _ in nb[i]):
                kind = 43                                                                  # NC%N: on a cyano group
            elif any(el[j] == "C" and double_to(j, ("O", "S")) for j, _ in nb[i]):
                kind = 40 if ALTERNATIVE == "amide40" else 10
            elif any(el[j] == "N" and top(j) == 2 for j, _ in nb[i]):
                kind = 8 if ALTERNATIVE == "nn8" else 10                                    # N-N=X: NN=C / NN=N, or a plain amine
            elif any(el[j] == "S" and len(nb[j]) == 4 for j, _ in nb[i]):
                kind = 43
            elif any(el[j] == "C" and top(j) >= 2 for j, _ in nb[i]):
                kind = 8 if ALTERNATIVE == "enamine8" else 40
            t[i] = kind
        else:
            return None
    for i in range(n):                           # hydrogen: by the atom it sits on
        if el[i] != "H":
            continue
        if len(nb[i]) != 1:
            return None
        j = nb[i][0][0]
        e = el[j]
        if e in ("C", "SI"):
            t[i] = 5
        elif e in ("S", "P"):
            t[i] = 71
        elif e == "N":
            t[i] = {8: 23, 10: 28, 40: 28, 43: 28, 9: 27, 39: 23, 34: 36, 54: 36, 55: 36, 56: 36, 58: 36, 81: 36,
                    48: 28, 62: 23, 67: 23, 68: 23}.get(t[j], 0)
        elif e == "O":
            if t[j] == 70:
                t[i] = 31
            else:
                x = [k for k, _ in nb[j] if k != i]
                if len(x) == 1:
                    x = x[0]
                    if el[x] == "C":
                        t[i] = 24 if double_to(x, ("O",)) else (29 if t[x] in (2, 3, 37, 30, 63, 64) else 21)
                    elif el[x] == "S":
                        t[i] = 33
                    elif el[x] == "P":
                        t[i] = 24
                    else:
                        t[i] = 21                                                          # HOR: on N-O-H and the rest
        if not t[i]:
            return None
    global LAST_INPUT_FORMAL
    LAST_INPUT_FORMAL = input_formal
    return t


def kekule_variants(n, bonds):
    """Bond orders (1 / 2 / 3) with the aromatic ("ar") bonds of the mol2 record resolved: every perfect matching of
    the aromatic atoms found by backtracking (at most 4 are tried), then all-single as a last resort."""
    base = [{"1": 1, "2": 2, "3": 3, "am": 1, "ar": 0}[o] for _, _, o in bonds]
    ar = [k for k, o in enumerate(base) if o == 0]
    if not ar:
        return [base]
    atoms = sorted({x for k in ar for x in bonds[k][:2]})
    # an aromatic atom that already carries a double bond outside the aromatic system takes no ring double bond
    taken = {x for a, b, o in bonds if o == "2" for x in (a, b)}
    need = [a for a in atoms if a not in taken]
    out = []

    def solve(idx, used, chosen):
        if len(out) >= 4:
            return
        while idx < len(need) and need[idx] in used:
            idx += 1
        if idx == len(need):
            out.append(list(chosen))
            return
        a = need[idx]
        for k in ar:
            x, y = bonds[k][:2]
            if a in (x, y):
                other = y if x == a else x
                if other not in used and other not in taken:
                    solve(idx + 1, used | {a, other}, chosen + [k])
        # (an atom that stays unmatched, e.g. a pyridine-like N in a non-benzenoid system, is allowed last)
        solve(idx + 1, used | {a}, chosen)

    solve(0, frozenset(), [])
    variants = []
    for chosen in out:
        v = list(base)
        for k in ar:
            v[k] = 2 if k in chosen else 1
        if v not in variants:
            variants.append(v)
    v = [1 if o == 0 else o for o in base]
    if v not in variants:
        variants.append(v)
    return variants


def read_mol2(path):
    """name -> dict(atoms=[(element, x, y, z, mol2_charge)], bonds=[(a, b, order_str)], line)."""
    mols = {}
    with open(path) as fh:
        lines = fh.read().split("\n")
    n = 0
    while n < len(lines):
        if lines[n].startswith("@<TRIPOS>MOLECULE"):
            name = lines[n + 1].strip()
            na, nb = (int(v) for v in lines[n + 2].split()[:2])
            m = n
            while not lines[m].startswith("@<TRIPOS>ATOM"):
                m += 1
            atoms = []
            for row in lines[m + 1:m + 1 + na]:
                c = row.split()
                atoms.append((c[5].split(".")[0].upper(), float(c[2]), float(c[3]), float(c[4]), float(c[8])))
            while not lines[m].startswith("@<TRIPOS>BOND"):
                m += 1
            bonds = []
            for row in lines[m + 1:m + 1 + nb]:
                c = row.split()
                bonds.append((int(c[1]) - 1, int(c[2]) - 1, c[3]))
            mols[name] = dict(atoms=atoms, bonds=bonds, line=n + 2)
            n = m
        n += 1
    return mols


def read_energies(path):
    out = {}
    with open(path) as fh:
        for ln, line in enumerate(fh, 1):
            if line.startswith("*") or not line.strip():
                continue
            c = line.split()
            out[c[0]] = (float(c[1]), ln)
    return out


def ethanol_geometry():
    """Staggered (anti) ethanol from textbook internal coordinates; 9 atoms C C O H H H H H H.
    Not a reference fixture (ethanol is not in the validation suite): it is BASELINE.json's
    config 1 input, so it has no published energy."""
    def place(a, b, c, r, theta_deg, phi_deg):
        # position of atom d bonded to c, angle (b,c,d)=theta, dihedral (a,b,c,d)=phi
        th, ph = np.radians(theta_deg), np.radians(phi_deg)
        bc = c - b
        bc /= np.linalg.norm(bc)
        nrm = np.cross(b - a, bc)
        nrm /= np.linalg.norm(nrm)
        m = np.cross(nrm, bc)
        d2 = np.array([-r * np.cos(th), r * np.sin(th) * np.cos(ph), r * np.sin(th) * np.sin(ph)])
        return c + d2[0] * bc + d2[1] * m + d2[2] * nrm
    c1 = np.array([0.0, 0.0, 0.0])
    c2 = np.array([1.52, 0.0, 0.0])
    dummy = np.array([0.0, 1.0, 0.0])
    o = place(dummy, c1, c2, 1.43, 109.5, 180.0)
    h_o = place(c1, c2, o, 0.97, 108.0, 180.0)
    h21 = place(dummy, c1, c2, 1.09, 109.5, 60.0)
    h22 = place(dummy, c1, c2, 1.09, 109.5, -60.0)
    h11 = place(o, c2, c1, 1.09, 109.5, 180.0)
    h12 = place(o, c2, c1, 1.09, 109.5, 60.0)
    h13 = place(o, c2, c1, 1.09, 109.5, -60.0)
    xyz = np.round(np.array([c1, c2, o, h11, h12, h13, h21, h22, h_o]), 4)
    types = [1, 1, 6, 5, 5, 5, 5, 5, 21]          # CR CR OR HC*5 HOR
    bonds = [(0, 1), (1, 2), (0, 3), (0, 4), (0, 5), (1, 6), (1, 7), (2, 8)]
    return xyz, types, bonds


def _unit(v):
    return v / np.linalg.norm(v)


def build_chain(backbone: str):
    """All-anti zig-zag chain from a backbone string of 'C'/'O' (e.g. 'CCCCOCC', trailing 'O' = alcohol).
    Hydrogens complete every carbon to 4 and a terminal oxygen to 2 neighbours.  Returns
    (xyz, types, bonds) with MMFF types CR=1, OR=6, HC=5, HOR=21.  Synthetic-system template."""
    blen = {("C", "C"): 1.52, ("C", "O"): 1.43, ("O", "C"): 1.43}
    ang = np.radians(110.5)
    heavy = [np.zeros(3)]
    # zig-zag in the xz-plane
    direction = np.array([np.sin(ang / 2), 0.0, np.cos(ang / 2)])
    for n in range(1, len(backbone)):
        r = blen[(backbone[n - 1], backbone[n])]
        d = direction * np.array([1.0, 1.0, 1.0 if n % 2 else -1.0])
        heavy.append(heavy[-1] + r * d)
    xyz = [h.copy() for h in heavy]
    types = [1 if el == "C" else 6 for el in backbone]
    bonds = [(n - 1, n) for n in range(1, len(backbone))]
    tet = np.radians(109.47 / 2)
    for n, el in enumerate(backbone):
        nb = [m for m in (n - 1, n + 1) if 0 <= m < len(backbone)]
        c = heavy[n]
        if el == "C" and len(nb) == 2:
            u1, u2 = _unit(heavy[nb[0]] - c), _unit(heavy[nb[1]] - c)
            b, nrm = -_unit(u1 + u2), _unit(np.cross(u1, u2))
            for sgn in (1, -1):
                xyz.append(c + 1.09 * (np.cos(tet) * b + sgn * np.sin(tet) * nrm))
                types.append(5)
                bonds.append((n, len(xyz) - 1))
        elif len(nb) == 1:
            u1 = _unit(heavy[nb[0]] - c)
            # frame: e1 in the backbone plane (xz), e2 = y
            e2 = np.array([0.0, 1.0, 0.0])
            e1 = _unit(np.cross(e2, u1))
            if el == "C":
                th = np.radians(109.47)
                for k in range(3):
                    phi = np.radians(60.0 + 120.0 * k)
                    dvec = np.cos(np.pi - th) * u1 * -1 + np.sin(th) * (np.cos(phi) * e1 + np.sin(phi) * e2)
                    xyz.append(c + 1.09 * _unit(dvec))
                    types.append(5)
                    bonds.append((n, len(xyz) - 1))
            else:   # terminal hydroxyl
                th = np.radians(108.0)
                xyz.append(c + 0.97 * _unit(np.cos(th) * u1 + np.sin(th) * e1))
                types.append(21)
                bonds.append((n, len(xyz) - 1))
    return np.round(np.array(xyz), 4), types, bonds


def water_geometry(tables):
    kb, r0 = tables.find_bond(31, 70, 0)
    ka, th0 = tables.find_angle(31, 70, 31, 0)
    r, th = float(r0), np.radians(float(th0))
    xyz = np.array([[0.0, 0.0, 0.0], [r * np.sin(th / 2), 0.0, r * np.cos(th / 2)],
                    [-r * np.sin(th / 2), 0.0, r * np.cos(th / 2)]])
    return xyz, [70, 31, 31], [(0, 1), (0, 2)]


def write_templates(tables):
    """Single-molecule templates for the synthetic benchmark systems (jmolecularenergy_b200/systems.py):
    parameters looked up once here so that nothing needs the .par tables at run time."""
    out = os.path.join(ROOT, "jmolecularenergy_b200", "data")
    os.makedirs(out, exist_ok=True)
    lines = []
    xyz, types, bonds = water_geometry(tables)
    parameterize(tables, xyz, types, bonds, name="WATER", source="template: MMFF94 water, types 70/31").save(
        os.path.join(out, "WATER.json"))
    for name, t, q in (("NA", 93, 1.0), ("CL", 90, -1.0)):
        parameterize(tables, [[0.0, 0.0, 0.0]], [t], [], formal=[q], name=name,
                     source=f"template: monatomic ion type {t}").save(os.path.join(out, f"{name}.json"))
    for name, backbone in (("DECANE", "C" * 10), ("DIGLYME_LIKE", "CCOCCOCCOCC"), ("OCTANOL", "CCCCCCCCO"),
                           ("HEXANE", "C" * 6), ("BUTANOL", "CCCCO")):
        xyz, types, bonds = build_chain(backbone)
        s = parameterize(tables, xyz, types, bonds, name=name, source=f"template: all-anti chain {backbone}")
        s.save(os.path.join(out, f"{name}.json"))
        e = py_oracle.energy_components(s)
        lines.append(f"TEMPLATE {name}: {s.n_atoms} atoms, {len(s.tor_i)} torsions, E={e['total']:.4f}")
    return lines


def pin_suite(tables, mol2_path, energies_path, prefix, hand_typed, kept, tolerance, report):
    """Type, parameterise and check every record of one validation suite; fixtures that pass are written as
    ``<prefix><name>.json``."""
    mols = read_mol2(mol2_path)
    energies = read_energies(energies_path)
    candidates = [(name, hand_typed[name], "hand-assigned") for name in hand_typed if name in mols]
    for name in mols:
        if name not in hand_typed and name in energies:
            guess = auto_types(mols[name]["atoms"], mols[name]["bonds"], allow_rings=True)
            if guess is not None:
                candidates.append((name, (guess, {}), "rule-assigned (auto_types)"))
    global ALTERNATIVE
    for name, (typing, ion_charges), how in candidates:
        m = mols[name]
        verdict = None
        for alt in (ALTERNATIVES if how.startswith("rule") else ("",)):
            ALTERNATIVE = alt
            if callable(typing):
                types = [typing(a[0]) for a in m["atoms"]]
            else:
                types = list(typing)
            formal = [ion_charges.get(a[0], 0.0) for a in m["atoms"]]
            input_formal = None
            xyz = [[a[1], a[2], a[3]] for a in m["atoms"]]
            bonds = [(a, b) for a, b, _ in m["bonds"]]
            aromatic = [o == "ar" for _, _, o in m["bonds"]]
            if how.startswith("rule"):           # the perception of auto_types (run again: it keeps the last result)
                types = auto_types(m["atoms"], m["bonds"], allow_rings=True)
                if types is None or (alt and types == list(typing)):
                    continue
                aromatic = [frozenset((a, b)) in LAST_AROMATIC_BONDS for a, b, _ in m["bonds"]]
                formal, input_formal = None, list(LAST_INPUT_FORMAL)      # derived as the reference does (formal_charges)
            expected, eline = energies[name]
            src = (f"{os.path.basename(mol2_path)}:{m['line']} / {os.path.basename(energies_path)}:{eline} (OPTIMOL); "
                   f"{how}{' [' + alt + ']' if alt else ''} types {types}")
            done = False
            # mol2 writes aromatic bonds as "ar"; the reference sees a Kekule structure (CDK): its alternatives are tried,
            # the published energy and charges decide
            for orders in kekule_variants(len(types), m["bonds"]):
                try:
                    s = assign_parameters(tables, xyz, types, bonds, orders, aromatic, formal, name=prefix + name,
                                          expected_energy=expected, source=src, input_formal=input_formal)
                except (ParameterNotTabulated, AssignmentError, KeyError) as exc:
                    verdict = verdict if verdict and verdict.startswith("FAIL") else f"SKIP {prefix}{name}: {exc}"
                    continue
                got = py_oracle.energy_components(s)
                diff = abs(got["total"] - expected)
                # mol2 USER_CHARGES are printed to 4 decimals: the bci charges must round to them
                qok = np.allclose(s.charge, [a[4] for a in m["atoms"]], atol=5.1e-5)
                # three levels: 1e-5, 1e-4, and - only with matching charges - half the reference's own criterion
                # (MMFF94Test.java:79-81 asserts 0.01): the empirical-rule molecules land there (float rounding of the rules)
                ok = diff <= 1e-4 or (qok and diff <= 5e-3)
                verdict = (f"{'KEEP' if ok and qok else 'FAIL'} {prefix}{name}: E={got['total']:.6f} "
                           f"published={expected:.5f} diff={got['total'] - expected:+.2e} charges_ok={qok}")
                if ok and qok:
                    s.save(os.path.join(HERE, f"{prefix}{name}.json"))
                    kept.append(prefix + name)
                    if diff > 1e-4:
                        tolerance[prefix + name] = 1e-2
                    elif diff > 1e-5:
                        tolerance[prefix + name] = 1e-4
                    done = True
                    break
            if done:
                break
        ALTERNATIVE = ""
        report.append(verdict)


def main():
    tables = MMFFTables(PAR_DIR)
    kept, report, tolerance = [], [], {}
    pin_suite(tables, MOL2, ENERGIES, "", HAND_TYPED, kept, tolerance, report)
    # the MMFF94s suite: same typing, the static out-of-plane / torsion tables (MMFF94.java:265,660)
    pin_suite(MMFFTables(PAR_DIR, mmff94s=True), MOL2.replace("MMFF94_", "MMFF94s_"), ENERGIES.replace("MMFF94.", "MMFF94s."),
              "S_", {}, kept, tolerance, report)
    # ethanol: config-1 input, no published value
    xyz, types, bonds = ethanol_geometry()
    s = parameterize(tables, xyz, types, bonds, name="ETHANOL",
                     source="built from internal coordinates by tests/golden/make_golden.py; "
                            "BASELINE.json config 1; no published energy")
    s.save(os.path.join(HERE, "ETHANOL.json"))
    os.makedirs(os.path.join(ROOT, "jmolecularenergy_b200", "data"), exist_ok=True)
    s.save(os.path.join(ROOT, "jmolecularenergy_b200", "data", "ETHANOL.json"))
    report.append(f"INFO ETHANOL: E={py_oracle.energy_components(s)['total']:.6f} "
                  f"terms: {len(s.bond_i)} bonds {len(s.angle_i)} angles {len(s.oop_i)} oop "
                  f"{len(s.tor_i)} torsions")
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as fh:
        json.dump({"pinned": kept, "unpinned": ["ETHANOL"], "tolerance": tolerance}, fh, indent=1)
        fh.write("\n")
    report += write_templates(tables)
    print("\n".join(report))


if __name__ == "__main__":
    main()
