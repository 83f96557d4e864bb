"""Rule-based MMFF94 atom typing from connectivity and bond orders.

The reference delegates typing to CDK (``Mmff.assignAtomTypes``, ``mmff/MMFF94.java:111-137``: third-party, needs a JVM).
This module is NOT a port of CDK: it is a rule set that covers the chemistry of the reference's two validation suites
well enough to reproduce the published OPTIMOL energy AND partial charges of 916 of their 1 026 records (681 of 761
MMFF94, 235 of 265 MMFF94s; ``tests/golden/make_golden.py`` keeps a record only when both agree) - acyclic functional
groups, saturated and small rings, benzenoid aromatics, pyridines, five-ring heteroaromatics with one lone-pair donor,
and the common ions (ammonium, carboxylate, alkoxide, thiolate, amidinium / guanidinium, pyridinium, imidazolium,
N-oxides, azides, sulfonamide anions).  Outside that it returns ``None`` rather than guess.  Where connectivity alone
cannot tell two types apart it takes its first choice; ``alternative`` selects the other reading (the fixture
generator tries them and lets the published values decide).  ``tests/test_atom_typing.py`` pins the rules against the
committed typings of all 916 records.  Host-side set-up only - nothing here runs on the hot path.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Set, Tuple

from .assign import find_rings

ALTERNATIVES = ("", "nn8", "enamine8", "amide40")


class Typing(NamedTuple):
    types: List[int]                 # MMFF numeric type per atom
    aromatic_bonds: Set[frozenset]   # bonds perceived as aromatic (frozenset of the two atom indices)
    input_formal: List[int]          # integer formal charges as the input structure would carry them (for formal_charges)


def perceive_types(atoms: Sequence[Sequence], bonds: Sequence[Tuple[int, int, str]], allow_rings: bool = True,
                   alternative: str = "") -> Optional[Typing]:
    """``atoms``: per atom a sequence whose first item is the upper-case element symbol; ``bonds``: (i, j, order) with
    order "1", "2", "3", "ar" or "am" (mol2 conventions).  Returns the typing or ``None`` when the molecule is outside
    the rules."""
    ALTERNATIVE = alternative
    result = {}
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
    result["arom"] = arom_bonds

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
    return Typing(t, set(result["arom"]), list(input_formal))


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


def system_from_structure(tables, elements: Sequence[str], xyz, bonds: Sequence[Tuple[int, int, str]], name: str = "",
                          alternative: str = ""):
    """Typed and parameterised ``MolecularSystem`` (what ``MMFF94.assignParameters`` leaves behind, ``MMFF94.java:
    111-219``) from elements, coordinates and bonds with mol2-style orders: ``perceive_types``, then the reference's
    assignment restated in ``assign.py`` (ring-dependent types, tables, empirical rules, formal and partial charges).
    Aromatic ("ar") bonds take the first Kekule structure that parameterises.  Raises ``ValueError`` when the molecule is
    outside the typing rules."""
    from .assign import AssignmentError, assign_parameters
    from .mmff_tables import ParameterNotTabulated
    atoms = [(e.upper(),) for e in elements]
    typing = perceive_types(atoms, bonds, allow_rings=True, alternative=alternative)
    if typing is None:
        raise ValueError("atom typing: the molecule is outside the rule set (see jmolecularenergy_b200/atom_typing.py)")
    pairs = [(a, b) for a, b, _ in bonds]
    aromatic = [frozenset((a, b)) in typing.aromatic_bonds for a, b, _ in bonds]
    last: Optional[Exception] = None
    for orders in kekule_variants(len(atoms), bonds):
        try:
            return assign_parameters(tables, xyz, typing.types, pairs, orders, aromatic, None, name=name,
                                     input_formal=typing.input_formal)
        except (ParameterNotTabulated, AssignmentError, KeyError) as exc:
            last = exc
    raise ValueError(f"parameter assignment failed for every Kekule structure: {last}")
