"""Table-hit parameter assignment: ``(types, bonds, coordinates)`` -> :class:`MolecularSystem`.

This is the acyclic, table-hit subset of ``MMFF94.assignParameters``
(``MMFF94.java:111-219``): atom types are given (the CDK typer is not available),
charges come from the bond-charge-increment table, every bonded parameter must be a
row of the ``.par`` tables (no empirical rules).  It is host-side marshalling for the
golden fixtures and the synthetic benchmark systems, not part of the GPU hot path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np

from .mmff_tables import MMFFTables
from .system import MolecularSystem, enumerate_terms


def parameterize(tables: MMFFTables, xyz, types: Sequence[int], bonds: Sequence[Tuple[int, int]],
                 formal: Optional[Sequence[float]] = None, bond_orders: Optional[Sequence[int]] = None,
                 name: str = "", expected_energy: Optional[float] = None, source: str = "",
                 charges: Optional[Sequence[float]] = None) -> MolecularSystem:
    types = [int(t) for t in types]
    n = len(types)
    bonds = [(int(a), int(b)) for a, b in bonds]
    orders = [1] * len(bonds) if bond_orders is None else [int(o) for o in bond_orders]
    formal = [0.0] * n if formal is None else [float(f) for f in formal]
    order_of = {}
    for (a, b), o in zip(bonds, orders):
        order_of[(a, b)] = order_of[(b, a)] = o

    linear = [1 if tables.geom_row(t).linear else 0 for t in types]
    if charges is None:
        charges = tables.partial_charges(types, bonds, formal, orders)

    def bt(a, b):
        return tables.bond_type(types[a], types[b], order_of[(a, b)])

    bond_kb, bond_r0 = [], []
    for a, b in bonds:
        kb, r0 = tables.find_bond(types[a], types[b], bt(a, b))
        bond_kb.append(kb)
        bond_r0.append(r0)

    angles, oops, torsions = enumerate_terms(types, linear, bonds)
    ka, th0, kijk, kkji = [], [], [], []
    for (i, j, k) in angles:
        at = bt(i, j) + bt(j, k)           # acyclic findAngleType (MMFF94.java:838-856)
        a_ka, a_th0 = tables.find_angle(types[i], types[j], types[k], at)
        s_ijk, s_kji = tables.find_stretch_bend(types[i], types[j], types[k], at, bt(i, j), bt(j, k))
        ka.append(a_ka)
        th0.append(a_th0)
        kijk.append(s_ijk)
        kkji.append(s_kji)
    # null out-of-plane rows contribute nothing (MMFF94OutOfPlaneComponent.java:124-126): dropped here
    oop_keep, koop = [], []
    for (i, j, k, l) in oops:
        v = tables.find_oop(types[i], types[j], types[k], types[l])
        if v is not None:
            oop_keep.append((i, j, k, l))
            koop.append(v)
    v1, v2, v3 = [], [], []
    for (i, j, k, l) in torsions:
        jk = bt(j, k)
        tt = jk
        if jk == 0 and order_of[(j, k)] == 1 and (bt(i, j) == 1 or bt(k, l) == 1):
            tt = 2                          # MMFF94.java:642-644
        a, b, c = tables.find_torsion(types[i], types[j], types[k], types[l], tt)
        v1.append(a)
        v2.append(b)
        v3.append(c)

    used = sorted(set(types))
    rows = [tables.vdw_row(t) for t in used]

    def col(seq, idx):
        return [t[idx] for t in seq]

    return MolecularSystem(
        xyz=np.asarray(xyz, dtype=np.float64).reshape(-1, 3), atom_type=types, charge=charges, linear=linear,
        vdw_type=used, vdw_alpha=[r.alpha for r in rows], vdw_N=[r.N for r in rows],
        vdw_A=[r.A for r in rows], vdw_G=[r.G for r in rows], vdw_DA=[r.DA for r in rows],
        bond_i=col(bonds, 0), bond_j=col(bonds, 1), bond_kb=bond_kb, bond_r0=bond_r0,
        angle_i=col(angles, 0), angle_j=col(angles, 1), angle_k=col(angles, 2),
        angle_ka=ka, angle_theta0=th0, stbn_kba_ijk=kijk, stbn_kba_kji=kkji,
        oop_i=col(oop_keep, 0), oop_j=col(oop_keep, 1), oop_k=col(oop_keep, 2), oop_l=col(oop_keep, 3),
        oop_koop=koop,
        tor_i=col(torsions, 0), tor_j=col(torsions, 1), tor_k=col(torsions, 2), tor_l=col(torsions, 3),
        tor_v1=v1, tor_v2=v2, tor_v3=v3,
        name=name, expected_energy=expected_energy, source=source)
