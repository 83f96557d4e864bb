"""TEST INFRASTRUCTURE ONLY - pure-Python restatement of the reference's energy path.

Nothing under ``jmolecularenergy_b200/`` may import this module; only ``tests/``,
``tests/golden/make_golden.py``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
leg may.  It exists as a second, independently written statement of the same algorithm
that the C++ oracle (``oracle/jme_oracle.cpp``) restates, for small molecules only
(plain Python loops, O(N^2) pair scan).

Parity status: pinned by the reference's own golden energies - see
``tests/test_oracle_golden.py`` (MMFF94.energies OPTIMOL column, 1e-5 kcal/mol).

Each function cites the reference lines it follows (paths relative to
``/root/reference/src/main/java/org/jme/forcefield``).
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np

from jmolecularenergy_b200.system import MolecularSystem, pair_separation

KCAL, KJ, JOULE = 4184.0, 1000.0, 1.0
UNIT_FACTORS = {"KCAL_PER_MOL": KCAL, "KJ_PER_MOL": KJ, "JOULE_PER_MOL": JOULE}


def convert(value: float, unit: str) -> float:
    """``ForceField.EnergyUnit.convertTo`` from kcal/mol (``ForceField.java:153-158``)."""
    if unit == "KCAL_PER_MOL":
        return value
    return (value * KCAL) / UNIT_FACTORS[unit]


def _dist(p, q) -> float:
    # javax.vecmath Point3d.distance: sqrt(dx*dx + dy*dy + dz*dz)
    dx, dy, dz = p[0] - q[0], p[1] - q[1], p[2] - q[2]
    return math.sqrt(dx * dx + dy * dy + dz * dz)


def _clamp(n: float) -> float:
    return 1.0 if n > 1 else (-1.0 if n < -1 else n)


def angle(i, j, k) -> float:
    """``GeometryUtils.calculateAngle`` (``GeometryUtils.java:43-51``), radians."""
    ux, uy, uz = i[0] - j[0], i[1] - j[1], i[2] - j[2]
    vx, vy, vz = k[0] - j[0], k[1] - j[1], k[2] - j[2]
    return math.acos(_clamp((ux * vx + uy * vy + uz * vz) / (_dist(i, j) * _dist(k, j))))


def _cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def _dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def _len(a):
    return math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def _normalize(a):
    n = 1.0 / math.sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])   # vecmath: multiply by 1/len
    return (a[0] * n, a[1] * n, a[2] * n)


def torsion_angle(i, j, k, l) -> float:
    """``GeometryUtils.calculateTorsionAngle`` (``GeometryUtils.java:61-71``), radians in [0, pi]."""
    ij = (i[0] - j[0], i[1] - j[1], i[2] - j[2])
    kj = (k[0] - j[0], k[1] - j[1], k[2] - j[2])
    jk = (-kj[0], -kj[1], -kj[2])
    lk = (l[0] - k[0], l[1] - k[1], l[2] - k[2])
    n1, n2 = _cross(ij, kj), _cross(jk, lk)
    return math.acos(_clamp(_dot(n1, n2) / (_len(n1) * _len(n2))))


def oop_angle(i, j, k, l) -> float:
    """``GeometryUtils.calculateOutOfPlaneAngle`` (``GeometryUtils.java:85-95``); asin unclamped."""
    ij = _normalize((i[0] - j[0], i[1] - j[1], i[2] - j[2]))
    kj = _normalize((k[0] - j[0], k[1] - j[1], k[2] - j[2]))
    perp = _cross(ij, kj)
    lj = _normalize((l[0] - j[0], l[1] - j[1], l[2] - j[2]))
    return math.asin(_dot(perp, lj) / math.sin(angle(i, j, k)))


def vdw_pair_constants(a, b) -> Tuple[float, float]:
    """R*ij and eps_ij of ``MMFF94VdwComponent.java:106-135``; ``a``/``b`` are
    ``(alpha, N, A, G, DA)`` with float32 members."""
    al_i, N_i, A_i, G_i, DA_i = a
    al_j, N_j, A_j, G_j, DA_j = b
    rii = float(A_i) * math.pow(float(al_i), 0.25)
    rjj = float(A_j) * math.pow(float(al_j), 0.25)
    B = 0.0 if (DA_i == 1 or DA_j == 1) else 0.2
    gamma = (rii - rjj) / (rii + rjj)
    rstar = 0.5 * (rii + rjj) * (1 + B * (1 - math.exp(-12.0 * math.pow(gamma, 2))))
    eps = 181.16 * float(G_i) * float(G_j) * float(al_i) * float(al_j)
    # alpha / N is a float/float division rounded to float (MMFF94VdwComponent.java:127)
    eps /= (math.sqrt(float(np.float32(al_i) / np.float32(N_i))) + math.sqrt(float(np.float32(al_j) / np.float32(N_j))))
    eps /= math.pow(rstar, 6)
    if DA_i + DA_j == 3:
        rstar *= 0.8
        eps *= 0.5
    return rstar, eps


def energy_components(s: MolecularSystem, cutoff: float = 0.0, dielectric: float = 1.0,
                      unit: str = "KCAL_PER_MOL") -> Dict[str, float]:
    """Seven component sums in the order of ``MMFF94.java:98`` (bond, angle, stretch-bend,
    out-of-plane, torsion, vdW, electrostatic) and their total (``MMFF94.java:390-392``)."""
    X = [tuple(float(v) for v in row) for row in s.xyz]
    rad = 180 / math.pi
    c_angle = math.radians(math.radians(143.9325))
    r0_of = {}
    e_bond = 0.0
    for a, b, kb, r0 in zip(s.bond_i, s.bond_j, s.bond_kb, s.bond_r0):
        kb, r0 = float(kb), float(r0)
        r0_of[(int(a), int(b))] = r0_of[(int(b), int(a))] = r0
        d = _dist(X[a], X[b]) - r0                                      # BondStretching :103-104
        e = .5 * 143.9325 * kb * d * d * (1 - 2 * d + (7.0 / 12.0) * 4 * d * d)   # :105
        e_bond += convert(e, unit)
    e_angle = e_stbn = 0.0
    for i, j, k, ka, th0, kijk, kkji in zip(s.angle_i, s.angle_j, s.angle_k, s.angle_ka, s.angle_theta0,
                                              s.stbn_kba_ijk, s.stbn_kba_kji):
        ang = angle(X[i], X[j], X[k]) * 180 / math.pi                   # AngleBending :118
        dth = ang - float(th0)
        if s.linear[j]:
            e = 143.9325 * float(ka) * (1 + math.cos(math.radians(ang)))            # :122
        else:
            e = c_angle * float(ka) * .5 * dth * dth * (1 - 0.006981317 * dth)   # :124
        e_angle += convert(e, unit)
        if not s.linear[j]:                                               # StretchBend :125-127
            drij = _dist(X[i], X[j]) - r0_of[(int(i), int(j))]
            drjk = _dist(X[j], X[k]) - r0_of[(int(j), int(k))]
            e = 2.51210 * (float(kijk) * drij + float(kkji) * drjk) * dth          # :134
            e_stbn += convert(e, unit)
    e_oop = 0.0
    for i, j, k, l, koop in zip(s.oop_i, s.oop_j, s.oop_k, s.oop_l, s.oop_koop):
        chi = oop_angle(X[i], X[j], X[k], X[l]) * 180 / math.pi          # OutOfPlane :127
        e_oop += convert(c_angle * .5 * float(koop) * chi * chi, unit)    # :128
    e_tor = 0.0
    for i, j, k, l, v1, v2, v3 in zip(s.tor_i, s.tor_j, s.tor_k, s.tor_l, s.tor_v1, s.tor_v2, s.tor_v3):
        phi = torsion_angle(X[i], X[j], X[k], X[l])                      # Torsion :125
        e = 0.5 * (float(v1) * (1 + math.cos(phi)) + float(v2) * (1 - math.cos(2 * phi))
                   + float(v3) * (1 + math.cos(3 * phi)))                 # :126
        e_tor += convert(e, unit)

    rows = {int(t): (s.vdw_alpha[n], s.vdw_N[n], s.vdw_A[n], s.vdw_G[n], int(s.vdw_DA[n]))
            for n, t in enumerate(s.vdw_type)}
    excluded, is14 = pair_separation(s.n_atoms, list(zip(s.bond_i.tolist(), s.bond_j.tolist())))
    e_vdw = e_ele = 0.0
    n_pairs = 0
    for a in range(s.n_atoms):
        for b in range(a + 1, s.n_atoms):
            if (a, b) in excluded:
                continue
            rstar, eps = vdw_pair_constants(rows[int(s.atom_type[a])], rows[int(s.atom_type[b])])
            r = _dist(X[a], X[b])
            if cutoff > 0 and r > cutoff:                                 # Vdw :137-139, Ele :139-141
                continue
            n_pairs += 1
            e = eps * math.pow((1.07 * rstar) / (r + 0.07 * rstar), 7)   # Vdw :140
            e = e * (((1.12 * math.pow(rstar, 7)) / (math.pow(r, 7) + 0.12 * math.pow(rstar, 7))) - 2)  # :141
            e_vdw += convert(e, unit)
            q = 332.0716 * float(s.charge[a]) * float(s.charge[b]) / (dielectric * (r + 0.05))   # Ele :142
            if (a, b) in is14:
                q *= 0.75                                                 # :143-145
            e_ele += convert(q, unit)
    comps = {"bond": e_bond, "angle": e_angle, "stretch_bend": e_stbn, "oop": e_oop,
             "torsion": e_tor, "vdw": e_vdw, "ele": e_ele}
    total = 0.0
    for key in ("bond", "angle", "stretch_bend", "oop", "torsion", "vdw", "ele"):
        total += comps[key]
    comps["total"] = total
    comps["pairs"] = n_pairs
    return comps
