"""Seeded synthetic systems for BASELINE.json's configs.

Every system is assembled from single-molecule templates under ``jmolecularenergy_b200/data``
(parameters looked up once from the reference's tables by ``tests/golden/make_golden.py``), so
nothing here needs the ``.par`` files.  Geometry: ``numpy.random.Generator(PCG64(seed))``,
coordinates rounded to 1e-4 Angstrom like mol2 records, non-periodic (the reference has no PBC).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .system import MolecularSystem

DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
DEFAULT_SEED = 20261019
DENSITY = 0.1 / 3.0          # molecules per cubic Angstrom (0.1 atoms/A^3 for water)

_BONDED = (("bond", ("bond_i", "bond_j"), ("bond_kb", "bond_r0")),
           ("angle", ("angle_i", "angle_j", "angle_k"), ("angle_ka", "angle_theta0", "stbn_kba_ijk", "stbn_kba_kji")),
           ("oop", ("oop_i", "oop_j", "oop_k", "oop_l"), ("oop_koop",)),
           ("tor", ("tor_i", "tor_j", "tor_k", "tor_l"), ("tor_v1", "tor_v2", "tor_v3")))


def template(name: str) -> MolecularSystem:
    return MolecularSystem.load(os.path.join(DATA_DIR, f"{name}.json"))


def _random_rotations(rng: np.random.Generator, n: int) -> np.ndarray:
    """n uniformly random rotation matrices (from unit quaternions)."""
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((n, 3, 3))
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - z * w); R[:, 0, 2] = 2 * (x * z + y * w)
    R[:, 1, 0] = 2 * (x * y + z * w); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - x * w)
    R[:, 2, 0] = 2 * (x * z - y * w); R[:, 2, 1] = 2 * (y * z + x * w); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def assemble(parts: Sequence[Tuple[MolecularSystem, np.ndarray]], name: str, source: str) -> MolecularSystem:
    """Concatenate copies: ``parts`` = [(template, coords (m, n_template_atoms, 3)), ...]."""
    cols = {k: [] for k in ("xyz", "atom_type", "charge", "linear")}
    for _, idx, par in _BONDED:
        for k in idx + par:
            cols[k] = []
    vdw = {}
    offset = 0
    for t, coords in parts:
        m = coords.shape[0]
        na = t.n_atoms
        if m == 0:
            continue
        cols["xyz"].append(coords.reshape(-1, 3))
        cols["atom_type"].append(np.tile(t.atom_type, m))
        cols["charge"].append(np.tile(t.charge, m))
        cols["linear"].append(np.tile(t.linear, m))
        shift = (offset + na * np.arange(m, dtype=np.int64))[:, None]
        for _, idx, par in _BONDED:
            for k in idx:
                cols[k].append((getattr(t, k)[None, :].astype(np.int64) + shift).reshape(-1))
            for k in par:
                cols[k].append(np.tile(getattr(t, k), m))
        for n, ty in enumerate(t.vdw_type.tolist()):
            vdw[ty] = (t.vdw_alpha[n], t.vdw_N[n], t.vdw_A[n], t.vdw_G[n], t.vdw_DA[n])
        offset += na * m
    cat = {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in cols.items()}
    types = sorted(vdw)
    return MolecularSystem(
        vdw_type=types, vdw_alpha=[vdw[t][0] for t in types], vdw_N=[vdw[t][1] for t in types],
        vdw_A=[vdw[t][2] for t in types], vdw_G=[vdw[t][3] for t in types], vdw_DA=[vdw[t][4] for t in types],
        name=name, source=source, **cat)


def _jittered_lattice(rng, n_side: int, spacing: float, jitter: float) -> np.ndarray:
    g = np.arange(n_side, dtype=np.float64)
    pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3)
    pts = (pts + 0.5) * spacing
    return pts + rng.uniform(-jitter, jitter, size=pts.shape)


def _place(rng, t: MolecularSystem, centres: np.ndarray, atom_jitter: float) -> np.ndarray:
    m = centres.shape[0]
    local = t.xyz - t.xyz.mean(axis=0)
    R = _random_rotations(rng, m)
    coords = np.einsum("mij,aj->mai", R, local) + centres[:, None, :]
    if atom_jitter > 0:
        coords = coords + rng.normal(0.0, atom_jitter, size=coords.shape)
    return np.round(coords, 4)


def water_box(n_waters: int = 1000, seed: int = DEFAULT_SEED) -> MolecularSystem:
    """Config 2 (n_waters=1000 -> 3 000 atoms, 4 495 500 pairs) and config 4 (1667 -> 5 001 atoms):
    rigid-geometry waters with N(0, 0.02 A) per-atom jitter, random orientations, on a jittered cubic
    lattice at 0.0334 molecules/A^3."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_side = int(np.ceil(n_waters ** (1.0 / 3.0) - 1e-9))
    spacing = (1.0 / DENSITY) ** (1.0 / 3.0)
    centres = _jittered_lattice(rng, n_side, spacing, 0.25)
    pick = np.sort(rng.permutation(centres.shape[0])[:n_waters])
    w = template("WATER")
    coords = _place(rng, w, centres[pick], 0.02)
    return assemble([(w, coords)], name=f"water-{3 * n_waters}",
                    source=f"systems.water_box(n_waters={n_waters}, seed={seed})")


def ion_box(n_side: int = 100, seed: int = DEFAULT_SEED, spacing: Optional[float] = None) -> MolecularSystem:
    """Config 5: n_side^3 monatomic ions, Na+ (type 93, q=+1) / Cl- (type 90, q=-1) alternating on a
    jittered cubic lattice (spacing (1/0.1)^(1/3) = 2.154 A, jitter U(-0.3, 0.3) A per axis); no bonds."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if spacing is None:
        spacing = 10.0 ** (1.0 / 3.0)
    g = np.arange(n_side, dtype=np.int64)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    parity = ((ix + iy + iz) % 2).reshape(-1)
    pts = (np.stack([ix, iy, iz], axis=-1).reshape(-1, 3) + 0.5) * spacing
    pts = np.round(pts + rng.uniform(-0.3, 0.3, size=pts.shape), 4)
    na, cl = template("NA"), template("CL")
    # interleave in lattice order so that neighbours in index are neighbours in space
    n = pts.shape[0]
    is_na = parity == 0
    rows = {93: (na.vdw_alpha[0], na.vdw_N[0], na.vdw_A[0], na.vdw_G[0], na.vdw_DA[0]),
            90: (cl.vdw_alpha[0], cl.vdw_N[0], cl.vdw_A[0], cl.vdw_G[0], cl.vdw_DA[0])}
    types = [90, 93]
    return MolecularSystem(
        xyz=pts, atom_type=np.where(is_na, 93, 90), charge=np.where(is_na, float(na.charge[0]), float(cl.charge[0])),
        linear=np.zeros(n, np.uint8),
        vdw_type=types, vdw_alpha=[rows[t][0] for t in types], vdw_N=[rows[t][1] for t in types],
        vdw_A=[rows[t][2] for t in types], vdw_G=[rows[t][3] for t in types], vdw_DA=[rows[t][4] for t in types],
        name=f"ions-{n}", source=f"systems.ion_box(n_side={n_side}, seed={seed})")


def solvated_chains(n_atoms_target: int = 24999, chain_fraction: float = 0.1, seed: int = DEFAULT_SEED,
                    exact_cutoff_pairs: int = 0, cutoff: float = 10.0) -> MolecularSystem:
    """Config 3: hand-typed alkane / ether / alcohol chains (types 1, 5, 6, 21 - table-hit parameters only)
    totalling about ``chain_fraction`` of the atoms, the rest water, in a cube at 0.1 atoms/A^3.
    Chains sit on a coarse lattice; waters fill the fine-lattice sites that keep 2.6 A clear of any
    chain atom.  ``exact_cutoff_pairs`` > 0 additionally moves that many water molecules so that one of
    their oxygens lies at EXACTLY ``cutoff`` (as computed in strict double) from another oxygen, to
    exercise the strict ``>`` of the cutoff test (MMFF94VdwComponent.java:137)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    names = ("DECANE", "DIGLYME_LIKE", "OCTANOL", "HEXANE", "BUTANOL")
    tmpl = [template(n) for n in names]
    side = (n_atoms_target / 0.1) ** (1.0 / 3.0)
    # chains: coarse lattice with 12.5 A spacing (longest template ~ 12.5 A)
    n_chain_atoms_target = int(chain_fraction * n_atoms_target)
    coarse = max(1, int(side // 12.5))
    sites = _jittered_lattice(rng, coarse, side / coarse, 0.5)
    sites = sites[rng.permutation(sites.shape[0])]
    parts: List[Tuple[MolecularSystem, np.ndarray]] = []
    used = 0
    per_kind = [[] for _ in tmpl]
    k = 0
    for sidx in range(sites.shape[0]):
        if used >= n_chain_atoms_target:
            break
        per_kind[k % len(tmpl)].append(sites[sidx])
        used += tmpl[k % len(tmpl)].n_atoms
        k += 1
    chain_xyz = []
    for t, cs in zip(tmpl, per_kind):
        cs = np.array(cs).reshape(-1, 3)
        coords = _place(rng, t, cs, 0.02) if cs.shape[0] else np.zeros((0, t.n_atoms, 3))
        parts.append((t, coords))
        chain_xyz.append(coords.reshape(-1, 3))
    chain_xyz = np.concatenate(chain_xyz) if chain_xyz else np.zeros((0, 3))
    n_chain = chain_xyz.shape[0]
    # waters
    n_waters = (n_atoms_target - n_chain) // 3
    spacing = (1.0 / DENSITY) ** (1.0 / 3.0)
    n_side = int(np.ceil(side / spacing))
    while True:
        centres = _jittered_lattice(rng, n_side, side / n_side, 0.25)
        if n_chain:
            from scipy.spatial import cKDTree
            d, _ = cKDTree(chain_xyz).query(centres, k=1)
            keep = np.where(d > 3.0)[0]
        else:
            keep = np.arange(centres.shape[0])
        if keep.shape[0] >= n_waters:
            break
        n_side += 1
    pick = np.sort(keep[rng.permutation(keep.shape[0])[:n_waters]])
    w = template("WATER")
    wcoords = _place(rng, w, centres[pick], 0.02)
    if exact_cutoff_pairs > 0:
        # move water m so that |O_m - O_partner| == cutoff exactly in strict double arithmetic:
        # put O_m at O_partner + (cutoff, 0, 0) on a 1e-4 grid where the subtraction is exact.
        m_idx = rng.permutation(n_waters)[:2 * exact_cutoff_pairs].reshape(-1, 2)
        for a, b in m_idx:
            # partner oxygen onto a multiple of 0.5 A in x (exactly representable), so that
            # (x + cutoff) - x == cutoff and sqrt(cutoff*cutoff + 0 + 0) == cutoff in strict double
            snap = np.round(wcoords[b, 0, 0] * 2.0) / 2.0
            wcoords[b, :, 0] = np.round(wcoords[b, :, 0] + (snap - wcoords[b, 0, 0]), 4)
            wcoords[b, 0, 0] = snap
            delta = wcoords[b, 0] + np.array([cutoff, 0.0, 0.0]) - wcoords[a, 0]
            wcoords[a] = np.round(wcoords[a] + delta, 4)
            wcoords[a, 0, 1:] = wcoords[b, 0, 1:]
            wcoords[a, 0, 0] = snap + cutoff
    parts.append((w, wcoords))
    return assemble(parts, name=f"solvated-{n_chain + 3 * n_waters}",
                    source=f"systems.solvated_chains(n_atoms_target={n_atoms_target}, chain_fraction={chain_fraction}, "
                           f"seed={seed}, exact_cutoff_pairs={exact_cutoff_pairs}, cutoff={cutoff})")


def ethanol() -> MolecularSystem:
    """Config 1: the 9-atom ethanol fixture (types CR/OR/HC/HOR; 8 bonds, 13 angles, 12 torsions, 15 pairs)."""
    path = os.path.join(DATA_DIR, "ETHANOL.json")
    return MolecularSystem.load(path)
