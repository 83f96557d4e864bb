#!/usr/bin/env python
"""Pair SLOTS evaluated per useful pair for candidate work decompositions of the cutoff path.

CPU only (numpy + scipy).  A 36^3 jittered ion lattice at 0.1 atoms/A^3 (or uniformly random positions with `rand`),
cutoff 10 A, cells of edge 5 A sorted (z, y, x) like the device sort.  Counts, for the exact pair set:
  * lists of (cluster of C consecutive sorted atoms) x (partner atom later in the order): entries * C / pairs;
  * dense (Ci x Cj) tiles of sorted atoms that hold at least one pair: tiles * Ci * Cj / pairs;
  * the share of the pairs in each row (dy, dz) of the half shell.
"""
import sys
from collections import Counter

import numpy as np
from scipy.spatial import cKDTree

rng = np.random.default_rng(1)
ns, sp, cut = 36, (1 / 0.1) ** (1 / 3), 10.0
g = np.arange(ns)
pts = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
xyz = (pts + 0.5) * sp + rng.uniform(-0.3, 0.3, size=pts.shape)
if len(sys.argv) > 1 and sys.argv[1] == "rand":
    xyz = rng.uniform(0, ns * sp, size=pts.shape)
n = len(xyz)
pairs = cKDTree(xyz).query_pairs(cut, output_type="ndarray")
print(f"n {n}  pairs {len(pairs)}  partners per atom {2 * len(pairs) / n:.1f}")


def sorted_positions(edge):
    c = np.floor(xyz / edge).astype(int)
    nx = c.max(0) + 1
    key = (c[:, 2] * nx[1] + c[:, 1]) * nx[0] + c[:, 0]
    order = np.lexsort((np.arange(n), key))
    pos = np.empty(n, int)
    pos[order] = np.arange(n)
    return pos, c


def cluster_lists(edge, C):
    pos, _ = sorted_positions(edge)
    pi, pj = pos[pairs[:, 0]], pos[pairs[:, 1]]
    lo, hi = np.minimum(pi, pj), np.maximum(pi, pj)
    same = (hi // C) == (lo // C)
    entries = len(np.unique((lo[~same] // C) * n + hi[~same]))
    print(f"cluster lists  edge {edge:4.2f}  C {C}: entries per cluster {entries / (n / C):6.1f}   slots per pair {C * entries / len(pairs):.2f}")


def dense_tiles(edge, ci, cj):
    pos, _ = sorted_positions(edge)
    pi, pj = pos[pairs[:, 0]], pos[pairs[:, 1]]
    lo, hi = np.minimum(pi, pj), np.maximum(pi, pj)
    tiles = len(np.unique((lo // ci) * n + (hi // cj)))
    print(f"dense tiles    edge {edge:4.2f}  {ci:2d} x {cj:2d}: slots per pair {tiles * ci * cj / len(pairs):.2f}")


for C in (1, 2, 4, 8):
    cluster_lists(5.0, C)
for ci, cj in ((32, 32), (8, 8), (8, 4), (4, 4)):
    dense_tiles(5.0, ci, cj)
pos, c = sorted_positions(5.0)
pi, pj = pos[pairs[:, 0]], pos[pairs[:, 1]]
lo = np.where(pi < pj, pairs[:, 0], pairs[:, 1])
hi = np.where(pi < pj, pairs[:, 1], pairs[:, 0])
cnt = Counter(zip((c[hi, 2] - c[lo, 2]).tolist(), (c[hi, 1] - c[lo, 1]).tolist()))
rows = [(0, 0), (0, 1), (0, 2)] + [(1, d) for d in range(-2, 3)] + [(2, d) for d in range(-2, 3)]
print("share of the pairs per half-shell row (dy, dz):", "  ".join(f"({y},{z}) {cnt[(z, y)] / len(pairs):.3f}" for z, y in rows))
