#!/usr/bin/env python
"""One all-pairs (no cutoff) energy + gradient evaluation of the 25 k-atom solvated system: the steady-state
regime of ap_kernel (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmolecularenergy_b200 import MMFF94, systems
s = systems.solvated_chains(24999)
ff = MMFF94(False, path="allpairs")
ff.assignParameters(s)
for _ in range(3):
    r = ff.calculateComponents(s, gradient=True)
print("pairs", r["pairs"], "E", r["total"])
