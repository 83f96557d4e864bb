#!/usr/bin/env python
"""Kernel-level timing of the nonbonded paths on a few systems (uses the library's profiling events)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from jmolecularenergy_b200 import systems, _lib
from jmolecularenergy_b200.forcefield import _GpuState

def run(name, s, cutoff, path, grad=True, reps=20):
    L = _lib.lib()
    st = _GpuState(s)
    st.check(L.jme_set_options(st.h, cutoff, 1.0, 0))
    st.check(L.jme_set_path(st.h, {"auto": 0, "allpairs": 1, "cells": 2}[path]))
    st.sync_coords(s.xyz)
    tot = C.c_double(); comps = (C.c_double * 7)(); g = np.zeros((s.n_atoms, 3))
    def ev():
        if grad: st.check(L.jme_energy_gradient(st.h, C.byref(tot), comps, g.ctypes.data))
        else: st.check(L.jme_energy(st.h, C.byref(tot), comps))
    for _ in range(3): ev()
    L.jme_set_profiling(st.h, 1)
    t0 = time.perf_counter()
    for _ in range(reps): ev()
    wall = (time.perf_counter() - t0) / reps
    nb = (C.c_float * reps)(); bl = (C.c_float * reps)(); n = C.c_int()
    L.jme_profile_read(st.h, nb, bl, reps, C.byref(n))
    pe = C.c_uint64(); pc = C.c_uint64(); L.jme_pair_stats(st.h, C.byref(pe), C.byref(pc))
    k = float(np.median(nb[:n.value]))
    flop = 49.0 if grad else 30.0
    print(f"{name:28s} atoms {s.n_atoms:8d} pairs {pe.value:12d} {'grad' if grad else 'energy'} kernel {k*1e3:9.1f} us  "
          f"{pe.value/(k*1e-3):.3e} pairs/s  {flop*pe.value/(k*1e-3)/1e12:6.2f} TFLOP/s  call wall {wall*1e6:8.1f} us")
    st.close()

if __name__ == "__main__":
    run("water3k allpairs", systems.water_box(1000), 0.0, "allpairs")
    run("water3k allpairs", systems.water_box(1000), 0.0, "allpairs", grad=False)
    run("water5k allpairs", systems.water_box(1667), 0.0, "allpairs")
    run("water5k allpairs", systems.water_box(1667), 0.0, "allpairs", grad=False)
    run("solvated25k allpairs nocut", systems.solvated_chains(24999), 0.0, "allpairs")
    run("solvated25k allpairs cut10", systems.solvated_chains(24999), 10.0, "allpairs")
    run("solvated25k cells cut10", systems.solvated_chains(24999), 10.0, "cells")
    run("ions125k cells cut10", systems.ion_box(50), 10.0, "cells")
