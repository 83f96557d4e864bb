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

if __name__ == "__main__" and len(sys.argv) == 1:
    run("water3k allpairs", systems.water_box(1000), 0.0, "allpairs")
    run("water3k allpairs", systems.water_box(1000), 0.0, "allpairs", grad=False)
    run("water5k allpairs", systems.water_box(1667), 0.0, "allpairs")
    run("water5k allpairs", systems.water_box(1667), 0.0, "allpairs", grad=False)
    run("solvated25k allpairs nocut", systems.solvated_chains(24999), 0.0, "allpairs")
    run("solvated25k allpairs cut10", systems.solvated_chains(24999), 10.0, "allpairs")
    run("solvated25k cells cut10", systems.solvated_chains(24999), 10.0, "cells")
    run("ions125k cells cut10", systems.ion_box(50), 10.0, "cells")


def opt5k(n_waters=1667, max_steps=0):
    """BASELINE config 4: the reference's GradientDescentMinimizer loop (natively restated, jme_gd_minimize) on a
    5 001-atom water box: energy calls per second (each call = one coordinate update + one full energy)."""
    from jmolecularenergy_b200 import MMFF94, NativeGradientDescentMinimizer
    s = systems.water_box(n_waters)
    ff = MMFF94(False)
    ff.assignParameters(s)
    m = NativeGradientDescentMinimizer(ff)
    t0 = time.perf_counter()
    m.minimize(s, max_steps, 0.01, 1e-4)
    dt = time.perf_counter() - t0
    pairs = ff.calculateComponents(s)["pairs"]
    print(f"opt5k: atoms {s.n_atoms} steps {m.step} energy calls {m.energyCalls} in {dt:.3f} s -> {m.energyCalls/dt:.0f} calls/s, "
          f"{m.energyCalls*pairs/dt:.3e} pair-energies/s; E {m.stepEnergies[0]:.3f} -> {m.stepEnergies[-1]:.3f}")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "opt":
    opt5k(max_steps=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
