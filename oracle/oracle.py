"""TEST INFRASTRUCTURE ONLY - ctypes front end of ``oracle/libjme_oracle.so`` (the C++ CPU
restatement of the reference; see the header of ``oracle/jme_oracle.cpp``).

Allowed importers: ``tests/``, ``__graft_entry__.smoke()``, ``bench.py`` (cpu_baseline and
``--impl reference`` legs).  The product package never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, Optional, Sequence

import numpy as np

from jmolecularenergy_b200._abi import (COMPONENTS, ENERGY_UNITS, JmeSystem, c_f64p, c_i32p, c_i64p, c_u8p,
                                        c_u64p, marshal, ptr)
from jmolecularenergy_b200.system import MolecularSystem

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libjme_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "jme_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE, "-s"] + (["-B"] if force else []), check=True)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_evaluate.restype = C.c_int
        L.oracle_evaluate.argtypes = [C.POINTER(JmeSystem), c_f64p, C.c_double, C.c_double, C.c_int,
                                      c_f64p, c_f64p, c_f64p, c_u64p, c_u64p]
        L.oracle_fd_gradient.restype = C.c_int
        L.oracle_fd_gradient.argtypes = [C.POINTER(JmeSystem), c_f64p, C.c_double, C.c_double, C.c_int,
                                         c_i64p, C.c_int64, C.c_double, c_f64p]
        L.oracle_pair_list.restype = C.c_int64
        L.oracle_pair_list.argtypes = [C.POINTER(JmeSystem), c_f64p, C.c_double, c_i32p, c_i32p, c_u8p, C.c_int64]
        L.oracle_enumerate_terms.restype = C.c_int
        L.oracle_enumerate_terms.argtypes = [C.c_int64, c_i32p, c_u8p, C.c_int64, c_i32p, c_i32p,
                                             c_i32p, c_i64p, c_i32p, c_i64p, c_i32p, c_i64p]
        L.oracle_cutoff_r2_threshold.restype = C.c_double
        L.oracle_cutoff_r2_threshold.argtypes = [C.c_double]
        L.oracle_fast_evaluate.restype = C.c_int
        L.oracle_fast_evaluate.argtypes = [C.POINTER(JmeSystem), c_f64p, C.c_double, C.c_double, C.c_int, C.c_int,
                                           C.c_int, c_f64p, c_f64p, c_f64p, c_u64p]
        L.oracle_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _xyz(s: MolecularSystem, xyz) -> np.ndarray:
    a = np.ascontiguousarray(s.xyz if xyz is None else xyz, dtype=np.float64).reshape(-1, 3)
    if a.shape[0] != s.n_atoms:
        raise ValueError("coordinate count differs from the system")
    return a


def evaluate(s: MolecularSystem, xyz=None, cutoff: float = 0.0, dielectric: float = 1.0,
             unit: str = "KCAL_PER_MOL", gradient: bool = False) -> Dict[str, object]:
    """Reference-faithful energy (and AD gradient): dict with the seven components, ``total``,
    ``pairs``, ``candidates`` and, when asked, ``gradient`` (N,3)."""
    js, keep = marshal(s)
    x = _xyz(s, xyz)
    comps = np.zeros(7)
    total = C.c_double()
    npairs, ncand = C.c_uint64(), C.c_uint64()
    grad = np.zeros((s.n_atoms, 3)) if gradient else None
    rc = lib().oracle_evaluate(C.byref(js), ptr(x), cutoff, dielectric, ENERGY_UNITS[unit], ptr(comps),
                               C.byref(total), ptr(grad) if gradient else None, C.byref(npairs), C.byref(ncand))
    if rc != 0:
        raise RuntimeError(f"oracle_evaluate failed: {rc}")
    out = {k: float(v) for k, v in zip(COMPONENTS, comps)}
    out.update(total=total.value, pairs=npairs.value, candidates=ncand.value)
    if gradient:
        out["gradient"] = grad
    return out


def fd_gradient(s: MolecularSystem, atoms: Sequence[int], xyz=None, cutoff: float = 0.0, dielectric: float = 1.0,
                unit: str = "KCAL_PER_MOL", h: float = 1e-6) -> np.ndarray:
    """``__float128`` central differences of the restated energy for the listed atoms -> (n,3)."""
    js, keep = marshal(s)
    x = _xyz(s, xyz)
    sel = np.ascontiguousarray(atoms, dtype=np.int64)
    out = np.zeros((sel.shape[0], 3))
    rc = lib().oracle_fd_gradient(C.byref(js), ptr(x), cutoff, dielectric, ENERGY_UNITS[unit], ptr(sel),
                                  sel.shape[0], h, ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle_fd_gradient failed: {rc}")
    return out


def pair_list(s: MolecularSystem, xyz=None, cutoff: float = 0.0):
    js, keep = marshal(s)
    x = _xyz(s, xyz)
    n = lib().oracle_pair_list(C.byref(js), ptr(x), cutoff, None, None, None, 0)
    pi, pj, f = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.uint8)
    if n:
        lib().oracle_pair_list(C.byref(js), ptr(x), cutoff, ptr(pi), ptr(pj), ptr(f), n)
    return pi, pj, f


def enumerate_terms(atom_type, linear, bond_i, bond_j):
    t = np.ascontiguousarray(atom_type, np.int32)
    lin = np.ascontiguousarray(linear, np.uint8)
    bi, bj = np.ascontiguousarray(bond_i, np.int32), np.ascontiguousarray(bond_j, np.int32)
    na, no, nt = C.c_int64(), C.c_int64(), C.c_int64()
    args = (t.shape[0], ptr(t), ptr(lin), bi.shape[0], ptr(bi), ptr(bj))
    lib().oracle_enumerate_terms(*args, None, C.byref(na), None, C.byref(no), None, C.byref(nt))
    A, O, T = (np.zeros((na.value, 3), np.int32), np.zeros((no.value, 4), np.int32),
               np.zeros((nt.value, 4), np.int32))
    lib().oracle_enumerate_terms(*args, ptr(A) if A.size else None, C.byref(na), ptr(O) if O.size else None,
                                 C.byref(no), ptr(T) if T.size else None, C.byref(nt))
    return A, O, T


def cutoff_r2_threshold(cutoff: float) -> float:
    return lib().oracle_cutoff_r2_threshold(cutoff)


def num_threads() -> int:
    return lib().oracle_num_threads()


def fast_evaluate(s: MolecularSystem, xyz=None, cutoff: float = 0.0, dielectric: float = 1.0,
                  unit: str = "KCAL_PER_MOL", gradient: bool = True, threads: int = 0,
                  faithful_constants: bool = False) -> Dict[str, object]:
    """Nonbonded-only CPU evaluation used as the timed CPU baseline and for large-system checks:
    OpenMP over atom rows, cell list when ``cutoff > 0``.  ``faithful_constants`` recomputes R*, eps
    per pair like the reference; otherwise a type-pair table is used (same values)."""
    js, keep = marshal(s)
    x = _xyz(s, xyz)
    comps = np.zeros(7)
    total = C.c_double()
    npairs = C.c_uint64()
    grad = np.zeros((s.n_atoms, 3)) if gradient else None
    rc = lib().oracle_fast_evaluate(C.byref(js), ptr(x), cutoff, dielectric, ENERGY_UNITS[unit], threads,
                                    1 if faithful_constants else 0, ptr(comps), C.byref(total),
                                    ptr(grad) if gradient else None, C.byref(npairs))
    if rc != 0:
        raise RuntimeError(f"oracle_fast_evaluate failed: {rc}")
    out = {k: float(v) for k, v in zip(COMPONENTS, comps)}
    out.update(total=total.value, pairs=npairs.value)
    if gradient:
        out["gradient"] = grad
    return out
