"""ctypes mirror of ``include/jme_b200.h`` (struct layout + marshalling of a MolecularSystem)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .system import MolecularSystem

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)
c_u64p = C.POINTER(C.c_uint64)
c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)

JME_OK = 0
JME_ERR_INVALID = -1
JME_ERR_CUDA = -2
JME_ERR_NOT_ASSIGNED = -3
JME_ERR_NOMEM = -4
JME_ERR_UNSUPPORTED = -5

ENERGY_UNITS = {"KCAL_PER_MOL": 0, "KJ_PER_MOL": 1, "JOULE_PER_MOL": 2}
COMPONENTS = ("bond", "angle", "stretch_bend", "oop", "torsion", "vdw", "ele")
PATHS = {"auto": 0, "allpairs": 1, "cells": 2, "cells-lists": 3, "cells-clusters": 4}


class JmeSystem(C.Structure):
    _fields_ = [
        ("n_atoms", C.c_int64),
        ("atom_type", c_i32p), ("charge", c_f64p), ("linear", c_u8p),
        ("n_vdw_types", C.c_int32),
        ("vdw_type", c_i32p), ("vdw_alpha", c_f32p), ("vdw_N", c_f32p), ("vdw_A", c_f32p), ("vdw_G", c_f32p),
        ("vdw_DA", c_u8p),
        ("n_bonds", C.c_int64),
        ("bond_i", c_i32p), ("bond_j", c_i32p), ("bond_kb", c_f32p), ("bond_r0", c_f32p),
        ("n_angles", C.c_int64),
        ("angle_i", c_i32p), ("angle_j", c_i32p), ("angle_k", c_i32p),
        ("angle_ka", c_f32p), ("angle_theta0", c_f32p), ("stbn_kba_ijk", c_f32p), ("stbn_kba_kji", c_f32p),
        ("n_oops", C.c_int64),
        ("oop_i", c_i32p), ("oop_j", c_i32p), ("oop_k", c_i32p), ("oop_l", c_i32p), ("oop_koop", c_f32p),
        ("n_torsions", C.c_int64),
        ("tor_i", c_i32p), ("tor_j", c_i32p), ("tor_k", c_i32p), ("tor_l", c_i32p),
        ("tor_v1", c_f32p), ("tor_v2", c_f32p), ("tor_v3", c_f32p),
    ]


_PTR = {np.dtype(np.int32): c_i32p, np.dtype(np.float32): c_f32p, np.dtype(np.float64): c_f64p,
        np.dtype(np.uint8): c_u8p, np.dtype(np.int64): c_i64p, np.dtype(np.uint64): c_u64p}


def ptr(a: np.ndarray):
    """ctypes pointer to a C-contiguous numpy array (the array must outlive the call)."""
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("array must be C-contiguous")
    return a.ctypes.data_as(_PTR[a.dtype])


def marshal(s: MolecularSystem):
    """Return ``(JmeSystem, keepalive)``; ``keepalive`` owns the arrays the struct points into."""
    keep = []

    def p(name):
        a = np.ascontiguousarray(getattr(s, name))
        keep.append(a)
        return ptr(a)

    js = JmeSystem()
    js.n_atoms = s.n_atoms
    js.atom_type, js.charge, js.linear = p("atom_type"), p("charge"), p("linear")
    js.n_vdw_types = int(s.vdw_type.shape[0])
    js.vdw_type, js.vdw_alpha, js.vdw_N = p("vdw_type"), p("vdw_alpha"), p("vdw_N")
    js.vdw_A, js.vdw_G, js.vdw_DA = p("vdw_A"), p("vdw_G"), p("vdw_DA")
    js.n_bonds = int(s.bond_i.shape[0])
    js.bond_i, js.bond_j, js.bond_kb, js.bond_r0 = p("bond_i"), p("bond_j"), p("bond_kb"), p("bond_r0")
    js.n_angles = int(s.angle_i.shape[0])
    js.angle_i, js.angle_j, js.angle_k = p("angle_i"), p("angle_j"), p("angle_k")
    js.angle_ka, js.angle_theta0 = p("angle_ka"), p("angle_theta0")
    js.stbn_kba_ijk, js.stbn_kba_kji = p("stbn_kba_ijk"), p("stbn_kba_kji")
    js.n_oops = int(s.oop_i.shape[0])
    js.oop_i, js.oop_j, js.oop_k, js.oop_l, js.oop_koop = p("oop_i"), p("oop_j"), p("oop_k"), p("oop_l"), p("oop_koop")
    js.n_torsions = int(s.tor_i.shape[0])
    js.tor_i, js.tor_j, js.tor_k, js.tor_l = p("tor_i"), p("tor_j"), p("tor_k"), p("tor_l")
    js.tor_v1, js.tor_v2, js.tor_v3 = p("tor_v1"), p("tor_v2"), p("tor_v3")
    return js, keep
