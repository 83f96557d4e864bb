"""Loader for ``libjme_b200.so`` (the C ABI of ``include/jme_b200.h``).

There is no fallback of any kind: if the shared library is missing or cannot be loaded the import
of anything that computes fails loudly, and on a machine without a CUDA device ``jme_create``
returns ``JME_ERR_CUDA``.
"""
from __future__ import annotations

import ctypes as C
import os

from ._abi import JmeSystem, c_f64p, c_i32p, c_i64p, c_u8p, c_u64p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# JME_B200_LIB lets kernel-tuning experiments load an alternative build of the SAME library
LIB_PATH = os.environ.get("JME_B200_LIB") or os.path.join(PKG_DIR, "libjme_b200.so")

EXPORTS = (
    "jme_abi_version", "jme_last_create_error", "jme_create", "jme_destroy", "jme_last_error",
    "jme_set_options", "jme_set_path", "jme_set_shard", "jme_set_coords", "jme_set_coord1", "jme_get_coords",
    "jme_energy", "jme_energy_gradient", "jme_pair_stats", "jme_pair_list", "jme_set_stream",
    "jme_set_coords_device", "jme_eval_device", "jme_eval_out_size", "jme_launch_count",
    "jme_measure_fp64_peak", "jme_enumerate_terms", "jme_exclusions", "jme_set_profiling", "jme_profile_read",
    "jme_gd_minimize", "jme_adam_minimize", "jme_active_path",
    "jme_peer_create", "jme_peer_chunk", "jme_peer_handle", "jme_peer_open", "jme_peer_step", "jme_peer_last_error",
    "jme_peer_destroy", "jme_energy_delta", "jme_set_incremental",
)

_lib = None


class JmeLibraryMissing(ImportError):
    pass


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise JmeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). jmolecularenergy_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    L.jme_abi_version.restype = C.c_int
    L.jme_last_create_error.restype = C.c_char_p
    L.jme_create.restype = C.c_int
    L.jme_create.argtypes = [C.POINTER(JmeSystem), C.POINTER(H)]
    L.jme_destroy.restype = None
    L.jme_destroy.argtypes = [H]
    L.jme_last_error.restype = C.c_char_p
    L.jme_last_error.argtypes = [H]
    L.jme_set_options.restype = C.c_int
    L.jme_set_options.argtypes = [H, C.c_double, C.c_double, C.c_int]
    L.jme_set_path.restype = C.c_int
    L.jme_set_path.argtypes = [H, C.c_int]
    L.jme_active_path.restype = C.c_int
    L.jme_active_path.argtypes = [H]
    L.jme_set_shard.restype = C.c_int
    L.jme_set_shard.argtypes = [H, C.c_int, C.c_int]
    L.jme_set_coords.restype = C.c_int
    L.jme_set_coords.argtypes = [H, C.c_void_p, C.c_int64]
    L.jme_set_coord1.restype = C.c_int
    L.jme_set_coord1.argtypes = [H, C.c_int64, C.c_int, C.c_double]
    L.jme_get_coords.restype = C.c_int
    L.jme_get_coords.argtypes = [H, c_f64p, C.c_int64]
    L.jme_energy.restype = C.c_int
    L.jme_energy.argtypes = [H, C.POINTER(C.c_double), c_f64p]
    L.jme_energy_gradient.restype = C.c_int
    L.jme_energy_gradient.argtypes = [H, C.POINTER(C.c_double), c_f64p, C.c_void_p]
    L.jme_pair_stats.restype = C.c_int
    L.jme_pair_stats.argtypes = [H, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.jme_pair_list.restype = C.c_int
    L.jme_pair_list.argtypes = [H, c_i32p, c_i32p, c_u8p, C.c_int64, C.POINTER(C.c_int64)]
    L.jme_set_stream.restype = C.c_int
    L.jme_set_stream.argtypes = [H, C.c_void_p]
    L.jme_set_coords_device.restype = C.c_int
    L.jme_set_coords_device.argtypes = [H, C.c_void_p, C.c_int64]
    L.jme_eval_device.restype = C.c_int
    L.jme_eval_device.argtypes = [H, C.c_int, C.c_void_p]
    L.jme_eval_out_size.restype = C.c_int64
    L.jme_eval_out_size.argtypes = [H, C.c_int]
    L.jme_launch_count.restype = C.c_uint64
    L.jme_launch_count.argtypes = [H]
    L.jme_measure_fp64_peak.restype = C.c_int
    L.jme_measure_fp64_peak.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.jme_set_profiling.restype = C.c_int
    L.jme_set_profiling.argtypes = [H, C.c_int]
    L.jme_profile_read.restype = C.c_int
    L.jme_profile_read.argtypes = [H, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int)]
    L.jme_gd_minimize.restype = C.c_int
    L.jme_gd_minimize.argtypes = [H, C.c_int, C.c_double, C.c_double, c_f64p, C.POINTER(C.c_int64), C.POINTER(C.c_int),
                                  c_f64p, C.c_int]
    L.jme_adam_minimize.restype = C.c_int
    L.jme_adam_minimize.argtypes = [H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, c_f64p,
                                    C.POINTER(C.c_int64), C.POINTER(C.c_int), c_f64p, C.c_int]
    L.jme_enumerate_terms.restype = C.c_int
    L.jme_enumerate_terms.argtypes = [C.c_int64, c_i32p, c_u8p, C.c_int64, c_i32p, c_i32p,
                                      c_i32p, c_i64p, c_i32p, c_i64p, c_i32p, c_i64p]
    L.jme_exclusions.restype = C.c_int
    L.jme_exclusions.argtypes = [C.c_int64, C.c_int64, c_i32p, c_i32p, c_i64p, c_i32p, c_u8p, C.c_int64, c_i64p]
    L.jme_energy_delta.restype = C.c_int
    L.jme_energy_delta.argtypes = [H, C.c_int64, C.c_int, C.c_double, c_f64p]
    L.jme_set_incremental.restype = C.c_int
    L.jme_set_incremental.argtypes = [H, C.c_int]
    L.jme_peer_create.restype = C.c_int
    L.jme_peer_create.argtypes = [C.c_int, C.c_int, C.c_int64, C.POINTER(H)]
    L.jme_peer_chunk.restype = C.c_int64
    L.jme_peer_chunk.argtypes = [H]
    L.jme_peer_handle.restype = C.c_int
    L.jme_peer_handle.argtypes = [H, C.c_void_p]
    L.jme_peer_open.restype = C.c_int
    L.jme_peer_open.argtypes = [H, C.c_void_p]
    L.jme_peer_step.restype = C.c_int
    L.jme_peer_step.argtypes = [H, H, C.c_void_p, C.c_void_p, C.c_void_p]
    L.jme_peer_last_error.restype = C.c_char_p
    L.jme_peer_last_error.argtypes = [H]
    L.jme_peer_destroy.restype = None
    L.jme_peer_destroy.argtypes = [H]
    _lib = L
    return L
