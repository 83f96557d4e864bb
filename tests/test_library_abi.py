"""No-GPU checks of the drop-in boundary: the shared library loads and exports every symbol
include/jme_b200.h declares; the product fails loudly without a CUDA device; host-side topology helpers
(no device needed) agree with the oracle's restatement of MMFF94.java:158-212,1106-1140."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    g.build()
    from jmolecularenergy_b200 import _lib
    return _lib.lib()


def test_every_declared_symbol_is_exported(L):
    header = open(os.path.join(ROOT, "include", "jme_b200.h")).read()
    declared = set(re.findall(r"\b(jme_[a-z0-9_]+)\s*\(", header))
    declared -= {"jme_handle"}
    from jmolecularenergy_b200 import _lib
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.jme_abi_version() == 1


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "jmolecularenergy_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "libjme_oracle" not in text, f


def test_fails_loudly_without_cuda(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from jmolecularenergy_b200 import MMFF94, JmeError
    s = load_golden("ETHANOL")
    ff = MMFF94(False)
    with pytest.raises(RuntimeError, match="MMFF94 parameters need to be assigned first"):
        ff.calculateEnergy(s)
    with pytest.raises(JmeError, match="no CPU fallback"):
        ff.assignParameters(s)


def test_host_term_enumeration_matches_oracle(L, oracle):
    from jmolecularenergy_b200._abi import ptr
    from jmolecularenergy_b200 import systems
    cases = [load_golden(n) for n in golden_names(pinned_only=False)] + [systems.template("DIGLYME_LIKE")]
    for s in cases:
        na, no, nt = C.c_int64(), C.c_int64(), C.c_int64()
        args = (s.n_atoms, ptr(s.atom_type), ptr(s.linear), len(s.bond_i), ptr(s.bond_i), ptr(s.bond_j))
        assert L.jme_enumerate_terms(*args, None, C.byref(na), None, C.byref(no), None, C.byref(nt)) == 0
        A, O, T = (np.zeros((max(1, na.value), 3), np.int32), np.zeros((max(1, no.value), 4), np.int32),
                   np.zeros((max(1, nt.value), 4), np.int32))
        assert L.jme_enumerate_terms(*args, ptr(A), C.byref(na), ptr(O), C.byref(no), ptr(T), C.byref(nt)) == 0
        rA, rO, rT = oracle.enumerate_terms(s.atom_type, s.linear, s.bond_i, s.bond_j)
        assert np.array_equal(A[:na.value], rA) and np.array_equal(O[:no.value], rO) and np.array_equal(T[:nt.value], rT)


def test_host_exclusions_match_pair_separation(L):
    from jmolecularenergy_b200._abi import ptr
    from jmolecularenergy_b200 import systems
    from jmolecularenergy_b200.system import pair_separation
    for s in (load_golden("CO08A"), systems.template("OCTANOL"), systems.water_box(5)):
        n = s.n_atoms
        cnt = C.c_int64()
        row = np.zeros(n + 1, np.int64)
        assert L.jme_exclusions(n, len(s.bond_i), ptr(s.bond_i), ptr(s.bond_j), ptr(row), None, None, 0, C.byref(cnt)) == 0
        partner, f14 = np.zeros(max(1, cnt.value), np.int32), np.zeros(max(1, cnt.value), np.uint8)
        assert L.jme_exclusions(n, len(s.bond_i), ptr(s.bond_i), ptr(s.bond_j), ptr(row), ptr(partner), ptr(f14),
                                cnt.value, C.byref(cnt)) == 0
        excl, p14 = pair_separation(n, list(zip(s.bond_i.tolist(), s.bond_j.tolist())))
        got_excl = {(a, int(partner[e])) for a in range(n) for e in range(row[a], row[a + 1]) if not f14[e] and partner[e] > a}
        got_14 = {(a, int(partner[e])) for a in range(n) for e in range(row[a], row[a + 1]) if f14[e] and partner[e] > a}
        assert got_excl == excl and got_14 == p14


def test_bad_arguments_are_rejected_without_a_device(L):
    assert L.jme_enumerate_terms(-1, None, None, 0, None, None, None, None, None, None, None, None) != 0
    from jmolecularenergy_b200._abi import JmeSystem
    h = C.c_void_p()
    js = JmeSystem()
    js.n_atoms = 3          # null arrays
    assert L.jme_create(C.byref(js), C.byref(h)) == -1
    assert b"null" in L.jme_last_create_error()
