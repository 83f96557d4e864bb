"""The product's per-term math (csrc/jme_math.cuh) and host marshalling (csrc/jme_host.cpp), compiled
for the CPU by tests/hostcheck, against the oracle: values to 1e-12 relative, analytic gradients against
the oracle's AD gradient to 1e-8 relative.  This is the no-GPU gate for the expressions the kernels run;
the GPU parity tests (-m gpu) then check the kernels themselves through the C ABI."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_names, load_golden
from jmolecularenergy_b200 import systems
from jmolecularenergy_b200._abi import COMPONENTS, JmeSystem, c_f64p, c_u64p, marshal, ptr

ALL = golden_names(pinned_only=False)


@pytest.fixture(scope="module")
def hostcheck():
    d = os.path.join(ROOT, "tests", "hostcheck")
    subprocess.run(["make", "-C", d, "-s"], check=True)
    lib = C.CDLL(os.path.join(d, "libhostcheck.so"))
    lib.hostcheck_evaluate.restype = C.c_int
    lib.hostcheck_evaluate.argtypes = [C.POINTER(JmeSystem), c_f64p, C.c_double, C.c_double, c_f64p, c_f64p, c_u64p]

    def run(s, xyz=None, cutoff=0.0, dielectric=1.0):
        js, keep = marshal(s)
        x = np.ascontiguousarray(s.xyz if xyz is None else xyz, dtype=np.float64)
        comps = np.zeros(7)
        grad = np.zeros((s.n_atoms, 3))
        n = C.c_uint64()
        rc = lib.hostcheck_evaluate(C.byref(js), ptr(x), cutoff, dielectric, ptr(comps), ptr(grad), C.byref(n))
        assert rc == 0
        out = dict(zip(COMPONENTS, comps.tolist()))
        out.update(total=float(sum(comps.tolist())), gradient=grad, pairs=n.value)
        return out
    return run


def check(a, b, gscale=None):
    for k in COMPONENTS:
        assert a[k] == pytest.approx(b[k], rel=1e-12, abs=1e-11), k
    assert a["pairs"] == b["pairs"]
    g = b["gradient"]
    floor = 1e-10 * max(1.0, np.abs(g).max())
    np.testing.assert_allclose(a["gradient"], g, rtol=1e-8, atol=floor)


@pytest.mark.parametrize("name", ALL)
def test_golden_molecules(hostcheck, oracle, name):
    s = load_golden(name)
    check(hostcheck(s), oracle.evaluate(s, gradient=True))


@pytest.mark.parametrize("name", ALL)
def test_golden_molecules_perturbed(hostcheck, oracle, name):
    s = load_golden(name)
    rng = np.random.default_rng(11)
    for _ in range(3):
        xyz = s.xyz + rng.normal(0, 0.08, size=s.xyz.shape)
        check(hostcheck(s, xyz), oracle.evaluate(s, xyz=xyz, gradient=True))


def test_chain_templates(hostcheck, oracle):
    rng = np.random.default_rng(5)
    for name in ("DECANE", "DIGLYME_LIKE", "OCTANOL"):
        s = systems.template(name)
        xyz = s.xyz + rng.normal(0, 0.05, size=s.xyz.shape)
        check(hostcheck(s, xyz), oracle.evaluate(s, xyz=xyz, gradient=True))


def test_linear_centre_form(hostcheck, oracle):
    """Ideally-linear centre (AngleBending :121-122, StretchBend :125-127): reuse formic acid with the
    carbon flagged linear - parameters are irrelevant to the form being exercised."""
    s = load_golden("KHDFRM11")
    s.linear[2] = 1
    rng = np.random.default_rng(3)
    xyz = s.xyz + rng.normal(0, 0.05, size=s.xyz.shape)
    check(hostcheck(s, xyz), oracle.evaluate(s, xyz=xyz, gradient=True))


@pytest.mark.parametrize("cutoff", [0.0, 6.0, 9.0])
def test_water_and_solvated_boxes(hostcheck, oracle, cutoff):
    for s in (systems.water_box(60), systems.solvated_chains(700, chain_fraction=0.35, exact_cutoff_pairs=4, cutoff=6.0)):
        check(hostcheck(s, cutoff=cutoff), oracle.evaluate(s, cutoff=cutoff, gradient=True))


def test_dielectric(hostcheck, oracle):
    s = systems.water_box(30)
    check(hostcheck(s, dielectric=3.5), oracle.evaluate(s, dielectric=3.5, gradient=True))
