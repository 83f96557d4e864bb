"""The rule-based atom typer (jmolecularenergy_b200/atom_typing.py) against the committed
typings of the validation-suite records it is known to get right: tests/golden/TYPING.json holds, for every rule-typed
fixture, the elements and bonds of the mol2 record, the alternative that was used and the types that reproduced the
published OPTIMOL energy and charges when the fixture was made (tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import load_golden
from jmolecularenergy_b200.atom_typing import ALTERNATIVES, kekule_variants, perceive_types, system_from_structure

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "TYPING.json")) as _fh:
    RECORDS = json.load(_fh)
PAR_DIR = "/root/reference/src/main/resources/org/jme/forcefield/mmff"


def test_dataset_size():
    assert len(RECORDS) >= 640
    assert sum(1 for r in RECORDS.values() if r["alternative"]) >= 5
    assert all(r["alternative"] in ALTERNATIVES for r in RECORDS.values())


@pytest.mark.parametrize("name", sorted(RECORDS))
def test_types_reproduced(name):
    r = RECORDS[name]
    atoms = [(e,) for e in r["elements"]]
    bonds = [(int(a), int(b), o) for a, b, o in r["bonds"]]
    got = perceive_types(atoms, bonds, allow_rings=True, alternative=r["alternative"])
    assert got is not None
    assert got.types == r["types"]
    assert len(got.input_formal) == len(atoms)
    # the fixture made from this typing carries the same types
    assert load_golden(name).atom_type.tolist() == r["types"]


def test_outside_the_rules_is_refused():
    # a five-coordinate phosphorus is not covered: None, never a guess
    atoms = [("P",)] + [("F",)] * 5
    bonds = [(0, k, "1") for k in range(1, 6)]
    assert perceive_types(atoms, bonds) is None
    # disconnected input
    assert perceive_types([("O",), ("H",), ("H",), ("NA",)], [(0, 1, "1"), (0, 2, "1")]) is None


def test_kekule_variants_of_benzene():
    bonds = [(k, (k + 1) % 6, "ar") for k in range(6)]
    v = kekule_variants(6, bonds)
    assert v[0] in ([1, 2, 1, 2, 1, 2], [2, 1, 2, 1, 2, 1])          # a perfect matching first
    assert v[-1] == [1] * 6                                           # all single as the last resort
    assert kekule_variants(2, [(0, 1, "2")]) == [[2]]                 # nothing aromatic: the orders as given


@pytest.mark.skipif(not os.path.isdir(PAR_DIR), reason="needs the reference's .par tables (build container only)")
def test_structure_to_system_matches_fixture():
    """elements + coordinates + bonds -> typed, parameterised system: equal to the committed fixture, parameter by
    parameter (acetate-like ion, an aromatic amine, a plain alkane: the first three rule-typed records by name)."""
    from jmolecularenergy_b200.mmff_tables import MMFFTables
    tables = MMFFTables(PAR_DIR)
    done = 0
    for name in sorted(RECORDS):
        r = RECORDS[name]
        if r["alternative"]:
            continue
        fx = load_golden(name)
        bonds = [(int(a), int(b), o) for a, b, o in r["bonds"]]
        s = system_from_structure(tables, r["elements"], fx.xyz, bonds, name=name)
        assert s.atom_type.tolist() == fx.atom_type.tolist()
        for k in ("charge", "bond_kb", "bond_r0", "angle_ka", "angle_theta0", "tor_v1", "tor_v2", "tor_v3"):
            np.testing.assert_array_equal(getattr(s, k), getattr(fx, k), err_msg=f"{name}.{k}")
        done += 1
        if done == 40:
            break
    assert done == 40
