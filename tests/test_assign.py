"""Host-side parameter assignment port (jmolecularenergy_b200/assign.py): ring perception,
the table lookups and the empirical rules.  The `.par` tables are the reference's (not shipped): the checks that need
them run where /root/reference exists (the build container) and are skipped elsewhere.  The real proof of this module
is tests/golden: 317 molecules of the reference's two validation suites, typed by rule, parameterised by this code,
reproduce their published energies and charges (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import load_golden
from jmolecularenergy_b200.assign import Assigner, assign_parameters, find_rings

PAR_DIR = "/root/reference/src/main/resources/org/jme/forcefield/mmff"
needs_tables = pytest.mark.skipif(not os.path.isdir(PAR_DIR), reason="the reference's .par tables are not on this machine")


def test_find_rings_counts():
    hexagon = [(k, (k + 1) % 6) for k in range(6)]
    assert [len(r) for r in find_rings(6, hexagon)] == [6]
    # naphthalene skeleton: two fused six-rings sharing the bond 0-5, and the ten-membered envelope
    naph = hexagon + [(0, 6), (6, 7), (7, 8), (8, 9), (9, 5)]
    assert sorted(len(r) for r in find_rings(10, naph, max_size=10)) == [6, 6, 10]
    # bicyclo[1.1.0]butane: two three-rings and the four-ring around them
    assert sorted(len(r) for r in find_rings(4, [(0, 1), (1, 2), (2, 0), (1, 3), (3, 2)])) == [3, 3, 4]
    assert find_rings(5, [(0, 1), (1, 2), (2, 3), (3, 4)]) == []


@needs_tables
def test_ethanol_reassigned_equals_the_fixture():
    from jmolecularenergy_b200.mmff_tables import MMFFTables
    s = load_golden("ETHANOL")
    bonds = list(zip(s.bond_i.tolist(), s.bond_j.tolist()))
    t = assign_parameters(MMFFTables(PAR_DIR), s.xyz, s.atom_type.tolist(), bonds, [1] * len(bonds))
    for k in ("charge", "bond_kb", "bond_r0", "angle_ka", "angle_theta0", "stbn_kba_ijk", "stbn_kba_kji", "tor_v1", "tor_v2", "tor_v3"):
        np.testing.assert_array_equal(getattr(t, k), getattr(s, k), err_msg=k)


@needs_tables
def test_ring_angle_and_torsion_types():
    from jmolecularenergy_b200.mmff_tables import MMFFTables
    tables = MMFFTables(PAR_DIR)
    # cyclobutane carbons (CR4R = 20) with two hydrogens each: ring angles are type 4, ring torsions type 4
    bonds = [(0, 1), (1, 2), (2, 3), (3, 0)] + [(c, 4 + 2 * c + h) for c in range(4) for h in range(2)]
    a = Assigner(tables, [20] * 4 + [5] * 8, bonds, [1] * len(bonds))
    assert a.angle_type(0, 1, 2) == 4 and a.angle_type(4, 0, 1) == 0
    for x, y in bonds:
        a.r0[(x, y)] = a.r0[(y, x)] = a.bond_parameters(x, y)[1]
    ka, th0 = a.angle_parameters(0, 1, 2)
    assert 80.0 < float(th0) < 95.0 and float(ka) > 0
    v = a.torsion_parameters(0, 1, 2, 3)
    assert len(v) == 3
    # cyclopropane (CR3R = 22): angle type 3
    bonds3 = [(0, 1), (1, 2), (2, 0)] + [(c, 3 + 2 * c + h) for c in range(3) for h in range(2)]
    b = Assigner(tables, [22] * 3 + [5] * 6, bonds3, [1] * len(bonds3))
    assert b.angle_type(0, 1, 2) == 3


@needs_tables
def test_formal_charge_sharing():
    """``calculateFormalCharge`` (MMFF94.java:404-564): shared negative charge of a carboxylate, the input's +1 averaged
    over the nitrogens of a guanidinium, fixed values of the ionic types, nothing for a nitro group."""
    from jmolecularenergy_b200.mmff_tables import MMFFTables
    tables = MMFFTables(PAR_DIR)
    # acetate: CH3 (1, 5 x 3) - C (41) - O (32) x 2
    bonds = [(0, 1), (1, 2), (1, 3), (0, 4), (0, 5), (0, 6)]
    a = Assigner(tables, [1, 41, 32, 32, 5, 5, 5], bonds, [1, 2, 1, 1, 1, 1])
    assert a.formal_charges() == [0.0, 0.0, -0.5, -0.5, 0.0, 0.0, 0.0]
    q = a.partial_charges(a.formal_charges())
    assert abs(float(np.sum(q)) + 1.0) < 1e-6                     # the partial charges carry the net charge
    # guanidinium C(NH2)3+: C 57, N 56 x 3, H 36 x 6; the input has +1 on the nitrogen drawn with the double bond
    gb = [(0, 1), (0, 2), (0, 3)] + [(1 + k, 4 + 2 * k + h) for k in range(3) for h in range(2)]
    g = Assigner(tables, [57, 56, 56, 56] + [36] * 6, gb, [2, 1, 1] + [1] * 6)
    f = g.formal_charges([0, 1, 0, 0] + [0] * 6)
    assert f[0] == 0.0 and f[1] == f[2] == f[3] == pytest.approx(1.0 / 3.0)
    assert abs(float(np.sum(g.partial_charges(f))) - 1.0) < 1e-6
    # tetramethylammonium nitrogen (34) is +1 whatever the input says; nitromethane's oxygens (32 on 45) stay neutral
    nb = [(0, 1), (0, 2), (0, 3), (0, 4)] + [(1 + c, 5 + 3 * c + h) for c in range(4) for h in range(3)]
    n4 = Assigner(tables, [34, 1, 1, 1, 1] + [5] * 12, nb, [1] * len(nb))
    assert n4.formal_charges()[0] == 1.0 and sum(n4.formal_charges()) == 1.0
    nitro = Assigner(tables, [1, 45, 32, 32, 5, 5, 5], [(0, 1), (1, 2), (1, 3), (0, 4), (0, 5), (0, 6)], [1, 2, 1, 1, 1, 1])
    assert nitro.formal_charges() == [0.0] * 7


def test_pinned_charged_molecules_carry_their_net_charge():
    """The fixtures typed by rule include cations and anions (ammonium, carboxylate, guanidinium, pyridinium, N-oxides
    ...): their bci charges - already checked against the suite's published charges when the fixture was made - must
    add up to an integer, and more than a hundred of the pinned molecules are ions."""
    from conftest import golden_names
    ions = 0
    for name in golden_names(pinned_only=True):
        s = load_golden(name)
        total = float(np.sum(s.charge))
        assert abs(total - round(total)) < 2e-3, name
        ions += int(round(total) != 0)
    assert ions >= 100
