"""oracle_fast_evaluate (OpenMP, type-pair table, cell list) against the reference-faithful evaluator."""
import numpy as np
import pytest

from jmolecularenergy_b200 import systems


@pytest.mark.parametrize("cutoff", [0.0, 8.0])
def test_fast_matches_faithful_on_water(oracle, cutoff):
    s = systems.water_box(120)
    a = oracle.evaluate(s, cutoff=cutoff, gradient=True)
    for faithful in (False, True):
        b = oracle.fast_evaluate(s, cutoff=cutoff, faithful_constants=faithful)
        assert b["pairs"] == a["pairs"]
        for k in ("bond", "angle", "stretch_bend", "vdw", "ele", "total"):
            assert b[k] == pytest.approx(a[k], rel=1e-12, abs=1e-10), k
        np.testing.assert_allclose(b["gradient"], a["gradient"], rtol=1e-9, atol=1e-9)


def test_fast_matches_faithful_on_solvated_chains_with_cutoff(oracle):
    s = systems.solvated_chains(1500, chain_fraction=0.3, exact_cutoff_pairs=5, cutoff=7.0)
    a = oracle.evaluate(s, cutoff=7.0, gradient=True)
    b = oracle.fast_evaluate(s, cutoff=7.0)
    assert b["pairs"] == a["pairs"]
    assert b["total"] == pytest.approx(a["total"], rel=1e-12)
    scale = np.abs(a["gradient"]).max()
    np.testing.assert_allclose(b["gradient"], a["gradient"], rtol=1e-9, atol=1e-12 * scale)


def test_exact_cutoff_pairs_are_kept(oracle):
    """systems.solvated_chains(exact_cutoff_pairs=k) plants k O...O pairs at r == cutoff exactly; the
    strict '>' (MMFF94VdwComponent.java:137) keeps them, one ulp less drops them."""
    s = systems.solvated_chains(900, chain_fraction=0.0, exact_cutoff_pairs=8, cutoff=7.0)
    i, j, _ = oracle.pair_list(s, cutoff=7.0)
    i2, j2, _ = oracle.pair_list(s, cutoff=float(np.nextafter(7.0, 0)))
    assert len(i) - len(i2) >= 8


def test_ion_box_threads_do_not_change_result(oracle):
    s = systems.ion_box(12)
    a = oracle.fast_evaluate(s, cutoff=6.0, threads=1)
    b = oracle.fast_evaluate(s, cutoff=6.0, threads=4)
    assert a["pairs"] == b["pairs"]
    assert a["total"] == pytest.approx(b["total"], rel=1e-13)
