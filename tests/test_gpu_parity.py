"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> libjme_b200.so), against
the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): total energy <= 1e-10 relative, per-atom gradient component
<= 1e-8 relative with an absolute floor of 1e-10 x the largest component; pair selection bit-exact."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_names, golden_tolerance, load_golden

pytestmark = pytest.mark.gpu

PINNED = golden_names()
ALL = golden_names(pinned_only=False)
E_RTOL = 1e-10
G_RTOL = 1e-8
# the two evaluation schemes of the cutoff path (per-atom lists / cluster lists with Newton's third law); "cells" picks by size
CELL_PATHS = ("cells-lists", "cells-clusters")


def gpu_ff(path="auto", cutoff=0.0, dielectric=1.0, unit=None):
    from jmolecularenergy_b200 import MMFF94, EnergyUnit
    ff = MMFF94(False, path=path)
    ff.setCutoffDistance(cutoff)
    ff.setDielectricConstant(dielectric)
    if unit is not None:
        ff.setEnergyUnit(EnergyUnit[unit])
    return ff


def assert_parity(got, ref, label=""):
    for k in ("bond", "angle", "stretch_bend", "oop", "torsion", "vdw", "ele"):
        # every component against ITS OWN magnitude (scaling by the total checked a small component next to
        # a large total loosely); the floor only guards components that are exactly or nearly zero
        scale = max(abs(ref[k]), 1e-3)
        assert abs(got[k] - ref[k]) <= E_RTOL * scale, (label, k, got[k], ref[k])
    assert abs(got["total"] - ref["total"]) <= E_RTOL * max(abs(ref["total"]), 1e-3), (label, got["total"], ref["total"])
    assert got["pairs"] == ref["pairs"], (label, got["pairs"], ref["pairs"])
    if "gradient" in ref:
        g = ref["gradient"]
        floor = 1e-10 * max(1.0, float(np.abs(g).max()))
        np.testing.assert_allclose(got["gradient"], g, rtol=G_RTOL, atol=floor, err_msg=label)


@pytest.mark.parametrize("name", PINNED)
def test_golden_energy_matches_published_value(name):
    """The reference's own criterion (MMFF94Test.java:79-81) through the GPU path, at 1e-5 (1e-4 for the fixtures
    MANIFEST.json lists) not 1e-2."""
    s = load_golden(name)
    ff = gpu_ff()
    ff.assignParameters(s)
    assert abs(ff.calculateEnergy(s) - s.expected_energy) <= golden_tolerance(name)


@pytest.mark.parametrize("name", ALL)
def test_golden_molecules_against_oracle(oracle, name):
    s = load_golden(name)
    ff = gpu_ff()
    ff.assignParameters(s)
    rng = np.random.default_rng(11)
    for trial in range(3):
        if trial:
            s.xyz[:] = s.xyz + rng.normal(0, 0.08, size=s.xyz.shape)
        assert_parity(ff.calculateComponents(s, gradient=True), oracle.evaluate(s, gradient=True), f"{name}#{trial}")


def test_linear_centre_form(oracle):
    s = load_golden("KHDFRM11")
    s.linear[2] = 1
    s.xyz[:] = s.xyz + np.random.default_rng(3).normal(0, 0.05, size=s.xyz.shape)
    ff = gpu_ff()
    ff.assignParameters(s)
    assert_parity(ff.calculateComponents(s, gradient=True), oracle.evaluate(s, gradient=True))


def test_chain_templates(oracle):
    from jmolecularenergy_b200 import systems
    rng = np.random.default_rng(5)
    for name in ("DECANE", "DIGLYME_LIKE", "OCTANOL"):
        s = systems.template(name)
        s.xyz[:] = s.xyz + rng.normal(0, 0.05, size=s.xyz.shape)
        ff = gpu_ff()
        ff.assignParameters(s)
        assert_parity(ff.calculateComponents(s, gradient=True), oracle.evaluate(s, gradient=True), name)


@pytest.mark.parametrize("n_waters", [1, 10, 11, 100, 341])
def test_water_boxes_all_pairs(oracle, n_waters):
    """Block-count edge cases of the tile decomposition: 3, 30, 33 (one atom over a block), 300, 1023 atoms."""
    from jmolecularenergy_b200 import systems
    s = systems.water_box(n_waters)
    ff = gpu_ff(path="allpairs")
    ff.assignParameters(s)
    assert_parity(ff.calculateComponents(s, gradient=True), oracle.evaluate(s, gradient=True), s.name)
    e_only = ff.calculateEnergy(s)
    assert e_only == pytest.approx(ff.calculateComponents(s, gradient=True)["total"], rel=1e-14)


def test_config2_water_3k(oracle):
    """BASELINE config 2: 1 000 waters, 4 495 500 pairs, full bonded + nonbonded energy/gradient."""
    from jmolecularenergy_b200 import systems
    s = systems.water_box(1000)
    ff = gpu_ff()
    ff.assignParameters(s)
    got = ff.calculateComponents(s, gradient=True)
    ref = oracle.evaluate(s, gradient=True)
    assert ref["pairs"] == 4495500
    assert_parity(got, ref, s.name)


@pytest.mark.parametrize("path", ["allpairs", *CELL_PATHS])
@pytest.mark.parametrize("cutoff", [6.0, 9.5])
def test_cutoff_paths_small(oracle, path, cutoff):
    from jmolecularenergy_b200 import systems
    s = systems.solvated_chains(2400, chain_fraction=0.3, exact_cutoff_pairs=12, cutoff=cutoff)
    ff = gpu_ff(path=path, cutoff=cutoff)
    ff.assignParameters(s)
    ref = oracle.evaluate(s, cutoff=cutoff, gradient=True)
    assert_parity(ff.calculateComponents(s, gradient=True), ref, f"{path}@{cutoff}")
    # bit-exact pair selection, including the planted r == cutoff pairs
    pi, pj, f14 = ff.pairList(s)
    oi, oj, of = oracle.pair_list(s, cutoff=cutoff)
    assert len(pi) == len(oi) == ref["pairs"]
    assert np.array_equal(pi, oi) and np.array_equal(pj, oj) and np.array_equal(f14, of)
    # one ulp below the cutoff the planted pairs drop out on both sides
    c2 = float(np.nextafter(cutoff, 0))
    ff.setCutoffDistance(c2)
    n2 = len(ff.pairList(s)[0])
    assert n2 == len(oracle.pair_list(s, cutoff=c2)[0])
    assert len(pi) - n2 >= 12


def test_pair_list_without_cutoff(oracle):
    s = load_golden("ETHANOL")
    ff = gpu_ff()
    ff.assignParameters(s)
    pi, pj, f14 = ff.pairList(s)
    oi, oj, of = oracle.pair_list(s)
    assert np.array_equal(pi, oi) and np.array_equal(pj, oj) and np.array_equal(f14, of)
    assert len(pi) == 15 and int(f14.sum()) == 12


@pytest.mark.parametrize("cutoff", [10.0, 12.0])
def test_config3_solvated_25k(oracle, cutoff):
    """BASELINE config 3: ~25 k atoms, cutoff 10 / 12 A, torsions + 1-4 scaling + exclusions + cell list,
    >= 100 pairs planted at exactly r == cutoff.  Checker: the multi-threaded oracle (itself checked against
    the reference-faithful evaluator in tests/test_oracle_fast.py)."""
    from jmolecularenergy_b200 import systems
    s = systems.solvated_chains(24999, exact_cutoff_pairs=100, cutoff=cutoff)
    ref = oracle.fast_evaluate(s, cutoff=cutoff)
    for path in ("auto", *CELL_PATHS):
        ff = gpu_ff(path=path, cutoff=cutoff)
        ff.assignParameters(s)
        got = ff.calculateComponents(s, gradient=True)
        assert_parity(got, ref, f"{s.name} {path}")
    # the all-pairs tiles with the cutoff predicate must select the same pairs and agree to rounding
    ff2 = gpu_ff(path="allpairs", cutoff=cutoff)
    ff2.assignParameters(s)
    got2 = ff2.calculateComponents(s, gradient=True)
    assert got2["pairs"] == got["pairs"]
    assert got2["total"] == pytest.approx(got["total"], rel=1e-12)
    np.testing.assert_allclose(got2["gradient"], got["gradient"], rtol=1e-9, atol=1e-10 * np.abs(got["gradient"]).max())


def test_config5_ion_box_125k(oracle):
    from jmolecularenergy_b200 import systems
    s = systems.ion_box(50)
    ref = oracle.fast_evaluate(s, cutoff=10.0)
    for path in CELL_PATHS:
        ff = gpu_ff(path=path, cutoff=10.0)
        ff.assignParameters(s)
        assert_parity(ff.calculateComponents(s, gradient=True), ref, f"{s.name} {path}")


def test_config5_ion_box_1m_properties(oracle):
    """BASELINE config 5 at full size (10^6 ions, cutoff 10 A): pair count and energy against the
    multi-threaded oracle, plus size-independent properties - zero net force, run-to-run bit identity,
    a 2-way shard sum equal to the unsharded result."""
    from jmolecularenergy_b200 import systems
    s = systems.ion_box(100)
    ref = oracle.fast_evaluate(s, cutoff=10.0, gradient=True)
    for path in ("auto", "cells-lists"):             # auto = the cluster lists at this size
        ff = gpu_ff(path=path, cutoff=10.0)
        ff.assignParameters(s)
        a = ff.calculateComponents(s, gradient=True)
        b = ff.calculateComponents(s, gradient=True)
        assert a["total"] == b["total"] and np.array_equal(a["gradient"], b["gradient"]), path
        net = np.abs(a["gradient"].sum(axis=0)).max()
        assert net <= 1e-9 * np.abs(a["gradient"]).sum()
        assert_parity(a, ref, f"{s.name} {path}")
        ff.release(s)


def test_units_and_dielectric(oracle):
    from jmolecularenergy_b200 import systems
    s = systems.water_box(40)
    for unit in ("KJ_PER_MOL", "JOULE_PER_MOL"):
        ff = gpu_ff(unit=unit, dielectric=2.5)
        ff.assignParameters(s)
        assert_parity(ff.calculateComponents(s, gradient=True),
                      oracle.evaluate(s, unit=unit, dielectric=2.5, gradient=True), unit)


def test_single_coordinate_update_equals_full_upload(oracle):
    from jmolecularenergy_b200 import systems
    s = systems.water_box(50)
    ff = gpu_ff()
    ff.assignParameters(s)
    ff.calculateEnergy(s)
    launches0 = ff.launchCount(s)
    s.xyz[17, 1] += 0.013
    e1 = ff.calculateEnergy(s)
    # one set_coord1 + nonbonded + bonded + reduce
    assert ff.launchCount(s) - launches0 == 4
    t = systems.water_box(50)
    t.xyz[:] = s.xyz
    ff2 = gpu_ff()
    ff2.assignParameters(t)
    assert e1 == ff2.calculateEnergy(t)
    assert e1 == pytest.approx(oracle.evaluate(s)["total"], rel=E_RTOL)


def test_removed_component_is_not_summed(oracle):
    s = load_golden("CO08A")
    ff = gpu_ff()
    ff.assignParameters(s)
    full = ff.calculateComponents(s)
    ff.removeEnergyComponent(ff.electrostaticComponent)
    assert ff.calculateEnergy(s) == pytest.approx(full["total"] - full["ele"], rel=1e-12)
    assert ff.vdwComponent.calculateEnergy(s) == pytest.approx(full["vdw"], rel=1e-14)


def test_errors():
    from jmolecularenergy_b200 import MMFF94, JmeError
    from jmolecularenergy_b200 import systems
    s = systems.water_box(3)
    ff = MMFF94(False)
    with pytest.raises(RuntimeError, match="MMFF94 parameters need to be assigned first"):
        ff.calculateEnergy(s)
    bad = systems.water_box(3)
    bad.atom_type[0] = 12          # no vdW row for chlorine in a water system
    with pytest.raises((JmeError, ValueError)):
        ff.assignParameters(bad)
    ff.assignParameters(s)
    ff.setDielectricConstant(0.0)
    with pytest.raises(JmeError):
        ff.calculateEnergy(s)


def test_two_way_shard_sums_to_unsharded(oracle):
    """Atom-row sharding emulated on one GPU: two handles with rank 0/1 of 2; the
    partial energies / gradients / pair counts must add up to the unsharded evaluation."""
    from jmolecularenergy_b200 import systems, _lib
    from jmolecularenergy_b200.forcefield import _GpuState
    import torch
    for s, cutoff, path in ((systems.water_box(200), 0.0, 1), (systems.solvated_chains(3000, chain_fraction=0.3), 8.0, 3),
                            (systems.solvated_chains(3000, chain_fraction=0.3), 8.0, 4), (systems.ion_box(24), 9.0, 4)):
        L = _lib.lib()
        full = None
        parts = []
        for rank, world in ((0, 1), (0, 2), (1, 2)):
            st = _GpuState(s)
            st.check(L.jme_set_options(st.h, cutoff, 1.0, 0))
            st.check(L.jme_set_path(st.h, path))
            st.check(L.jme_set_shard(st.h, rank, world))
            st.sync_coords(s.xyz)
            n_out = L.jme_eval_out_size(st.h, 1)
            out = torch.zeros(n_out, dtype=torch.float64, device="cuda")
            st.check(L.jme_set_stream(st.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            st.check(L.jme_eval_device(st.h, 1, C.c_void_p(out.data_ptr())))
            torch.cuda.synchronize()
            if world == 1:
                full = out.cpu().numpy()
            else:
                parts.append(out.cpu().numpy())
            st.close()
        summed = parts[0] + parts[1]
        assert summed[8] == full[8]
        np.testing.assert_allclose(summed[:8], full[:8], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(summed[10:], full[10:], rtol=1e-9, atol=1e-9 * np.abs(full[10:]).max())


def test_peer_exchange_single_rank_equals_plain_evaluation(oracle):
    """The owner-mode exchange over peer memory (include/jme_b200.h: jme_peer_*) with a world of one: push to and pull
    from the own exported block.  Header and gradient must equal a plain device evaluation bit for bit, step after step
    (the flags count steps); the multi-rank case is tests/test_multi_gpu.py."""
    from jmolecularenergy_b200 import systems, _lib
    from jmolecularenergy_b200.forcefield import _GpuState
    import torch
    L = _lib.lib()
    for s, cutoff, path in ((systems.water_box(100), 0.0, 1), (systems.solvated_chains(3001, chain_fraction=0.3), 8.0, 4)):
        n = s.n_atoms
        st = _GpuState(s)
        st.check(L.jme_set_options(st.h, cutoff, 1.0, 0))
        st.check(L.jme_set_path(st.h, path))
        st.check(L.jme_set_stream(st.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        st.sync_coords(s.xyz)
        total = C.c_double()
        st.check(L.jme_energy(st.h, C.byref(total), None))             # sizes the lists of the cutoff path
        plain = torch.zeros(int(L.jme_eval_out_size(st.h, 1)), dtype=torch.float64, device="cuda")
        xyz = torch.from_numpy(s.xyz).cuda()
        st.check(L.jme_set_coords_device(st.h, C.c_void_p(xyz.data_ptr()), n))
        st.check(L.jme_eval_device(st.h, 1, C.c_void_p(plain.data_ptr())))
        comm = C.c_void_p()
        assert L.jme_peer_create(0, 1, n, C.byref(comm)) == 0
        chunk = int(L.jme_peer_chunk(comm))
        assert chunk >= n and chunk % 2 == 0
        handle = (C.c_ubyte * 64)()
        assert L.jme_peer_handle(comm, handle) == 0
        assert L.jme_peer_open(comm, handle) == 0                      # a world of one maps nothing
        own = torch.zeros(3 * chunk, dtype=torch.float64, device="cuda")
        own[:3 * n] = xyz.reshape(-1)
        grad = torch.empty(3 * chunk, dtype=torch.float64, device="cuda")
        head = torch.empty(10, dtype=torch.float64, device="cuda")
        for _ in range(3):
            grad.fill_(float("nan"))
            head.fill_(float("nan"))
            st.check(L.jme_peer_step(comm, st.h, C.c_void_p(own.data_ptr()), C.c_void_p(grad.data_ptr()), C.c_void_p(head.data_ptr())))
            torch.cuda.synchronize()
            assert torch.equal(head, plain[:10])
            assert torch.equal(grad[:3 * n], plain[10:10 + 3 * n])
        ref = oracle.evaluate(s, cutoff=cutoff, gradient=False) if n <= 1000 else oracle.fast_evaluate(s, cutoff=cutoff, gradient=False)
        assert abs(float(head[0].item()) - ref["total"]) <= 1e-10 * max(1.0, abs(ref["total"]))
        L.jme_peer_destroy(comm)
        st.close()


def test_energy_delta_equals_difference_of_full_evaluations(oracle):
    """``jme_energy_delta`` (the energy change of a single-coordinate move from the moved atom's terms only) against the
    difference of two full evaluations - the quantity the reference's minimizers compute - per component: molecules with
    every bonded term type and ions (golden fixtures), a water box without and with a cutoff, other units / dielectric.
    The oracle's full energy after the moves pins the accumulated result."""
    from jmolecularenergy_b200 import MMFF94, EnergyUnit, systems
    rng = np.random.default_rng(11)
    cases = [(load_golden(n), 0.0) for n in ("ETHANOL", "CO08A", "KHDFRM11", "GUANCH01", "CUVFOO", "ARGIND11") if n in set(golden_names(pinned_only=False))]
    cases += [(systems.water_box(60), 0.0), (systems.water_box(60), 6.0), (systems.solvated_chains(900, chain_fraction=0.4), 7.5)]
    assert len(cases) >= 6
    for s, cutoff in cases:
        s = s.copy() if hasattr(s, "copy") else s
        for unit, eps in ((EnergyUnit.KCAL_PER_MOL, 1.0), (EnergyUnit.KJ_PER_MOL, 2.5)):
            ff = MMFF94(False)
            ff.setCutoffDistance(cutoff)
            ff.setDielectricConstant(eps)
            ff.setEnergyUnit(unit)
            ff.assignParameters(s)
            before = ff.calculateComponents(s)
            running = dict(before)
            for _ in range(12):
                a, ax = int(rng.integers(s.n_atoms)), int(rng.integers(3))
                value = float(s.xyz[a, ax] + rng.uniform(-0.25, 0.25))
                d = ff.energyChangeOfMove(s, a, ax, value)
                after = ff.calculateComponents(s)                 # full evaluation of the moved geometry
                scale = max(1.0, abs(after["total"]))
                for k in ("total", "bond", "angle", "stretch_bend", "oop", "torsion", "vdw", "ele"):
                    assert abs(d[k] - (after[k] - before[k])) <= 2e-10 * scale, (s.name, cutoff, k, d[k], after[k] - before[k])
                    running[k] += d[k]
                before = after
            assert abs(running["total"] - before["total"]) <= 1e-9 * max(1.0, abs(before["total"]))
            if unit == EnergyUnit.KCAL_PER_MOL:
                ref = oracle.evaluate(s, cutoff=cutoff, dielectric=eps, gradient=False)
                assert abs(before["total"] - ref["total"]) <= 1e-10 * max(1.0, abs(ref["total"]))
            ff.release(s)


def test_energy_delta_on_the_cell_paths():
    """jme_energy_delta interleaved with full evaluations on both cutoff schemes (per-atom lists, cluster lists): a move
    made by the delta kernel must invalidate the cell sort, and a move ACROSS a cell boundary or the cutoff of many
    partners must still add up."""
    from jmolecularenergy_b200 import MMFF94, systems
    rng = np.random.default_rng(23)
    for path in ("cells-lists", "cells-clusters"):
        s = systems.ion_box(14)
        ff = MMFF94(False, path=path)
        ff.setCutoffDistance(9.0)
        ff.assignParameters(s)
        before = ff.calculateComponents(s)
        for k in range(10):
            a, ax = int(rng.integers(s.n_atoms)), int(rng.integers(3))
            step = 6.0 if k % 3 == 0 else 0.3                     # every third move jumps over a cell (edge 4.5 A)
            d = ff.energyChangeOfMove(s, a, ax, float(s.xyz[a, ax] + step))
            after = ff.calculateComponents(s)
            scale = max(1.0, abs(after["total"]))
            for key in ("total", "vdw", "ele"):
                assert abs(d[key] - (after[key] - before[key])) <= 5e-10 * scale, (path, k, key, d[key], after[key] - before[key])
            before = after
        ff.release(s)


def test_incremental_minimizers_follow_the_full_evaluation_loops():
    """The native minimizer loops with ``incremental=True`` (every trial answered by jme_energy_delta) against the same
    loops with a full evaluation per trial: same number of steps and energy calls, step energies equal to rounding."""
    from jmolecularenergy_b200 import MMFF94, NativeAdamMinimizer, NativeGradientDescentMinimizer, systems
    from jmolecularenergy_b200 import EnergyUnit
    builders = [(lambda: systems.water_box(8), 0.0, EnergyUnit.KCAL_PER_MOL),
                (lambda: load_golden("CO08A"), 0.0, EnergyUnit.KJ_PER_MOL),          # torsions, out-of-plane terms, 1-4 pairs
                (lambda: systems.water_box(20), 5.0, EnergyUnit.KCAL_PER_MOL)]      # with a cutoff
    makers = (lambda ff, inc: NativeGradientDescentMinimizer(ff, incremental=inc),
              lambda ff, inc: NativeAdamMinimizer(ff, 0.9, 0.999, 1e-8, incremental=inc))
    for build, cutoff, unit in builders:
      for make in makers:
        runs = []
        for inc in (False, True, "device"):
            s = build()
            s.xyz += np.random.default_rng(5).normal(0, 0.05, s.xyz.shape)
            ff = MMFF94(False)
            ff.setCutoffDistance(cutoff)
            ff.setEnergyUnit(unit)
            ff.assignParameters(s)
            m = make(ff, inc)
            m.minimize(s, 4, 0.02, 1e-9)
            final = ff.calculateEnergy(s)
            runs.append((m.getStep() if hasattr(m, "getStep") else m.step, m.energyCalls, list(m.stepEnergies), final, s.xyz.copy()))
            ff.release(s)
        (st0, c0, e0, f0, x0) = runs[0]
        for (st1, c1, e1, f1, x1) in runs[1:]:                     # jme_energy_delta per trial; the whole loop in one kernel
            assert st0 == st1 and len(e0) == len(e1)
            assert abs(c0 - c1) <= 1                               # (the incremental loops make one full evaluation, at the start)
            np.testing.assert_allclose(e1, e0, rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(x1, x0, rtol=0, atol=1e-9)
            assert abs(e1[-1] - f1) <= 1e-8 * max(1.0, abs(f1))   # the running energy of the incremental loop is the real one


def test_incremental_force_field_behind_the_unchanged_minimizer():
    """``MMFF94.setIncremental(True)``: the reference-style Python minimizer loop (one ``calculateEnergy`` after every
    single-coordinate move) gets its energies from jme_energy_delta; same trajectory as with full evaluations, and a
    change of more than one coordinate, of an option, or a gradient call falls back to a full evaluation."""
    from jmolecularenergy_b200 import GradientDescentMinimizer, MMFF94, systems
    runs = []
    for inc in (False, True):
        s = systems.water_box(6)
        s.xyz += np.random.default_rng(3).normal(0, 0.05, s.xyz.shape)
        ff = MMFF94(False)
        ff.setIncremental(inc)
        ff.assignParameters(s)
        energies = []
        m = GradientDescentMinimizer(ff)
        m.setOnStepConsumer(lambda k, e: energies.append(e))
        launches0 = ff.launchCount(s)
        m.minimize(s, 3, 0.02, 1e-9)
        runs.append((energies, s.xyz.copy(), ff.launchCount(s) - launches0))
        full = MMFF94(False)
        full.assignParameters(s)
        assert abs(full.calculateEnergy(s) - ff.calculateEnergy(s)) <= 1e-9
        # more than one coordinate moved: a full evaluation again
        s.xyz[0] += 0.01
        assert abs(full.calculateEnergy(s) - ff.calculateEnergy(s)) <= 1e-9
        # one coordinate, other dielectric constant: the cache must not be used
        s.xyz[1, 2] += 0.01
        ff.setDielectricConstant(2.0)
        full.setDielectricConstant(2.0)
        assert abs(full.calculateEnergy(s) - ff.calculateEnergy(s)) <= 1e-9
        full.release(s)
        ff.release(s)
    (e0, x0, l0), (e1, x1, l1) = runs
    assert len(e0) == len(e1) and len(e0) >= 2
    np.testing.assert_allclose(e1, e0, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(x1, x0, rtol=0, atol=1e-9)
    assert l1 < l0                                                 # one small kernel per trial instead of an evaluation


def test_gradient_descent_minimizer_matches_oracle_driven_loop(oracle):
    """The reference's GradientDescentMinimizer loop (restated in jmolecularenergy_b200/minimizer.py) driven
    by the GPU force field vs the same loop driven by the oracle energy: same accepted energies."""
    from jmolecularenergy_b200 import GradientDescentMinimizer, systems

    class OracleFF:
        def calculateEnergy(self, c):
            return oracle.evaluate(c)["total"]

    a, b = systems.water_box(4), systems.water_box(4)
    ff = gpu_ff()
    ff.assignParameters(a)
    log_a, log_b = [], []
    m1 = GradientDescentMinimizer(ff)
    m1.setOnStepConsumer(lambda step, e: log_a.append(e))
    m1.minimize(a, 3, 0.01, 1e-4)
    m2 = GradientDescentMinimizer(OracleFF())
    m2.setOnStepConsumer(lambda step, e: log_b.append(e))
    m2.minimize(b, 3, 0.01, 1e-4)
    assert m1.energyCalls == m2.energyCalls == 1 + 4 * 36
    np.testing.assert_allclose(log_a, log_b, rtol=1e-9)
    np.testing.assert_allclose(a.xyz, b.xyz, atol=1e-9)
    assert log_a[-1] < log_a[0]


def test_native_gradient_descent_equals_python_loop():
    """jme_gd_minimize (the loop run inside the library) and the Python mirror of GradientDescentMinimizer make the
    same moves: same number of energy calls, same energies per step, same final coordinates."""
    from jmolecularenergy_b200 import GradientDescentMinimizer, NativeGradientDescentMinimizer, systems
    a, b = systems.water_box(5), systems.water_box(5)
    fa, fb = gpu_ff(), gpu_ff()
    fa.assignParameters(a)
    fb.assignParameters(b)
    log_a = []
    m1 = GradientDescentMinimizer(fa)
    m1.setOnStepConsumer(lambda step, e: log_a.append(e))
    m1.minimize(a, 2, 0.01, 1e-4)
    m2 = NativeGradientDescentMinimizer(fb)
    m2.minimize(b, 2, 0.01, 1e-4)
    assert m1.energyCalls == m2.energyCalls == 1 + 3 * 45
    assert m1.getStep() == m2.getStep() == 2
    np.testing.assert_allclose(m2.stepEnergies, log_a, rtol=1e-13)
    np.testing.assert_allclose(a.xyz, b.xyz, rtol=0, atol=1e-13)
    # the handle's device coordinates follow the moves: a fresh evaluation agrees with the logged energy
    assert fb.calculateEnergy(b) == pytest.approx(m2.stepEnergies[-1], rel=1e-12)


def test_adam_minimizer_runs_and_matches_oracle_driven_loop(oracle):
    from jmolecularenergy_b200 import AdamMinimizer, systems

    class OracleFF:
        def calculateEnergy(self, c):
            return oracle.evaluate(c)["total"]

    a, b = systems.water_box(2), systems.water_box(2)
    ff = gpu_ff()
    ff.assignParameters(a)
    la, lb = [], []
    m1 = AdamMinimizer(ff, 0.9, 0.999, 1e-8)
    m1.setOnStepConsumer(lambda step, e: la.append(e))
    m1.minimize(a, 2, 0.01, 1e-6)
    m2 = AdamMinimizer(OracleFF(), 0.9, 0.999, 1e-8)
    m2.setOnStepConsumer(lambda step, e: lb.append(e))
    m2.minimize(b, 2, 0.01, 1e-6)
    assert m1.energyCalls == m2.energyCalls
    np.testing.assert_allclose(la, lb, rtol=1e-9)


def test_native_adam_equals_python_loop():
    """jme_adam_minimize and the Python mirror of AdamMinimizer (AdamMinimizer.java:62-120) make the same moves."""
    from jmolecularenergy_b200 import AdamMinimizer, NativeAdamMinimizer, systems
    a, b = systems.water_box(4), systems.water_box(4)
    fa, fb = gpu_ff(), gpu_ff()
    fa.assignParameters(a)
    fb.assignParameters(b)
    log_a = []
    m1 = AdamMinimizer(fa, 0.9, 0.999, 1e-8)
    m1.setOnStepConsumer(lambda step, e: log_a.append(e))
    m1.minimize(a, 2, 0.01, 1e-6)
    m2 = NativeAdamMinimizer(fb, 0.9, 0.999, 1e-8)
    m2.minimize(b, 2, 0.01, 1e-6)
    assert m1.energyCalls == m2.energyCalls == 1 + 3 * 2 * 36
    assert m1.getStep() == m2.getStep() == 2
    np.testing.assert_allclose(m2.stepEnergies, log_a, rtol=1e-13)
    np.testing.assert_allclose(a.xyz, b.xyz, rtol=0, atol=1e-13)
    assert fb.calculateEnergy(b) == pytest.approx(m2.stepEnergies[-1], rel=1e-12)


def test_config4_opt5k_native_loop_against_oracle_driven_loop(oracle):
    """BASELINE config 4 at size: the first 50 energy calls of the reference's GradientDescentMinimizer loop on the
    5 001-atom water box, native loop on the GPU vs the same loop driven by the CPU oracle."""
    from jmolecularenergy_b200 import NativeGradientDescentMinimizer, systems
    s = systems.water_box(1667)
    assert s.n_atoms == 5001
    ff = gpu_ff()
    ff.assignParameters(s)
    # the oracle-driven restatement of GradientDescentMinimizer.java:49-105, stopped after 50 energy calls
    xyz = s.xyz.copy()
    ref_s = systems.water_box(1667)

    def e_ref():
        ref_s.xyz[:] = xyz
        return oracle.fast_evaluate(ref_s, cutoff=0.0, gradient=False)["total"]

    calls = 1
    last = e_ref()
    trace = [last]
    step_size = 0.01
    for a in range(s.n_atoms):
        for i in range(3):
            if calls >= 50:
                break
            d = min(abs(step_size * step_size), .3)
            xyz[a, i] += d
            e = e_ref()
            calls += 1
            trace.append(e)
            xyz[a, i] += -d           # initialising sweep: every move is undone
        if calls >= 50:
            break
    # the GPU: same moves through jme_set_coord1 + jme_energy (what jme_gd_minimize does per call)
    st = ff._state(s)
    st.sync_options(ff, ff.path)
    got = [ff.calculateEnergy(s)]
    L = st.L
    tot = C.c_double()
    k = 1
    cur = s.xyz.copy()
    for a in range(s.n_atoms):
        for i in range(3):
            if k >= 50:
                break
            d = min(abs(step_size * step_size), .3)
            cur[a, i] += d
            st.check(L.jme_set_coord1(st.h, a, i, float(cur[a, i])))
            st.check(L.jme_energy(st.h, C.byref(tot), None))
            got.append(tot.value)
            cur[a, i] += -d
            st.check(L.jme_set_coord1(st.h, a, i, float(cur[a, i])))
            k += 1
        if k >= 50:
            break
    np.testing.assert_allclose(got, trace, rtol=1e-10)
    # and the native loop itself reproduces the first logged energy and makes 1 + 2 * 3N calls for one step
    m = NativeGradientDescentMinimizer(ff)
    m.minimize(s, 0, step_size, 1e-4)
    assert m.energyCalls == 1
    assert m.stepEnergies[0] == pytest.approx(trace[0], rel=1e-10)


@pytest.mark.parametrize("n_waters", [42, 43, 44, 86])
def test_superblock_boundaries(oracle, n_waters):
    """126 / 129 / 132 / 258 atoms: just below and above one and two 128-atom superblocks (padding chains,
    self tiles, split-k CTAs), with and without a cutoff on the all-pairs path."""
    from jmolecularenergy_b200 import systems
    s = systems.water_box(n_waters)
    for cutoff in (0.0, 6.5):
        ff = gpu_ff(path="allpairs", cutoff=cutoff)
        ff.assignParameters(s)
        assert_parity(ff.calculateComponents(s, gradient=True), oracle.evaluate(s, cutoff=cutoff, gradient=True), f"{s.name}@{cutoff}")
        pi, pj, f14 = ff.pairList(s)
        oi, oj, of = oracle.pair_list(s, cutoff=cutoff)
        assert np.array_equal(pi, oi) and np.array_equal(pj, oj) and np.array_equal(f14, of)


def test_fp64_peak_probe():
    from jmolecularenergy_b200 import measure_fp64_peak
    tflops, mhz = measure_fp64_peak()
    assert 5.0 < tflops < 80.0


def _star_in_ion_bath(n_leaves=70, n_side=9):
    """An artificial hub bonded to ``n_leaves`` atoms inside an ion box: every atom of the star has more than
    64 excluded partners (hub: its 1-2 partners; leaf: hub + the other leaves as 1-3), which is longer than the
    exclusion buffer the cell-path pair kernel keeps in shared memory."""
    from jmolecularenergy_b200 import systems
    from jmolecularenergy_b200.system import MolecularSystem
    bath = systems.ion_box(n_side)
    rng = np.random.default_rng(23)
    centre = bath.xyz.mean(axis=0)
    d = rng.normal(size=(n_leaves, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    star = np.vstack([centre[None, :], centre[None, :] + 1.9 * d])
    keep = np.linalg.norm(bath.xyz - centre, axis=1) > 3.5
    xyz = np.vstack([star, bath.xyz[keep]])
    n_star = n_leaves + 1
    return MolecularSystem(
        xyz=xyz, atom_type=np.concatenate([np.full(n_star, 90), bath.atom_type[keep]]),
        charge=np.concatenate([np.full(n_star, -0.05), bath.charge[keep]]),
        linear=np.zeros(xyz.shape[0], np.uint8),
        vdw_type=bath.vdw_type, vdw_alpha=bath.vdw_alpha, vdw_N=bath.vdw_N, vdw_A=bath.vdw_A, vdw_G=bath.vdw_G,
        vdw_DA=bath.vdw_DA,
        bond_i=np.zeros(n_leaves, np.int32), bond_j=np.arange(1, n_star, dtype=np.int32),
        bond_kb=np.full(n_leaves, 3.0, np.float32), bond_r0=np.full(n_leaves, 1.9, np.float32),
        name=f"star{n_leaves}-in-ions")


@pytest.mark.parametrize("path", ["allpairs", *CELL_PATHS])
def test_exclusion_rows_longer_than_the_shared_buffer(oracle, path):
    s = _star_in_ion_bath()
    ff = gpu_ff(path=path, cutoff=9.0)
    ff.assignParameters(s)
    ref = oracle.evaluate(s, cutoff=9.0, gradient=True)
    assert_parity(ff.calculateComponents(s, gradient=True), ref, f"{path} star")
    pi, pj, f14 = ff.pairList(s)
    oi, oj, of = oracle.pair_list(s, cutoff=9.0)
    assert np.array_equal(pi, oi) and np.array_equal(pj, oj) and np.array_equal(f14, of)


@pytest.mark.parametrize("cells", CELL_PATHS)
def test_cell_path_coincident_atoms_and_long_lists(oracle, cells):
    """Two ions on the same point (r = 0: finite buffered energy, no gradient direction) and a dense cluster
    whose survivor lists span several 128-entry chunks, through the cell path."""
    from jmolecularenergy_b200 import systems
    s = systems.ion_box(12, spacing=1.7)           # ~2x liquid density: ~850 partners within 10 A
    s.xyz[701] = s.xyz[64]
    ff = gpu_ff(path=cells, cutoff=10.0)
    ff.assignParameters(s)
    got = ff.calculateComponents(s, gradient=True)
    ref = oracle.evaluate(s, cutoff=10.0, gradient=True)
    assert np.isfinite(got["total"]) and np.all(np.isfinite(got["gradient"]))
    # the CPU restatement divides 0 by 0 for the direction of the coincident pair: compare everything else
    g_ref = ref.pop("gradient")
    assert_parity(got, ref, "coincident")
    ok = np.isfinite(g_ref).all(axis=1)
    assert ok.sum() >= s.n_atoms - 2
    floor = 1e-10 * float(np.abs(g_ref[ok]).max())
    np.testing.assert_allclose(got["gradient"][ok], g_ref[ok], rtol=G_RTOL, atol=floor)


@pytest.mark.parametrize("cells", CELL_PATHS)
def test_cell_path_grows_its_lists_for_dense_systems(oracle, cells):
    """More partners inside the cutoff sphere than the initial survivor-list capacity (1024): the host API grows
    the lists and re-evaluates; a device-side caller of a fresh handle sees a poisoned (NaN) total, never a
    silently truncated sum."""
    from jmolecularenergy_b200 import systems
    from jmolecularenergy_b200.distributed import ShardedMMFF94, unpack
    s = systems.ion_box(13, spacing=1.7)           # 0.2 atoms/A^3: ~1450 partners within 12 A
    ev = ShardedMMFF94(s, cutoff=12.0, path=cells, rank=0, world=1)
    ev.set_coords(s.xyz)
    raw = ev.evaluate(gradient=True)
    ev.close()
    assert np.isnan(raw["total"])
    ff = gpu_ff(path=cells, cutoff=12.0)
    ff.assignParameters(s)
    got = ff.calculateComponents(s, gradient=True)
    ref = oracle.evaluate(s, cutoff=12.0, gradient=True)
    assert_parity(got, ref, "dense")
    pi, pj, f14 = ff.pairList(s)
    assert len(pi) == ref["pairs"]


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6, 7, 8])
def test_random_systems_both_cutoff_paths(oracle, seed):
    """Randomised sweep over sizes, densities and cutoffs (group / chunk / trip boundaries of the cell path fall
    differently every time): cell path == all-pairs path == oracle, pair counts included."""
    from jmolecularenergy_b200 import systems
    rng = np.random.default_rng(1000 + seed)
    kind = seed % 3
    if kind == 0:
        s = systems.ion_box(int(rng.integers(5, 15)), seed=seed, spacing=float(rng.uniform(1.8, 3.2)))
    elif kind == 1:
        s = systems.water_box(int(rng.integers(40, 900)), seed=seed)
    else:
        s = systems.solvated_chains(int(rng.integers(1500, 5000)), chain_fraction=float(rng.uniform(0.1, 0.5)), seed=seed)
    s.xyz[:] = s.xyz + rng.normal(0, 0.05, size=s.xyz.shape)
    cutoff = float(rng.uniform(3.5, 11.0))
    ref = oracle.evaluate(s, cutoff=cutoff, gradient=True) if s.n_atoms <= 1200 else \
        oracle.fast_evaluate(s, cutoff=cutoff, gradient=True)
    for path in (*CELL_PATHS, "allpairs"):
        ff = gpu_ff(path=path, cutoff=cutoff)
        ff.assignParameters(s)
        got = ff.calculateComponents(s, gradient=True)
        assert got["pairs"] == ref["pairs"], (path, seed, got["pairs"], ref["pairs"])
        for k in ("vdw", "ele"):
            scale = max(abs(ref[k]), 1e-3)
            assert abs(got[k] - ref[k]) <= E_RTOL * scale, (path, seed, k, got[k], ref[k])
        if s.n_atoms <= 1200:
            assert_parity(got, ref, f"{path} seed {seed}")
        e_only = ff.calculateEnergy(s)
        assert e_only == pytest.approx(got["total"], rel=1e-13)


def test_cell_path_non_finite_coordinates_poison_the_total_and_return():
    """A coordinate that is +-inf or NaN has no cell: the cell path must neither hang in the grid set-up nor read
    out of bounds - it returns with a NaN total (the edge-growth loop used to spin for ever)."""
    from jmolecularenergy_b200 import systems
    for bad in (np.inf, -np.inf, np.nan, -np.nan):
        for cells in CELL_PATHS:
            s = systems.water_box(120)
            s.xyz[17, 1] = bad
            ff = gpu_ff(path=cells, cutoff=7.0)
            ff.assignParameters(s)
            got = ff.calculateComponents(s, gradient=True)
            assert np.isnan(got["total"]), (bad, cells)
            # and the handle is still usable afterwards
            s.xyz[17, 1] = 3.25
            assert np.isfinite(ff.calculateEnergy(s))


@pytest.mark.parametrize("cutoff,x1,x2", [(6.0, 2.9999999999999996, 9.0), (10.0, 4.999999999999999, 15.0)])
def test_cell_boundary_pair_at_exactly_the_cutoff(oracle, cutoff, x1, x2):
    """Two atoms whose double-rounded distance EQUALS the cutoff (kept: the reference drops a pair only if
    distance > cutoff) and whose cell coordinates straddle cell boundaries: the cell path, the all-pairs
    path and the oracle select the same pair set."""
    from jmolecularenergy_b200 import systems
    s = systems.ion_box(6)
    s.xyz[:] = s.xyz - s.xyz.min(axis=0)            # the bounding box starts at the origin, like the advisor's example
    s.xyz[0] = (0.0, 0.0, 0.0)
    s.xyz[1] = (x1, 0.0, 0.0)
    s.xyz[2] = (x2, 0.0, 0.0)
    assert np.sqrt((x2 - x1) ** 2) == cutoff
    oi, oj, of = oracle.pair_list(s, cutoff=cutoff)
    assert np.any((oi == 1) & (oj == 2))
    for path in (*CELL_PATHS, "allpairs"):
        ff = gpu_ff(path=path, cutoff=cutoff)
        ff.assignParameters(s)
        pi, pj, f14 = ff.pairList(s)
        assert np.array_equal(pi, oi) and np.array_equal(pj, oj), path
        assert_parity(ff.calculateComponents(s, gradient=True), oracle.evaluate(s, cutoff=cutoff, gradient=True), path)


def test_cell_path_many_atom_types_table_outside_shared_memory(oracle):
    """A type-pair table too large to sit next to the force windows (80+ types used to fail with a generic
    CUDA error): the cutoff path reads it through the cache hierarchy instead."""
    from jmolecularenergy_b200 import systems
    from jmolecularenergy_b200.system import MolecularSystem
    base = systems.ion_box(9)
    rng = np.random.default_rng(4)
    types = [t for t in list(range(1, 83)) + list(range(87, 100))][:90]
    s = MolecularSystem(
        xyz=base.xyz, atom_type=rng.choice(types, size=base.n_atoms), charge=base.charge, linear=base.linear,
        vdw_type=types, vdw_alpha=rng.uniform(0.5, 3.0, len(types)).astype(np.float32),
        vdw_N=rng.uniform(1.5, 4.5, len(types)).astype(np.float32), vdw_A=rng.uniform(2.5, 4.2, len(types)).astype(np.float32),
        vdw_G=rng.uniform(0.8, 1.4, len(types)).astype(np.float32), vdw_DA=rng.integers(0, 3, len(types)).astype(np.uint8),
        name="ions-90-types")
    ref = oracle.evaluate(s, cutoff=8.0, gradient=True)
    for path in (*CELL_PATHS, "allpairs"):
        ff = gpu_ff(path=path, cutoff=8.0)
        ff.assignParameters(s)
        assert_parity(ff.calculateComponents(s, gradient=True), ref, path)
