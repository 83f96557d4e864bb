"""Development check of the cutoff path on a GPU box: energies, gradient, pair count and pair SET of the cluster
path (JME_CLUSTER = 4, 2) and the round-1 per-atom lists (0) against the CPU oracle.  Not a test, not a bench."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmolecularenergy_b200 import MMFF94, systems  # noqa: E402
from oracle import oracle  # noqa: E402


def run(s, cutoff, cluster, ref, label):
    os.environ["JME_CLUSTER"] = str(cluster)
    ff = MMFF94(False, path="cells")
    ff.setCutoffDistance(cutoff)
    ff.assignParameters(s)
    t0 = time.perf_counter()
    try:
        got = ff.calculateComponents(s, gradient=True)
    except Exception as e:  # noqa: BLE001
        print(f"  {label} C={cluster}: FAILED {e}")
        return
    dt = time.perf_counter() - t0
    g, gr = got["gradient"], ref["gradient"]
    scale = max(1.0, float(np.abs(gr).max()))
    gerr = float(np.abs(g - gr).max()) / scale
    bad = int((np.abs(g - gr) > 1e-8 * np.abs(gr) + 1e-10 * scale).sum())
    e_only = ff.calculateComponents(s, gradient=False)
    print(f"  {label} C={cluster}: pairs {got['pairs']} (ref {ref['pairs']}) dE_rel {abs(got['total'] - ref['total']) / max(1e-3, abs(ref['total'])):.2e} "
          f"vdw {abs(got['vdw'] - ref['vdw']):.2e} ele {abs(got['ele'] - ref['ele']):.2e} grad_err {gerr:.2e} bad {bad} "
          f"E-only diff {abs(e_only['total'] - got['total']):.2e} cand {got['candidates']} ({dt * 1e3:.1f} ms first call)")
    if bad:
        idx = np.argwhere(np.abs(g - gr) > 1e-8 * np.abs(gr) + 1e-10 * scale)[:5]
        for a, c in idx:
            print(f"      atom {a} comp {c}: got {g[a, c]:.12e} ref {gr[a, c]:.12e}")
    g2 = ff.calculateComponents(s, gradient=True)
    print(f"      repeat bit-identical: {np.array_equal(g2['gradient'], g) and g2['total'] == got['total']}")
    ff.release(s)


def main():
    cases = [("water100", systems.water_box(100), 7.0), ("water1000", systems.water_box(1000), 9.5),
             ("solv6k", systems.solvated_chains(6000, exact_cutoff_pairs=50, cutoff=10.0), 10.0),
             ("ions27k", systems.ion_box(30), 10.0)]
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        cases.append(("solv25k", systems.solvated_chains(24999, exact_cutoff_pairs=100, cutoff=10.0), 10.0))
        cases.append(("ions125k", systems.ion_box(50), 10.0))
    for label, s, cutoff in cases:
        ref = oracle.fast_evaluate(s, cutoff=cutoff, gradient=True, threads=os.cpu_count())
        print(f"{label}: n={s.n_atoms} cutoff={cutoff} ref E={ref['total']:.6f} pairs={ref['pairs']}")
        for cluster in (4, 2, 0):
            run(s, cutoff, cluster, ref, label)


if __name__ == "__main__":
    main()
