"""World-size-2 CPU test (gloo) of the multi-GPU wrapper: rank bookkeeping, the packed
[energies, pair counts, gradient] buffer and the single all-reduce of SURVEY.md section 8e.

No CUDA call is possible here, so each rank's local evaluation is an oracle-backed stand-in that produces
a genuine partial result: rank r owns the atom rows [N r/W, N (r+1)/W) of the pair matrix (pairs (i<j) go
to the owner of i) and an equal slice of the bonded components.  The GPU shard logic itself
(jme_set_shard) is covered by tests/test_gpu_parity.py::test_two_way_shard_sums_to_unsharded."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partial_with_oracle(system, xyz, rank, world, gradient):
    """Partial packed buffer of one rank: the vdW / electrostatic energy and pair count of the pairs whose
    first atom lies in this rank's row block (recomputed pair by pair from the oracle's pair list), and a
    1/world share of every other additive quantity of one full oracle evaluation."""
    from oracle import oracle
    from oracle import py_oracle
    n = system.n_atoms
    lo, hi = n * rank // world, n * (rank + 1) // world
    s = system.with_coordinates(xyz)
    full = oracle.evaluate(s, gradient=True)
    pi, pj, f14 = oracle.pair_list(s)
    rows = {int(t): (s.vdw_alpha[k], s.vdw_N[k], s.vdw_A[k], s.vdw_G[k], int(s.vdw_DA[k])) for k, t in enumerate(s.vdw_type)}
    out = np.zeros(10 + 3 * n)
    mine = (pi >= lo) & (pi < hi)
    ev = ee = 0.0
    for a, b, f in zip(pi[mine].tolist(), pj[mine].tolist(), f14[mine].tolist()):
        rstar, eps = py_oracle.vdw_pair_constants(rows[int(s.atom_type[a])], rows[int(s.atom_type[b])])
        r = py_oracle._dist(tuple(xyz[a]), tuple(xyz[b]))
        e = eps * (1.07 * rstar / (r + 0.07 * rstar)) ** 7 * ((1.12 * rstar ** 7) / (r ** 7 + 0.12 * rstar ** 7) - 2)
        q = 332.0716 * s.charge[a] * s.charge[b] / (r + 0.05) * (0.75 if f else 1.0)
        ev += e
        ee += q
    out[6], out[7] = ev, ee
    for c, k in enumerate(("bond", "angle", "stretch_bend", "oop", "torsion")):
        out[1 + c] = full[k] / world
    out[0] = out[1:8].sum()
    out[8] = out[9] = int(mine.sum())
    out[10:] = (full["gradient"] / world).reshape(-1)
    return out


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from jmolecularenergy_b200 import systems
        from jmolecularenergy_b200.distributed import ShardedMMFF94
        s = systems.water_box(9)
        ev = ShardedMMFF94(s, local_eval=_partial_with_oracle)
        assert (ev.rank, ev.world) == (rank, world)
        ev.set_coords(s.xyz)
        res = ev.evaluate(gradient=True)
        # owner mode: supply only the own coordinate rows, receive only the own gradient rows
        xyz_full, out_full, grad_own = ev.new_owner_buffers()
        a0, a1 = ev.owned_range()
        own = torch.zeros(3 * ev.chunk, dtype=torch.float64)
        own[:3 * (a1 - a0)] = torch.from_numpy(s.xyz[a0:a1].reshape(-1))
        ev.set_coords_owned(own, xyz_full)
        ev.evaluate_owned_into(out_full, grad_own)
        owner = (a0, a1, float(out_full[0]), int(round(float(out_full[8]))), grad_own[:3 * (a1 - a0)].numpy().reshape(-1, 3).copy())
        q.put((rank, res["total"], res["vdw"], res["ele"], res["pairs"], res["gradient"], owner))
    finally:
        dist.destroy_process_group()


def test_two_rank_all_reduce_matches_single_process(oracle):
    from jmolecularenergy_b200 import systems
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = systems.water_box(9)
    ref = oracle.evaluate(s, gradient=True)
    covered = np.zeros(s.n_atoms, bool)
    for rank, total, vdw, ele, pairs, grad, (a0, a1, o_total, o_pairs, o_grad) in got:
        assert o_total == pytest.approx(ref["total"], rel=1e-11) and o_pairs == ref["pairs"]
        np.testing.assert_allclose(o_grad, ref["gradient"][a0:a1], rtol=1e-11, atol=1e-12)
        assert not covered[a0:a1].any()
        covered[a0:a1] = True
        assert total == pytest.approx(ref["total"], rel=1e-11)
        assert vdw == pytest.approx(ref["vdw"], rel=1e-11)
        assert ele == pytest.approx(ref["ele"], rel=1e-11)
        assert pairs == ref["pairs"]
        np.testing.assert_allclose(grad, ref["gradient"], rtol=1e-11, atol=1e-12)
    assert got[0][1] == got[1][1]          # every rank holds the same reduced result
    assert covered.all()                   # owner mode: the ranks' rows tile the atoms exactly once


def test_rank_validation():
    from jmolecularenergy_b200 import systems
    from jmolecularenergy_b200.distributed import ShardedMMFF94
    with pytest.raises(ValueError):
        ShardedMMFF94(systems.water_box(2), rank=2, world=2, local_eval=lambda *a: None)
