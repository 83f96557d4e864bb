"""Worker of tests/test_multi_gpu.py: one rank per GPU under torchrun (NCCL).  Every rank evaluates its shard of a
system through libjme_b200.so (owner mode, first with the collectives - all-gather of the coordinates, all-reduce of the
energies, reduce-scatter of the gradient - then over NVLink peer memory, jme_peer_step) and compares what it ends up with - the job-wide energies / pair count and the gradient rows of the atoms
it owns - with the CPU oracle.  Prints one JSON line per rank and exits non-zero on a mismatch."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    from jmolecularenergy_b200 import systems
    from jmolecularenergy_b200.distributed import ShardedMMFF94, unpack
    from oracle import oracle
    cases = [("solvated-6k cluster lists", systems.solvated_chains(6000, exact_cutoff_pairs=40, cutoff=9.0), 9.0, "cells-clusters"),
             ("ions-27k per-atom lists", systems.ion_box(30), 10.0, "cells-lists"),
             ("ions-27k cluster lists", systems.ion_box(30), 10.0, "cells-clusters"),
             ("water-900 all pairs", systems.water_box(300), 0.0, "allpairs")]
    worst = {"energy": 0.0, "gradient": 0.0}
    ok = True
    for label, s, cutoff, path in cases:
        ref = oracle.fast_evaluate(s, cutoff=cutoff, gradient=True)
        n = s.n_atoms
        ev = ShardedMMFF94(s, cutoff=cutoff, path=path, rank=rank, world=world)
        a0, a1 = ev.owned_range()
        xyz_full, out_full, grad_own = ev.new_owner_buffers()
        xyz_own = torch.zeros(3 * ev.chunk, dtype=torch.float64, device=dev)
        xyz_own[:3 * (a1 - a0)] = torch.from_numpy(s.xyz[a0:a1].reshape(-1)).to(dev)
        ev.set_coords_owned(xyz_own, xyz_full)
        ev.prime()                              # sizes the lists of the cutoff path (host entry point)
        for _ in range(2):                      # twice: the second evaluation replays the captured graph
            ev.set_coords_owned(xyz_own, xyz_full)
            ev.evaluate_owned_into(out_full, grad_own)
        torch.cuda.synchronize()
        # the same step over NVLink peer memory (jme_peer_step) must reproduce the collectives' result: energies to
        # rounding (another order of the sum over ranks), and itself bit for bit when repeated
        nccl_head, nccl_grad = out_full[:10].clone(), grad_own.clone()
        ev.open_peer_exchange()
        peer_runs = []
        for _ in range(3):
            out_full.zero_()
            grad_own.zero_()
            ev.step_owned(xyz_own, out_full, grad_own)
            torch.cuda.synchronize()
            peer_runs.append((out_full[:10].clone(), grad_own.clone()))
        peer_same = all(torch.equal(peer_runs[0][0], h) and torch.equal(peer_runs[0][1], g_) for h, g_ in peer_runs[1:])
        scale = float(nccl_grad.abs().max().item()) + 1e-300
        peer_vs_nccl = max(float(((peer_runs[0][0][:8] - nccl_head[:8]).abs() / nccl_head[:8].abs().clamp_min(1e-3)).max().item()),
                           float((peer_runs[0][1] - nccl_grad).abs().max().item()) / scale)
        peer_ok = peer_same and peer_vs_nccl <= 1e-12 and float(peer_runs[0][0][8].item()) == float(nccl_head[8].item())
        got = unpack(out_full.cpu().numpy(), n, False)
        g = grad_own.cpu().numpy()[:3 * (a1 - a0)].reshape(-1, 3)
        e_err = max(abs(got[k] - ref[k]) / max(abs(ref[k]), abs(ref["total"]), 1e-3)
                    for k in ("total", "bond", "angle", "stretch_bend", "oop", "torsion", "vdw", "ele"))
        gr = ref["gradient"][a0:a1]
        floor = 1e-10 * max(1.0, float(np.abs(ref["gradient"]).max()))
        g_err = float(np.max(np.abs(g - gr) / (np.abs(gr) + floor / 1e-8))) if a1 > a0 else 0.0
        case_ok = e_err <= 1e-10 and g_err <= 1e-8 and got["pairs"] == ref["pairs"] and peer_ok
        ok = ok and case_ok
        worst["energy"] = max(worst["energy"], e_err)
        worst["gradient"] = max(worst["gradient"], g_err)
        print(json.dumps({"rank": rank, "case": label, "ok": case_ok, "energy_rel_err": e_err, "gradient_rel_err": g_err,
                          "pairs": got["pairs"], "pairs_ref": ref["pairs"], "peer_repeat_identical": peer_same,
                          "peer_vs_collectives": peer_vs_nccl}), flush=True)
        ev.close()
    flag = torch.tensor([0.0 if ok else 1.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 0.0 else 1)


if __name__ == "__main__":
    main()
