"""Energy-only owner mode for the optimizer loop ("for the minimizer energy-only path just 8 doubles"):
per call one coordinate moves on every rank (jme_set_coord1), every rank evaluates the energy of its shard, and ONE
all-reduce of the 10-double header gives everybody the energy.  Measures calls per second on BASELINE config 4 (5 001
atoms) under torchrun at any world size, next to the single-GPU figure, to show where sharding stops paying."""
import ctypes as C
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmolecularenergy_b200 import systems  # noqa: E402
from jmolecularenergy_b200.distributed import ShardedMMFF94  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    s = systems.water_box(1667)
    ev = ShardedMMFF94(s, cutoff=0.0, path="allpairs", rank=rank, world=world)
    ev.set_coords(s.xyz)
    out = ev.new_output(False)
    L, h = ev.L, ev.state.h
    calls = 2000
    xyz = s.xyz.copy()

    def loop(n):
        for k in range(n):
            a, ax = k % s.n_atoms, k % 3
            L.jme_set_coord1(h, a, ax, float(xyz[a, ax] + 1e-4 * ((k & 1) * 2 - 1)))
            ev.evaluate_into(out, False)          # shard energy + all-reduce of 10 doubles
    loop(100)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    loop(calls)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"workload": "opt5k energy-only owner mode", "n_gpus": world, "atoms": s.n_atoms, "calls": calls,
                          "energy_calls_per_s": calls / dt, "us_per_call": 1e6 * dt / calls,
                          "energy": float(out[0].item())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ev.close()


if __name__ == "__main__":
    main()
