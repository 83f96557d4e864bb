"""Development: step time of the two cutoff paths (JME_CLUSTER = 4 cluster lists / 0 per-atom lists) over system size,
to place the AUTO switch.  Device-resident coordinates, CUDA events, L2 flushed between steps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmolecularenergy_b200 import systems  # noqa: E402
from jmolecularenergy_b200.distributed import ShardedMMFF94  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream(dev))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cases = [("ions%d" % (k ** 3), systems.ion_box(k)) for k in (20, 30, 40, 50, 60, 80)]
    cases.insert(2, ("solvated25k", systems.solvated_chains(24999, exact_cutoff_pairs=100, cutoff=10.0)))
    for label, s in cases:
        row = [f"{label:14s} n={s.n_atoms:7d}"]
        for cluster in (4, 2, 0):
            os.environ["JME_CLUSTER"] = str(cluster)
            ev = ShardedMMFF94(s, cutoff=10.0, path="cells", rank=0, world=1)
            xyz = torch.from_numpy(s.xyz).to(dev)
            out = ev.new_output(True)
            ev.set_coords(xyz)
            ev.prime()
            for _ in range(5):
                ev.set_coords(xyz)
                ev.evaluate_into(out, True)
            torch.cuda.synchronize()
            ts = []
            for k in range(20):
                flush.fill_(k)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                ev.set_coords(xyz)
                ev.evaluate_into(out, True)
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            row.append(f"C={cluster}: {np.median(ts) * 1e3:8.1f} us")
            ev.close()
        print("  ".join(row), flush=True)


if __name__ == "__main__":
    main()
