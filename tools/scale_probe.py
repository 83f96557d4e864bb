#!/usr/bin/env python
"""Phase timing of one owner-mode step under torchrun (config 5): all-gather | local evaluation | all-reduce of the
header | reduce-scatter of the gradient, CUDA events on the launching stream, mean over the timed steps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from jmolecularenergy_b200 import systems
from jmolecularenergy_b200.distributed import ShardedMMFF94

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
os.environ.pop("NCCL_DEBUG", None)
dist.init_process_group("nccl", device_id=dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
s = systems.ion_box(100)
ev = ShardedMMFF94(s, cutoff=10.0, path="cells", rank=rank, world=world)
xyz_full, out_full, grad_own = ev.new_owner_buffers()
a0, a1 = ev.owned_range()
own = torch.zeros(3 * ev.chunk, dtype=torch.float64, device=dev)
own[:3 * (a1 - a0)] = torch.from_numpy(s.xyz[a0:a1].reshape(-1)).to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
K = 40
names = ["all_gather", "set_coords+eval", "all_reduce(header)", "reduce_scatter"]
acc = [0.0] * 4
for k in range(K + 5):
    flush.fill_(k & 0xff)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    e[0].record()
    dist.all_gather_into_tensor(xyz_full, own)
    e[1].record()
    ev.set_coords(xyz_full[:3 * ev.n].view(ev.n, 3))
    ev.L.jme_eval_device(ev.state.h, 1, out_full.data_ptr())
    e[2].record()
    dist.all_reduce(out_full[:10])
    e[3].record()
    dist.reduce_scatter_tensor(grad_own, out_full[10:])
    e[4].record()
    torch.cuda.synchronize()
    if k >= 5:
        for i in range(4):
            acc[i] += e[i].elapsed_time(e[i + 1])
t = torch.tensor(acc, dtype=torch.float64, device=dev) / K
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("world", world, " ".join(f"{n}={v:.3f}ms" for n, v in zip(names, t.tolist())), f"sum={float(t.sum()):.3f}ms")
dist.barrier()
dist.destroy_process_group()
