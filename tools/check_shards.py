"""Development: shards of the cutoff path emulated on ONE GPU (handles with rank r of W) - the partial results must add
up to the unsharded evaluation.  Prints per-rank totals / pair counts so that a poisoned (NaN) shard is visible."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from jmolecularenergy_b200 import _lib, systems  # noqa: E402
from jmolecularenergy_b200.forcefield import _GpuState  # noqa: E402


def evaluate(s, cutoff, path, rank, world, reps=2):
    L = _lib.lib()
    st = _GpuState(s)
    st.check(L.jme_set_options(st.h, cutoff, 1.0, 0))
    st.check(L.jme_set_path(st.h, path))
    st.check(L.jme_set_shard(st.h, rank, world))
    xyz = torch.from_numpy(s.xyz).cuda()
    out = torch.zeros(L.jme_eval_out_size(st.h, 1), dtype=torch.float64, device="cuda")
    st.check(L.jme_set_stream(st.h, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    st.check(L.jme_set_coords_device(st.h, C.c_void_p(xyz.data_ptr()), s.n_atoms))
    tot = C.c_double()
    st.check(L.jme_energy(st.h, C.byref(tot), None))        # host entry point: sizes the lists
    for _ in range(reps):
        st.check(L.jme_set_coords_device(st.h, C.c_void_p(xyz.data_ptr()), s.n_atoms))
        st.check(L.jme_eval_device(st.h, 1, C.c_void_p(out.data_ptr())))
    torch.cuda.synchronize()
    r = out.cpu().numpy()
    st.close()
    return r


def main():
    torch.cuda.set_device(0)
    torch.cuda.set_stream(torch.cuda.Stream())
    cases = [("ions27k", systems.ion_box(30), 10.0), ("solv6k", systems.solvated_chains(6000), 9.0),
             ("ions125k", systems.ion_box(50), 10.0)]
    for label, s, cutoff in cases:
        for path in (4, 3):
            full = evaluate(s, cutoff, path, 0, 1)
            for world in (2, 3, 8):
                parts = [evaluate(s, cutoff, path, r, world) for r in range(world)]
                tot = sum(parts)
                gerr = np.abs(tot[10:] - full[10:]).max() / max(1.0, np.abs(full[10:]).max())
                print(f"{label} path {path} world {world}: pairs {int(tot[8])} vs {int(full[8])}  E {tot[0]:.6f} vs {full[0]:.6f}  "
                      f"grad err {gerr:.2e}  per-rank totals {[float(p[0]) for p in parts]} pairs {[int(p[8]) for p in parts]}", flush=True)


if __name__ == "__main__":
    main()
