"""Multi-GPU evaluation: one process per GPU, atom-row sharding, one all-reduce.

Every rank holds the whole system (coordinates, charges, types, pair table: 36 MB at 10^6 atoms) and
evaluates its share of the nonbonded rows and bonded terms (``jme_set_shard``); the partial results are
ONE packed device buffer ``[total, 7 components, pairs, candidates, gradient(3N)]`` that is summed over
ranks with ``torch.distributed.all_reduce`` (NCCL over NVLink on GPUs).  This is the only collective of the
path - the pair interactions themselves are independent.  ``torch`` is plumbing here: device memory,
streams and the process group.

``local_eval`` can be injected so that the rank bookkeeping and the reduction are testable on CPU with the
``gloo`` backend (tests/test_distributed_gloo.py), where no CUDA library call is possible.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Dict, Optional

import numpy as np

from ._abi import COMPONENTS, ENERGY_UNITS, PATHS
from .system import MolecularSystem

OUT_HEADER = 10          # kOutHeader in csrc/jme_kernels.cuh


def unpack(out: np.ndarray, n_atoms: int, gradient: bool) -> Dict[str, object]:
    res: Dict[str, object] = {"total": float(out[0]), "pairs": int(round(float(out[8]))),
                              "candidates": int(round(float(out[9])))}
    res.update({k: float(v) for k, v in zip(COMPONENTS, out[1:8])})
    if gradient:
        res["gradient"] = np.asarray(out[OUT_HEADER:OUT_HEADER + 3 * n_atoms]).reshape(n_atoms, 3)
    return res


class ShardedMMFF94:
    """Device-resident, sharded evaluator.  ``rank``/``world`` default to the initialised process group."""

    def __init__(self, system: MolecularSystem, cutoff: float = 0.0, dielectric: float = 1.0,
                 unit: str = "KCAL_PER_MOL", path: str = "auto", rank: Optional[int] = None,
                 world: Optional[int] = None, local_eval: Optional[Callable] = None, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.system = system
        self.n = system.n_atoms
        if rank is None or world is None:
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(), dist.get_world_size()
            else:
                rank, world = 0, 1
        if not (0 <= rank < world):
            raise ValueError("need 0 <= rank < world")
        self.rank, self.world = rank, world
        self.local_eval = local_eval
        self.state = None
        if local_eval is None:
            from . import _lib
            from .forcefield import _GpuState
            self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
            self.state = st = _GpuState(system)
            L = self.L = _lib.lib()
            st.check(L.jme_set_options(st.h, cutoff, dielectric, ENERGY_UNITS[unit]))
            st.check(L.jme_set_path(st.h, PATHS[path]))
            st.check(L.jme_set_shard(st.h, rank, world))
            self.n_out_grad = int(L.jme_eval_out_size(st.h, 1))
            self._bind_stream()
        else:
            self.device = torch.device("cpu")
            self.n_out_grad = OUT_HEADER + 3 * self.n

    def _bind_stream(self):
        st = self.state
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        st.check(self.L.jme_set_stream(st.h, C.c_void_p(stream)))

    def new_output(self, gradient: bool = True):
        n = self.n_out_grad if gradient else OUT_HEADER
        return self.torch.zeros(n, dtype=self.torch.float64, device=self.device)

    def set_coords(self, xyz) -> None:
        """``xyz``: host array (N,3) or a CUDA float64 tensor (N,3) already on this rank's device."""
        torch = self.torch
        if self.local_eval is not None:
            self._xyz_host = np.ascontiguousarray(xyz.cpu().numpy() if isinstance(xyz, torch.Tensor) else xyz,
                                                  dtype=np.float64).reshape(-1, 3)
            return
        st = self.state
        if isinstance(xyz, torch.Tensor) and xyz.is_cuda:
            t = xyz.contiguous()
            if t.dtype != torch.float64 or t.numel() != 3 * self.n:
                raise ValueError("need a float64 CUDA tensor with 3N elements")
            self._bind_stream()
            st.check(self.L.jme_set_coords_device(st.h, C.c_void_p(t.data_ptr()), self.n))
            self._keep = t
        else:
            a = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1)
            st.check(self.L.jme_set_coords(st.h, a.ctypes.data, self.n))

    def evaluate_into(self, out, gradient: bool = True) -> None:
        """Enqueue the local evaluation and the all-reduce on the current stream; no host sync."""
        if self.local_eval is not None:
            part = self.local_eval(self.system, self._xyz_host, self.rank, self.world, gradient)
            out.copy_(self.torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64)))
        else:
            st = self.state
            st.check(self.L.jme_eval_device(st.h, 1 if gradient else 0, C.c_void_p(out.data_ptr())))
        if self.world > 1:
            self.dist.all_reduce(out, op=self.dist.ReduceOp.SUM)

    def evaluate(self, gradient: bool = True) -> Dict[str, object]:
        out = self.new_output(gradient)
        self.evaluate_into(out, gradient)
        if out.is_cuda:
            self.torch.cuda.synchronize(self.device)
        return unpack(out.cpu().numpy(), self.n, gradient)

    def launch_count(self) -> int:
        return 0 if self.state is None else int(self.L.jme_launch_count(self.state.h))

    def close(self) -> None:
        if self.state is not None:
            self.state.close()
            self.state = None
