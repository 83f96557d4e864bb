"""Multi-GPU evaluation: one process per GPU, atom-row sharding, one reduction.

Every rank holds the whole system (coordinates, charges, types, pair table: 36 MB at 10^6 atoms) and
evaluates its share of the nonbonded rows and bonded terms (``jme_set_shard``); the partial results are
ONE packed device buffer ``[total, 7 components, pairs, candidates, gradient(3N)]`` that is summed over
ranks with ``torch.distributed.all_reduce`` (NCCL over NVLink on GPUs).  This is the only collective of the
path - the pair interactions themselves are independent.  ``torch`` is plumbing here: device memory,
streams and the process group.

``local_eval`` can be injected so that the rank bookkeeping and the reduction are testable on CPU with the
``gloo`` backend (tests/test_distributed_gloo.py), where no CUDA library call is possible.

Two result layouts: replicated (``evaluate_into``: every rank ends with everything, one all-reduce of 24 N + 80
bytes) and owner mode (``set_coords_owned`` / ``evaluate_owned_into``: a rank supplies and receives only the rows of
the atoms it owns - all-gather of the coordinates, all-reduce of the 10-double header, reduce-scatter of the
gradient) - or, after ``open_peer_exchange``, ``step_owned`` with the same meaning and no collective: the ranks store
their coordinate slices into each other's buffers and read each other's partial gradient rows over NVLink peer memory
(``jme_peer_*``, csrc/jme_peer.cuh).  bench.py uses that for N > 1 (``--exchange nccl`` for the collectives).
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
        self._peer = None
        self.peer_error = ""
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

    def prime(self) -> None:
        """One host-API energy evaluation of the coordinates set last (``jme_energy``): it is the host entry points
        that size the cutoff path's lists to the system (a device-side evaluation whose lists overflow reports a NaN
        total and leaves them as they are).  Call once after the first ``set_coords`` of a device-resident loop."""
        if self.local_eval is not None:
            return
        st = self.state
        total = C.c_double()
        st.check(self.L.jme_energy(st.h, C.byref(total), None))
        self._bind_stream()

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

    # ---- owner mode: every rank owns a contiguous range of atoms (by original index) ---------------------
    # Inputs: a rank supplies only the coordinates of its own atoms; one all-gather over NVLink gives every rank
    # the whole geometry (the cell build needs it).  Outputs: the energies are all-reduced (10 doubles), the
    # gradient is reduce-scattered, so a rank ends with the gradient rows of the atoms it owns.  Per rank this
    # moves 1/world of the bytes of the replicated mode over PCIe and half the bytes over NVLink.
    @property
    def chunk(self) -> int:
        """Atoms per rank in owner mode (even; the last ranks' chunks are padded) - ``jme_peer_chunk``."""
        return ((self.n + self.world - 1) // self.world + 1) // 2 * 2

    def owned_range(self):
        a0 = min(self.n, self.rank * self.chunk)
        return a0, min(self.n, a0 + self.chunk)

    def new_owner_buffers(self):
        """(xyz_full, out_full, grad_own): padded device buffers for ``set_coords_owned`` / ``evaluate_owned_into``."""
        t, c, w = self.torch, self.chunk, self.world
        xyz_full = t.zeros(3 * c * w, dtype=t.float64, device=self.device)
        out_full = t.zeros(OUT_HEADER + 3 * c * w, dtype=t.float64, device=self.device)
        grad_own = t.zeros(3 * c, dtype=t.float64, device=self.device)
        return xyz_full, out_full, grad_own

    def set_coords_owned(self, xyz_own, xyz_full) -> None:
        """``xyz_own``: float64 tensor with the 3 * chunk coordinates of this rank's atoms (padding ignored),
        on this rank's device.  Gathers everybody's slice into ``xyz_full`` and hands it to the library."""
        if self.world > 1:
            self.dist.all_gather_into_tensor(xyz_full, xyz_own)
        else:
            xyz_full.copy_(xyz_own)
        self.set_coords(xyz_full[:3 * self.n].view(self.n, 3))

    def evaluate_owned_into(self, out_full, grad_own) -> None:
        """Local evaluation into ``out_full``; afterwards ``out_full[:10]`` holds the job-wide energies / pair
        counts on every rank and ``grad_own`` the gradient rows of this rank's atoms.  No host sync."""
        n_out = self.n_out_grad
        if self.local_eval is not None:
            part = self.local_eval(self.system, self._xyz_host, self.rank, self.world, True)
            out_full[:n_out].copy_(self.torch.from_numpy(np.ascontiguousarray(part, dtype=np.float64)))
        else:
            st = self.state
            st.check(self.L.jme_eval_device(st.h, 1, C.c_void_p(out_full.data_ptr())))
        grad_full = out_full[OUT_HEADER:]
        if self.world > 1:
            self.dist.all_reduce(out_full[:OUT_HEADER], op=self.dist.ReduceOp.SUM)
            if out_full.is_cuda:
                self.dist.reduce_scatter_tensor(grad_own, grad_full, op=self.dist.ReduceOp.SUM)
            else:                       # gloo has no reduce-scatter: reduce everything, keep the own rows
                self.dist.all_reduce(grad_full, op=self.dist.ReduceOp.SUM)
                grad_own.copy_(grad_full[3 * self.chunk * self.rank:3 * self.chunk * (self.rank + 1)])
        else:
            grad_own.copy_(grad_full)

    # ---- owner mode over NVLink peer memory (include/jme_b200.h: jme_peer_*; csrc/jme_peer.cuh) -----------
    def open_peer_exchange(self) -> bool:
        """Create this rank's exported block, swap the 64-byte IPC handles through the process group (the only use
        of NCCL on this path) and map the peers' blocks.  Collective: every rank calls it.  Returns False - on every
        rank, with ``peer_error`` set - when some rank could not map a peer (no peer access between two of the GPUs,
        IPC not permitted); ``step_owned`` then keeps using the collectives."""
        if self.local_eval is not None or self.world == 1:
            return False
        if self._peer is not None:
            return True
        torch, dist, L = self.torch, self.dist, self.L
        comm = C.c_void_p()
        self.peer_error = ""
        rc = L.jme_peer_create(self.rank, self.world, self.n, C.byref(comm))
        mine = (C.c_ubyte * 64)()
        if rc == 0:
            rc = L.jme_peer_handle(comm, mine)
        if rc != 0:
            self.peer_error = f"jme_peer_create / jme_peer_handle failed ({rc})"
        t = torch.tensor(list(mine), dtype=torch.uint8, device=self.device)
        every = torch.empty(64 * self.world, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(every, t)
        if rc == 0:
            if int(L.jme_peer_chunk(comm)) != self.chunk:
                raise RuntimeError("chunk size mismatch between distributed.py and jme_peer_chunk")
            buf = (C.c_ubyte * (64 * self.world)).from_buffer_copy(every.cpu().numpy().tobytes())
            rc = L.jme_peer_open(comm, buf)
            if rc != 0:
                self.peer_error = f"jme_peer_open failed ({rc}): {L.jme_peer_last_error(comm).decode()}"
        ok = torch.tensor([1.0 if rc == 0 else 0.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)    # also: everybody has mapped everybody before the first store
        if ok.item() != 1.0:
            if comm:
                L.jme_peer_destroy(comm)
            self.peer_error = self.peer_error or "another rank could not map its peers"
            return False
        self._peer = comm
        return True

    def close_peer_exchange(self) -> None:
        """Back to the collectives (every rank must do it at the same point of its loop)."""
        if self._peer is not None:
            self.torch.cuda.synchronize(self.device)
            self.dist.barrier()                      # nobody unmaps a block a peer still reads
            self.L.jme_peer_destroy(self._peer)
            self._peer = None

    def step_owned(self, xyz_own, out_full, grad_own, xyz_full=None) -> None:
        """One owner-mode step: ``xyz_own`` (3 * chunk doubles on this rank's device) in, ``out_full[:10]`` = the
        job-wide header and ``grad_own`` (3 * chunk) out.  Over peer memory when ``open_peer_exchange`` was called,
        else through the process group (all-gather, all-reduce, reduce-scatter; needs ``xyz_full``)."""
        if self._peer is not None:
            self._bind_stream()
            self.state.check(self.L.jme_peer_step(self._peer, self.state.h, C.c_void_p(xyz_own.data_ptr()),
                                                  C.c_void_p(grad_own.data_ptr()), C.c_void_p(out_full.data_ptr())))
            self._keep = (xyz_own, out_full, grad_own)
        else:
            self.set_coords_owned(xyz_own, xyz_full)
            self.evaluate_owned_into(out_full, grad_own)

    def evaluate(self, gradient: bool = True) -> Dict[str, object]:
        out = self.new_output(gradient)
        self.evaluate_into(out, gradient)
        if out.is_cuda:
            self.torch.cuda.synchronize(self.device)
        return unpack(out.cpu().numpy(), self.n, gradient)

    def launch_count(self) -> int:
        return 0 if self.state is None else int(self.L.jme_launch_count(self.state.h))

    def close(self) -> None:
        """Collective when the peer exchange is open (a barrier: nobody frees a block that a peer's last pull still reads)."""
        self.close_peer_exchange()
        if self.state is not None:
            self.state.close()
            self.state = None
