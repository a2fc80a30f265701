"""Shot-parallel gradient over the GPUs of one box: one process per GPU (torch.distributed for the plumbing), shots split
into contiguous groups exactly like the reference's distribsrcs (src/utils/utils.jl:28-45), per-rank total gradients summed
with one NCCL all-reduce per gradient array on the engine's own stream (swb_sim_allreduce_total_gradient), misfits summed
alongside.  The reference has no counterpart: its shot loop is sequential (src/apis/gradient.jl:115-134) and its
`:threadpersrc` mode is broken.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .hostprep import distribsrcs


def shot_group(nshots: int, world_size: int, rank: int) -> range:
    """the shots rank `rank` owns: distribsrcs(nshots, world_size)[rank] (empty if there are fewer shots than ranks)"""
    groups = distribsrcs(nshots, world_size)
    return groups[rank] if rank < len(groups) else range(0)


class ShotParallel:
    """Process-group wrapper.  `torch.distributed` must already be initialised (backend nccl on GPUs, gloo in CPU tests)."""

    def __init__(self, device: Optional[int] = None):
        import torch.distributed as dist

        self.dist = dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.device = device
        self._comm = None

    # ---- host-side reduction (any backend; used for the misfit and by the CPU tests) -----------------------------
    def allreduce_host(self, grads: Dict[str, np.ndarray], misfit: float = 0.0) -> Tuple[Dict[str, np.ndarray], float]:
        import torch

        if self.world == 1:
            return grads, misfit
        on_gpu = self.dist.get_backend() == "nccl"
        out = {}
        for k in sorted(grads):
            t = torch.from_numpy(np.ascontiguousarray(grads[k]))
            t = t.cuda(self.device) if on_gpu else t
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            out[k] = np.asfortranarray(t.cpu().numpy().reshape(grads[k].shape))
        m = torch.tensor([misfit], dtype=torch.float64)
        m = m.cuda(self.device) if on_gpu else m
        self.dist.all_reduce(m, op=self.dist.ReduceOp.SUM)
        return out, float(m.item())

    # ---- device-side reduction of the engine's total gradient (NCCL through libswb200) ---------------------------
    def _ensure_comm(self):
        import torch

        if self._comm is not None or self.world == 1:
            return
        lib = _lib.load()
        raw = (C.c_ubyte * 128)()
        if self.rank == 0:
            _lib.check(lib.swb_comm_unique_id(raw))
        idbuf = torch.tensor(list(raw), dtype=torch.uint8)
        if self.dist.get_backend() == "nccl":
            idbuf = idbuf.cuda(self.device)
        self.dist.broadcast(idbuf, 0)
        raw = (C.c_ubyte * 128)(*idbuf.cpu().tolist())
        comm = C.c_void_p()
        _lib.check(lib.swb_comm_create(raw, self.world, self.rank, self.device or 0, C.byref(comm)))
        self._comm = comm

    def allreduce_total_gradient(self, wavesim) -> None:
        if self.world == 1:
            return
        self._ensure_comm()
        _lib.check(_lib.load().swb_sim_allreduce_total_gradient(wavesim._h, self._comm))

    def close(self) -> None:
        if self._comm is not None:
            _lib.load().swb_comm_destroy(self._comm)
            self._comm = None


def swgradient_sharded(wavesim, matprop, shots: Sequence, misfit: Sequence, sp: ShotParallel):
    """swgradient! over all shots with the shot loop sharded across the ranks of `sp`; every rank returns the full result."""
    from .api import _run_swgradient

    group = shot_group(len(shots), sp.world, sp.rank)
    local_misfit = _run_swgradient(wavesim, matprop, shots, misfit, shot_indices=list(group))
    sp.allreduce_total_gradient(wavesim)
    grad = wavesim.get_total_gradient()
    if wavesim.gradparams.compute_misfit:
        _, tot = sp.allreduce_host({}, float(local_misfit))
        return grad, wavesim.T(tot)
    return grad
