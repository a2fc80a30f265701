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

    def allreduce_max(self, value: float) -> float:
        import torch

        if self.world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64)
        t = t.cuda(self.device) if self.dist.get_backend() == "nccl" else t
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

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


# ---------------------------------------------------------------------------------------------------------
# z-slab domain decomposition of one 3D acoustic constant-density forward simulation (BASELINE config 5)
# ---------------------------------------------------------------------------------------------------------
def slab_range(nz: int, world_size: int, rank: int) -> range:
    """owned planes (0-based, along the last axis) of rank `rank`: contiguous, the first nz % world_size ranks own one more"""
    base, rem = divmod(nz, world_size)
    k0 = rank * base + min(rank, rem)
    return range(k0, k0 + base + (1 if rank < rem else 0))


def slab_local_planes(nz: int, world_size: int, rank: int) -> range:
    """planes a rank stores: its owned ones plus one ghost plane per interior face"""
    own = slab_range(nz, world_size, rank)
    return range(own.start - (1 if rank > 0 else 0), own.stop + (1 if rank < world_size - 1 else 0))


def slab_route_points(idx0: np.ndarray, nz: int, world_size: int, rank: int) -> np.ndarray:
    """indices (into the point list) of the points whose 0-based plane idx0[:, -1] this rank owns"""
    own = slab_range(nz, world_size, rank)
    k = idx0[:, -1]
    return np.nonzero((k >= own.start) & (k < own.stop))[0]


class SlabForward3D:
    """swforward! of ONE 3D acoustic CD shot on a grid cut into z slabs, one per rank (forward only).

    Every rank passes the global parameters and only ITS slab of the velocity model (`vp_local`, planes
    slab_local_planes(...), ghost planes included).  Sources / receivers are given with global positions; each is
    bound on the rank that owns its plane (the others use a muted dummy so that the engine's shot contract holds),
    and the seismograms are summed over the ranks at the end (every trace is non-zero on exactly one rank).
    The local sims exchange one plane per interior face and time step over NCCL (swb_sim_set_slab)."""

    def __init__(self, params, vp_local: np.ndarray, sp: "ShotParallel", runparams=None, vp_max_global: Optional[float] = None):
        from . import api
        from .types import InputParametersAcoustic, RunParameters, VpAcousticCDMaterialProperties

        assert len(params.gridsize) == 3, "z-slab decomposition is for 3D grids"
        self.sp, self.params = sp, params
        self.nz = params.gridsize[2]
        self.loc = slab_local_planes(self.nz, sp.world, sp.rank)
        self.own = slab_range(self.nz, sp.world, sp.rank)
        assert vp_local.shape == (params.gridsize[0], params.gridsize[1], len(self.loc)), "vp_local must hold this rank's planes (ghost planes included)"
        T = params.dtype.type
        lp = InputParametersAcoustic(params.ntimesteps, params.dt, (params.gridsize[0], params.gridsize[1], len(self.loc)), params.gridspacing, params.boundcond,
                                     dtype=params.dtype)
        rp = runparams or RunParameters(parall="B200", device=sp.device or 0)
        matprop = VpAcousticCDMaterialProperties(np.asfortranarray(vp_local))
        self.sim = api.build_wavesim(lp, matprop, runparams=rp)
        self.sim.set_wavesim_matprop(matprop)
        vmax = float(np.max(vp_local)) if vp_max_global is None else float(vp_max_global)
        if vp_max_global is None and sp.world > 1:
            vmax = sp.allreduce_max(vmax)
        self.sim._vp_max = T(vmax)  # the C-PML profiles depend on the global maximum velocity (acou_init_bc.jl:13-17)
        lower, upper = (sp.rank - 1 if sp.rank > 0 else -1), (sp.rank + 1 if sp.rank < sp.world - 1 else -1)
        if sp.world > 1:
            sp._ensure_comm()
        _lib.check(_lib.load().swb_sim_set_slab(self.sim._h, sp._comm, lower, upper))
        self.T = T

    def forward(self, shot) -> np.ndarray:
        """runs the shot and returns the full seismogram matrix (nt, nrec) on every rank"""
        from . import hostprep
        from .types import ScalarReceivers, ScalarShot, ScalarSources

        T, sim, sp = self.T, self.sim, self.sp
        spacing = self.params.gridspacing
        nt = self.params.ntimesteps
        gsrc = hostprep.find_nearest_grid_points(shot.srcs.positions, spacing, T) - 1  # 0-based global indices
        grec = hostprep.find_nearest_grid_points(shot.recs.positions, spacing, T) - 1
        isrc = slab_route_points(gsrc, self.nz, sp.world, sp.rank)
        irec = slab_route_points(grec, self.nz, sp.world, sp.rank)
        koff = self.loc.start

        def local_positions(gidx, sel):
            if len(sel) == 0:  # dummy point in the middle of the first owned plane (source: zero wavelet; receiver: trace discarded)
                g = np.array([[self.params.gridsize[0] // 2, self.params.gridsize[1] // 2, self.own.start]])
            else:
                g = gidx[sel]
            loc = g.astype(np.float64)
            loc[:, 2] -= koff
            return np.asfortranarray((loc * np.asarray(spacing, dtype=np.float64)[None, :]).astype(T))

        tf = np.asfortranarray(shot.srcs.tf[:, isrc].astype(T)) if len(isrc) else np.zeros((nt, 1), dtype=T, order="F")
        lshot = ScalarShot(srcs=ScalarSources(local_positions(gsrc, isrc), tf, shot.srcs.domfreq),
                           recs=ScalarReceivers(local_positions(grec, irec), nt, dtype=self.params.dtype))
        sim.init_shot(lshot)
        sim.swforward_1shot(lshot)
        full = np.zeros((nt, grec.shape[0]), dtype=self.params.dtype, order="F")
        if len(irec):
            full[:, irec] = lshot.recs.seismograms
        if sp.world > 1:
            full = sp.allreduce_host({"s": full})[0]["s"]
        shot.recs.seismograms[...] = full
        return full

    def exchange_mode(self) -> str:
        return "grouped ncclSend/ncclRecv of one plane per interior face on the engine's stream after every step" if self.sp.world > 1 else "none (single slab)"

    def close(self):
        self.sim.close()
