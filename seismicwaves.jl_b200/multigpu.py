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

    def allgather_bytes(self, payload: bytes) -> list:
        """every rank's `payload` (equal lengths), in rank order"""
        import torch

        if self.world == 1:
            return [payload]
        on_gpu = self.dist.get_backend() == "nccl"
        t = torch.tensor(list(payload), dtype=torch.uint8)
        t = t.cuda(self.device) if on_gpu else t
        outs = [torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(outs, t)
        return [bytes(o.cpu().tolist()) for o in outs]

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


class _SlabSim:
    """One slab of a z-slab decomposition: the slab-local simulation (owned planes + one ghost plane per interior face) and the routing
    of sources / receivers to the slab that owns their plane."""

    def __init__(self, params, vp_local: np.ndarray, world: int, rank: int, device: int, runparams=None, vp_max_global: Optional[float] = None):
        from . import api
        from .types import InputParametersAcoustic, RunParameters, VpAcousticCDMaterialProperties

        assert len(params.gridsize) == 3, "z-slab decomposition is for 3D grids"
        self.params, self.world, self.rank = params, world, rank
        self.nz = params.gridsize[2]
        self.loc = slab_local_planes(self.nz, world, rank)
        self.own = slab_range(self.nz, world, rank)
        assert vp_local.shape == (params.gridsize[0], params.gridsize[1], len(self.loc)), "vp_local must hold this rank's planes (ghost planes included)"
        self.T = params.dtype.type
        lp = InputParametersAcoustic(params.ntimesteps, params.dt, (params.gridsize[0], params.gridsize[1], len(self.loc)), params.gridspacing, params.boundcond,
                                     dtype=params.dtype)
        rp = runparams or RunParameters(parall="B200", device=device)
        matprop = VpAcousticCDMaterialProperties(np.asfortranarray(vp_local))
        self.sim = api.build_wavesim(lp, matprop, runparams=rp)
        self.sim.set_wavesim_matprop(matprop)
        self.vmax_local = float(np.max(vp_local))
        if vp_max_global is not None:
            self.set_vmax(vp_max_global)
        self.lower, self.upper = (rank - 1 if rank > 0 else -1), (rank + 1 if rank < world - 1 else -1)

    def set_vmax(self, vmax: float) -> None:
        self.sim._vp_max = self.T(vmax)  # the C-PML profiles depend on the global maximum velocity (acou_init_bc.jl:13-17)

    def set_slab(self, comm) -> None:
        _lib.check(_lib.load().swb_sim_set_slab(self.sim._h, comm, self.lower, self.upper))

    def export_handle(self) -> "_lib.swb_slab_handle":
        h = _lib.swb_slab_handle()
        _lib.check(_lib.load().swb_sim_slab_export(self.sim._h, C.byref(h)))
        return h

    def connect(self, lower, upper) -> None:
        _lib.check(_lib.load().swb_sim_slab_connect(self.sim._h, None if lower is None else C.byref(lower), None if upper is None else C.byref(upper)))

    def local_shot(self, shot):
        """(slab-local shot, indices of the receivers this slab owns)"""
        from . import hostprep
        from .types import ScalarReceivers, ScalarShot, ScalarSources

        T = self.T
        spacing = self.params.gridspacing
        nt = self.params.ntimesteps
        gsrc = hostprep.find_nearest_grid_points(shot.srcs.positions, spacing, T) - 1  # 0-based global indices
        grec = hostprep.find_nearest_grid_points(shot.recs.positions, spacing, T) - 1
        isrc = slab_route_points(gsrc, self.nz, self.world, self.rank)
        irec = slab_route_points(grec, self.nz, self.world, self.rank)
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
        return lshot, irec, grec.shape[0]

    def run(self, lshot, irec, nrec_total, barrier=None) -> np.ndarray:
        """runs the local shot; returns the (nt, nrec_total) matrix holding this slab's traces (zeros elsewhere).
        `barrier` (slabs driven by threads of one process on one device): every slab binds its shot -- which may allocate, and
        cudaMalloc / cudaFree wait for the whole device -- before any slab parks its stream on a neighbour's flag."""
        try:
            self.sim.init_shot(lshot)
            if barrier is not None:
                self.sim._bind(lshot)
        finally:
            if barrier is not None:
                barrier.wait()
        self.sim.swforward_1shot(lshot)
        full = np.zeros((self.params.ntimesteps, nrec_total), dtype=self.params.dtype, order="F")
        if len(irec):
            full[:, irec] = lshot.recs.seismograms
        return full

    def close(self):
        self.sim.close()


class SlabForward3D:
    """swforward! of ONE 3D acoustic CD shot on a grid cut into z slabs, one per rank (forward only).

    Every rank passes the global parameters and only ITS slab of the velocity model (`vp_local`, planes
    slab_local_planes(...), ghost planes included).  Sources / receivers are given with global positions; each is
    bound on the rank that owns its plane (the others use a muted dummy so that the engine's shot contract holds),
    and the seismograms are summed over the ranks at the end (every trace is non-zero on exactly one rank).
    exchange = "p2p" (default): the step kernels store the boundary planes straight into the neighbours' ghost planes over NVLink
    (IPC-mapped peer memory, per-step flags; swb_sim_slab_export / swb_sim_slab_connect); "nccl": grouped ncclSend / ncclRecv of
    one plane per interior face after every step."""

    def __init__(self, params, vp_local: np.ndarray, sp: "ShotParallel", runparams=None, vp_max_global: Optional[float] = None, exchange: str = "p2p"):
        import os

        self.sp, self.params = sp, params
        self.slab = _SlabSim(params, vp_local, sp.world, sp.rank, sp.device or 0, runparams, vp_max_global)
        self.sim, self.loc, self.own, self.T, self.nz = self.slab.sim, self.slab.loc, self.slab.own, self.slab.T, self.slab.nz
        if vp_max_global is None:
            self.slab.set_vmax(sp.allreduce_max(self.slab.vmax_local) if sp.world > 1 else self.slab.vmax_local)
        exchange = os.environ.get("SWB_SLAB_EXCHANGE", exchange)
        assert exchange in ("p2p", "nccl")
        self.exchange = exchange if sp.world > 1 else "none"
        if sp.world > 1 and exchange == "nccl":
            sp._ensure_comm()
        self.slab.set_slab(sp._comm if exchange == "nccl" else None)
        if sp.world > 1 and exchange == "p2p":
            handles = sp.allgather_bytes(bytes(self.slab.export_handle()))
            hs = [_lib.swb_slab_handle.from_buffer_copy(b) for b in handles]
            self.slab.connect(hs[sp.rank - 1] if sp.rank > 0 else None, hs[sp.rank + 1] if sp.rank < sp.world - 1 else None)

    def forward(self, shot) -> np.ndarray:
        """runs the shot and returns the full seismogram matrix (nt, nrec) on every rank"""
        sp = self.sp
        lshot, irec, nrec = self.slab.local_shot(shot)
        full = self.slab.run(lshot, irec, nrec)
        if sp.world > 1:
            full = sp.allreduce_host({"s": full})[0]["s"]
        shot.recs.seismograms[...] = full
        return full

    def exchange_mode(self) -> str:
        return {"p2p": "boundary planes stored by the step kernels straight into the neighbours' ghost planes (IPC-mapped peer memory over NVLink), "
                       "stream-ordered 32-bit flags keep neighbouring slabs within one step; no collective in the time loop",
                "nccl": "grouped ncclSend/ncclRecv of one plane per interior face on the engine's stream after every step",
                "none": "none (single slab)"}[self.exchange]

    def close(self):
        self.slab.close()


class SlabForwardLocal:
    """The same decomposition driven from ONE process: `nslabs` slab simulations on the given devices (all on device 0 by default), one host
    thread per slab for the duration of a shot (a slab's forward call blocks until its neighbours have caught up), the halo exchange
    through peer memory only.  This is the shape of the Julia integration (one process, one task per GPU); on a single GPU it runs
    the whole exchange protocol -- ghost planes, peer stores, flags -- without needing a second device."""

    def __init__(self, params, vp: np.ndarray, nslabs: int, devices: Optional[Sequence[int]] = None, runparams_for=None):
        from .types import RunParameters

        self.params, self.n = params, nslabs
        nz = params.gridsize[2]
        devices = list(devices) if devices is not None else [0] * nslabs
        vmax = float(np.max(vp))
        self.slabs = []
        for r in range(nslabs):
            loc = slab_local_planes(nz, nslabs, r)
            rp = runparams_for(devices[r]) if runparams_for is not None else RunParameters(parall="B200", device=devices[r], erroronPPW=False)
            self.slabs.append(_SlabSim(params, np.asfortranarray(vp[:, :, loc.start:loc.stop]), nslabs, r, devices[r], rp, vmax))
        for s in self.slabs:
            s.set_slab(None)
        hs = [s.export_handle() for s in self.slabs] if nslabs > 1 else []
        for r, s in enumerate(self.slabs):
            if nslabs > 1:
                s.connect(hs[r - 1] if r > 0 else None, hs[r + 1] if r < nslabs - 1 else None)

    def forward(self, shot) -> np.ndarray:
        import threading

        prepared = [s.local_shot(shot) for s in self.slabs]
        out, err = [None] * self.n, [None] * self.n
        barrier = threading.Barrier(self.n)

        def work(r):
            try:
                out[r] = self.slabs[r].run(*prepared[r], barrier=barrier)
            except Exception as e:  # noqa: BLE001
                err[r] = e

        threads = [threading.Thread(target=work, args=(r,)) for r in range(self.n)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in err:
            if e is not None:
                raise e
        full = out[0]
        for o in out[1:]:
            full = full + o
        shot.recs.seismograms[...] = full
        return full

    def close(self):
        for s in self.slabs:
            s.close()
