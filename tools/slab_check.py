"""z-slab decomposition helpers shared by bench.py (--gpus N > 1), tests/slab_worker.py and tools/bench_slab.py; every function is
called by all ranks of an initialised torch.distributed (nccl) group, one rank per GPU.

bitwise_twin : decomposed forward runs of small seeded 3D grids against the same shots on the single-GPU engine (rank 0): the per-cell
               arithmetic is identical, so the gathered traces must agree bit for bit.
throughput   : device-timed forward run of a large layered model, each rank building only its own slab."""
from __future__ import annotations

import math

import numpy as np

TWINS = [(np.float32, False, (70, 52, 90), 6, True), (np.float64, False, (64, 40, 61), 5, False), (np.float32, True, (140, 36, 75), 7, True)]


def bitwise_twin(S, torch, dist, world, rank, local, twins=TWINS, nt=160, exchange="p2p"):
    from swb200.multigpu import ShotParallel, SlabForward3D, slab_local_planes

    ok, detail = True, []
    nt0 = nt
    for dtype, fast, n, halo, freetop in twins:
        T = np.dtype(dtype).type
        rng = np.random.default_rng(5)
        nz0 = n[2]
        n = (n[0], n[1], max(n[2], (2 * halo + 6) * world))  # every slab-local grid must still hold the reference's minimum of 2 halo + 3 planes
        nx, ny, nz = n
        nt = int(round(nt0 * nz / nz0))  # ... and the wavefield still has to reach every receiver
        h = 10.0
        vp = 1800.0 + 1500.0 * (np.arange(nz) / (nz - 1))[None, None, :] + rng.normal(0, 30.0, size=n)
        vp = np.asfortranarray(vp.astype(T))
        dt = 0.9 * h / (float(vp.max()) * np.sqrt(3.0))
        bc = S.CPMLBoundaryConditionParameters(halo=halo, rcoef=T(1e-4), freeboundtop=freetop)
        params = S.InputParametersAcoustic(nt, T(dt), n, (T(h),) * 3, bc, dtype=np.dtype(T))
        t = np.arange(nt) * dt
        f0 = 10.0
        tf = np.zeros((nt, 2), dtype=T, order="F")
        tf[:, 0] = 1000.0 * S.rickerstf(t, 1.2 / f0, f0)
        tf[:, 1] = 700.0 * S.rickerstf(t, 1.3 / f0, f0)
        ext = [(n[d] - 1) * h for d in range(3)]
        spos = np.array([[0.5 * ext[0], 0.45 * ext[1], 0.30 * ext[2]], [0.4 * ext[0], 0.55 * ext[1], 0.72 * ext[2]]], dtype=T)  # one source per half
        nrec = 14
        rpos = np.zeros((nrec, 3), dtype=T)
        rpos[:, 0] = np.linspace(0.15, 0.85, nrec) * ext[0]
        rpos[:, 1] = np.linspace(0.8, 0.2, nrec) * ext[1]
        rpos[:, 2] = np.linspace(0.05, 0.95, nrec) * ext[2]  # receivers in every slab, some inside the C-PML strips

        def shot():
            return S.ScalarShot(srcs=S.ScalarSources(spos.copy(), tf.copy(), T(f0)), recs=S.ScalarReceivers(rpos.copy(), nt, dtype=np.dtype(T)))

        sp = ShotParallel(device=local)
        rp = S.RunParameters(parall="B200", device=local, erroronPPW=False, fast_f32=fast)
        loc = slab_local_planes(nz, world, rank)
        slab = SlabForward3D(params, np.asfortranarray(vp[:, :, loc.start:loc.stop]), sp, runparams=rp, exchange=exchange)
        got = slab.forward(shot())
        got2 = slab.forward(shot())  # a second shot on the same sims: replays whatever the first one set up
        slab.close()
        if rank == 0:
            ref_shot = shot()
            S.swforward(params, S.VpAcousticCDMaterialProperties(vp), [ref_shot], runparams=rp)
            ref = ref_shot.recs.seismograms
            same = bool(np.array_equal(got, ref) and np.array_equal(got2, ref))
            live = int(np.count_nonzero(np.max(np.abs(ref), axis=0)))
            detail.append({"exchange": exchange, "dtype": np.dtype(dtype).name, "fast_f32": fast, "grid": list(n), "nt": nt, "bitwise_equal": same, "live_traces": f"{live}/{nrec}"})
            ok = ok and same and live == nrec
        sp.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    return bool(int(flag.item()) == 1), detail


def throughput(S, torch, dist, world, rank, local, grid=(2048, 2048, 1024), nt=100, halo=20, fast_f32=True, reps=2, exchange="p2p"):
    from swb200.multigpu import ShotParallel, SlabForward3D, slab_local_planes

    T = np.float32
    nx, ny, nz = grid
    h = 10.0
    loc = slab_local_planes(nz, world, rank)
    depth = (np.arange(loc.start, loc.stop, dtype=np.float64) / (nz - 1)).astype(T)
    vp = np.empty((nx, ny, len(loc)), dtype=T, order="F")
    vp[...] = (1500.0 + 3000.0 * depth)[None, None, :]
    vmax = 4500.0
    dt = 0.99 * h / (vmax * math.sqrt(3.0))
    bc = S.CPMLBoundaryConditionParameters(halo=halo, rcoef=T(1e-4), freeboundtop=True)
    params = S.InputParametersAcoustic(nt, T(dt), (nx, ny, nz), (T(h),) * 3, bc, dtype=np.dtype(T))
    sp = ShotParallel(device=local)
    rp = S.RunParameters(parall="B200", device=local, erroronPPW=False, fast_f32=bool(fast_f32))
    slab = SlabForward3D(params, vp, sp, runparams=rp, vp_max_global=vmax, exchange=exchange)
    f0 = 8.0
    t = np.arange(nt) * dt
    tf = np.asfortranarray((1000.0 * S.rickerstf(t, 1.2 / f0, f0)).astype(T).reshape(nt, 1))
    ext = [(grid[d] - 1) * h for d in range(3)]
    spos = np.array([[0.5 * ext[0], 0.5 * ext[1], 2 * h]], dtype=T)
    nrec = 64
    rpos = np.zeros((nrec, 3), dtype=T)
    rpos[:, 0] = np.linspace(0.1, 0.9, nrec) * ext[0]
    rpos[:, 1] = 0.5 * ext[1]
    rpos[:, 2] = 3 * h

    def shot():
        return S.ScalarShot(srcs=S.ScalarSources(spos.copy(), tf.copy(), T(f0)), recs=S.ScalarReceivers(rpos.copy(), nt, dtype=np.dtype(T)))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream_ptr = S._lib.C.c_void_p()
    S._lib.check(S._lib.load().swb_sim_stream(slab.sim._h, S._lib.C.byref(stream_ptr)))
    ext_stream = torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", local))
    best = None
    for rep in range(reps + 1):
        s = shot()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext_stream)
        slab.forward(s)
        e1.record(ext_stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if rep > 0:
            best = ms if best is None else min(best, ms)
    tt = torch.tensor([best], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    sec = float(tt.item()) * 1e-3
    cells = float(nx) * ny * nz
    out = {"workload": f"C5 family: 3D acoustic CD {nx}x{ny}x{nz} Float32 forward-only, z slabs over {world} GPU(s), nt={nt}, halo {halo}, free surface", "n_gpus": world,
           "seconds": sec, "ms_per_step": 1e3 * sec / nt, "value": cells * nt / sec / 1e9, "unit": "Gcell-updates/s", "per_gpu_Gcell_per_s": cells * nt / sec / 1e9 / world,
           "local_planes": len(loc), "halo_exchange_bytes_per_step_per_face": nx * ny * 4 * 2, "device_GB_per_gpu": slab.sim.device_bytes() / 1e9,
           "exchange": slab.exchange_mode(), "algorithmic_GBps_per_gpu": 16.0 * cells * nt / sec / 1e9 / world,
           "timing": f"CUDA events on the engine's stream around the whole forward call (host binding, nt steps with halo exchange, seismogram download), max over ranks, best of {reps}"}
    slab.close()
    sp.close()
    return out
