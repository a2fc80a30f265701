#!/usr/bin/env python
"""tools/bench_slab.py -- forward-only throughput of the z-slab decomposition (BASELINE config 5 family), one rank per GPU:

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_slab.py --grid 2048 2048 1024 --nt 50

Each rank builds only its slab of the layered velocity model.  Rank 0 prints one JSON line: whole-job Gcell-updates/s
(global cells x nt / max-over-ranks time of the timed forward run)."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, nargs=3, default=[2048, 2048, 1024])
    ap.add_argument("--nt", type=int, default=50)
    ap.add_argument("--halo", type=int, default=20)
    ap.add_argument("--fast-f32", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist

    import swb200 as S
    from swb200.multigpu import ShotParallel, SlabForward3D, slab_local_planes

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if "RANK" in os.environ:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    T = np.float32
    nx, ny, nz = args.grid
    h = 10.0
    loc = slab_local_planes(nz, world, rank)
    depth = (np.arange(loc.start, loc.stop, dtype=np.float64) / (nz - 1)).astype(T)
    vp = np.empty((nx, ny, len(loc)), dtype=T, order="F")
    vp[...] = (1500.0 + 3000.0 * depth)[None, None, :]
    vmax = 4500.0
    dt = 0.99 * h / (vmax * math.sqrt(3.0))
    nt = args.nt
    bc = S.CPMLBoundaryConditionParameters(halo=args.halo, rcoef=T(1e-4), freeboundtop=True)
    params = S.InputParametersAcoustic(nt, T(dt), (nx, ny, nz), (T(h),) * 3, bc, dtype=np.dtype(T))
    sp = ShotParallel(device=local)
    rp = S.RunParameters(parall="B200", device=local, erroronPPW=False, fast_f32=bool(args.fast_f32))
    slab = SlabForward3D(params, vp, sp, runparams=rp, vp_max_global=vmax)
    f0 = 8.0
    t = np.arange(nt) * dt
    tf = np.asfortranarray((1000.0 * S.rickerstf(t, 1.2 / f0, f0)).astype(T).reshape(nt, 1))
    ext = [(args.grid[d] - 1) * h for d in range(3)]
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

    times = []
    for rep in range(3):
        s = shot()
        barrier()
        t0 = time.perf_counter()
        slab.forward(s)
        barrier()
        times.append(time.perf_counter() - t0)
    tt = torch.tensor([min(times[1:])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(tt.item())
        cells = float(nx) * ny * nz
        print(json.dumps({"workload": f"3D acoustic CD {nx}x{ny}x{nz} f32 forward-only, z slabs over {world} GPU(s), nt={nt}", "n_gpus": world,
                          "seconds": sec, "ms_per_step": 1e3 * sec / nt, "Gcell_per_s": cells * nt / sec / 1e9, "per_gpu_Gcell_per_s": cells * nt / sec / 1e9 / world,
                          "local_planes": len(loc), "halo_exchange_bytes_per_step_per_face": nx * ny * 4 * 2, "device_GB": slab.sim.device_bytes() / 1e9,
                          "timing": "wall clock of the whole forward call (host binding, nt steps with NCCL halo exchange, seismogram gather), max over ranks, best of 2"}),
              flush=True)
    slab.close()
    sp.close()
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
