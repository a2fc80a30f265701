"""Worker of the z-slab decomposition test: launched with torch.distributed.run, one rank per GPU.

Every rank builds the same seeded global model, keeps its slab, runs the decomposed forward simulation
(SlabForward3D, NCCL halo exchange) and rank 0 compares the gathered seismograms with the same shot computed on a
single GPU by the ordinary engine: the per-cell arithmetic is identical, so the traces must agree bit for bit."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import swb200 as S
    from swb200.multigpu import ShotParallel, SlabForward3D, slab_local_planes

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for dtype, fast, n, halo, freetop in [(np.float32, False, (70, 52, 90), 6, True), (np.float64, False, (64, 40, 61), 5, False), (np.float32, True, (140, 36, 75), 7, True)]:
        T = np.dtype(dtype).type
        rng = np.random.default_rng(5)
        nx, ny, nz = n
        h = 10.0
        vp = 1800.0 + 1500.0 * (np.arange(nz) / (nz - 1))[None, None, :] + rng.normal(0, 30.0, size=n)
        vp = np.asfortranarray(vp.astype(T))
        dt = 0.9 * h / (float(vp.max()) * np.sqrt(3.0))
        nt = 160
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
        slab = SlabForward3D(params, np.asfortranarray(vp[:, :, loc.start:loc.stop]), sp, runparams=rp)
        got = slab.forward(shot())
        slab.close()
        if rank == 0:
            ref_shot = shot()
            S.swforward(params, S.VpAcousticCDMaterialProperties(vp), [ref_shot], runparams=rp)
            ref = ref_shot.recs.seismograms
            same = np.array_equal(got, ref)
            nz_tr = int(np.count_nonzero(np.max(np.abs(ref), axis=0)))
            print(f"slab test {np.dtype(dtype).name} fast={fast} n={n} world={world}: bitwise equal = {same}, live traces = {nz_tr}/{nrec}", flush=True)
            ok = ok and same and nz_tr == nrec
        sp.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
