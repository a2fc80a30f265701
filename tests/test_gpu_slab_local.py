"""z-slab decomposition with the halo exchange through peer memory (swb_sim_slab_export / swb_sim_slab_connect), driven from ONE process:
several slab simulations on the single GPU of the box, one host thread per slab.  The whole exchange protocol runs -- ghost planes left
to the neighbour, boundary planes stored by the step kernels into the neighbour's memory, stream-ordered flags keeping the slabs within
one step of each other, epochs across shots -- and the gathered seismograms must equal the undivided simulation bit for bit (the per-cell
arithmetic is identical).  The multi-GPU, one-process-per-GPU form (IPC handles) is tests/test_gpu_slab.py / bench.py --gpus N."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _problem(dtype, n, halo, freetop, nt=140):
    import swb200 as S

    T = np.dtype(dtype).type
    rng = np.random.default_rng(11)
    nx, ny, nz = n
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
    spos = np.array([[0.5 * ext[0], 0.45 * ext[1], 0.22 * ext[2]], [0.4 * ext[0], 0.55 * ext[1], 0.71 * ext[2]]], dtype=T)
    nrec = 15
    rpos = np.zeros((nrec, 3), dtype=T)
    rpos[:, 0] = np.linspace(0.15, 0.85, nrec) * ext[0]
    rpos[:, 1] = np.linspace(0.8, 0.2, nrec) * ext[1]
    rpos[:, 2] = np.linspace(0.04, 0.96, nrec) * ext[2]  # receivers in every slab, on boundary planes and inside the C-PML strips

    def shot(scale=1.0):
        return S.ScalarShot(srcs=S.ScalarSources(spos.copy(), (scale * tf).astype(T), T(f0)), recs=S.ScalarReceivers(rpos.copy(), nt, dtype=np.dtype(T)))

    return params, vp, shot


@pytest.mark.parametrize("dtype,fast,n,halo,freetop,nslabs", [(np.float32, False, (70, 52, 91), 6, True, 3), (np.float64, False, (64, 40, 61), 5, False, 2),
                                                              (np.float32, True, (140, 36, 75), 7, True, 4)])
def test_peer_memory_slabs_on_one_gpu_match_the_undivided_run_bitwise(dtype, fast, n, halo, freetop, nslabs):
    import swb200 as S
    from swb200.multigpu import SlabForwardLocal

    params, vp, shot = _problem(dtype, n, halo, freetop)
    rp = S.RunParameters(parall="B200", device=0, erroronPPW=False, fast_f32=fast)
    ref = []
    for scale in (1.0, 0.5):
        s = shot(scale)
        S.swforward(params, S.VpAcousticCDMaterialProperties(vp), [s], runparams=rp)
        ref.append(s.recs.seismograms.copy())
    assert np.all(np.max(np.abs(ref[0]), axis=0) > 0)
    group = SlabForwardLocal(params, vp, nslabs, runparams_for=lambda dev: S.RunParameters(parall="B200", device=dev, erroronPPW=False, fast_f32=fast))
    try:
        for k, scale in enumerate((1.0, 0.5)):  # the second shot exercises the epoch hand-over between shots
            got = group.forward(shot(scale))
            assert np.array_equal(got, ref[k]), (k, float(np.max(np.abs(got - ref[k]))))
    finally:
        group.close()
