"""GPU parity tests for the acoustic hot path: CUDA engine (through the C ABI) vs the CPU oracle on identical
seeded inputs.  Tolerances are the north-star ones (rel-L2 <= 1e-5 Float64, 1e-4 Float32); the
reference-faithful arithmetic paths are additionally expected to agree to ~1 ulp per step, which the
tight bounds below pin."""
import numpy as np
import pytest

from cases import (acoustic_case, make_observed, oracle_forward, oracle_gradient, product_inputs, rel_l2, tol)

pytestmark = pytest.mark.gpu


def _forward_product(case, **kw):
    import swb200 as S

    params, matprop, shots, _, runparams, _ = product_inputs(case, **kw)
    snaps = S.swforward(params, matprop, shots, runparams=runparams)
    return [s.recs.seismograms for s in shots], snaps


def _gradient_product(case, observed, **kw):
    import swb200 as S

    params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=observed, **kw)
    res = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    return res, [s.recs.seismograms for s in shots]


FWD_CASES = [
    ("acoustic_cd", (64, 56), np.float64, True),
    ("acoustic_cd", (64, 56), np.float64, False),
    ("acoustic_cd", (61, 53), np.float32, True),
    ("acoustic_cd", (30, 26, 28), np.float64, True),
    ("acoustic_cd", (30, 26, 28), np.float32, False),
    ("acoustic_vd", (64, 56), np.float64, True),
    ("acoustic_vd", (64, 56), np.float64, False),
    ("acoustic_vd", (67, 59), np.float32, True),
]


@pytest.mark.parametrize("kind,n,dtype,freetop", FWD_CASES)
@pytest.mark.parametrize("fused", [False, True])
def test_forward_seismograms_match_oracle(kind, n, dtype, freetop, fused):
    halo = 6 if len(n) == 3 else 8
    case = acoustic_case(kind=kind, n=n, nt=150 if len(n) == 2 else 60, halo=halo, freetop=freetop, dtype=dtype, seed=len(n) * 7 + int(freetop))
    ref, _ = oracle_forward(case)
    got, _ = _forward_product(case, fused=fused)
    for r, g in zip(ref, got):
        assert g.dtype == np.dtype(dtype)
        assert np.max(np.abs(r)) > 0
        err = rel_l2(g, r)
        assert err <= tol(dtype), err
        # reference-faithful arithmetic: expect agreement to rounding level, far below the tolerance
        assert err <= (1e-12 if dtype == np.float64 else 1e-6), err


def test_forward_fast_f32_within_tolerance():
    case = acoustic_case(kind="acoustic_vd", n=(96, 80), nt=300, halo=10, dtype=np.float32, seed=5)
    ref, _ = oracle_forward(case)
    got, _ = _forward_product(case, fast_f32=True)
    for r, g in zip(ref, got):
        assert rel_l2(g, r) <= tol(np.float32)
    case = acoustic_case(kind="acoustic_cd", n=(96, 80), nt=300, halo=10, dtype=np.float32, seed=6)
    ref, _ = oracle_forward(case)
    got, _ = _forward_product(case, fast_f32=True)
    for r, g in zip(ref, got):
        assert rel_l2(g, r) <= tol(np.float32)


GRAD_CASES = [
    ("acoustic_cd", (48, 44), np.float64, 1),
    ("acoustic_cd", (48, 44), np.float64, 7),
    ("acoustic_cd", (48, 44), np.float64, 10),   # nt % check_freq == 0
    ("acoustic_cd", (48, 44), np.float32, 9),
    ("acoustic_cd", (24, 22, 26), np.float64, 6),
    ("acoustic_vd", (48, 44), np.float64, 1),
    ("acoustic_vd", (48, 44), np.float64, 7),
    ("acoustic_vd", (48, 44), np.float64, 10),
    ("acoustic_vd", (51, 47), np.float32, 9),
]


@pytest.mark.parametrize("kind,n,dtype,check_freq", GRAD_CASES)
@pytest.mark.parametrize("fused", [False, True])
def test_gradient_and_misfit_match_oracle(kind, n, dtype, check_freq, fused):
    halo = 5 if len(n) == 3 else 6
    nt = 100 if len(n) == 2 else 50
    case = acoustic_case(kind=kind, n=n, nt=nt, halo=halo, dtype=dtype, seed=11 + check_freq)
    syn, _ = oracle_forward(case)
    observed = make_observed(case, syn)
    (gref, mref), sref, _ = oracle_gradient(case, observed, check_freq=check_freq, mute_src=3, mute_rec=2)
    (ggot, mgot), sgot = _gradient_product(case, observed, check_freq=check_freq, mute_src=3, mute_rec=2, fused=fused)
    assert set(ggot) == set(gref)
    for k in gref:
        assert np.max(np.abs(gref[k])) > 0
        err = rel_l2(ggot[k], gref[k])
        assert err <= tol(dtype), (k, err)
        assert err <= (1e-11 if dtype == np.float64 else 2e-5), (k, err)
    assert abs(float(mgot) - float(mref)) <= tol(dtype) * abs(float(mref))
    for r, g in zip(sref, sgot):
        assert rel_l2(g, r) <= tol(dtype)


@pytest.mark.parametrize("kind", ["acoustic_cd", "acoustic_vd"])
def test_gradient_host_misfit_path_windows_and_diag_invcov(kind):
    """pluggable-misfit path (forward -> host adjoint source -> adjoint) with windows and a diagonal covariance"""
    case = acoustic_case(kind=kind, n=(48, 44), nt=100, halo=6, dtype=np.float64, seed=3, windows=True, diag_invcov=True)
    syn, _ = oracle_forward(case)
    observed = make_observed(case, syn)
    (gref, mref), _, _ = oracle_gradient(case, observed, check_freq=8)
    (ggot, mgot), _ = _gradient_product(case, observed, check_freq=8)
    for k in gref:
        assert rel_l2(ggot[k], gref[k]) <= 1e-11
    assert abs(float(mgot) - float(mref)) <= 1e-10 * abs(float(mref))


def test_vd_harmonic_interpolation_gradient():
    case = acoustic_case(kind="acoustic_vd", n=(48, 44), nt=90, halo=6, dtype=np.float64, seed=21)
    syn, _ = oracle_forward(case)
    observed = make_observed(case, syn)
    (gref, _), _, _ = oracle_gradient(case, observed, check_freq=9, interp_method="harmonic")
    (ggot, _), _ = _gradient_product(case, observed, check_freq=9, interp_method="harmonic")
    for k in gref:
        assert rel_l2(ggot[k], gref[k]) <= 1e-10, k


def test_checkpointed_equals_non_checkpointed_gradient():
    """reference test: test/test_gradient_acoustic_constant_density.jl:194-312 (check_freq = floor(sqrt(nt)))"""
    case = acoustic_case(kind="acoustic_cd", n=(64, 60), nt=144, halo=8, dtype=np.float64, seed=2)
    syn, _ = oracle_forward(case)
    observed = make_observed(case, syn)
    (g1, m1), _ = _gradient_product(case, observed, check_freq=1)
    (g2, m2), _ = _gradient_product(case, observed, check_freq=12)
    assert np.array_equal(g1["vp"], g2["vp"])  # re-forwarding is deterministic: bit-identical
    assert m1 == m2


def test_snapshots_match_oracle():
    case = acoustic_case(kind="acoustic_cd", n=(48, 44), nt=60, halo=6, dtype=np.float64, seed=4, nshots=1)
    _, snaps_ref = oracle_forward(case, snapevery=20)
    _, snaps = _forward_product(case, snapevery=20)
    assert sorted(snaps[0].keys()) == [20, 40, 60]
    for it in (20, 40, 60):
        assert rel_l2(snaps[0][it]["pcur"], snaps_ref[0][it]) <= 1e-12


def test_example_c1_full_size():
    """BASELINE config 1: examples/simple_example_acoustic.jl verbatim (300x280, F64, nt=1500, 3 shots, free top,
    mute 5/2, check_freq=1) -- seismograms, misfit and gradient against the oracle."""
    import swb200 as S
    from oracle import oracle as O

    nt, dt, nx, nz, dh = 1500, 0.001, 300, 280, 8.0
    velmod = np.zeros((nx, nz), order="F")
    velmod[:, :] = 2000.0 + 12.0 * np.arange(nz)[None, :]
    t = np.arange(nt) * dt
    ixsrc = np.round(np.linspace(32, nx - 31, 3)).astype(int)
    ixrec = np.round(np.linspace(30, nx - 29, 10)).astype(int)
    f0 = 12.0
    stf = (1000.0 * O.rickerstf(t, 1.20 / f0, f0)).reshape(nt, 1)
    posrecs = np.zeros((10, 2))
    posrecs[:, 0] = (ixrec - 1) * dh
    posrecs[:, 1] = 2 * dh

    def mkshots_product():
        out = []
        for i in range(3):
            ps = np.array([[(ixsrc[i] - 1) * dh, (nz - 40) * dh]])
            out.append(S.ScalarShot(srcs=S.ScalarSources(ps, stf.copy(), f0), recs=S.ScalarReceivers(posrecs.copy(), nt)))
        return out

    def mkshots_oracle():
        return [O.ScalarShot(src_positions=np.array([[(ixsrc[i] - 1) * dh, (nz - 40) * dh]]), src_tf=np.asfortranarray(stf.copy()), domfreq=f0,
                             rec_positions=posrecs.copy()) for i in range(3)]

    bc = S.CPMLBoundaryConditionParameters(halo=20, rcoef=0.0001, freeboundtop=True)
    params = S.InputParametersAcoustic(nt, dt, (nx, nz), (dh, dh), bc)
    rp = S.RunParameters(parall="B200")
    shots = mkshots_product()
    S.swforward(params, S.VpAcousticCDMaterialProperties(velmod), shots, runparams=rp)
    oparams = O.Params(nt=nt, dt=dt, gridsize=(nx, nz), spacing=(dh, dh), halo=20, rcoef=0.0001, freetop=True)
    oshots = mkshots_oracle()
    O.swforward(O.build_wavesim("acoustic_cd", oparams), [velmod], oshots)
    for a, b in zip(shots, oshots):
        assert rel_l2(a.recs.seismograms, b.seismograms) <= 1e-12
    newvel = velmod - 0.2
    newvel[29:40, 32:44] *= 0.9
    observed = [s.recs.seismograms.copy(order="F") for s in shots]
    gp = S.GradParameters(mute_radius_src=5, mute_radius_rec=2, compute_misfit=True)
    grad, mis = S.swgradient(params, S.VpAcousticCDMaterialProperties(newvel), shots, [S.L2Misfit(observed=o) for o in observed], runparams=rp, gradparams=gp)
    osim = O.build_wavesim("acoustic_cd", oparams, gradient=True, check_freq=1)
    ograd, omis = O.swgradient(osim, [np.asfortranarray(newvel)], oshots, [O.L2Misfit(observed=o) for o in observed], mute_radius_src=5, mute_radius_rec=2,
                               compute_misfit=True)
    assert rel_l2(grad["vp"], ograd["vp"]) <= 1e-10
    assert abs(mis - omis) <= 1e-10 * abs(omis)


# ---- fused engine on grids that span many tiles (interior fast-path tiles + C-PML / edge tiles) ----------------


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,freetop", [((400, 150), True), ((261, 203), False)])
def test_vd_fused_multi_tile_forward_and_gradient(dtype, n, freetop):
    case = acoustic_case(kind="acoustic_vd", n=n, nt=160, halo=9, freetop=freetop, dtype=dtype, seed=31, nshots=2, nsrc=2, nrec=24)
    ref, _ = oracle_forward(case)
    got, _ = _forward_product(case, fused=True)
    for r, g in zip(ref, got):
        assert np.max(np.abs(r)) > 0
        assert rel_l2(g, r) <= (1e-12 if dtype == np.float64 else 1e-6)
    observed = make_observed(case, ref)
    for cf in (13, 1):
        (gref, mref), _, _ = oracle_gradient(case, observed, check_freq=cf, mute_src=3, mute_rec=1)
        (ggot, mgot), _ = _gradient_product(case, observed, check_freq=cf, mute_src=3, mute_rec=1, fused=True)
        for k in gref:
            assert np.max(np.abs(gref[k])) > 0
            assert rel_l2(ggot[k], gref[k]) <= (1e-11 if dtype == np.float64 else 2e-5), (k, cf)
        assert abs(float(mgot) - float(mref)) <= tol(dtype) * abs(float(mref))


def test_vd_fused_equals_unfused_bitwise():
    """the fused single-launch step and the one-launch-per-reference-kernel path perform the same operations"""
    case = acoustic_case(kind="acoustic_vd", n=(300, 170), nt=120, halo=10, dtype=np.float32, seed=8, nshots=1, nrec=16)
    a, _ = _forward_product(case, fused=True)
    b, _ = _forward_product(case, fused=False)
    assert np.array_equal(a[0], b[0])


@pytest.mark.parametrize("dtype,n,freetop,halo", [(np.float32, (1100, 610), True, 20), (np.float64, (777, 421), False, 20), (np.float32, (515, 300), False, 12)])
def test_vd_fused_equals_unfused_bitwise_production_halo(dtype, n, freetop, halo):
    """many interior tiles, full-width C-PML strips, partial last tiles: TMA staging and the plain / special chunk split of the
    edge tiles reproduce the one-launch-per-reference-kernel path bit for bit (seismograms, both gradients, misfit)"""
    case = acoustic_case(kind="acoustic_vd", n=n, nt=90, halo=halo, freetop=freetop, dtype=dtype, seed=37, nshots=1, nsrc=2, nrec=20)
    ext_x = (n[0] - 1) * case["h"]
    for sh in case["shots"]:  # sources and receivers close together (and receivers also inside the strips), so that the wave arrives within nt steps
        sh["src_positions"][:, 0] = np.array([0.42, 0.61])[: sh["src_positions"].shape[0]] * ext_x
        sh["src_positions"][:, 1] = 33.0 * case["h"]
        sh["rec_positions"][:, 0] = np.linspace(0.02, 0.98, sh["rec_positions"].shape[0]) * ext_x
        sh["rec_positions"][:, 1] = 29.0 * case["h"]
    a, _ = _forward_product(case, fused=True)
    b, _ = _forward_product(case, fused=False)
    assert np.max(np.abs(a[0])) > 0
    assert np.array_equal(a[0], b[0])
    observed = make_observed(case, a)
    (ga, ma), _ = _gradient_product(case, observed, check_freq=9, fused=True)
    (gb, mb), _ = _gradient_product(case, observed, check_freq=9, fused=False)
    for k in ga:
        assert np.max(np.abs(ga[k])) > 0
        assert np.array_equal(ga[k], gb[k]), k
    assert ma == mb


def test_vd_fused_snapshots_match_oracle():
    case = acoustic_case(kind="acoustic_vd", n=(200, 90), nt=60, halo=6, dtype=np.float64, seed=4, nshots=1)
    _, snaps_ref = oracle_forward(case, snapevery=20)
    _, snaps = _forward_product(case, snapevery=20, fused=True)
    assert sorted(snaps[0].keys()) == [20, 40, 60]
    for it in (20, 40, 60):
        assert rel_l2(snaps[0][it]["pcur"], snaps_ref[0][it]) <= 1e-12


# ---- fused constant-density engine (2D and 3D register-queue march) on grids spanning many tiles / z chunks -----------------


CD_FUSED_CASES = [
    ((523, 301), True, 9, 150),        # 2D: several warps per row, row length not a multiple of the vector width
    ((260, 140), False, 7, 150),
    ((140, 37, 150), True, 6, 70),     # 3D: two x tiles, partial y tile, three z chunks (zc = 64)
    ((67, 30, 41), False, 5, 60),
]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,freetop,halo,nt", CD_FUSED_CASES)
def test_cd_fused_multi_tile_forward_and_gradient(dtype, n, freetop, halo, nt):
    case = acoustic_case(kind="acoustic_cd", n=n, nt=nt, halo=halo, freetop=freetop, dtype=dtype, seed=41 + len(n), nshots=2, nsrc=2, nrec=12)
    rng = np.random.default_rng(7)
    for s in case["shots"]:  # sources close enough to the receivers for the wave to arrive within nt steps; one inside a C-PML strip
        s["src_positions"][:, -1] = rng.uniform(12, 22, size=2) * case["h"]
        s["src_positions"][0, 0] = 3.0 * case["h"]
        s["rec_positions"][0, 0] = (n[0] - 3) * case["h"]
    ref, _ = oracle_forward(case)
    got, _ = _forward_product(case, fused=True)
    for r, g in zip(ref, got):
        assert np.max(np.abs(r)) > 0
        assert rel_l2(g, r) <= (1e-12 if dtype == np.float64 else 1e-6)
    observed = make_observed(case, ref)
    for cf in ((11, 1) if len(n) == 2 else (9,)):
        (gref, mref), _, _ = oracle_gradient(case, observed, check_freq=cf, mute_src=3, mute_rec=1)
        (ggot, mgot), _ = _gradient_product(case, observed, check_freq=cf, mute_src=3, mute_rec=1, fused=True)
        for k in gref:
            assert np.max(np.abs(gref[k])) > 0
            assert rel_l2(ggot[k], gref[k]) <= (1e-11 if dtype == np.float64 else 2e-5), (k, cf)
        assert abs(float(mgot) - float(mref)) <= tol(dtype) * abs(float(mref))


@pytest.mark.parametrize("n", [(300, 170), (70, 45, 80)])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_cd_fused_equals_unfused_bitwise(n, dtype):
    """the fused single-launch step and the one-launch-per-reference-kernel path perform the same operations"""
    case = acoustic_case(kind="acoustic_cd", n=n, nt=90, halo=6, dtype=dtype, seed=8, nshots=1, nrec=16)
    a, _ = _forward_product(case, fused=True)
    b, _ = _forward_product(case, fused=False)
    assert np.array_equal(a[0], b[0])
    syn, _ = oracle_forward(case)
    observed = make_observed(case, syn)
    for cf in (1, 2, 3, 8):
        (ga, ma), _ = _gradient_product(case, observed, check_freq=cf, fused=True)
        (gb, mb), _ = _gradient_product(case, observed, check_freq=cf, fused=False)
        assert np.array_equal(ga["vp"], gb["vp"]), cf
        assert ma == mb


def test_cd_fused_points_on_faces_and_fast_f32():
    """sources / receivers on the grid faces (cells the stencil never updates) and the pure-Float32 arithmetic mode"""
    case = acoustic_case(kind="acoustic_cd", n=(200, 120), nt=120, halo=8, dtype=np.float32, seed=9, nshots=1, nsrc=2, nrec=6)
    h = case["h"]
    case["shots"][0]["rec_positions"][0, :] = (0.0, 40 * h)           # i = 1 face
    case["shots"][0]["rec_positions"][1, :] = (199 * h, 50 * h)       # i = nx face
    case["shots"][0]["rec_positions"][2, :] = (100 * h, 0.0)          # j = 1 face
    case["shots"][0]["src_positions"][1, :] = (60 * h, 0.0)           # a source on the top face
    ref, _ = oracle_forward(case)
    got, _ = _forward_product(case, fused=True)
    assert rel_l2(got[0], ref[0]) <= 1e-6
    fast, _ = _forward_product(case, fused=True, fast_f32=True)
    assert rel_l2(fast[0], ref[0]) <= tol(np.float32)
    observed = make_observed(case, ref)
    (gref, _), _, _ = oracle_gradient(case, observed, check_freq=7)
    (ggot, _), _ = _gradient_product(case, observed, check_freq=7, fused=True, fast_f32=True)
    assert rel_l2(ggot["vp"], gref["vp"]) <= tol(np.float32)


@pytest.mark.parametrize("freetop", [True, False])
def test_cd3d_fused_points_inside_the_rim_boxes(freetop):
    """3D: sources / receivers inside the C-PML strips -- the cells of the z-marched x / y strip boxes, of their corner columns and of
    the per-vector z planes -- reach the kernels through per-CTA lists whose numbering differs per rim form: the fused engine must
    agree bit for bit with the one-launch-per-reference-kernel path (which has no such lists), and with the oracle"""
    n, halo = (70, 45, 80), 6
    case = acoustic_case(kind="acoustic_cd", n=n, nt=140, halo=halo, dtype=np.float32, seed=12, nshots=1, nsrc=3, nrec=12, freetop=freetop)
    h = case["h"]
    idx = [(2, 20, 40), (66, 22, 30), (35, 3, 50), (30, 41, 20), (2, 2, 41), (67, 42, 60), (33, 21, 76), (3, 3, 77), (34, 20, 79), (0, 10, 33), (40, 44, 35), (36, 2, 3)]
    for r, (i, j, k) in enumerate(idx):  # x strips, y strips, x-y corner columns, the lower z strip and its corners, faces, the top planes
        case["shots"][0]["rec_positions"][r, :] = (i * h, j * h, k * h)
    case["shots"][0]["src_positions"][1, :] = (30 * h, 4 * h, 45 * h)   # a source inside a y strip
    case["shots"][0]["src_positions"][2, :] = (64 * h, 25 * h, 38 * h)  # and one inside an x strip
    a, _ = _forward_product(case, fused=True)
    b, _ = _forward_product(case, fused=False)
    assert np.array_equal(a[0], b[0])
    assert np.count_nonzero(np.abs(a[0]).max(axis=0) > 0) >= 9  # the strip receivers see the wavefield (faces stay zero)
    ref, _ = oracle_forward(case)
    assert rel_l2(a[0], ref[0]) <= 1e-6
    observed = make_observed(case, ref)
    for cf in (1, 7):
        (ga, ma), _ = _gradient_product(case, observed, check_freq=cf, fused=True)
        (gb, mb), _ = _gradient_product(case, observed, check_freq=cf, fused=False)
        assert np.array_equal(ga["vp"], gb["vp"]), cf
        assert ma == mb


@pytest.mark.parametrize("kind,n", [("acoustic_vd", (300, 170)), ("acoustic_cd", (300, 170)), ("acoustic_cd", (70, 45, 80))])
def test_eager_launches_equal_graph_replay(kind, n):
    """SWB_FLAG_NO_GRAPH: the same launch sequence enqueued eagerly (two shots: the second one replays the captured graphs)"""
    case = acoustic_case(kind=kind, n=n, nt=80, halo=10, dtype=np.float32, seed=9, nshots=2, nrec=12)
    for sh in case["shots"]:  # shallow sources: the wave reaches the receivers near the top within nt steps
        sh["src_positions"][:, -1] = 14.0 * case["h"]
    res = {}
    for graphs in (True, False):
        seis, _ = _forward_product(case, graphs=graphs)
        observed = make_observed(case, seis)
        res[graphs] = (seis, _gradient_product(case, observed, check_freq=9, graphs=graphs)[0])
    for a, b in zip(res[True][0], res[False][0]):
        assert np.max(np.abs(a)) > 0
        assert np.array_equal(a, b)
    (ga, ma), (gb, mb) = res[True][1], res[False][1]
    for k in ga:
        assert np.array_equal(ga[k], gb[k]), k
    assert ma == mb


@pytest.mark.parametrize("kind,n", [("acoustic_cd", (61, 53)), ("acoustic_cd", (30, 26, 28)), ("acoustic_vd", (67, 59))])
def test_snapshots_from_the_graph_captured_sweep(kind, n):
    """savesnapshot! (snapshotter.jl:33-41) inside the CUDA-graph replay: device-resident slots drained to pinned host memory on a side
    stream.  Same fields as the oracle's snapshots, bit-identical to the eager (no-graph) path, unchanged seismograms, and a second shot
    replays the captured sweep into the same slots."""
    import swb200 as S

    nt = 60 if len(n) == 2 else 40
    case = acoustic_case(kind=kind, n=n, nt=nt, halo=5, dtype=np.float32, seed=8, nshots=2)
    ref_seis, snaps_ref = oracle_forward(case, snapevery=10)
    params, matprop, shots, _, runparams, _ = product_inputs(case, snapevery=10)
    ws = S.build_wavesim(params, matprop, runparams=runparams)
    snaps = S.swforward(ws, matprop, shots)
    ws.close()
    seis_plain, _ = _forward_product(case)
    _, snaps_eager = _forward_product(case, snapevery=10, graphs=False)
    assert len(snaps) == 2
    for s in range(2):
        assert sorted(snaps[s].keys()) == list(range(10, nt + 1, 10))
        assert np.array_equal(shots[s].recs.seismograms, seis_plain[s])
        for it in snaps[s]:
            assert np.max(np.abs(snaps[s][it]["pcur"])) > 0
            assert rel_l2(snaps[s][it]["pcur"], snaps_ref[s][it]) <= 1e-6
            assert np.array_equal(snaps[s][it]["pcur"], snaps_eager[s][it]["pcur"])
            if kind == "acoustic_vd":
                for c in range(2):
                    assert np.max(np.abs(snaps[s][it]["vcur"][c])) > 0
                    assert np.array_equal(snaps[s][it]["vcur"][c], snaps_eager[s][it]["vcur"][c])
    assert not np.array_equal(snaps[0][nt]["pcur"], snaps[1][nt]["pcur"])  # the second shot really overwrote the slots
