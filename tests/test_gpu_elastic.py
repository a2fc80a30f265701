"""GPU parity tests for the elastic P-SV path: CUDA engine (through the C ABI) vs the CPU oracle on identical seeded inputs.
Tolerances: north-star rel-L2 <= 1e-5 (Float64) / 1e-4 (Float32); the reference-faithful arithmetic is expected to agree to
rounding level, which the tighter bounds pin."""
import numpy as np
import pytest

from cases import rel_l2, tol
from elastic_cases import elastic_case, make_observed, oracle_forward, oracle_gradient

pytestmark = pytest.mark.gpu


def product_inputs(case, observed=None, check_freq=1, mute_src=0, mute_rec=0, fast_f32=False, snapevery=None, fused=True, graphs=True):
    import swb200 as S

    T = case["dtype"].type
    bc = S.CPMLBoundaryConditionParameters(halo=case["halo"], rcoef=T(case["rcoef"]), freeboundtop=case["freetop"])
    params = S.InputParametersElastic(case["nt"], T(case["dt"]), case["n"], (T(case["h"]), T(case["h"])), bc, dtype=case["dtype"])
    matprop = S.ElasticIsoMaterialProperties(case["rho"], case["lam"], case["mu"])
    shots = []
    for s in case["shots"]:
        recs = S.VectorReceivers(s["rec_positions"].astype(T), case["nt"], dtype=case["dtype"])
        if s["kind"] == "momten":
            mts = [S.MomentTensor2D(*[T(v) for v in row]) for row in s["momtens"]]
            srcs = S.MomentTensorSources(s["src_positions"].astype(T), s["src_tf"].astype(T), mts, T(s["domfreq"]))
            shots.append(S.MomentTensorShot(srcs=srcs, recs=recs))
        else:
            srcs = S.ExternalForceSources(s["src_positions"].astype(T), s["src_tf"].astype(T), T(s["domfreq"]))
            shots.append(S.ExternalForceShot(srcs=srcs, recs=recs))
    runparams = S.RunParameters(parall="B200", fast_f32=fast_f32, fused=fused, graphs=graphs, snapevery=snapevery, erroronPPW=False)
    gradparams = S.GradParameters(mute_radius_src=mute_src, mute_radius_rec=mute_rec, compute_misfit=True, check_freq=check_freq)
    misfit = [S.L2Misfit(observed=o) for o in observed] if observed is not None else None
    return params, matprop, shots, misfit, runparams, gradparams


@pytest.mark.parametrize("kind", ["momten", "extforce"])
@pytest.mark.parametrize("dtype,freetop,n", [(np.float64, True, (96, 80)), (np.float64, False, (83, 71)), (np.float32, True, (90, 77))])
@pytest.mark.parametrize("fused", [False, True])
def test_forward_seismograms_match_oracle(kind, dtype, freetop, n, fused):
    import swb200 as S

    case = elastic_case(n=n, nt=150, halo=7, freetop=freetop, dtype=dtype, kind=kind, nshots=2, nsrc=2, nrec=5, seed=13)
    ref, _ = oracle_forward(case)
    params, matprop, shots, _, runparams, _ = product_inputs(case, fused=fused)
    S.swforward(params, matprop, shots, runparams=runparams)
    for r, sh in zip(ref, shots):
        g = sh.recs.seismograms
        assert g.dtype == np.dtype(dtype) and np.max(np.abs(r)) > 0
        err = rel_l2(g, r)
        assert err <= tol(dtype), err
        assert err <= (1e-12 if dtype == np.float64 else 2e-6), err


@pytest.mark.parametrize("kind,dtype,check_freq", [("momten", np.float64, 1), ("momten", np.float64, 7), ("extforce", np.float64, 10), ("extforce", np.float32, 9),
                                                   ("momten", np.float32, 1)])
@pytest.mark.parametrize("fused", [False, True])
def test_gradient_and_misfit_match_oracle(kind, dtype, check_freq, fused):
    import swb200 as S

    case = elastic_case(n=(72, 64), nt=100, halo=6, dtype=dtype, kind=kind, nshots=2, nrec=4, seed=17 + check_freq)
    syn, _ = oracle_forward(case)
    obs = make_observed(case, syn)
    (gref, mref), sref = oracle_gradient(case, obs, check_freq=check_freq, mute_src=3, mute_rec=1)
    params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=obs, check_freq=check_freq, mute_src=3, mute_rec=1, fused=fused)
    ggot, mgot = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    assert set(ggot) == {"rho", "lambda", "mu"}
    for k in gref:
        assert np.max(np.abs(gref[k])) > 0
        err = rel_l2(ggot[k], gref[k])
        assert err <= tol(dtype), (k, err)
        assert err <= (1e-10 if dtype == np.float64 else 5e-5), (k, err)
    assert abs(float(mgot) - float(mref)) <= tol(dtype) * abs(float(mref))
    for r, sh in zip(sref, shots):
        assert rel_l2(sh.recs.seismograms, r) <= tol(dtype)


def test_checkpointed_equals_non_checkpointed_gradient():
    """reference test: test/test_gradient_elastic_homogeneous.jl:62-119"""
    import swb200 as S

    case = elastic_case(n=(80, 64), nt=121, halo=6, kind="extforce", seed=4, nrec=3)
    syn, _ = oracle_forward(case)
    obs = make_observed(case, syn)
    out = []
    for cf in (1, 11):
        params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=obs, check_freq=cf)
        out.append(S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams))
    for k in ("rho", "lambda", "mu"):
        assert np.array_equal(out[0][0][k], out[1][0][k])
    assert out[0][1] == out[1][1]


def test_fast_f32_within_tolerance():
    import swb200 as S

    case = elastic_case(n=(96, 80), nt=300, halo=8, dtype=np.float32, kind="momten", seed=6)
    ref, _ = oracle_forward(case)
    params, matprop, shots, _, runparams, _ = product_inputs(case, fast_f32=True)
    S.swforward(params, matprop, shots, runparams=runparams)
    assert rel_l2(shots[0].recs.seismograms, ref[0]) <= tol(np.float32)


def test_nearest_grid_point_sources_and_snapshots():
    """sincinterp = false path (ela_models.jl:31-45) and the ucur / σ snapshots (ela_forward.jl:64-67)"""
    import swb200 as S
    from oracle import oracle as O
    from elastic_cases import matprops, oracle_shots, params_oracle

    case = elastic_case(n=(64, 60), nt=60, halo=6, kind="extforce", seed=8, nrec=3)
    sim = O.build_wavesim("elastic_iso", params_oracle(case), sincinterp=False)
    oshots = oracle_shots(case)
    snaps_ref = O.swforward(sim, matprops(case), oshots, snapevery=20)
    params, matprop, shots, _, runparams, _ = product_inputs(case, snapevery=20)
    ws = S.build_wavesim(params, matprop, runparams=runparams, sincinterp=False)
    snaps = S.swforward(ws, matprop, shots)
    assert rel_l2(shots[0].recs.seismograms, oshots[0].seismograms) <= 1e-12
    assert sorted(snaps[0].keys()) == [20, 40, 60]
    for it in (20, 40, 60):
        for c in range(2):
            assert rel_l2(snaps[0][it]["ucur"][c], snaps_ref[0][it]["ucur"][c]) <= 1e-12
        for c in range(3):
            assert rel_l2(snaps[0][it]["σ"][c], snaps_ref[0][it]["sigma"][c]) <= 1e-12


# ---- fused engine (stresses on chip) on grids that span many tiles: interior fast-path tiles + C-PML / edge / free-surface tiles ----


@pytest.mark.parametrize("kind", ["momten", "extforce"])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("n,freetop", [((420, 150), True), ((301, 97), False)])
def test_fused_multi_tile_forward_and_gradient(kind, dtype, n, freetop):
    import swb200 as S

    case = elastic_case(n=n, nt=130, halo=9, freetop=freetop, dtype=dtype, kind=kind, nshots=1, nsrc=2, nrec=6, seed=23)
    for s in case["shots"]:  # a source close to the receivers so that the wave arrives within nt steps
        s["src_positions"][:, 1] = np.array([14.3, 19.8])[: s["src_positions"].shape[0]] * case["h"]
    ref, _ = oracle_forward(case)
    params, matprop, shots, _, runparams, _ = product_inputs(case, fused=True)
    S.swforward(params, matprop, shots, runparams=runparams)
    for r, sh in zip(ref, shots):
        assert np.max(np.abs(r)) > 0
        assert rel_l2(sh.recs.seismograms, r) <= (1e-12 if dtype == np.float64 else 2e-6)
    obs = make_observed(case, ref)
    for cf in (11, 1):
        (gref, mref), _ = oracle_gradient(case, obs, check_freq=cf, mute_src=3, mute_rec=1)
        params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=obs, check_freq=cf, mute_src=3, mute_rec=1, fused=True)
        ggot, mgot = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
        for k in gref:
            assert np.max(np.abs(gref[k])) > 0
            assert rel_l2(ggot[k], gref[k]) <= (1e-10 if dtype == np.float64 else 5e-5), (k, cf)
        assert abs(float(mgot) - float(mref)) <= tol(dtype) * abs(float(mref))


@pytest.mark.parametrize("kind", ["momten", "extforce"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_equals_unfused_bitwise(kind, dtype):
    """the fused step (stresses on chip, halo recomputed) performs the operations of the four-sweep path"""
    import swb200 as S

    case = elastic_case(n=(300, 140), nt=100, halo=8, dtype=dtype, kind=kind, nshots=1, nsrc=2, nrec=5, seed=29)
    for s in case["shots"]:
        s["src_positions"][:, 1] = np.array([15.2, 21.7])[: s["src_positions"].shape[0]] * case["h"]
    out = {}
    for fused in (True, False):
        params, matprop, shots, _, runparams, _ = product_inputs(case, fused=fused)
        S.swforward(params, matprop, shots, runparams=runparams)
        out[fused] = shots[0].recs.seismograms.copy()
    assert np.max(np.abs(out[True])) > 0
    assert np.array_equal(out[True], out[False])
    obs = make_observed(case, [out[True]])
    grads = {}
    for fused in (True, False):
        for cf in (1, 2, 3, 9):
            params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=obs, check_freq=cf, fused=fused)
            grads[(fused, cf)] = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    for cf in (1, 2, 3, 9):
        (ga, ma), (gb, mb) = grads[(True, cf)], grads[(False, cf)]
        for k in ga:
            assert np.array_equal(ga[k], gb[k]), (k, cf)
        assert ma == mb


@pytest.mark.parametrize("dtype,n,freetop,halo", [(np.float32, (1100, 610), True, 20), (np.float64, (777, 421), False, 20), (np.float32, (515, 300), False, 12)])
def test_fused_equals_unfused_bitwise_production_halo(dtype, n, freetop, halo):
    """many interior tiles, full-width C-PML strips (halo as in BASELINE's configurations), partial last tiles: the plain / special
    vector split of the edge tiles and the correlation fused into the adjoint launch reproduce the four-sweep path bit for bit"""
    import swb200 as S

    case = elastic_case(n=n, nt=70, halo=halo, freetop=freetop, dtype=dtype, kind="momten", nshots=1, nsrc=2, nrec=7, seed=31)
    for s in case["shots"]:
        ext_x = (n[0] - 1) * case["h"]
        s["src_positions"][:, 0] = np.array([0.43, 0.58])[: s["src_positions"].shape[0]] * ext_x
        s["src_positions"][:, 1] = np.array([31.2, 40.7])[: s["src_positions"].shape[0]] * case["h"]
        s["rec_positions"][:, 0] = np.linspace(0.31, 0.69, s["rec_positions"].shape[0]) * ext_x
        s["rec_positions"][:, 1] = 27.4 * case["h"]
    out, grads = {}, {}
    for fused in (True, False):
        params, matprop, shots, _, runparams, _ = product_inputs(case, fused=fused)
        S.swforward(params, matprop, shots, runparams=runparams)
        out[fused] = shots[0].recs.seismograms.copy()
    assert np.max(np.abs(out[True])) > 0
    assert np.array_equal(out[True], out[False])
    obs = make_observed(case, [out[True]])
    for fused in (True, False):
        params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=obs, check_freq=8, fused=fused)
        grads[fused] = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    (ga, ma), (gb, mb) = grads[True], grads[False]
    for k in ga:
        assert np.max(np.abs(ga[k])) > 0
        assert np.array_equal(ga[k], gb[k]), k
    assert ma == mb


@pytest.mark.parametrize("kind", ["momten", "extforce"])
def test_eager_launches_equal_graph_replay(kind):
    """SWB_FLAG_NO_GRAPH: the same launch sequence (side-stream receiver sums, fused correlation) enqueued eagerly"""
    import swb200 as S

    case = elastic_case(n=(300, 140), nt=60, halo=8, dtype=np.float32, kind=kind, nshots=2, nsrc=2, nrec=5, seed=33)
    for s in case["shots"]:
        s["src_positions"][:, 1] = np.array([15.2, 21.7])[: s["src_positions"].shape[0]] * case["h"]
    res = {}
    for graphs in (True, False):
        params, matprop, shots, _, runparams, _ = product_inputs(case, graphs=graphs)
        S.swforward(params, matprop, shots, runparams=runparams)
        seis = [sh.recs.seismograms.copy() for sh in shots]
        obs = make_observed(case, seis)
        params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=obs, check_freq=7, graphs=graphs)
        res[graphs] = (seis, S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams))
    for a, b in zip(res[True][0], res[False][0]):
        assert np.max(np.abs(a)) > 0
        assert np.array_equal(a, b)
    (ga, ma), (gb, mb) = res[True][1], res[False][1]
    for k in ga:
        assert np.array_equal(ga[k], gb[k]), k
    assert ma == mb


def test_moment_tensor_source_near_the_edge_falls_back_to_the_step_by_step_path():
    """A sinc-spread moment-tensor source within a few cells of a lateral edge touches stress cells the reference never updates
    (they accumulate the injection, elastic2D_iso_xPU.jl:187-199).  The fused engine cannot keep such a stress on chip: the shot is
    run on the four-sweep path automatically (no error, no flag), and the next ordinary shot is fused again."""
    import swb200 as S

    case = elastic_case(n=(72, 64), nt=110, halo=6, dtype=np.float64, kind="momten", nshots=2, nrec=4, seed=31)
    h = case["h"]
    case["shots"][0]["src_positions"][:, 0] = (case["n"][0] - 3) * h + 0.31 * h   # 2.7 cells from the right edge: the sinc stencil reaches column nx
    case["shots"][0]["src_positions"][:, 1] = 0.55 * (case["n"][1] - 1) * h
    ref, _ = oracle_forward(case)
    params, matprop, shots, _, runparams, _ = product_inputs(case)
    ws = S.build_wavesim(params, matprop, runparams=runparams)
    l0 = S._lib.load().swb_launch_count()
    S.swforward(ws, matprop, shots[:1])
    l1 = S._lib.load().swb_launch_count()
    S.swforward(ws, matprop, shots[1:])
    l2 = S._lib.load().swb_launch_count()
    ws.close()
    for r, sh in zip(ref, shots):
        assert np.max(np.abs(r)) > 0
        assert rel_l2(sh.recs.seismograms, r) <= 1e-12
    assert (l1 - l0) >= 3 * case["nt"] > (l2 - l1)   # four-sweep path (>= 3 launches per step) vs fused path
    # the gradient of the near-edge shot takes the same route
    obs = make_observed(case, ref)
    (gref, mref), _ = oracle_gradient(case, obs, check_freq=7, mute_src=2)
    params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=obs, check_freq=7, mute_src=2)
    ggot, mgot = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    for k in gref:
        assert rel_l2(ggot[k], gref[k]) <= 1e-10, k
    assert abs(float(mgot) - float(mref)) <= 1e-10 * abs(float(mref))
