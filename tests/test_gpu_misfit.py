"""Misfit handling on the B200 path (src/inversion/misfits/L2Misfit.jl:24-95, src/apis/misfit.jl:26-104):
the device-side L2 adjoint source with windows and a diagonal inverse covariance against the host (pluggable-misfit) path and the
oracle, and swmisfit! through the public API including the reference's shot-1 quirk."""
import numpy as np
import pytest

import cases
import elastic_cases as EC
from cases import rel_l2

pytestmark = pytest.mark.gpu


def _acoustic(case, observed, force_host, check_freq=8, dense=False):
    import swb200 as S

    params, matprop, shots, misfit, runparams, gradparams = cases.product_inputs(case, observed=observed, check_freq=check_freq)
    if dense:  # a dense (nt, nt) matrix with the same diagonal: not eligible for the device path
        for m in misfit:
            m.invcov = np.diag(m.invcov) + 0.0
    ws = S.build_wavesim(params, matprop, runparams=runparams, gradparams=gradparams, gradient=True)
    ws.force_host_misfit = force_host
    try:
        return S.swgradient(ws, matprop, shots, misfit)
    finally:
        ws.close()


@pytest.mark.parametrize("kind,n", [("acoustic_cd", (48, 44)), ("acoustic_vd", (48, 44)), ("acoustic_cd", (24, 22, 26))])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_device_l2_with_windows_and_diagonal_invcov_equals_host_path(kind, n, dtype):
    case = cases.acoustic_case(kind=kind, n=n, nt=100 if len(n) == 2 else 50, halo=6 if len(n) == 2 else 5, dtype=dtype, seed=3, windows=True, diag_invcov=True)
    syn, _ = cases.oracle_forward(case)
    observed = cases.make_observed(case, syn)
    (gd, md), (gh, mh) = _acoustic(case, observed, False), _acoustic(case, observed, True)
    for k in gd:
        assert np.max(np.abs(gd[k])) > 0
        assert np.array_equal(gd[k], gh[k]), k  # same operations in the same order: bit for bit
    assert md == mh
    (gref, mref), _, _ = cases.oracle_gradient(case, observed, check_freq=8)
    for k in gref:
        assert rel_l2(gd[k], gref[k]) <= (1e-11 if dtype == np.float64 else 2e-5), k
    assert abs(float(md) - float(mref)) <= (1e-10 if dtype == np.float64 else 1e-5) * abs(float(mref))


def test_dense_invcov_takes_the_host_path_and_agrees():
    case = cases.acoustic_case(kind="acoustic_vd", n=(48, 44), nt=100, halo=6, dtype=np.float64, seed=3, windows=True, diag_invcov=True)
    syn, _ = cases.oracle_forward(case)
    observed = cases.make_observed(case, syn)
    (g1, m1), (g2, m2) = _acoustic(case, observed, False), _acoustic(case, observed, False, dense=True)
    for k in g1:
        assert rel_l2(g2[k], g1[k]) <= 1e-12, k  # dense mat-vec sums nt terms (all but one zero): rounding-level agreement
    assert abs(m1 - m2) <= 1e-12 * abs(m1)


def test_device_l2_elastic_windows_diag():
    import swb200 as S
    from test_gpu_elastic import product_inputs as ela_inputs

    case = EC.elastic_case(n=(72, 64), nt=100, halo=6, dtype=np.float64, kind="momten", nshots=1, nrec=4, seed=23)
    syn, _ = EC.oracle_forward(case)
    obs = EC.make_observed(case, syn)
    nt = case["nt"]
    kw = dict(windows=[(10, 50), (55, 97)], invcov=1.0 + 0.5 * np.sin(np.arange(nt) * 0.1))
    out = []
    for force_host in (False, True):
        params, matprop, shots, _, runparams, gradparams = ela_inputs(case, observed=obs, check_freq=9)
        misfit = [S.L2Misfit(observed=o, **kw) for o in obs]
        ws = S.build_wavesim(params, matprop, runparams=runparams, gradparams=gradparams, gradient=True)
        ws.force_host_misfit = force_host
        out.append(S.swgradient(ws, matprop, shots, misfit))
        ws.close()
    (gd, md), (gh, mh) = out
    for k in ("rho", "lambda", "mu"):
        assert np.max(np.abs(gd[k])) > 0
        assert np.array_equal(gd[k], gh[k]), k
    assert md == mh
    # against the oracle with the same misfit
    from oracle import oracle as O

    sim = O.build_wavesim("elastic_iso", EC.params_oracle(case), gradient=True, check_freq=9)
    oshots = EC.oracle_shots(case)
    omis = [O.L2Misfit(observed=o, **kw) for o in obs]
    gref, mref = O.swgradient(sim, EC.matprops(case), oshots, omis, compute_misfit=True)
    for k in ("rho", "lambda", "mu"):
        assert rel_l2(gd[k], gref[k]) <= 1e-10, k
    assert abs(float(md) - float(mref)) <= 1e-10 * abs(float(mref))


@pytest.mark.parametrize("kind", ["acoustic_cd", "acoustic_vd"])
def test_swmisfit_through_the_b200_path(kind):
    """swmisfit! (misfit.jl:26-104): forward run + misfit; the reference's loop `for s in length(shots)` evaluates shot 1 only
    (misfit.jl:97-100) -- reproduced by default, the sum over all shots on request; both against the oracle"""
    import swb200 as S
    from oracle import oracle as O

    case = cases.acoustic_case(kind=kind, n=(64, 56), nt=120, halo=8, dtype=np.float64, seed=9, nshots=3, windows=True, diag_invcov=True)
    syn, _ = cases.oracle_forward(case)
    observed = cases.make_observed(case, syn)
    params, matprop, shots, misfit, runparams, _ = cases.product_inputs(case, observed=observed)
    compat = S.swmisfit(params, matprop, shots, misfit, runparams=runparams)
    for s, r in zip(shots, syn):
        assert rel_l2(s.recs.seismograms, r) <= 1e-12
    full = S.swmisfit(params, matprop, shots, misfit, runparams=runparams, reference_compat=False)
    sim = O.build_wavesim(case["kind"], cases.case_params_oracle(case))
    omis = [O.L2Misfit(observed=obs, **cases.misfit_kwargs(case, case["nt"])) for obs in observed]
    ref_compat = O.swmisfit(sim, cases.matprop_list(case), cases.oracle_shots(case), omis, reference_bug=True)
    ref_full = O.swmisfit(sim, cases.matprop_list(case), cases.oracle_shots(case), omis, reference_bug=False)
    assert ref_full > ref_compat > 0
    assert abs(compat - ref_compat) <= 1e-10 * ref_compat
    assert abs(full - ref_full) <= 1e-10 * ref_full
    # a WaveSimulation built once can be reused (misfit.jl:62-104)
    ws = S.build_wavesim(params, matprop, runparams=runparams)
    assert S.swmisfit(ws, matprop, shots, misfit) == compat
    ws.close()
