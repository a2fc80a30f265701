"""1D acoustic simulations (acoustic1D_xPU.jl:1-125, acoustic1D_VD_xPU.jl:1-140) on the B200 path: the per-reference-kernel engines run a
1D grid as the single-row case of their 2D kernels.  Seismograms, snapshots, gradients (with and without checkpointing) and misfits
against the CPU oracle; reference-side analogue: the 1D test sets of test/test_gradient_acoustic_constant_density.jl:18-120 and
test/test_gradient_acoustic_variable_density.jl:18-91."""
import numpy as np
import pytest

import cases
from cases import rel_l2, tol

pytestmark = pytest.mark.gpu


def _case(kind, dtype, freetop, seed=2, nt=320):
    return cases.acoustic_case(kind=kind, n=(301,), nt=nt, halo=10, freetop=freetop, dtype=dtype, nshots=2, nsrc=2, nrec=4, seed=seed, f0=10.0)


@pytest.mark.parametrize("kind", ["acoustic_cd", "acoustic_vd"])
@pytest.mark.parametrize("dtype,freetop", [(np.float64, False), (np.float64, True), (np.float32, False)])
def test_1d_forward_and_snapshots_match_oracle(kind, dtype, freetop):
    import swb200 as S

    case = _case(kind, dtype, freetop)
    ref, snaps_ref = cases.oracle_forward(case, snapevery=80)
    params, matprop, shots, _, runparams, _ = cases.product_inputs(case, snapevery=80)
    snaps = S.swforward(params, matprop, shots, runparams=runparams)
    for s, (r, sh) in enumerate(zip(ref, shots)):
        assert np.max(np.abs(r)) > 0
        err = rel_l2(sh.recs.seismograms, r)
        assert err <= (1e-12 if dtype == np.float64 else 1e-6), err
        for it in (80, 160, 240, 320):
            assert rel_l2(snaps[s][it]["pcur"], snaps_ref[s][it]) <= (1e-12 if dtype == np.float64 else 1e-6)


@pytest.mark.parametrize("kind", ["acoustic_cd", "acoustic_vd"])
@pytest.mark.parametrize("dtype,check_freq,fast", [(np.float64, 1, False), (np.float64, 7, False), (np.float32, 9, False), (np.float32, 9, True)])
def test_1d_gradient_and_misfit_match_oracle(kind, dtype, check_freq, fast):
    import swb200 as S

    case = _case(kind, dtype, False, seed=5)
    syn, _ = cases.oracle_forward(case)
    observed = cases.make_observed(case, syn)
    (gref, mref), sref, _ = cases.oracle_gradient(case, observed, check_freq=check_freq, mute_src=3, mute_rec=2)
    params, matprop, shots, misfit, runparams, gradparams = cases.product_inputs(case, observed=observed, check_freq=check_freq, mute_src=3, mute_rec=2, fast_f32=fast)
    ggot, mgot = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    assert set(ggot) == set(gref)
    for k in gref:
        assert ggot[k].shape == (301,) and np.max(np.abs(gref[k])) > 0
        err = rel_l2(ggot[k], gref[k])
        assert err <= tol(dtype), (k, err)
        if not fast:
            assert err <= (1e-11 if dtype == np.float64 else 2e-5), (k, err)
    assert abs(float(mgot) - float(mref)) <= tol(dtype) * abs(float(mref))


def test_1d_checkpointed_equals_non_checkpointed():
    """test_gradient_acoustic_constant_density.jl:194-230 (1D): the gradients with and without checkpointing agree"""
    import swb200 as S

    for kind in ("acoustic_cd", "acoustic_vd"):
        case = _case(kind, np.float64, False, seed=7, nt=300)
        syn, _ = cases.oracle_forward(case)
        observed = cases.make_observed(case, syn)
        out = []
        for cf in (1, 11):
            params, matprop, shots, misfit, runparams, gradparams = cases.product_inputs(case, observed=observed, check_freq=cf)
            out.append(S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams))
        for k in out[0][0]:
            assert np.array_equal(out[0][0][k], out[1][0][k]), (kind, k)
        assert out[0][1] == out[1][1]
