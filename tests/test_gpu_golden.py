"""The CUDA path against the committed golden vectors of the oracle (tests/golden/oracle_v1.npz), through the public API / C ABI:
north-star tolerances (relative L2 <= 1e-5 Float64, 1e-4 Float32), met with orders of magnitude to spare."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

from cases import acoustic_case, make_observed, product_inputs, rel_l2, tol  # noqa: E402
from elastic_cases import elastic_case  # noqa: E402
from elastic_cases import make_observed as ela_make_observed  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fused", [True, False])
def test_acoustic_matches_golden(fused):
    import make_golden
    import swb200 as S

    ref = np.load(os.path.join(HERE, "golden", "oracle_v1.npz"))
    for k, (kind, n, dtype, freetop) in enumerate(make_golden.ACOUSTIC):
        case = acoustic_case(kind=kind, n=n, nt=60, halo=6, freetop=freetop, dtype=dtype, seed=100 + k, nshots=1, nsrc=2, nrec=4)
        for sh in case["shots"]:
            sh["src_positions"][:, -1] = 9.0 * case["h"]
        seis_ref = ref[f"acou{k}_seis"]
        params, matprop, shots, _, runparams, _ = product_inputs(case, fused=fused)
        S.swforward(params, matprop, shots, runparams=runparams)
        assert rel_l2(shots[0].recs.seismograms, seis_ref) <= tol(dtype) * 1e-2, (k, "seismograms")
        obs = make_observed(case, [seis_ref])
        params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=obs, check_freq=7, mute_src=2, mute_rec=1, fused=fused)
        grad, mis = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
        for name, g in grad.items():
            assert rel_l2(g, ref[f"acou{k}_grad_{name}"]) <= tol(dtype), (k, name)
        assert abs(float(mis) - float(ref[f"acou{k}_misfit"])) <= tol(dtype) * abs(float(ref[f"acou{k}_misfit"])), k


@pytest.mark.parametrize("fused", [True, False])
def test_elastic_matches_golden(fused):
    import make_golden
    import swb200 as S
    from test_gpu_elastic import product_inputs as ela_inputs

    ref = np.load(os.path.join(HERE, "golden", "oracle_v1.npz"))
    for k, (kind, dtype, freetop) in enumerate(make_golden.ELASTIC):
        case = elastic_case(n=(56, 44), nt=60, halo=6, freetop=freetop, dtype=dtype, kind=kind, nshots=1, nsrc=1, nrec=3, seed=200 + k)
        for sh in case["shots"]:
            sh["src_positions"][:, 1] = 11.3 * case["h"]
        seis_ref = ref[f"ela{k}_seis"]
        params, matprop, shots, _, runparams, _ = ela_inputs(case, fused=fused)
        S.swforward(params, matprop, shots, runparams=runparams)
        assert rel_l2(shots[0].recs.seismograms, seis_ref) <= tol(dtype) * 1e-1, (k, "seismograms")
        obs = ela_make_observed(case, [seis_ref])
        params, matprop, shots, misfit, runparams, gradparams = ela_inputs(case, observed=obs, check_freq=7, mute_src=2, mute_rec=1, fused=fused)
        grad, mis = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
        for name, g in grad.items():
            assert rel_l2(g, ref[f"ela{k}_grad_{name}"]) <= tol(dtype), (k, name)
        assert abs(float(mis) - float(ref[f"ela{k}_misfit"])) <= tol(dtype) * abs(float(ref[f"ela{k}_misfit"])), k
