"""Parity of the arithmetic modes bench.py actually times (SWB_FLAG_FAST_F32 = Float32 storage + Float32 arithmetic, and the
default Float64-intermediate mode) against the CPU oracle: seismograms, misfit and every gradient component, on down-scaled
twins of C2 / C3 / C4 (tests/twins.py: halo 20, free surface, check_freq = isqrt(nt), zero observed data, the benchmark's own
source / receiver layout) and on the full-size C2 shot itself (4096^2, nt = 1000, check_freq = 31; OpenMP oracle, ~2 min of host
time).  Reference-side analogue: test/test_gradient_acoustic_variable_density.jl:93-179 (gradients with and without
checkpointing on a VD model), test/test_gradient_elastic_homogeneous.jl:62-119.

Tolerance = north_star's: relative L2 <= 1e-4 in Float32 (1e-5 in Float64)."""
import json
import os
import time

import numpy as np
import pytest

import cases
import elastic_cases as EC
import twins
from cases import rel_l2, tol

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _omp_oracle():
    from oracle import oracle as O

    O.use_openmp(True)
    yield
    O.use_openmp(False)


def _acoustic_gradient(case, fast_f32, check_freq):
    import swb200 as S

    observed = [np.zeros((case["nt"], sh["rec_positions"].shape[0]), dtype=case["dtype"], order="F") for sh in case["shots"]]
    params, matprop, shots, misfit, runparams, gradparams = cases.product_inputs(case, observed=observed, fast_f32=fast_f32, check_freq=check_freq, mute_src=3)
    grad, mis = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    return grad, float(mis), [s.recs.seismograms for s in shots], observed


def _report(name, errs):
    print(f"\n[parity] {name}: " + ", ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):  # on the GPU box: keep the measured errors next to the other evidence
        with open(os.path.join(out, "parity_errors.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, **errs}) + "\n")


@pytest.mark.parametrize("fast_f32", [True, False])
@pytest.mark.parametrize("check_freq", [None, 1])
def test_c2_twin_vd_gradient(fast_f32, check_freq):
    case = twins.c2_twin(n=512, nt=600, nrec=512)
    cf = check_freq or case["check_freq"]
    grad, mis, seis, observed = _acoustic_gradient(case, fast_f32, cf)
    (gref, mref), sref, _ = cases.oracle_gradient(case, observed, check_freq=cf, mute_src=3)
    errs = {"seis": rel_l2(seis[0], sref[0]), "vp": rel_l2(grad["vp"], gref["vp"]), "rho": rel_l2(grad["rho"], gref["rho"]),
            "misfit": abs(mis - float(mref)) / abs(float(mref))}
    _report(f"C2 twin 512^2 nt=600 cf={cf} fast_f32={fast_f32}", errs)
    assert np.max(np.abs(gref["vp"])) > 0 and np.max(np.abs(gref["rho"])) > 0 and float(mref) > 0
    for k, e in errs.items():
        assert e <= tol(np.float32), (k, e)


@pytest.mark.parametrize("fast_f32", [True, False])
def test_c4_twin_cd3d_gradient(fast_f32):
    case = twins.c4_twin(n=(96, 88, 104), nt=120)
    cf = case["check_freq"]
    grad, mis, seis, observed = _acoustic_gradient(case, fast_f32, cf)
    (gref, mref), sref, _ = cases.oracle_gradient(case, observed, check_freq=cf, mute_src=3)
    errs = {"seis": rel_l2(seis[0], sref[0]), "vp": rel_l2(grad["vp"], gref["vp"]), "misfit": abs(mis - float(mref)) / abs(float(mref))}
    _report(f"C4 twin 96x88x104 nt=120 cf={cf} fast_f32={fast_f32}", errs)
    assert np.max(np.abs(gref["vp"])) > 0
    for k, e in errs.items():
        assert e <= tol(np.float32), (k, e)


@pytest.mark.parametrize("dtype,fast_f32", [(np.float32, True), (np.float32, False), (np.float64, False)])
def test_c3_twin_elastic_gradient(dtype, fast_f32):
    import swb200 as S
    from test_gpu_elastic import product_inputs as ela_inputs

    case = twins.c3_twin(n=(512, 256), nt=400, dtype=dtype)
    cf = case["check_freq"]
    observed = [np.zeros((case["nt"], 2, case["shots"][0]["rec_positions"].shape[0]), dtype=dtype, order="F")]
    params, matprop, shots, misfit, runparams, gradparams = ela_inputs(case, observed=observed, check_freq=cf, mute_src=3, fast_f32=fast_f32)
    grad, mis = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    (gref, mref), sref = EC.oracle_gradient(case, observed, check_freq=cf, mute_src=3)
    errs = {"seis": rel_l2(shots[0].recs.seismograms, sref[0]), "misfit": abs(float(mis) - float(mref)) / abs(float(mref))}
    errs.update({k: rel_l2(grad[k], gref[k]) for k in ("rho", "lambda", "mu")})
    _report(f"C3 twin 512x256 nt=400 cf={cf} {np.dtype(dtype).name} fast_f32={fast_f32}", errs)
    for k in ("rho", "lambda", "mu"):
        assert np.max(np.abs(gref[k])) > 0
    for k, e in errs.items():
        assert e <= tol(dtype), (k, e)


def test_c2_full_size_shot_matches_oracle():
    """The benchmarked shot itself: bench.c2_problem() at 4096^2, nt = 1000, check_freq = 31, source 32 of 64, 512 receivers, in the
    benchmarked arithmetic (fast_f32) -- seismograms, misfit, vp and rho gradients against the OpenMP oracle.  The default
    (promoted) mode is compared with the same oracle run."""
    case = twins.c2_twin(n=4096, nt=1000, nrec=512, shot_index=32, f0=8.0)
    cf = case["check_freq"]
    assert cf == 31
    t0 = time.time()
    observed = [np.zeros((case["nt"], 512), dtype=np.float32, order="F")]
    (gref, mref), sref, _ = cases.oracle_gradient(case, observed, check_freq=cf, mute_src=3)
    t_oracle = time.time() - t0
    assert np.max(np.abs(gref["vp"])) > 0 and float(mref) > 0
    for fast in (True, False):
        grad, mis, seis, _ = _acoustic_gradient(case, fast, cf)
        errs = {"seis": rel_l2(seis[0], sref[0]), "vp": rel_l2(grad["vp"], gref["vp"]), "rho": rel_l2(grad["rho"], gref["rho"]),
                "misfit": abs(mis - float(mref)) / abs(float(mref)), "oracle_seconds": t_oracle}
        _report(f"C2 FULL SIZE 4096^2 nt=1000 cf=31 fast_f32={fast}", errs)
        for k in ("seis", "vp", "rho", "misfit"):
            assert errs[k] <= tol(np.float32), (fast, k, errs[k])
