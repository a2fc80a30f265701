"""Size-independent properties at BASELINE.json's full grid sizes (the CPU oracle is unaffordable there): checkpointed ==
non-checkpointed gradients (the reference's own criterion, test/test_gradient_*.jl), exact linearity of the forward operator in the
source amplitude for power-of-two factors, receiver-permutation symmetry (test/test_forward_constant_density.jl:42-266) and
fused == one-launch-per-reference-kernel seismograms.  Few time steps: the properties do not depend on nt."""
import numpy as np
import pytest

from cases import acoustic_case, make_observed, product_inputs
from elastic_cases import elastic_case

pytestmark = pytest.mark.gpu


def _near_surface_geometry(case, nrec_depth=3.0, src_depth=12.0):
    """sources and receivers within a few cells of each other near the top centre, so that a few dozen steps produce a normal-range
    signal at every receiver (the finite-difference front advances about one cell per step)"""
    n, h = case["n"], case["h"]
    for sh in case["shots"]:
        ns, nr = sh["src_positions"].shape[0], sh["rec_positions"].shape[0]
        sh["src_positions"][:, 0] = (n[0] // 2 + np.linspace(-2.0, 2.0, ns)) * h
        sh["src_positions"][:, -1] = src_depth * h
        sh["rec_positions"][:, 0] = (n[0] // 2 + np.linspace(-7.0, 7.0, nr)) * h
        sh["rec_positions"][:, -1] = nrec_depth * h
        for d in range(1, len(n) - 1):  # 3D: middle axis
            sh["src_positions"][:, d] = (n[d] // 2) * h
            sh["rec_positions"][:, d] = (n[d] // 2 + np.linspace(-3.0, 3.0, nr)) * h


def _acoustic_forward(case, **kw):
    import swb200 as S

    params, matprop, shots, _, runparams, _ = product_inputs(case, **kw)
    S.swforward(params, matprop, shots, runparams=runparams)
    return [s.recs.seismograms.copy() for s in shots]


def _acoustic_gradient(case, observed, **kw):
    import swb200 as S

    params, matprop, shots, misfit, runparams, gradparams = product_inputs(case, observed=observed, **kw)
    return S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)


@pytest.mark.parametrize("kind,n,nt", [("acoustic_vd", (4096, 4096), 64), ("acoustic_cd", (4096, 4096), 64), ("acoustic_cd", (768, 768, 768), 24)])
def test_acoustic_full_size_properties(kind, n, nt):
    """C2 (2D VD 4096^2), C1 physics at 4096^2 and C4 (3D CD 768^3), Float32, halo 20"""
    case = acoustic_case(kind=kind, n=n, nt=nt, halo=20, freetop=True, dtype=np.float32, seed=41, nshots=1, nsrc=2, nrec=12, f0=14.0)
    _near_surface_geometry(case, src_depth=6.0 if len(n) == 3 else 12.0)
    base = _acoustic_forward(case)[0]
    assert np.max(np.abs(base)) > 0 and np.all(np.isfinite(base))
    # the fused engine performs the operations of the one-launch-per-reference-kernel path, at full size too
    assert np.array_equal(_acoustic_forward(case, fused=False)[0], base)
    # linearity: a source 4x as strong gives exactly 4x the seismograms (power of two: every product and sum scales exactly)
    c4 = dict(case, shots=[dict(sh, src_tf=4.0 * sh["src_tf"]) for sh in case["shots"]])
    assert np.array_equal(_acoustic_forward(c4)[0], np.float32(4.0) * base)
    # receiver permutation: the traces follow their receivers
    perm = np.random.default_rng(5).permutation(base.shape[1])
    cp = dict(case, shots=[dict(sh, rec_positions=sh["rec_positions"][perm]) for sh in case["shots"]])
    assert np.array_equal(_acoustic_forward(cp)[0], base[:, perm])
    # checkpointed == non-checkpointed gradient and misfit
    observed = make_observed(case, [base])
    (g1, m1), (g7, m7) = _acoustic_gradient(case, observed, check_freq=1), _acoustic_gradient(case, observed, check_freq=7)
    for k in g1:
        assert np.max(np.abs(g1[k])) > 0 and np.all(np.isfinite(g1[k]))
        assert np.array_equal(g1[k], g7[k]), k
    assert m1 == m7


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_elastic_full_size_properties(dtype):
    """C3 (2D elastic P-SV 4096 x 2048, halo 20, free surface, off-grid moment-tensor source)"""
    import swb200 as S
    from test_gpu_elastic import product_inputs as ela_inputs

    case = elastic_case(n=(4096, 2048), nt=48, halo=20, freetop=True, dtype=dtype, kind="momten", nshots=1, nsrc=1, nrec=8, seed=43)
    _near_surface_geometry(case, nrec_depth=3.3, src_depth=9.4)

    def forward(c, fused=True):
        params, matprop, shots, _, runparams, _ = ela_inputs(c, fused=fused)
        S.swforward(params, matprop, shots, runparams=runparams)
        return shots[0].recs.seismograms.copy()

    base = forward(case)
    assert np.max(np.abs(base)) > 0 and np.all(np.isfinite(base))
    assert np.array_equal(forward(case, fused=False), base)  # stresses on chip == four sweeps, at full size
    c2 = dict(case, shots=[dict(sh, src_tf=2.0 * sh["src_tf"]) for sh in case["shots"]])
    assert np.array_equal(forward(c2), dtype(2.0) * base)
    observed = [np.zeros_like(base, order="F")]
    grads = {}
    for cf in (1, 6):
        params, matprop, shots, misfit, runparams, gradparams = ela_inputs(case, observed=observed, check_freq=cf)
        grads[cf] = S.swgradient(params, matprop, shots, misfit, runparams=runparams, gradparams=gradparams)
    (g1, m1), (g6, m6) = grads[1], grads[6]
    for k in g1:
        assert np.max(np.abs(g1[k])) > 0 and np.all(np.isfinite(g1[k]))
        assert np.array_equal(g1[k], g6[k]), k
    assert m1 == m6
