"""Known-answer checks of the elastic P-SV oracle (CPU): the reference's own criteria
(test/test_gradient_elastic_homogeneous.jl:19-119: non-zero rho/lambda/mu gradients, checkpointed == non-checkpointed)
plus physical sanity checks that pin the restatement (arrival times, adjoint-vs-finite-difference directional derivative)."""
import numpy as np
import pytest

from elastic_cases import elastic_case, make_observed, matprops, oracle_forward, oracle_gradient
from cases import rel_l2
from oracle import oracle_elastic as OE


def test_sinc_coefficients_partition_of_unity_and_on_grid_delta():
    T = np.float64
    # a position exactly on a grid node of an unshifted grid collapses to a single unit coefficient
    idx, co = OE.coeffsinc1d(50.0, 5.0, 100, 4, 6.31, 0.0, False, 0.0, 495.0, T)
    assert idx == [11] and abs(co[0] - 1.0) < 1e-12
    # off-grid: the band-limited delta sums to ~1
    idx, co = OE.coeffsinc1d(52.3, 5.0, 100, 4, 6.31, 0.0, False, 0.0, 495.0, T)
    assert len(idx) >= 8 and abs(sum(co) - 1.0) < 2e-2
    # mirrored over the left boundary with sign flip (stress sources near the free surface)
    idx_m, co_m = OE.coeffsinc1d(6.1, 5.0, 100, 4, 6.31, 0.0, True, 0.0, 495.0, T)
    assert min(idx_m) >= 1


def test_forward_p_wave_arrival_time_homogeneous():
    case = elastic_case(n=(140, 120), nt=260, halo=10, freetop=False, homogeneous=True, kind="momten", nrec=3, seed=3, ongrid=True)
    case["shots"][0]["momtens"][:] = [[1e10, 1e10, 0.0]]  # explosive source: pure P
    seis, _ = oracle_forward(case)
    s = case["shots"][0]
    vp, f0 = 2000.0, s["domfreq"]
    for r in range(s["rec_positions"].shape[0]):
        d = np.linalg.norm(s["rec_positions"][r] - s["src_positions"][0])
        amp = np.sqrt(seis[0][:, 0, r] ** 2 + seis[0][:, 1, r] ** 2)
        t_peak = np.argmax(amp) * case["dt"]
        t_expected = d / vp + 1.2 / f0
        assert abs(t_peak - t_expected) < 0.6 / f0, (t_peak, t_expected)
    assert np.isfinite(seis[0]).all()


@pytest.mark.parametrize("kind", ["momten", "extforce"])
def test_gradient_checkpoint_equivalence_and_nonzero(kind):
    case = elastic_case(n=(80, 64), nt=100, halo=6, kind=kind, seed=5, nrec=3)
    syn, _ = oracle_forward(case)
    obs = make_observed(case, syn)
    (g1, m1), s1 = oracle_gradient(case, obs, check_freq=1)
    (g2, m2), s2 = oracle_gradient(case, obs, check_freq=10)
    (g3, m3), _ = oracle_gradient(case, obs, check_freq=7)
    for k in ("rho", "lambda", "mu"):
        assert np.max(np.abs(g1[k])) > 0
        assert np.array_equal(g1[k], g2[k]) and np.array_equal(g1[k], g3[k])
    assert m1 == m2 == m3 and m1 > 0
    assert np.array_equal(s1[0], s2[0])


def test_gradient_matches_finite_difference_directional_derivative():
    """The correlated gradient is the derivative of the L2 misfit: compare <grad, dm> with a centred finite difference."""
    case = elastic_case(n=(64, 56), nt=110, halo=6, freetop=False, kind="extforce", seed=9, nrec=3)
    syn, _ = oracle_forward(case)
    obs = make_observed(case, syn)
    (g, m0), _ = oracle_gradient(case, obs, check_freq=1)
    rng = np.random.default_rng(0)
    nx, nz = case["n"]
    bump = np.zeros(case["n"])
    bump[20:44, 18:40] = 1.0
    for name, key, scale in (("mu", "mu", 2e-3), ("lambda", "lam", 2e-3), ("rho", "rho", 2e-3)):
        dm = np.asfortranarray(bump * scale * case[key].mean())
        mis = []
        for sgn in (+1, -1):
            mp = {k: case[k].copy(order="F") for k in ("rho", "lam", "mu")}
            mp[key] = np.asfortranarray(mp[key] + sgn * dm)
            (_, mm), _ = oracle_gradient(case, obs, check_freq=1, matprop=[mp["rho"], mp["lam"], mp["mu"]])
            mis.append(mm)
        fd = (mis[0] - mis[1]) / 2.0
        ad = float(np.sum(g[name] * dm))
        assert abs(fd - ad) <= 0.08 * abs(fd), (name, fd, ad)


def test_float32_close_to_float64():
    c64 = elastic_case(n=(72, 60), nt=90, halo=6, dtype=np.float64, seed=2)
    c32 = elastic_case(n=(72, 60), nt=90, halo=6, dtype=np.float32, seed=2)
    s64, _ = oracle_forward(c64)
    s32, _ = oracle_forward(c32)
    assert s32[0].dtype == np.float32
    assert rel_l2(s32[0], s64[0]) < 5e-3
