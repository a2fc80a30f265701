"""Seeded synthetic problems shared by the oracle tests and the GPU parity tests (test infrastructure).

A case is a plain dict of numpy arrays and scalars; `to_oracle` / `to_product` turn it into the inputs of the
CPU oracle (oracle/oracle.py) and of the product's reference-style API (swb200).
"""
from __future__ import annotations

import numpy as np

from oracle import oracle as O


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a64, b64 = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b64.ravel())
    num = np.linalg.norm((a64 - b64).ravel())
    return float(num / den) if den > 0 else float(num)


def tol(dtype) -> float:
    """north_star tolerance: relative L2 <= 1e-5 in Float64, 1e-4 in Float32."""
    return 1e-4 if np.dtype(dtype) == np.float32 else 1e-5


def _positions(rng, nsrc, extent, lo_frac=0.25, hi_frac=0.75, ongrid=None):
    N = len(extent)
    p = np.zeros((nsrc, N))
    for d in range(N):
        p[:, d] = rng.uniform(lo_frac * extent[d], hi_frac * extent[d], size=nsrc)
    return p


def acoustic_case(kind="acoustic_cd", n=(64, 56), nt=120, halo=8, freetop=True, dtype=np.float64, nshots=2, nsrc=2, nrec=5, seed=0, h=10.0,
                  f0=12.0, observed="perturbed", windows=False, diag_invcov=False):
    """Heterogeneous layered model + seeded noise; sources in the interior, receivers near the top."""
    rng = np.random.default_rng(seed)
    N = len(n)
    T = np.dtype(dtype).type
    depth = np.arange(n[-1]) / max(n[-1] - 1, 1)
    shape_b = (1,) * (N - 1) + (n[-1],)
    vp = 2000.0 + 1500.0 * depth.reshape(shape_b) + rng.normal(0, 40.0, size=n)
    vp = np.clip(vp, 1500.0, 3800.0)
    rho = 310.0 * vp**0.25
    vmax = float(vp.max())
    cfl = (6.0 / 7.0 if kind == "acoustic_vd" else 1.0) * h / (vmax * np.sqrt(N))
    dt = 0.9 * cfl
    extent = [(n[d] - 1) * h for d in range(N)]
    t = np.arange(nt) * dt
    shots = []
    for s in range(nshots):
        sp = _positions(rng, nsrc, extent, 0.3, 0.7)
        rp = _positions(rng, nrec, extent, 0.2, 0.8)
        rp[:, -1] = rng.uniform(2 * h, 5 * h, size=nrec) if freetop else rp[:, -1]
        if s == 0 and nsrc >= 2:
            sp[1, :] = sp[0, :]  # two sources in the same cell: exercises the deterministic injection order
        tf = np.zeros((nt, nsrc))
        for k in range(nsrc):
            t0 = 1.2 / f0 + 0.01 * k
            tf[:, k] = (1000.0 * (1 + 0.3 * k)) * (O.rickerstf(t, t0, f0) if kind == "acoustic_cd" else O.gaussderivstf(t, t0, f0))
        shots.append(dict(src_positions=sp, src_tf=tf, domfreq=f0, rec_positions=rp))
    case = dict(kind=kind, n=tuple(n), nt=nt, dt=dt, h=h, halo=halo, rcoef=1e-4, freetop=freetop, dtype=np.dtype(dtype), vp=np.asfortranarray(vp.astype(T)),
                rho=np.asfortranarray(rho.astype(T)), shots=shots, observed=observed, windows=windows, diag_invcov=diag_invcov, seed=seed)
    return case


def case_params_oracle(case):
    N = len(case["n"])
    return O.Params(nt=case["nt"], dt=case["dt"], gridsize=case["n"], spacing=(case["h"],) * N, halo=case["halo"], rcoef=case["rcoef"],
                    freetop=case["freetop"], dtype=case["dtype"].type)


def matprop_list(case):
    return [case["vp"]] if case["kind"] == "acoustic_cd" else [case["vp"], case["rho"]]


def oracle_shots(case):
    T = case["dtype"].type
    return [O.ScalarShot(src_positions=np.asfortranarray(s["src_positions"].astype(T)), src_tf=np.asfortranarray(s["src_tf"].astype(T)),
                         domfreq=T(s["domfreq"]), rec_positions=np.asfortranarray(s["rec_positions"].astype(T))) for s in case["shots"]]


def make_observed(case, seismograms):
    """Observed data for the gradient tests, derived from the oracle's synthetics so both sides see identical arrays."""
    T = case["dtype"].type
    rng = np.random.default_rng(case["seed"] + 1000)
    out = []
    for seis in seismograms:
        if case["observed"] == "zeros":
            out.append(np.zeros_like(seis, order="F"))
        else:
            scale = float(np.max(np.abs(seis))) or 1.0
            out.append(np.asfortranarray((0.8 * seis + 0.05 * scale * rng.standard_normal(seis.shape)).astype(T)))
    return out


def misfit_kwargs(case, nt):
    kw = {}
    T = case["dtype"].type
    if case["windows"]:
        kw["windows"] = [(max(1, nt // 10), nt // 2), (nt // 2 + 5, nt - 3)]
    if case["diag_invcov"]:
        kw["invcov"] = (1.0 + 0.5 * np.sin(np.arange(nt) * 0.1)).astype(T)
    return kw


def oracle_forward(case, snapevery=None):
    sim = O.build_wavesim(case["kind"], case_params_oracle(case))
    shots = oracle_shots(case)
    snaps = O.swforward(sim, matprop_list(case), shots, snapevery=snapevery)
    return [s.seismograms for s in shots], snaps


def oracle_gradient(case, observed, check_freq=1, mute_src=0, mute_rec=0, compute_misfit=True, interp_method=None):
    kw = {}
    if interp_method is not None:
        kw["interp_method"] = interp_method
    sim = O.build_wavesim(case["kind"], case_params_oracle(case), gradient=True, check_freq=check_freq, **kw)
    shots = oracle_shots(case)
    mis = [O.L2Misfit(observed=obs, **misfit_kwargs(case, case["nt"])) for obs in observed]
    res = O.swgradient(sim, matprop_list(case), shots, mis, mute_radius_src=mute_src, mute_radius_rec=mute_rec, compute_misfit=compute_misfit)
    return res, [s.seismograms for s in shots], sim


# ---- product side -----------------------------------------------------------------------------------------


def product_inputs(case, observed=None, fast_f32=False, fused=True, check_freq=1, mute_src=0, mute_rec=0, compute_misfit=True, interp_method="arithmetic",
                   snapevery=None, graphs=True):
    import swb200 as S

    T = case["dtype"].type
    N = len(case["n"])
    bc = S.CPMLBoundaryConditionParameters(halo=case["halo"], rcoef=T(case["rcoef"]), freeboundtop=case["freetop"])
    params = S.InputParametersAcoustic(case["nt"], T(case["dt"]), case["n"], tuple(T(case["h"]) for _ in range(N)), bc, dtype=case["dtype"])
    if case["kind"] == "acoustic_cd":
        matprop = S.VpAcousticCDMaterialProperties(case["vp"])
    else:
        matprop = S.VpRhoAcousticVDMaterialProperties(case["vp"], case["rho"], interp_method=interp_method)
    shots = []
    for s in case["shots"]:
        srcs = S.ScalarSources(s["src_positions"].astype(T), s["src_tf"].astype(T), T(s["domfreq"]))
        recs = S.ScalarReceivers(s["rec_positions"].astype(T), case["nt"], dtype=case["dtype"])
        shots.append(S.ScalarShot(srcs=srcs, recs=recs))
    runparams = S.RunParameters(parall="B200", fast_f32=fast_f32, fused=fused, graphs=graphs, snapevery=snapevery, erroronPPW=False)
    gradparams = S.GradParameters(mute_radius_src=mute_src, mute_radius_rec=mute_rec, compute_misfit=compute_misfit, check_freq=check_freq)
    misfit = None
    if observed is not None:
        misfit = [S.L2Misfit(observed=obs, **misfit_kwargs(case, case["nt"])) for obs in observed]
    return params, matprop, shots, misfit, runparams, gradparams
