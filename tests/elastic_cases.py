"""Seeded elastic P-SV problems shared by the oracle tests and the GPU parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O
from oracle import oracle_elastic as OE


def elastic_case(n=(96, 80), nt=140, halo=8, freetop=True, dtype=np.float64, kind="momten", nshots=1, nsrc=1, nrec=4, seed=0, h=5.0, f0=14.0,
                 homogeneous=False, ongrid=False):
    rng = np.random.default_rng(seed)
    T = np.dtype(dtype).type
    nx, nz = n
    depth = np.arange(nz) / max(nz - 1, 1)
    vp = 2000.0 + (0.0 if homogeneous else 600.0) * depth[None, :] + (0.0 if homogeneous else 1.0) * rng.normal(0, 20.0, size=n)
    vs = vp / np.sqrt(3.0)
    rho = 2100.0 + (0.0 if homogeneous else 1.0) * rng.normal(0, 15.0, size=n)
    mu = vs**2 * rho
    lam = vp**2 * rho - 2 * mu
    vmax = float(vp.max())
    dt = 0.85 * h / (vmax * np.sqrt(2.0)) * 6.0 / 7.0
    extent = [(nx - 1) * h, (nz - 1) * h]
    t = np.arange(nt) * dt
    shots = []
    for s in range(nshots):
        sp = np.zeros((nsrc, 2))
        sp[:, 0] = rng.uniform(0.35 * extent[0], 0.65 * extent[0], size=nsrc)
        sp[:, 1] = rng.uniform(0.35 * extent[1], 0.6 * extent[1], size=nsrc)
        rp = np.zeros((nrec, 2))
        rp[:, 0] = rng.uniform(0.25 * extent[0], 0.75 * extent[0], size=nrec)
        rp[:, 1] = rng.uniform(2.2 * h, 6.3 * h, size=nrec) if freetop else rng.uniform(0.25 * extent[1], 0.4 * extent[1], size=nrec)
        if ongrid:
            sp = np.round(sp / h) * h
            rp = np.round(rp / h) * h
        if kind == "momten":
            tf = np.zeros((nt, nsrc))
            for k in range(nsrc):
                tf[:, k] = O.rickerstf(t, 1.2 / f0 + 0.004 * k, f0)
            mt = np.tile(np.array([[5e10, 4e10, 0.89e10]]), (nsrc, 1)) * (1 + 0.1 * np.arange(nsrc))[:, None]
            shots.append(dict(kind=kind, src_positions=sp, src_tf=tf, momtens=mt, domfreq=f0, rec_positions=rp))
        else:
            tf = np.zeros((nt, 2, nsrc))
            for k in range(nsrc):
                tf[:, 0, k] = 2e9 * O.rickerstf(t, 1.2 / f0, f0)
                tf[:, 1, k] = 3e9 * O.rickerstf(t, 1.2 / f0 + 0.003, f0)
            shots.append(dict(kind=kind, src_positions=sp, src_tf=tf, domfreq=f0, rec_positions=rp))
    return dict(kind="elastic_iso", n=tuple(n), nt=nt, dt=dt, h=h, halo=halo, rcoef=1e-4, freetop=freetop, dtype=np.dtype(dtype), rho=np.asfortranarray(rho.astype(T)),
                lam=np.asfortranarray(lam.astype(T)), mu=np.asfortranarray(mu.astype(T)), shots=shots, seed=seed)


def params_oracle(case):
    return O.Params(nt=case["nt"], dt=case["dt"], gridsize=case["n"], spacing=(case["h"], case["h"]), halo=case["halo"], rcoef=case["rcoef"],
                    freetop=case["freetop"], dtype=case["dtype"].type)


def oracle_shots(case):
    T = case["dtype"].type
    out = []
    for s in case["shots"]:
        sp, rp = np.asfortranarray(s["src_positions"].astype(T)), np.asfortranarray(s["rec_positions"].astype(T))
        tf = np.asfortranarray(s["src_tf"].astype(T))
        if s["kind"] == "momten":
            out.append(OE.MomentTensorShot(src_positions=sp, src_tf=tf, momtens=s["momtens"].astype(T), domfreq=T(s["domfreq"]), rec_positions=rp))
        else:
            out.append(OE.ExternalForceShot(src_positions=sp, src_tf=tf, domfreq=T(s["domfreq"]), rec_positions=rp))
    return out


def matprops(case):
    return [case["rho"], case["lam"], case["mu"]]


def oracle_forward(case, sincinterp=True, snapevery=None):
    sim = O.build_wavesim("elastic_iso", params_oracle(case), sincinterp=sincinterp)
    shots = oracle_shots(case)
    snaps = O.swforward(sim, matprops(case), shots, snapevery=snapevery)
    return [s.seismograms for s in shots], snaps


def make_observed(case, seismograms):
    T = case["dtype"].type
    rng = np.random.default_rng(case["seed"] + 77)
    out = []
    for seis in seismograms:
        scale = float(np.max(np.abs(seis))) or 1.0
        out.append(np.asfortranarray((0.7 * seis + 0.05 * scale * rng.standard_normal(seis.shape)).astype(T)))
    return out


def oracle_gradient(case, observed, check_freq=1, mute_src=0, mute_rec=0, sincinterp=True, matprop=None):
    sim = O.build_wavesim("elastic_iso", params_oracle(case), gradient=True, check_freq=check_freq, sincinterp=sincinterp)
    shots = oracle_shots(case)
    mis = [O.L2Misfit(observed=o) for o in observed]
    res = O.swgradient(sim, matprop or matprops(case), shots, mis, mute_radius_src=mute_src, mute_radius_rec=mute_rec, compute_misfit=True)
    return res, [s.seismograms for s in shots]
