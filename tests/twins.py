"""Down-scaled twins of the BASELINE.json configurations C2 (2D VD 4096^2 F32), C3 (elastic 4096x2048) and C4 (3D CD 768^3):
the same model family, source / receiver layout, boundary set-up (C-PML halo 20, free surface), checkpoint rule
(check_freq = isqrt(nt)) and observed data (zeros) as the benchmarked workloads, at sizes the CPU oracle
finishes in seconds.  Test infrastructure: cases are the plain dicts tests/cases.py and tests/elastic_cases.py use.

The source frequency is raised with the inverse of the down-scaling so that the wavefield crosses a comparable part of the
(smaller) grid within the (fewer) time steps."""
from __future__ import annotations

import math

import numpy as np

from oracle import oracle as O


def c2_twin(n=512, nt=600, nrec=512, shot_index=32, f0=None, dtype=np.float32):
    """bench.c2_problem at grid size n: layered + seeded vp, Gardner density, VD CFL rule, one source at z = 2h, nrec receivers at z = 3h."""
    import bench

    prob = bench.c2_problem(n=n, nt=nt, nrec=nrec)
    T = np.dtype(dtype).type
    h = prob["h"]
    f0 = float(f0 if f0 is not None else prob["f0"] * min(8.0, 4096.0 / n))
    t = prob["t"]
    tf = (1000.0 * O.gaussderivstf(t, 2.0 / f0, f0)).reshape(nt, 1)
    sp = np.array([[prob["xs"][shot_index % len(prob["xs"])], 2 * h]])
    rp = np.stack([prob["xr"], np.full_like(prob["xr"], 3 * h)], axis=1)
    return dict(kind="acoustic_vd", n=(n, n), nt=nt, dt=prob["dt"], h=h, halo=prob["halo"], rcoef=1e-4, freetop=True, dtype=np.dtype(dtype),
                vp=np.asfortranarray(prob["vp"].astype(T)), rho=np.asfortranarray(prob["rho"].astype(T)),
                shots=[dict(src_positions=sp, src_tf=tf, domfreq=f0, rec_positions=rp)], observed="zeros", windows=False, diag_invcov=False, seed=1234,
                check_freq=max(2, math.isqrt(nt)))


def c4_twin(n=(96, 96, 96), nt=120, nrec_side=8, f0=30.0, dtype=np.float32):
    """vp = 1500 + 3000 z/zmax + seed-1236 perturbation, dt = 0.99 h / (vmax sqrt 3), source centre-top, surface receiver grid."""
    T = np.dtype(dtype).type
    h = 10.0
    rng = np.random.default_rng(1236)
    depth = np.arange(n[2], dtype=np.float64) / (n[2] - 1)
    vp = np.clip(1500.0 + 3000.0 * depth[None, None, :] + rng.normal(0.0, 30.0, size=n), 1400.0, 4700.0)
    vp = np.asfortranarray(vp.astype(T))
    dt = 0.99 * h / (float(vp.max()) * math.sqrt(3.0))
    t = np.arange(nt) * dt
    tf = (1000.0 * O.rickerstf(t, 1.2 / f0, f0)).reshape(nt, 1)
    ext = [(n[d] - 1) * h for d in range(3)]
    sp = np.array([[0.5 * ext[0], 0.5 * ext[1], 2 * h]])
    gx, gy = np.meshgrid(np.linspace(0.25, 0.75, nrec_side) * ext[0], np.linspace(0.25, 0.75, nrec_side) * ext[1], indexing="ij")
    rp = np.stack([gx.ravel(), gy.ravel(), np.full(gx.size, 3 * h)], axis=1)
    return dict(kind="acoustic_cd", n=tuple(n), nt=nt, dt=dt, h=h, halo=20, rcoef=1e-4, freetop=True, dtype=np.dtype(dtype), vp=vp,
                rho=np.asfortranarray((310.0 * vp.astype(np.float64) ** 0.25).astype(T)),
                shots=[dict(src_positions=sp, src_tf=tf, domfreq=f0, rec_positions=rp)], observed="zeros", windows=False, diag_invcov=False, seed=1236,
                check_freq=max(2, math.isqrt(nt)))


def c3_twin(n=(512, 256), nt=400, nrec=10, f0=40.0, dtype=np.float32):
    """h = 4.5 m, vp = 2000 + h (j-1) + seed-1235 perturbation <= 2 %, vs = vp / sqrt 3, rho = 2100, dt = 0.99 x the 7/6-CFL limit,
    one off-grid moment-tensor source (+0.124 m), vector receivers at z = 3h - 0.324 m."""
    T = np.dtype(dtype).type
    h = 4.5
    nx, nz = n
    rng = np.random.default_rng(1235)
    vp = (2000.0 + h * np.arange(nz, dtype=np.float64))[None, :] * (1.0 + np.clip(rng.normal(0.0, 0.007, size=n), -0.02, 0.02))
    vs = vp / math.sqrt(3.0)
    rho = np.full(n, 2100.0)
    mu = vs**2 * rho
    lam = vp**2 * rho - 2.0 * mu
    dt = 0.99 * (6.0 / 7.0) * h / (float(vp.max()) * math.sqrt(2.0))
    t = np.arange(nt) * dt
    tf = O.rickerstf(t, 1.2 / f0, f0).reshape(nt, 1)
    ext = [(nx - 1) * h, (nz - 1) * h]
    sp = np.array([[0.5 * ext[0] + 0.124, 0.3 * ext[1] + 0.124]])
    rp = np.stack([np.linspace(0.2, 0.8, nrec) * ext[0] - 0.324, np.full(nrec, 3 * h - 0.324)], axis=1)
    mt = np.array([[5e10, 5e10, 0.89e10]])
    return dict(kind="elastic_iso", n=tuple(n), nt=nt, dt=dt, h=h, halo=20, rcoef=1e-4, freetop=True, dtype=np.dtype(dtype),
                rho=np.asfortranarray(rho.astype(T)), lam=np.asfortranarray(lam.astype(T)), mu=np.asfortranarray(mu.astype(T)),
                shots=[dict(kind="momten", src_positions=sp, src_tf=tf, momtens=mt, domfreq=f0, rec_positions=rp)], seed=1235,
                check_freq=max(2, math.isqrt(nt)))
