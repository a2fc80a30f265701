"""Fixtures re-enacted from the reference's test/utils/setup_models.jl (test infrastructure)."""
import numpy as np

from oracle import oracle as O


def trapz(t, y):
    return float(np.sum((t[1:] - t[:-1]) * (y[1:] + y[:-1]) / 2))


def setup_constant_vel_cpml(nt, dt, n, d, c0, f0, halo, rcoef, dtype=np.float64):
    """setup_models.jl:24-50 (1D), 78-103 (2D), 155-179 (3D): source at centre, receiver at L/3."""
    N = len(n)
    L = [(n[k] - 1) * d[k] for k in range(N)]
    params = O.Params(nt=nt, dt=dt, gridsize=tuple(n), spacing=tuple(d), halo=halo, rcoef=rcoef, freetop=False, dtype=dtype)
    t0 = 2 / f0
    times = np.arange(nt) * dt
    possrcs = np.array([[L[k] / 2 for k in range(N)]])
    posrecs = np.array([[L[0] / 3] + [L[k] / 2 for k in range(1, N)]])
    srctf = np.asfortranarray(O.rickerstf(times, t0, f0).reshape(nt, 1))
    shot = O.ScalarShot(src_positions=possrcs, src_tf=srctf, domfreq=f0, rec_positions=posrecs)
    misfit = O.L2Misfit(observed=srctf.copy(order="F"), invcov=None)
    vp = np.asfortranarray(c0 * np.ones(n))
    return params, [shot], [misfit], vp


def setup_constant_vel_rho_cpml(nt, dt, n, d, c0, rho0, t0, f0, halo, rcoef, dtype=np.float64):
    """setup_models.jl:52-76 (1D), 105-129 (2D)."""
    N = len(n)
    L = [(n[k] - 1) * d[k] for k in range(N)]
    params = O.Params(nt=nt, dt=dt, gridsize=tuple(n), spacing=tuple(d), halo=halo, rcoef=rcoef, freetop=False, dtype=dtype)
    times = np.arange(nt) * dt
    possrcs = np.array([[L[k] / 2 for k in range(N)]])
    posrecs = np.array([[L[0] / 3] + [L[k] / 2 for k in range(1, N)]])
    srctf = np.asfortranarray(O.gaussderivstf(times, t0, f0).reshape(nt, 1))
    shot = O.ScalarShot(src_positions=possrcs, src_tf=srctf, domfreq=f0, rec_positions=posrecs)
    misfit = O.L2Misfit(observed=srctf.copy(order="F"), invcov=None)
    vp = np.asfortranarray(c0 * np.ones(n))
    rho = np.asfortranarray(rho0 * np.ones(n))
    return params, [shot], [misfit], vp, rho


def analytical_cd(N, c0, dt, nt, shot):
    """setup_models.jl:181-199 (1D), 221-239 (2D), 261-280 (3D)."""
    times = dt + np.arange(nt) * dt
    dist = float(np.linalg.norm(shot.src_positions[0] - shot.rec_positions[0]))
    src = c0**2 * shot.src_tf[:, 0]
    G = np.zeros(nt)
    for it in range(nt):
        if times[it] - dist / c0 >= 0:
            if N == 1:
                G[it] = 1.0 / (2 * c0)
            elif N == 2:
                G[it] = 1.0 / (2 * np.pi * c0**2 * np.sqrt(times[it] ** 2 - dist**2 / c0**2))
            else:
                G[it] = 1.0 / (4 * np.pi * c0**2 * dist)
                break
    Gc = np.convolve(G, src * dt if N < 3 else src)[:nt]
    return times, Gc


def analytical_vd(N, c0, rho, dt, nt, t0, f0, shot):
    """setup_models.jl:201-219 (1D), 241-259 (2D)."""
    times = dt / 2 + np.arange(nt) * dt
    dist = float(np.linalg.norm(shot.src_positions[0] - shot.rec_positions[0]))
    src = (c0**2 * rho) * O.rickerstf(times, t0, f0)
    G = np.zeros(nt)
    for it in range(nt):
        if times[it] - dist / c0 >= 0:
            G[it] = 1.0 / (2 * c0) if N == 1 else 1.0 / (2 * np.pi * c0**2 * np.sqrt(times[it] ** 2 - dist**2 / c0**2))
    Gc = np.convolve(G, src * dt)[:nt]
    return times, Gc
