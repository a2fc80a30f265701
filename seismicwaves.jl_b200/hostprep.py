"""Host-side preparation that the reference performs in Julia before handing arrays to a backend:
C-PML coefficient profiles, nearest-grid-point lookup, source-time-function scaling, Kaiser-sinc
spreading.  All of it is O(halo) / O(nt * nsrc) work; the O(N) steps (material factors, gradient
post-processing) run on the device inside libswb200.

In a Julia deployment these steps stay in the reference's own code; this module is the
Python stand-in so that the same C ABI can be driven end-to-end from tests and bench.py.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

# ---- source time functions (src/utils/utils.jl:6-20) ------------------------------------------------


def rickerstf(t, t0, f0):
    a = (np.pi * f0 * (t - t0)) ** 2
    return (1 - 2 * a) * np.exp(-a)


def gaussderivstf(t, t0, f0):
    return (t - t0) * np.exp(-((np.pi * f0 * (t - t0)) ** 2))


def gaussstf(t, t0, f0):
    return -np.exp(-((np.pi * f0 * (t - t0)) ** 2)) / (2 * (np.pi * f0) ** 2)


def distribsrcs(nsrc: int, nw: int) -> List[range]:
    """src/utils/utils.jl:28-45: contiguous shot groups, the first `nsrc % nw` workers get one extra."""
    if nsrc >= nw:
        sizes = [nsrc // nw + (1 if k < nsrc % nw else 0) for k in range(nw)]
    else:
        sizes = [1] * nsrc
    groups, first = [], 0
    for s in sizes:
        groups.append(range(first, first + s))
        first += s
    return groups


# ---- C-PML coefficient profiles (src/models/cpmlcoeffs.jl:18-100) -------------------------------------


def _lerp_range(start: float, stop: float, n: int) -> np.ndarray:
    # Base.LinRange element j (0-based): (1 - j/d)*start + (j/d)*stop with d = max(n-1, 1)
    d = max(n - 1, 1)
    t = np.arange(n, dtype=np.float64) / d
    return (1 - t) * start + t * stop


def _profile(halo: int, dt, d0, alpha_max, half: bool, T) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    size = halo if half else halo + 1
    shift = 0.5 if half else 0.0
    dist = _lerp_range(shift, size + shift - 1, size) if size > 0 else np.zeros(0)
    scale = float(halo) if halo != 0 else 1.0
    out = []
    for nd in (dist[::-1] / scale, dist / scale):
        d = float(d0) * (nd * nd)  # normdist .^ npower with npower = 2
        alpha = float(alpha_max) * (1.0 - nd)
        b = np.exp(-(d / 1.0 + alpha) * float(dt))
        with np.errstate(invalid="ignore", divide="ignore"):
            a = d * (b - 1.0) / (1.0 * (d + 1.0 * alpha))
        out.append((a.astype(T), b.astype(T)))
    (a_l, b_l), (a_r, b_r) = out
    return a_l, a_r, b_l, b_r


def cpml_coefficients_axis(vel_max, dt, halo: int, rcoef, thickness, f0, dtype):
    """compute_CPML_coefficientsAxis! (cpmlcoeffs.jl:18-41): returns (a, a_h, b, b_h) as vcat(left, right)."""
    T = np.dtype(dtype).type
    vel_max, dt, rcoef, thickness, f0 = T(vel_max), T(dt), T(rcoef), T(thickness), T(f0)
    alpha_max = T(T(np.pi) * f0)
    npower = T(2.0)
    if halo == 0:
        d0 = T(0.0)
    else:
        num = T(T(-(npower + T(1))) * vel_max) * T(np.log(rcoef))
        d0 = T(np.float64(num) / (2.0 * np.float64(thickness)))
    a_l, a_r, b_l, b_r = _profile(halo, dt, d0, alpha_max, False, T)
    ah_l, ah_r, bh_l, bh_r = _profile(halo, dt, d0, alpha_max, True, T)
    return np.concatenate([a_l, a_r]), np.concatenate([ah_l, ah_r]), np.concatenate([b_l, b_r]), np.concatenate([bh_l, bh_r])


def init_bdc(vel_max, dt, halo: int, rcoef, spacing: Sequence, freeboundtop: bool, domfreq, dtype):
    """init_bdc! (src/models/acoustic/acou_init_bc.jl:7-40, src/models/elastic/ela_init_bc.jl:7-40):
    one (a, a_h, b, b_h) tuple per axis, free-surface override on the first half of the last axis."""
    T = np.dtype(dtype).type
    axes = []
    for n in range(len(spacing)):
        axes.append(list(cpml_coefficients_axis(vel_max, dt, halo, rcoef, T(T(spacing[n]) * T(halo)), domfreq, T)))
    if freeboundtop and axes:
        a, a_h, b, b_h = axes[-1]
        a[: len(a) // 2] = T(0.0)
        a_h[: len(a_h) // 2] = T(0.0)
        b[: len(b) // 2] = T(1.0)
        b_h[: len(b_h) // 2] = T(1.0)
    return axes


# ---- positions and scaling ---------------------------------------------------------------------------


def find_nearest_grid_points(positions: np.ndarray, spacing: Sequence, dtype) -> np.ndarray:
    """src/utils/utils.jl:47-58: round(pos/spacing + 1, RoundNearestTiesUp) in T; (npos, N) int64, 1-based, column-major."""
    T = np.dtype(dtype).type
    pos = np.asarray(positions, dtype=T)
    out = np.zeros(pos.shape, dtype=np.int64, order="F")
    for d in range(pos.shape[1]):
        tmp = (pos[:, d] / T(spacing[d]) + T(1)).astype(T).astype(np.float64)
        fl = np.floor(tmp)
        out[:, d] = (fl + (tmp - fl >= 0.5)).astype(np.int64)
    return out


def scale_stf_acoustic_cd(tf: np.ndarray, spacing: Sequence, dt, vp: np.ndarray, possrcs: np.ndarray, dtype) -> np.ndarray:
    """possrcrec_scaletf (src/models/acoustic/acou_forward.jl:6-20): tf ./ prod(spacing) .* dt^2 .* vp[src]^2."""
    T = np.dtype(dtype).type
    prod = T(np.prod(np.array(spacing, dtype=T), dtype=T))
    dt = T(dt)
    out = np.asfortranarray(((np.asarray(tf, dtype=T) / prod).astype(T) * T(dt * dt)).astype(T))
    for s in range(out.shape[1]):
        v = vp[tuple(possrcs[s, :] - 1)]
        out[:, s] = (out[:, s] * T(v * v)).astype(T)
    return out


def scale_stf_acoustic_vd(tf: np.ndarray, spacing: Sequence, dt, vp: np.ndarray, rho: np.ndarray, possrcs: np.ndarray, dtype) -> np.ndarray:
    """possrcrec_scaletf (src/models/acoustic/acou_forward.jl:67-81): tf ./ prod(spacing) .* dt .* (vp[src]^2 * rho[src])."""
    T = np.dtype(dtype).type
    prod = T(np.prod(np.array(spacing, dtype=T), dtype=T))
    out = np.asfortranarray(((np.asarray(tf, dtype=T) / prod).astype(T) * T(dt)).astype(T))
    for s in range(out.shape[1]):
        q = tuple(possrcs[s, :] - 1)
        v = vp[q]
        out[:, s] = (out[:, s] * T(T(v * v) * rho[q])).astype(T)
    return out


# ---- Kaiser-windowed sinc spreading (src/utils/utils.jl:71-214) -----------------------------------------


def _kaiser(x, r, beta, T):
    """kaiser(x, r, β) = besseli(0, β sqrt(1 - (x/r)^2)) / besseli(0, β) inside [-r, r], else 0.0 (utils.jl:71).
    The reference evaluates it in T (SpecialFunctions.besseli has Float32 methods); scipy's i0 is evaluated in double and rounded to T."""
    from scipy.special import i0

    if -r <= x <= r:
        q = T(x / r)
        arg = T(beta * T(np.sqrt(T(T(1) - T(q * q)))))
        return T(T(i0(float(arg))) / T(i0(float(beta))))
    return 0.0


def _sinc(x, T):
    x = T(x)
    if x == 0:
        return T(1)
    px = T(T(np.pi) * x)
    return T(np.sin(px) / px)


def coeffsinc1d(x0, dx, nx: int, r: int, beta, xstart, mirror: bool, xbl, xbr, dtype):
    """coeffsinc1D (utils.jl:105-154).  Returns (idxs, coeffs) with 1-based indices, duplicates summed, in ascending index
    order (the reference collects the keys of a Dict, whose order is unspecified and irrelevant beyond rounding)."""
    T = np.dtype(dtype).type
    x0, dx, xstart, xbl, xbr, beta = T(x0), T(dx), T(xstart), T(xbl), T(xbr), T(beta)
    xs = (xstart + np.arange(nx, dtype=np.float64) * np.float64(dx)).astype(T)  # range(xstart; length, step): exact products rounded to T

    def findnearest(x):
        return int(np.argmin(np.abs(T(x) - xs))) + 1

    i0_ = findnearest(x0)
    pts = []
    for idx in range(i0_ - r - 1, i0_ + r + 2):
        xcurr = T(T(T(idx - 1) * dx) + xstart)
        coe = _kaiser(T(xcurr - x0), T(T(r) * dx), beta, T) * _sinc(T(T(xcurr - x0) / dx), T)
        if not abs(float(coe)) <= 1e-15:
            pts.append((idx, T(coe)))
    acc = {}
    for idx, coe in pts:
        xcurr = T(T(T(idx - 1) * dx) + xstart)
        if xcurr < xbl:
            j = findnearest(T(xbl + T(xbl - xcurr)))
            c = T(-coe) if mirror else coe
        elif xcurr > xbr:
            j = findnearest(T(xbr - T(xcurr - xbr)))
            c = T(-coe) if mirror else coe
        else:
            j, c = idx, coe
        acc[j] = T(acc[j] + c) if j in acc else T(c)
    keys = sorted(acc)
    return keys, [acc[k] for k in keys]


def spread_positions(gridsize: Sequence[int], spacing: Sequence, positions: np.ndarray, shift: Sequence, mirror: bool, freesurfposition: str,
                     dtype, r: int = 4, beta: float = 6.31):
    """spread_positions (utils.jl:168-214): per position, tensor product of the 1-D band-limited deltas.
    Returns (list of (npts, N) int arrays [1-based], list of (npts,) coefficient arrays)."""
    T = np.dtype(dtype).type
    N = len(gridsize)
    if freesurfposition == "halfgridin":
        fs = [T(T(s) / T(2)) for s in spacing]
    elif freesurfposition == "ongridbound":
        fs = [T(0)] * N
    else:
        raise ValueError(f"spread_positions(): Wrong keyword argument freesurfposition {freesurfposition}")
    extent = [T(T(spacing[d]) * T(gridsize[d] - 1)) for d in range(N)]
    idxs, coefs = [], []
    for p in range(positions.shape[0]):
        per_dim = [coeffsinc1d(positions[p, d], spacing[d], gridsize[d], r, T(beta), shift[d], mirror, fs[d], extent[d], T) for d in range(N)]
        lens = [len(pd[0]) for pd in per_dim]
        tot = int(np.prod(lens))
        ij = np.zeros((tot, N), dtype=np.int64, order="F")
        cf = np.zeros(tot, dtype=T)
        # Iterators.product: first dimension fastest
        for flat in range(tot):
            rem = flat
            c = T(1)
            for d in range(N):
                k = rem % lens[d]
                rem //= lens[d]
                ij[flat, d] = per_dim[d][0][k]
                c = T(c * per_dim[d][1][k])
            cf[flat] = c
        idxs.append(ij)
        coefs.append(cf)
    return idxs, coefs
