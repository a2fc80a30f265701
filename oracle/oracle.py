"""
ORACLE -- test infrastructure, NOT product code.

CPU restatement of SeismicWaves.jl's finite-difference hot path (reference at /root/reference,
v0.9.0).  The time-stepping kernels live in C (oracle/swref*.h, compiled by oracle/Makefile);
this module restates the host side that feeds them: CPML coefficient profiles, source/receiver
scaling, the LinearCheckpointer schedule, the L2 misfit, gradient post-processing and the shot loop.
Every function cites the reference file:line it follows.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product (seismicwaves.jl_b200/) never does.

Parity status: the reference is Julia-only and Julia is not installed here, and the reference ships
no golden vectors, so this oracle is pinned through the reference's own known-answer tests
(tests/test_oracle_*.py re-enact them: analytic Green's functions within 1 %, checkpointed ==
non-checkpointed gradients, receiver permutation symmetry, exact interpolation values).  Absolute
values on heterogeneous / free-surface models are NOT pinned by any reference test: "parity
unpinned" for those beyond line-by-line review.  Third-party arithmetic restated from its published
behaviour: Interpolations.jl BSpline(Linear()) at half indices (2-/4-point mean, Float64 weights),
SpecialFunctions.besseli(0, .) (scipy.special.i0), Base.LinRange (lerp).

Array layout everywhere: Julia column-major -> numpy arrays with order="F".
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# --------------------------------------------------------------------------------------
# C library loading
# --------------------------------------------------------------------------------------


def build(force: bool = False) -> None:
    """Compile oracle/libswref.so and libswref_omp.so with the committed Makefile."""
    # always defer to make: it rebuilds only when a source is newer than the libraries
    subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))


_LIBS: Dict[bool, C.CDLL] = {}


def lib(omp: bool = False) -> C.CDLL:
    if omp not in _LIBS:
        build()
        _LIBS[omp] = C.CDLL(os.path.join(_HERE, "libswref_omp.so" if omp else "libswref.so"))
    return _LIBS[omp]


_USE_OMP = False


def use_openmp(flag: bool) -> None:
    """Select the OpenMP build (CPU-baseline timing) or the serial build (default, parity)."""
    global _USE_OMP
    _USE_OMP = bool(flag)


def set_threads(n: int) -> int:
    """Thread count of the OpenMP build (bench.py's CPU legs); returns the count in effect."""
    L = lib(True)
    L.swref_set_num_threads(int(n))
    return int(L.swref_get_max_threads())


def _L() -> C.CDLL:
    return lib(_USE_OMP)


def _sfx(dtype) -> str:
    return "_f32" if np.dtype(dtype) == np.float32 else "_f64"


def _creal(dtype):
    return C.c_float if np.dtype(dtype) == np.float32 else C.c_double


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["F_CONTIGUOUS"] or a.ndim <= 1, "oracle arrays must be column-major"
    return C.c_void_p(a.ctypes.data)


def _parr(arrs: Sequence[Optional[np.ndarray]], n: int):
    out = (C.c_void_p * n)()
    for i in range(n):
        out[i] = arrs[i].ctypes.data if i < len(arrs) and arrs[i] is not None else None
    return out


def zeros(shape, dtype) -> np.ndarray:
    return np.zeros(tuple(int(s) for s in shape), dtype=dtype, order="F")


# --------------------------------------------------------------------------------------
# FD coefficients -- src/utils/fdgen.jl:11-47 (Fornberg 1998, SIAM Rev. 40.3)
# --------------------------------------------------------------------------------------


def fornberg(xx: Sequence[float], m: int) -> np.ndarray:
    """Weights of the m-th derivative at 0 on nodes xx (fdgen.jl:11-42), Float64."""
    x = sorted(float(v) for v in xx)
    z = 0.0
    n = len(x) - 1
    c = np.zeros((len(x), m + 1))
    c1 = 1.0
    c4 = x[0] - z
    c[0, 0] = 1.0
    for i in range(1, n + 1):
        mn = min(i, m)
        c2 = 1.0
        c5 = c4
        c4 = x[i] - z
        for j in range(0, i):
            c3 = x[i] - x[j]
            c2 = c2 * c3
            if j == i - 1:
                for k in range(mn, 0, -1):
                    c[i, k] = c1 * (k * c[i - 1, k - 1] - c5 * c[i - 1, k]) / c2
                c[i, 0] = -c1 * c5 * c[i - 1, 0] / c2
            for k in range(mn, 0, -1):
                c[j, k] = (c4 * c[j, k] - k * c[j, k - 1]) / c3
            c[j, 0] = c4 * c[j, 0] / c3
        c1 = c2
    return c[:, -1].copy()


def fdcoeffs(deriv: int, order: int) -> np.ndarray:
    """fdgen.jl:44-47."""
    nnn = order + deriv - 1
    return fornberg([k - (nnn / 2 + 0.5) for k in range(1, nnn + 1)], deriv)


# --------------------------------------------------------------------------------------
# Small Base-Julia restatements
# --------------------------------------------------------------------------------------


def linrange(start: float, stop: float, length: int) -> np.ndarray:
    """Base.LinRange getindex: lerpi(j, d, a, b) = (1 - j/d)*a + (j/d)*b, d = max(len-1, 1)."""
    d = max(length - 1, 1)
    out = np.empty(length, dtype=np.float64)
    for j in range(length):
        t = j / d
        out[j] = (1 - t) * start + t * stop
    return out


def round_ties_up(x: np.ndarray) -> np.ndarray:
    """round(Int, x, RoundNearestTiesUp) = floor(x + 1/2) done exactly."""
    f = np.floor(x)
    return (f + (x - f >= 0.5)).astype(np.int64)


def rickerstf(t, t0, f0):
    """src/utils/utils.jl:6."""
    return (1 - 2 * (np.pi * f0 * (t - t0)) ** 2) * np.exp(-((np.pi * f0 * (t - t0)) ** 2))


def gaussderivstf(t, t0, f0):
    """src/utils/utils.jl:13."""
    return (t - t0) * np.exp(-((np.pi * f0 * (t - t0)) ** 2))


def gaussstf(t, t0, f0):
    """src/utils/utils.jl:20."""
    return -np.exp(-((np.pi * f0 * (t - t0)) ** 2)) / (2 * (np.pi * f0) ** 2)


def distribsrcs(nsrc: int, nw: int) -> List[range]:
    """src/utils/utils.jl:28-45 -- contiguous groups, first `nsrc mod nw` workers get one extra."""
    if nsrc >= nw:
        dis = nsrc // nw
        sizes = [dis] * nw
        for k in range(nsrc % nw):
            sizes[k] += 1
    else:
        sizes = [1] * nsrc
    out, start = [], 0
    for s in sizes:
        out.append(range(start, start + s))
        start += s
    return out


# --------------------------------------------------------------------------------------
# CPML coefficients -- src/models/cpmlcoeffs.jl:18-100, src/models/acoustic/acou_init_bc.jl:7-40
# --------------------------------------------------------------------------------------


@dataclass
class CPMLAxis:
    a: np.ndarray
    a_h: np.ndarray
    b: np.ndarray
    b_h: np.ndarray


def _calc_Kab_CPML_axis(halo: int, dt, npower, d0, alpha_max, onwhere: str, T):
    """cpmlcoeffs.jl:43-100 with K_max_pml === nothing (K = 1.0)."""
    Kab_size = halo
    if onwhere == "halfgrd":
        shift = 0.5
    elif onwhere == "ongrd":
        shift = 0.0
        Kab_size += 1
    else:
        raise ValueError("Wrong onwhere parameter!")
    dist = linrange(0 + shift, Kab_size + shift - 1, Kab_size) if Kab_size > 0 else np.zeros(0)
    if halo != 0:
        nl = dist[::-1] / halo
        nr = dist / halo
    else:
        nl = dist[::-1].copy()
        nr = dist.copy()
    # scalars dt, npower, d0, alpha_max are of type T; promoted to Float64 against the Float64 vectors
    dt64, d064, al64 = float(dt), float(d0), float(alpha_max)
    assert float(npower) == 2.0

    def side(nd):
        d = d064 * (nd * nd)
        alpha = al64 * (1.0 - nd)
        b = np.exp(-(d / 1.0 + alpha) * dt64)
        with np.errstate(invalid="ignore", divide="ignore"):
            a = d * (b - 1.0) / (1.0 * (d + 1.0 * alpha))
        return a.astype(T), b.astype(T)

    a_l, b_l = side(nl)
    a_r, b_r = side(nr)
    return a_l, a_r, b_l, b_r


def compute_cpml_axis(vel_max, dt, halo: int, rcoef, thickness, f0, T) -> CPMLAxis:
    """compute_CPML_coefficientsAxis! -- cpmlcoeffs.jl:18-41.  All scalar inputs are of type T."""
    T = np.dtype(T).type
    vel_max, dt, rcoef, thickness, f0 = T(vel_max), T(dt), T(rcoef), T(thickness), T(f0)
    alpha_max = T(T(np.pi) * f0)  # convert(T, π * f0): π promotes to T first
    npower = T(2.0)
    if halo == 0:
        d0 = T(0.0)
    else:
        # -(npower + 1) * vel_max * log(rcoef) / (2.0 * thickness): T arithmetic until the Float64 divisor
        num = T(T(-(npower + T(1))) * vel_max) * T(np.log(rcoef))
        d0 = T(np.float64(num) / (2.0 * np.float64(thickness)))
    a_l, a_r, b_l, b_r = _calc_Kab_CPML_axis(halo, dt, npower, d0, alpha_max, "ongrd", T)
    a_hl, a_hr, b_hl, b_hr = _calc_Kab_CPML_axis(halo, dt, npower, d0, alpha_max, "halfgrd", T)
    return CPMLAxis(
        a=np.concatenate([a_l, a_r]),
        a_h=np.concatenate([a_hl, a_hr]),
        b=np.concatenate([b_l, b_r]),
        b_h=np.concatenate([b_hl, b_hr]),
    )


def init_bdc(vel_max, dt, halo: int, rcoef, spacing: Sequence, freetop: bool, domfreq, T) -> List[CPMLAxis]:
    """init_bdc! -- acou_init_bc.jl:7-40 / ela_init_bc.jl:7-40."""
    T = np.dtype(T).type
    axes = []
    for n in range(len(spacing)):
        thickness = T(T(spacing[n]) * T(halo))
        axes.append(compute_cpml_axis(vel_max, dt, halo, rcoef, thickness, domfreq, T))
    if freetop and len(axes) >= 1:
        last = axes[-1]
        last.a[: len(last.a) // 2] = T(0.0)
        last.a_h[: len(last.a_h) // 2] = T(0.0)
        last.b[: len(last.b) // 2] = T(1.0)
        last.b_h[: len(last.b_h) // 2] = T(1.0)
    return axes


# --------------------------------------------------------------------------------------
# Positions -- src/utils/utils.jl:47-58
# --------------------------------------------------------------------------------------


def find_nearest_grid_points(positions: np.ndarray, spacing: Sequence, T) -> np.ndarray:
    """idx = round(pos/spacing + 1, RoundNearestTiesUp), arithmetic in T.  Returns (npos, N) int64, 1-based."""
    T = np.dtype(T).type
    pos = np.asarray(positions, dtype=T)
    out = np.zeros(pos.shape, dtype=np.int64, order="F")
    for d in range(pos.shape[1]):
        tmp = (pos[:, d] / T(spacing[d]) + T(1)).astype(T)
        out[:, d] = round_ties_up(tmp.astype(np.float64))
    return out


# --------------------------------------------------------------------------------------
# Interpolation onto staggered points and its transpose -- src/utils/interpolations.jl:1-47
# --------------------------------------------------------------------------------------


def interp_arith(m: np.ndarray, dims: Sequence[int]) -> np.ndarray:
    """ArithmeticAverageInterpolation: BSpline(Linear()) evaluated at half indices along `dims`
    (0-based here) with Float64 weights 0.5 -> result is Float64 (interpolations.jl:35-38).
    Nested evaluation, first axis innermost."""
    out = np.asarray(m, dtype=np.float64)
    for d in sorted(dims):
        n = out.shape[d]
        lo = np.take(out, range(0, n - 1), axis=d)
        hi = np.take(out, range(1, n), axis=d)
        out = 0.5 * lo + 0.5 * hi
    return np.asfortranarray(out)


def interp_harm(m: np.ndarray, dims: Sequence[int]) -> np.ndarray:
    """HarmonicAverageInterpolation (interpolations.jl:40-43): 1 ./ itp(1 ./ m)."""
    T = m.dtype.type
    return np.asfortranarray(1.0 / interp_arith((T(1) / m).astype(m.dtype), dims))


def interp(method: str, m: np.ndarray, dims: Sequence[int]) -> np.ndarray:
    return interp_arith(m, dims) if method == "arithmetic" else interp_harm(m, dims)


def dfdm(method: str, m: np.ndarray, idxs, dims: Sequence[int]) -> np.ndarray:
    """∂f∂m (interpolations.jl:45-47)."""
    T = m.dtype.type
    shape = tuple(s - (1 if d in dims else 0) for d, s in enumerate(m.shape))
    if method == "arithmetic":
        return (np.ones(shape, dtype=m.dtype) / T(2 ** len(dims))).astype(m.dtype)
    itp = interp_harm(m, dims)  # Float64
    return (itp**2) / (m[idxs].astype(m.dtype) ** 2) / (2 ** len(dims))


def back_interp(method: str, m: np.ndarray, g_interp: np.ndarray, dims: Sequence[int]) -> np.ndarray:
    """back_interp (interpolations.jl:14-28); dims 0-based.  Permutations in binary counting order."""
    T = m.dtype.type
    res = zeros(m.shape, m.dtype)
    nd = len(dims)
    for code in range(2**nd):
        bits = [(code >> (nd - 1 - k)) & 1 for k in range(nd)]  # bitstring order: most significant first
        pp = [0] * m.ndim
        for k, d in enumerate(dims):
            pp[d] = bits[k]
        idxs = tuple(slice(pp[d], m.shape[d] - 1 + pp[d]) if d in dims else slice(0, m.shape[d]) for d in range(m.ndim))
        contrib = g_interp * dfdm(method, m, idxs, dims)
        res[idxs] = (res[idxs] + contrib).astype(m.dtype)
    return res


# --------------------------------------------------------------------------------------
# Gradient muting -- src/utils/mute_grad.jl:3-80
# --------------------------------------------------------------------------------------


def _jl_div(x: float, y: float) -> float:
    """Base.div for floats: round((x - rem(x, y)) / y) with exact rem."""
    return float(np.round((x - math.fmod(x, y)) / y))


def mutearoundpoint(arr: np.ndarray, pt: np.ndarray, spacing: Sequence, radiuspx: int) -> None:
    if radiuspx == 0:
        return
    assert radiuspx > 0
    T = arr.dtype.type
    N = arr.ndim
    extent = [T(T(spacing[d]) * T(arr.shape[d] - 1)) for d in range(N)]
    rmax = T(T(radiuspx) * max(T(s) for s in spacing))
    lo, hi = [], []
    for d in range(N):
        if not (T(0) <= T(pt[d]) <= extent[d]):
            raise ValueError(f"mutearoundpoint!(): The point lies outside the grid on dimension {d + 1} at position {pt[d]}.")
        res = T(_jl_div(float(T(pt[d])), float(T(spacing[d]))))
        ijk = int(math.floor(res)) + 1
        lo.append(ijk - radiuspx)
        hi.append(ijk + radiuspx)
    import itertools

    for idx in itertools.product(*[range(lo[d], hi[d] + 1) for d in reversed(range(N))]):
        idx = idx[::-1]
        if not all(1 <= idx[d] <= arr.shape[d] for d in range(N)):
            continue
        cur = [float(T(T(idx[d] - 1) * T(spacing[d]))) for d in range(N)]  # stored into a Float64 vector
        r = math.sqrt(sum((float(T(pt[d])) - cur[d]) ** 2 for d in range(N)))
        if r <= float(rmax):
            att = r / float(rmax)
            q = tuple(i - 1 for i in idx)
            arr[q] = T(float(arr[q]) * att)


def mutearoundmultiplepoints(arr: np.ndarray, pts: np.ndarray, spacing: Sequence, radiuspx: int) -> None:
    for i in range(pts.shape[0]):
        mutearoundpoint(arr, pts[i, :], spacing, radiuspx)


# --------------------------------------------------------------------------------------
# L2 misfit -- src/inversion/misfits/L2Misfit.jl:24-95
# --------------------------------------------------------------------------------------


@dataclass
class L2Misfit:
    observed: np.ndarray  # (nt, nrec) or (nt, ndim, nrec)
    invcov: Optional[np.ndarray] = None  # (nt, nt) dense, or 1-D diagonal, or None = identity
    windows: List[Tuple[int, int]] = field(default_factory=list)  # 1-based inclusive

    def _residuals(self, seis: np.ndarray) -> np.ndarray:
        res = (seis - self.observed).astype(seis.dtype)
        if len(self.windows) > 0:
            mask = np.zeros(res.shape[0], dtype=seis.dtype)
            for (a, b) in self.windows:
                mask[a - 1 : b] = 1
            res = (mask.reshape((-1,) + (1,) * (res.ndim - 1)) * res).astype(seis.dtype)
        return res

    def _apply_invcov(self, r2d: np.ndarray) -> np.ndarray:
        if self.invcov is None:
            return r2d.copy()
        ic = np.asarray(self.invcov)
        if ic.ndim == 1:
            return (ic.reshape(-1, 1) * r2d).astype(r2d.dtype)
        return (ic @ r2d).astype(r2d.dtype)

    def calcmisfit(self, seis: np.ndarray):
        """dot(res, invcov, res)/2 (L2Misfit.jl:24-60)."""
        res = self._residuals(seis)
        if res.ndim == 2:
            return np.sum(res * self._apply_invcov(res), dtype=np.float64) / 2
        tot = 0.0
        for i in range(res.shape[1]):
            r = np.asfortranarray(res[:, i, :])
            tot += np.sum(r * self._apply_invcov(r), dtype=np.float64)
        return tot / 2

    def dchi_du(self, seis: np.ndarray) -> np.ndarray:
        """∂χ_∂u = invcov * (mask .* (syn - obs)) (L2Misfit.jl:64-95)."""
        res = self._residuals(seis)
        if res.ndim == 2:
            return np.asfortranarray(self._apply_invcov(res))
        out = np.empty_like(res, order="F")
        for d in range(res.shape[1]):
            out[:, d, :] = self._apply_invcov(np.asfortranarray(res[:, d, :]))
        return out


# --------------------------------------------------------------------------------------
# LinearCheckpointer -- src/utils/checkpointers.jl:1-107 (fields are dict name -> list of arrays)
# --------------------------------------------------------------------------------------

Field = List[np.ndarray]  # a ScalarVariableField is a 1-list, a MultiVariableField an n-list


def _fcopy(dst: Field, src: Field) -> None:
    for d, s in zip(dst, src):
        np.copyto(d, s)


def _fzero(f: Field) -> Field:
    return [np.zeros_like(a, order="F") for a in f]


class LinearCheckpointer:
    def __init__(self, nt: int, check_freq: int, checkpointed: Dict[str, Field], buffered: List[str], widths: Dict[str, int]):
        assert check_freq < nt, "Checkpointing frequency must be smaller than the number of timesteps!"
        self.nt, self.check_freq = nt, check_freq
        self.last_checkpoint = (nt // check_freq) * check_freq
        self.curr_checkpoint = self.last_checkpoint
        self.widths = dict(widths)
        self.checkpoints: Dict[int, Dict[str, Field]] = {}
        for it in range(0, nt + 2):
            if it % check_freq == 0:
                for name, fld in checkpointed.items():
                    w = self.widths.get(name, 1)
                    for itw in range(it, it - w, -1):
                        self.checkpoints.setdefault(itw, {})[name] = _fzero(fld)
        self.buffers: Dict[str, List[Field]] = {name: [_fzero(checkpointed[name]) for _ in range(check_freq + 1)] for name in buffered}
        self.n_refwd = 0  # bookkeeping only (cell-update accounting)

    def savecheckpoint(self, name: str, fld: Field, it: int) -> None:
        w = self.widths.get(name, 1)
        for itw in range(it, it + w):
            if itw % self.check_freq == 0:
                _fcopy(self.checkpoints[it][name], fld)
        if name in self.buffers and it >= self.last_checkpoint:
            _fcopy(self.buffers[name][it - self.last_checkpoint], fld)

    def isbuffered(self, name: str, it: int) -> bool:
        return name in self.buffers and (self.curr_checkpoint <= it <= self.curr_checkpoint + self.check_freq)

    def ischeckpointed(self, name: str, it: int) -> bool:
        return it in self.checkpoints and name in self.checkpoints[it]

    def issaved(self, name: str, it: int) -> bool:
        return self.isbuffered(name, it) or self.ischeckpointed(name, it)

    def getsaved(self, name: str, it: int) -> Field:
        if self.ischeckpointed(name, it):
            return self.checkpoints[it][name]
        assert self.isbuffered(name, it)
        return self.buffers[name][it - self.curr_checkpoint]

    def initrecover(self) -> None:
        old = self.curr_checkpoint
        self.curr_checkpoint -= self.check_freq
        for name, buf in self.buffers.items():
            _fcopy(buf[0], self.checkpoints[self.curr_checkpoint][name])
            _fcopy(buf[-1], self.checkpoints[old][name])

    def recover(self, recoverfun) -> None:
        start = self.curr_checkpoint + 1
        end = self.curr_checkpoint + self.check_freq - 1
        for it in range(start, end + 1):
            for name, fld in recoverfun(it):
                _fcopy(self.buffers[name][it - start + 1], fld)
            self.n_refwd += 1

    def reset(self) -> None:
        self.curr_checkpoint = self.last_checkpoint


# --------------------------------------------------------------------------------------
# Problem description shared by the acoustic drivers
# --------------------------------------------------------------------------------------


@dataclass
class ScalarShot:
    src_positions: np.ndarray  # (nsrc, N) in metres
    src_tf: np.ndarray  # (nt, nsrc)
    domfreq: float
    rec_positions: np.ndarray  # (nrec, N)
    seismograms: Optional[np.ndarray] = None  # (nt, nrec), filled by forward


@dataclass
class Params:
    nt: int
    dt: float
    gridsize: Tuple[int, ...]
    spacing: Tuple[float, ...]
    halo: int = 20
    rcoef: float = 0.0001
    freetop: bool = True
    dtype: type = np.float64


class _GeomCD(C.Structure):
    pass


def _make_geom_cd(T):
    R = _creal(T)

    class G(C.Structure):
        _fields_ = [
            ("ndim", C.c_int),
            ("n", C.c_long * 3),
            ("inv_d", R * 3),
            ("halo", C.c_long),
            ("a", C.c_void_p * 3),
            ("b", C.c_void_p * 3),
            ("a_h", C.c_void_p * 3),
            ("b_h", C.c_void_p * 3),
            ("c_d1o2", C.c_double * 2),
            ("c_d2o2", C.c_double * 3),
        ]

    return G


def _make_geom_vd(T):
    R = _creal(T)

    class G(C.Structure):
        _fields_ = [
            ("ndim", C.c_int),
            ("n", C.c_long * 2),
            ("inv_d", R * 2),
            ("halo", C.c_long),
            ("a", C.c_void_p * 2),
            ("b", C.c_void_p * 2),
            ("a_h", C.c_void_p * 2),
            ("b_h", C.c_void_p * 2),
            ("c_d1o4", C.c_double * 4),
        ]

    return G


# --------------------------------------------------------------------------------------
# Acoustic constant density -- acou_models.jl:68-226, acou_forward.jl:6-62, acou_gradient.jl:4-95
# --------------------------------------------------------------------------------------


class AcousticCDSim:
    def __init__(self, params: Params, gradient: bool = False, check_freq: int = 1):
        self.p = params
        T = self.T = np.dtype(params.dtype).type
        N = self.N = len(params.gridsize)
        n = self.n = tuple(int(v) for v in params.gridsize)
        halo = params.halo
        ns_cpml = n[:-1] if params.freetop else n
        assert all(v >= 2 * halo + 3 for v in ns_cpml), "Number grid points in the dimensions with C-PML boundaries must be at least 2*halo+3!"
        self.dt = T(params.dt)
        self.spacing = tuple(T(s) for s in params.spacing)
        self.gradient = gradient
        self.f: Dict[str, Field] = {}
        self.f["fact"] = [zeros(n, T)]
        for nm in ("pold", "pcur", "pnew"):
            self.f[nm] = [zeros(n, T)]
        if gradient:
            for nm in ("grad_vp", "adjold", "adjcur", "adjnew"):
                self.f[nm] = [zeros(n, T)]
        psi_shape = lambda i: tuple(2 * halo if j == i else n[j] for j in range(N))
        xi_shape = lambda i: tuple(2 * (halo + 1) if j == i else n[j] for j in range(N))
        self.f["psi"] = [zeros(psi_shape(i), T) for i in range(N)]
        self.f["xi"] = [zeros(xi_shape(i), T) for i in range(N)]
        if gradient:
            self.f["psi_adj"] = [zeros(psi_shape(i), T) for i in range(N)]
            self.f["xi_adj"] = [zeros(xi_shape(i), T) for i in range(N)]
        self.vp = zeros(n, T)
        self.cpml: List[CPMLAxis] = []
        self.ckpt = None
        if gradient:
            self.ckpt = LinearCheckpointer(params.nt, check_freq, {k: self.f[k] for k in ("pcur", "psi", "xi")}, ["pcur"], {"pcur": 2})
            self.ckpt.savecheckpoint("pcur", self.f["pold"], -1)
            self.ckpt.savecheckpoint("pcur", self.f["pcur"], 0)
            self.ckpt.savecheckpoint("psi", self.f["psi"], 0)
            self.ckpt.savecheckpoint("xi", self.f["xi"], 0)
        self.c1 = fdcoeffs(1, 2)
        self.c2 = fdcoeffs(2, 2)
        self.cell_updates = 0

    # update_matprop! / precompute_fact! -- acou_models.jl:53-60
    def set_matprop(self, vp: np.ndarray) -> None:
        T = self.T
        assert vp.shape == self.n and np.all(vp > 0)
        np.copyto(self.vp, np.asarray(vp, dtype=T))
        dt2 = T(self.dt * self.dt)
        np.copyto(self.f["fact"][0], (dt2 * (self.vp * self.vp).astype(T)).astype(T))

    def init_shot(self, shot: ScalarShot) -> None:
        self.cpml = init_bdc(self.vp.max(), self.dt, self.p.halo, self.p.rcoef, self.spacing, self.p.freetop, shot.domfreq, self.T)

    def _geom(self):
        G = _make_geom_cd(self.T)
        g = G()
        g.ndim = self.N
        for d in range(3):
            g.n[d] = self.n[d] if d < self.N else 1
            g.inv_d[d] = float(self.T(1) / self.spacing[d]) if d < self.N else 0.0
        g.halo = self.p.halo
        for d in range(self.N):
            g.a[d] = self.cpml[d].a.ctypes.data
            g.b[d] = self.cpml[d].b.ctypes.data
            g.a_h[d] = self.cpml[d].a_h.ctypes.data
            g.b_h[d] = self.cpml[d].b_h.ctypes.data
        for k in range(2):
            g.c_d1o2[k] = self.c1[k]
        for k in range(3):
            g.c_d2o2[k] = self.c2[k]
        return g

    def reset(self) -> None:
        for name, fld in self.f.items():
            if name != "fact":
                for a in fld:
                    a[...] = 0
        if self.ckpt is not None:
            self.ckpt.reset()

    # possrcrec_scaletf -- acou_forward.jl:6-20
    def possrcrec_scaletf(self, shot: ScalarShot):
        T = self.T
        possrcs = find_nearest_grid_points(shot.src_positions, self.spacing, T)
        posrecs = find_nearest_grid_points(shot.rec_positions, self.spacing, T)
        prod = T(np.prod(np.array(self.spacing, dtype=T), dtype=T))
        tf = np.asarray(shot.src_tf, dtype=T)
        scal = np.asfortranarray(((tf / prod).astype(T) * T(self.dt * self.dt)).astype(T))
        for s in range(scal.shape[1]):
            v = self.vp[tuple(possrcs[s, :] - 1)]
            scal[:, s] = (scal[:, s] * T(v * v)).astype(T)
        return possrcs, posrecs, scal

    def _step(self, names, psi, xi, possrcs, tf, posrecs, traces, it, g) -> None:
        """forward_onestep_CPML! / adjoint_onestep_CPML! -- acoustic2D_xPU.jl:78-169 (3D: acoustic3D_xPU.jl:96-219)."""
        L, sx, T = _L(), _sfx(self.T), self.T
        o, c, nw = names
        pold, pcur, pnew = self.f[o][0], self.f[c][0], self.f[nw][0]
        getattr(L, "acou_cd_update_psi" + sx)(C.byref(g), _p(pcur), _parr(self.f[psi], 3))
        getattr(L, "acou_cd_update_p" + sx)(C.byref(g), _p(pold), _p(pcur), _p(pnew), _p(self.f["fact"][0]), _parr(self.f[psi], 3), _parr(self.f[xi], 3))
        nn = (C.c_long * 3)(*[self.n[d] if d < self.N else 1 for d in range(3)])
        nt = tf.shape[0]
        getattr(L, "acou_inject" + sx)(_p(pnew), nn, self.N, _p(possrcs), C.c_long(possrcs.shape[0]), _p(tf), C.c_long(nt), C.c_long(it))
        if traces is not None:
            getattr(L, "acou_record" + sx)(_p(pnew), nn, self.N, _p(posrecs), C.c_long(posrecs.shape[0]), _p(traces), C.c_long(traces.shape[0]), C.c_long(it))
        # field handle rotation (acoustic2D_xPU.jl:121-124): pnew ends up aliasing pold
        self.f[o] = self.f[c]
        self.f[c] = self.f[nw]
        self.f[nw] = self.f[o]
        self.cell_updates += int(np.prod(self.n))

    # swforward_1shot! -- acou_forward.jl:22-62
    def forward_1shot(self, shot: ScalarShot, snapevery: Optional[int] = None):
        possrcs, posrecs, tf = self.possrcrec_scaletf(shot)
        traces = zeros((self.p.nt, posrecs.shape[0]), self.T)
        self.reset()
        g = self._geom()
        snaps = {}
        for it in range(1, self.p.nt + 1):
            self._step(("pold", "pcur", "pnew"), "psi", "xi", possrcs, tf, posrecs, traces, it, g)
            if snapevery is not None and it % snapevery == 0:
                snaps[it] = self.f["pcur"][0].copy(order="F")
        shot.seismograms = traces
        return snaps

    # swgradient_1shot! -- acou_gradient.jl:4-95
    def gradient_1shot(self, shot: ScalarShot, misfit, mute_radius_src: int = 0, mute_radius_rec: int = 0) -> Dict[str, np.ndarray]:
        T, L, sx = self.T, _L(), _sfx(self.T)
        ck = self.ckpt
        possrcs, posrecs, tf = self.possrcrec_scaletf(shot)
        nt = self.p.nt
        traces = zeros((nt, posrecs.shape[0]), T)
        self.reset()
        g = self._geom()
        fwd = ("pold", "pcur", "pnew")
        for it in range(1, nt + 1):
            self._step(fwd, "psi", "xi", possrcs, tf, posrecs, traces, it, g)
            ck.savecheckpoint("pcur", self.f["pcur"], it)
            ck.savecheckpoint("psi", self.f["psi"], it)
            ck.savecheckpoint("xi", self.f["xi"], it)
        shot.seismograms = traces
        adjsrc = np.asfortranarray((-misfit.dchi_du(traces)).astype(T))
        nn = (C.c_long * 3)(*[self.n[d] if d < self.N else 1 for d in range(3)])
        getattr(L, "acou_prescale" + sx)(_p(adjsrc), C.c_long(nt), nn, self.N, _p(posrecs), C.c_long(posrecs.shape[0]), _p(self.f["fact"][0]))
        ncell = int(np.prod(self.n))
        corr = getattr(L, "acou_cd_correlate" + sx)
        for it in range(nt, 0, -1):
            self._step(("adjold", "adjcur", "adjnew"), "psi_adj", "xi_adj", posrecs, adjsrc, None, None, it, g)
            if not ck.issaved("pcur", it - 2):
                ck.initrecover()
                _fcopy(self.f["pold"], ck.getsaved("pcur", ck.curr_checkpoint - 1))
                _fcopy(self.f["pcur"], ck.getsaved("pcur", ck.curr_checkpoint))
                _fcopy(self.f["psi"], ck.getsaved("psi", ck.curr_checkpoint))
                _fcopy(self.f["xi"], ck.getsaved("xi", ck.curr_checkpoint))

                def rec(recit):
                    self._step(fwd, "psi", "xi", possrcs, tf, None, None, recit, g)
                    return [("pcur", self.f["pcur"])]

                ck.recover(rec)
            pcur_corr = ck.getsaved("pcur", it - 2)[0]
            pold_corr = ck.getsaved("pcur", it - 1)[0]
            pveryold_corr = ck.getsaved("pcur", it)[0]
            corr(_p(self.f["grad_vp"][0]), _p(self.f["adjcur"][0]), _p(pcur_corr), _p(pold_corr), _p(pveryold_corr), _creal(T)(float(self.dt)), C.c_size_t(ncell))
        gradient = self.f["grad_vp"][0].copy(order="F")
        mutearoundmultiplepoints(gradient, np.asarray(shot.src_positions, dtype=T), self.spacing, mute_radius_src)
        mutearoundmultiplepoints(gradient, np.asarray(shot.rec_positions, dtype=T), self.spacing, mute_radius_rec)
        vp3 = ((self.vp * self.vp).astype(T) * self.vp).astype(T)  # vp .^ 3 = literal_pow -> x*x*x
        gradient = ((T(2.0) / vp3).astype(T) * gradient).astype(T)
        return {"vp": np.asfortranarray(gradient)}


# --------------------------------------------------------------------------------------
# Acoustic variable density -- acou_models.jl:283-452, acou_forward.jl:67-125, acou_gradient.jl:97-203
# --------------------------------------------------------------------------------------


class AcousticVDSim:
    def __init__(self, params: Params, gradient: bool = False, check_freq: int = 1, interp_method: str = "arithmetic"):
        self.p = params
        T = self.T = np.dtype(params.dtype).type
        N = self.N = len(params.gridsize)
        assert N in (1, 2)
        n = self.n = tuple(int(v) for v in params.gridsize)
        halo = params.halo
        ns_cpml = n[:-1] if params.freetop else n
        assert all(v >= 2 * halo + 3 for v in ns_cpml)
        self.dt = T(params.dt)
        self.spacing = tuple(T(s) for s in params.spacing)
        self.interp_method = interp_method
        stag = lambda i: tuple(n[j] - (1 if i == j else 0) for j in range(N))
        self.f: Dict[str, Field] = {}
        self.f["fact_m0"] = [zeros(n, T)]
        self.f["fact_m1_stag"] = [zeros(stag(i), T) for i in range(N)]
        self.f["pcur"] = [zeros(n, T)]
        self.f["vcur"] = [zeros(stag(i), T) for i in range(N)]
        if gradient:
            self.f["grad_m0"] = [zeros(n, T)]
            self.f["grad_m1_stag"] = [zeros(stag(i), T) for i in range(N)]
            self.f["adjpcur"] = [zeros(n, T)]
            self.f["adjvcur"] = [zeros(stag(i), T) for i in range(N)]
        psi_shape = lambda i: tuple(2 * halo if j == i else n[j] for j in range(N))
        xi_shape = lambda i: tuple(2 * (halo + 1) if j == i else n[j] for j in range(N))
        self.f["psi"] = [zeros(psi_shape(i), T) for i in range(N)]
        self.f["xi"] = [zeros(xi_shape(i), T) for i in range(N)]
        if gradient:
            self.f["psi_adj"] = [zeros(psi_shape(i), T) for i in range(N)]
            self.f["xi_adj"] = [zeros(xi_shape(i), T) for i in range(N)]
        self.vp = zeros(n, T)
        self.rho = zeros(n, T)
        self.cpml: List[CPMLAxis] = []
        self.ckpt = None
        if gradient:
            self.ckpt = LinearCheckpointer(params.nt, check_freq, {k: self.f[k] for k in ("pcur", "vcur", "psi", "xi")}, ["pcur"], {"pcur": 1})
            for k in ("pcur", "vcur", "psi", "xi"):
                self.ckpt.savecheckpoint(k, self.f[k], 0)
        self.c4 = fdcoeffs(1, 4)
        self.cell_updates = 0

    # update_matprop! / precompute_fact! -- acou_models.jl:283-300
    def set_matprop(self, vp: np.ndarray, rho: np.ndarray) -> None:
        T = self.T
        assert vp.shape == self.n and rho.shape == self.n and np.all(vp > 0) and np.all(rho > 0)
        np.copyto(self.vp, np.asarray(vp, dtype=T))
        np.copyto(self.rho, np.asarray(rho, dtype=T))
        v2 = (self.vp * self.vp).astype(T)
        np.copyto(self.f["fact_m0"][0], ((v2 * self.rho).astype(T) * self.dt).astype(T))
        inv_rho = (T(1) / self.rho).astype(T)
        for i in range(self.N):
            m1 = interp(self.interp_method, inv_rho, [i])  # Float64
            np.copyto(self.f["fact_m1_stag"][i], (m1 * np.float64(self.dt)).astype(T))

    def init_shot(self, shot: ScalarShot) -> None:
        self.cpml = init_bdc(self.vp.max(), self.dt, self.p.halo, self.p.rcoef, self.spacing, self.p.freetop, shot.domfreq, self.T)

    def _geom(self):
        G = _make_geom_vd(self.T)
        g = G()
        g.ndim = self.N
        for d in range(2):
            g.n[d] = self.n[d] if d < self.N else 1
            g.inv_d[d] = float(self.T(1) / self.spacing[d]) if d < self.N else 0.0
        g.halo = self.p.halo
        for d in range(self.N):
            g.a[d] = self.cpml[d].a.ctypes.data
            g.b[d] = self.cpml[d].b.ctypes.data
            g.a_h[d] = self.cpml[d].a_h.ctypes.data
            g.b_h[d] = self.cpml[d].b_h.ctypes.data
        for k in range(4):
            g.c_d1o4[k] = self.c4[k]
        return g

    def reset(self) -> None:
        for name, fld in self.f.items():
            if name not in ("fact_m0", "fact_m1_stag"):
                for a in fld:
                    a[...] = 0
        if self.ckpt is not None:
            self.ckpt.reset()

    # possrcrec_scaletf -- acou_forward.jl:67-81
    def possrcrec_scaletf(self, shot: ScalarShot):
        T = self.T
        possrcs = find_nearest_grid_points(shot.src_positions, self.spacing, T)
        posrecs = find_nearest_grid_points(shot.rec_positions, self.spacing, T)
        prod = T(np.prod(np.array(self.spacing, dtype=T), dtype=T))
        tf = np.asarray(shot.src_tf, dtype=T)
        scal = np.asfortranarray(((tf / prod).astype(T) * self.dt).astype(T))
        for s in range(scal.shape[1]):
            q = tuple(possrcs[s, :] - 1)
            v = self.vp[q]
            scal[:, s] = (scal[:, s] * T(T(v * v) * self.rho[q])).astype(T)
        return possrcs, posrecs, scal

    def _nn(self):
        return (C.c_long * 3)(*[self.n[d] if d < self.N else 1 for d in range(3)])

    def _fwd_step(self, possrcs, tf, posrecs, traces, it, g) -> None:
        """forward_onestep_CPML! -- acoustic2D_VD_xPU.jl:91-137: p, inject, v, record."""
        L, sx = _L(), _sfx(self.T)
        p, v = self.f["pcur"][0], self.f["vcur"]
        getattr(L, "acou_vd_update_p" + sx)(C.byref(g), _p(p), _p(v[0]), _p(v[1]) if self.N > 1 else None, _p(self.f["fact_m0"][0]), _parr(self.f["xi"], 2))
        getattr(L, "acou_inject" + sx)(_p(p), self._nn(), self.N, _p(possrcs), C.c_long(possrcs.shape[0]), _p(tf), C.c_long(tf.shape[0]), C.c_long(it))
        m1 = self.f["fact_m1_stag"]
        getattr(L, "acou_vd_update_v" + sx)(C.byref(g), _p(p), _p(v[0]), _p(v[1]) if self.N > 1 else None, _p(m1[0]), _p(m1[1]) if self.N > 1 else None, _parr(self.f["psi"], 2))
        if traces is not None:
            getattr(L, "acou_record" + sx)(_p(p), self._nn(), self.N, _p(posrecs), C.c_long(posrecs.shape[0]), _p(traces), C.c_long(traces.shape[0]), C.c_long(it))
        self.cell_updates += int(np.prod(self.n))

    def _adj_step(self, posrecs, adjsrc, it, g) -> None:
        """adjoint_onestep_CPML! -- acoustic2D_VD_xPU.jl:139-178: v, p, inject."""
        L, sx = _L(), _sfx(self.T)
        p, v = self.f["adjpcur"][0], self.f["adjvcur"]
        m1 = self.f["fact_m1_stag"]
        getattr(L, "acou_vd_update_v" + sx)(C.byref(g), _p(p), _p(v[0]), _p(v[1]) if self.N > 1 else None, _p(m1[0]), _p(m1[1]) if self.N > 1 else None, _parr(self.f["psi_adj"], 2))
        getattr(L, "acou_vd_update_p" + sx)(C.byref(g), _p(p), _p(v[0]), _p(v[1]) if self.N > 1 else None, _p(self.f["fact_m0"][0]), _parr(self.f["xi_adj"], 2))
        getattr(L, "acou_inject" + sx)(_p(p), self._nn(), self.N, _p(posrecs), C.c_long(posrecs.shape[0]), _p(adjsrc), C.c_long(adjsrc.shape[0]), C.c_long(it))
        self.cell_updates += int(np.prod(self.n))

    # swforward_1shot! -- acou_forward.jl:83-125
    def forward_1shot(self, shot: ScalarShot, snapevery: Optional[int] = None):
        possrcs, posrecs, tf = self.possrcrec_scaletf(shot)
        traces = zeros((self.p.nt, posrecs.shape[0]), self.T)
        self.reset()
        g = self._geom()
        snaps = {}
        for it in range(1, self.p.nt + 1):
            self._fwd_step(possrcs, tf, posrecs, traces, it, g)
            if snapevery is not None and it % snapevery == 0:
                snaps[it] = self.f["pcur"][0].copy(order="F")
        shot.seismograms = traces
        return snaps

    # swgradient_1shot! -- acou_gradient.jl:97-203
    def gradient_1shot(self, shot: ScalarShot, misfit, mute_radius_src: int = 0, mute_radius_rec: int = 0) -> Dict[str, np.ndarray]:
        T, L, sx = self.T, _L(), _sfx(self.T)
        ck = self.ckpt
        possrcs, posrecs, tf = self.possrcrec_scaletf(shot)
        nt = self.p.nt
        traces = zeros((nt, posrecs.shape[0]), T)
        self.reset()
        g = self._geom()
        for it in range(1, nt + 1):
            self._fwd_step(possrcs, tf, posrecs, traces, it, g)
            for k in ("pcur", "vcur", "psi", "xi"):
                ck.savecheckpoint(k, self.f[k], it)
        shot.seismograms = traces
        adjsrc = np.asfortranarray((-misfit.dchi_du(traces)).astype(T))
        getattr(L, "acou_prescale" + sx)(_p(adjsrc), C.c_long(nt), self._nn(), self.N, _p(posrecs), C.c_long(posrecs.shape[0]), _p(self.f["fact_m0"][0]))
        ncell = int(np.prod(self.n))
        for it in range(nt, 0, -1):
            self._adj_step(posrecs, adjsrc, it, g)
            if not ck.issaved("pcur", it - 1):
                ck.initrecover()
                for k in ("pcur", "vcur", "psi", "xi"):
                    _fcopy(self.f[k], ck.getsaved(k, ck.curr_checkpoint))

                def rec(recit):
                    self._fwd_step(possrcs, tf, None, None, recit, g)
                    return [("pcur", self.f["pcur"])]

                ck.recover(rec)
            p_it = ck.getsaved("pcur", it)[0]
            p_itm1 = ck.getsaved("pcur", it - 1)[0]
            getattr(L, "acou_vd_correlate_m0" + sx)(_p(self.f["grad_m0"][0]), _p(self.f["adjpcur"][0]), _p(p_it), _p(p_itm1), _creal(T)(float(self.dt)), C.c_size_t(ncell))
            gm1, av = self.f["grad_m1_stag"], self.f["adjvcur"]
            getattr(L, "acou_vd_correlate_m1" + sx)(C.byref(g), _p(gm1[0]), _p(gm1[1]) if self.N > 1 else None, _p(av[0]), _p(av[1]) if self.N > 1 else None, _p(p_it))
        gradient_m0 = self.f["grad_m0"][0].copy(order="F")
        gradient_m1 = zeros(self.n, T)
        inv_rho = (T(1) / self.rho).astype(T)
        for i in range(self.N):
            gradient_m1 = (gradient_m1 + back_interp(self.interp_method, inv_rho, self.f["grad_m1_stag"][i], [i])).astype(T)
        srcp = np.asarray(shot.src_positions, dtype=T)
        recp = np.asarray(shot.rec_positions, dtype=T)
        mutearoundmultiplepoints(gradient_m0, srcp, self.spacing, mute_radius_src)
        mutearoundmultiplepoints(gradient_m1, srcp, self.spacing, mute_radius_src)
        mutearoundmultiplepoints(gradient_m0, recp, self.spacing, mute_radius_rec)
        mutearoundmultiplepoints(gradient_m1, recp, self.spacing, mute_radius_rec)
        vp, rho = self.vp, self.rho
        vp2 = (vp * vp).astype(T)
        vp3 = (vp2 * vp).astype(T)
        rho2 = (rho * rho).astype(T)
        g_vp = (((-T(2.0)) * gradient_m0).astype(T) / (vp3 * rho).astype(T)).astype(T)
        g_rho = (((-gradient_m0) / (vp2 * rho2).astype(T)).astype(T) - (gradient_m1 / rho).astype(T)).astype(T)
        return {"vp": np.asfortranarray(g_vp), "rho": np.asfortranarray(g_rho)}


# --------------------------------------------------------------------------------------
# Shot loops -- src/apis/forward.jl:71-115, src/apis/gradient.jl:93-136, src/apis/misfit.jl:83-104
# --------------------------------------------------------------------------------------


def build_wavesim(kind: str, params: Params, gradient: bool = False, check_freq: int = 1, **kw):
    if kind == "acoustic_cd":
        return AcousticCDSim(params, gradient=gradient, check_freq=check_freq)
    if kind == "acoustic_vd":
        return AcousticVDSim(params, gradient=gradient, check_freq=check_freq, **kw)
    if kind == "elastic_iso":
        from . import oracle_elastic  # noqa: WPS433  (kept separate; same "test infrastructure" status)

        return oracle_elastic.ElasticIsoSim(params, gradient=gradient, check_freq=check_freq, **kw)
    raise ValueError(kind)


def swforward(sim, matprop: Sequence[np.ndarray], shots: Sequence, snapevery: Optional[int] = None):
    sim.set_matprop(*matprop)
    snaps = []
    for shot in shots:
        sim.init_shot(shot)
        snaps.append(sim.forward_1shot(shot, snapevery=snapevery))
    return snaps if snapevery is not None else None


def swgradient(sim, matprop: Sequence[np.ndarray], shots: Sequence, misfits: Sequence, mute_radius_src: int = 0, mute_radius_rec: int = 0,
               compute_misfit: bool = False):
    sim.set_matprop(*matprop)
    tot: Dict[str, np.ndarray] = {}
    totmis = 0.0
    for shot, mis in zip(shots, misfits):
        sim.init_shot(shot)
        cur = sim.gradient_1shot(shot, mis, mute_radius_src=mute_radius_src, mute_radius_rec=mute_radius_rec)
        for k, v in cur.items():
            tot[k] = v.copy(order="F") if k not in tot else (tot[k] + v).astype(v.dtype)
        if compute_misfit:
            totmis += mis.calcmisfit(shot.seismograms)
    return (tot, totmis) if compute_misfit else tot


def swmisfit(sim, matprop: Sequence[np.ndarray], shots: Sequence, misfits: Sequence, reference_bug: bool = False):
    """run_swmisfit! (misfit.jl:83-104).  The reference loops `for s in length(shots)` and always uses
    shots[1]/misfit[1]; reference_bug=True reproduces that, False sums all shots."""
    swforward(sim, matprop, shots)
    if reference_bug:
        return misfits[0].calcmisfit(shots[0].seismograms)
    return sum(m.calcmisfit(s.seismograms) for s, m in zip(shots, misfits))
