"""A second, independently written restatement of the reference's stencil arithmetic (test infrastructure), used to break common-mode
errors between the C oracle (oracle/swref_*.h: hand-expanded kernels) and the CUDA kernels, which share an author.

Instead of expanding each kernel by hand it INTERPRETS the reference's code generators at run time: `dn` follows ∂ⁿ_ (src/utils/fdgen.jl:65-135:
Fornberg weights, index sets of fdidxs, the left / right boundary-check cascade that drops out-of-range taps, n-ary `+` evaluated left to
right, one multiplication by _Δ^deriv), `dtilde` follows ∂̃_ (fdgen.jl:137-161) and `dtilde2` follows ∂̃²_ (fdgen.jl:163-193); the step
functions then read like the @parallel_indices kernels of acoustic2D_xPU.jl:1-62,78-126 and acoustic2D_VD_xPU.jl:1-137.  Plain Python
loops over 1-based indices, Float64 only, tiny grids only."""
from __future__ import annotations

import numpy as np


def fornberg(x, m):
    """fdgen.jl:11-41 (Fornberg 1998), derivative m at 0 on nodes x"""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    c = np.zeros((n, m + 1))
    c1, c4 = 1.0, x[0]
    c[0, 0] = 1.0
    for i in range(1, n):
        mn = min(i, m)
        c2, c5, c4 = 1.0, c4, x[i]
        for j in range(i):
            c3 = x[i] - x[j]
            c2 *= c3
            if j == i - 1:
                for s in range(mn, 0, -1):
                    c[i, s] = c1 * (s * c[i - 1, s - 1] - c5 * c[i - 1, s]) / c2
                c[i, 0] = -c1 * c5 * c[i - 1, 0] / c2
            for s in range(mn, 0, -1):
                c[j, s] = (c4 * c[j, s] - s * c[j, s - 1]) / c3
            c[j, 0] = c4 * c[j, 0] / c3
        c1 = c2
    return c[:, m]


def fdcoeffs(deriv, order):
    nnn = order + deriv - 1
    return fornberg(np.arange(1, nnn + 1) - (nnn / 2 + 0.5), deriv)


def fdoffsets(deriv, order):
    nnn = order + deriv - 1
    grid = np.arange(1, nnn + 1) - (nnn / 2 + 0.5)
    if (deriv + order) % 2 == 1:
        grid = grid + 0.5
    return [int(g) for g in grid]


class A1:
    """1-based view of a numpy array: A1(a)[i, j] == a[i-1, j-1] (Julia indexing)"""

    def __init__(self, a):
        self.a = a

    def __getitem__(self, I):
        return self.a[tuple(i - 1 for i in I)]

    def __setitem__(self, I, v):
        self.a[tuple(i - 1 for i in I)] = v

    def size(self, dim):
        return self.a.shape[dim - 1]


def _sum_lr(terms):
    s = terms[0]
    for t in terms[1:]:
        s = s + t
    return s


def dn(A: A1, dim, I, inv, deriv=1, order=2, bdcheck=True):
    """∂ⁿ_ without mirroring: the value of the generated expression at index tuple I (1-based)"""
    co, offs = fdcoeffs(deriv, order), fdoffsets(deriv, order)
    Is = [tuple(I[k] + (o if k == dim - 1 else 0) for k in range(len(I))) for o in offs]
    scale = inv if deriv == 1 else (inv * inv if deriv == 2 else inv**deriv)  # Julia lowers x^2 to x*x
    full = lambda sel: _sum_lr([co[i] * A[Is[i]] for i in sel]) * scale
    w, n = len(co), A.size(dim)
    is_ = [t[dim - 1] for t in Is]
    if not bdcheck or (1 <= is_[0] and is_[-1] <= n):
        return full(range(w))
    for state in range(1, w):  # checks_recursive: taps 1..state fall off the left end
        if is_[state - 1] < 1 and 1 <= is_[state] <= is_[-1] <= n:
            return full(range(state, w))
    for state in range(w, 1, -1):  # right_checks_recursive: taps state..w fall off the right end
        if 1 <= is_[0] <= is_[state - 2] <= n and is_[state - 1] > n:
            return full(range(state - 1))
    return 0.0


def dtilde(A: A1, a, b, psi: A1, dim, I, inv, halo, halfgrid, order):
    """∂̃_: first derivative with the C-PML memory variable psi updated in place (fdgen.jl:137-161); a, b 0-based numpy vectors"""
    d = dn(A, dim, I, inv, deriv=1, order=order)
    plusone = 0 if halfgrid else 1
    idim = I[dim - 1] + plusone
    ndim = A.size(dim) + plusone
    at = lambda k: tuple(k if q == dim - 1 else I[q] for q in range(len(I)))
    if idim <= halo + plusone:
        psi[at(idim)] = b[idim - 1] * psi[at(idim)] + a[idim - 1] * d
        return d + psi[at(idim)]
    if idim >= ndim - halo:
        ii = idim - (ndim - halo) + 1 + (halo + plusone)
        psi[at(ii)] = b[ii - 1] * psi[at(ii)] + a[ii - 1] * d
        return d + psi[at(ii)]
    return d


def dtilde2(A: A1, a, b, psi: A1, xi: A1, dim, I, inv, halo, order=2):
    """∂̃²_: second derivative with psi read (no bounds check) and xi updated in place (fdgen.jl:163-193)"""
    d2 = dn(A, dim, I, inv, deriv=2, order=order)
    idim, ndim = I[dim - 1], A.size(dim)
    at = lambda k: tuple(k if q == dim - 1 else I[q] for q in range(len(I)))
    if idim <= halo:
        dpsi = dn(psi, dim, at(idim - 1), inv, deriv=1, order=order, bdcheck=False)
        xi[at(idim)] = b[idim - 1] * xi[at(idim)] + a[idim - 1] * (d2 + dpsi)
        return d2 + dpsi + xi[at(idim)]
    if idim >= ndim - halo + 1:
        ii = idim - (ndim - halo) + 1 + halo
        dpsi = dn(psi, dim, at(ii - 1), inv, deriv=1, order=order, bdcheck=False)
        xi[at(ii)] = b[ii - 1] * xi[at(ii)] + a[ii - 1] * (d2 + dpsi)
        return d2 + dpsi + xi[at(ii)]
    return d2


# ---- acoustic2D_VD_xPU.jl -------------------------------------------------------------------------------------------------
def vd_forward_step(st, cp, possrcs, tf, posrecs, traces, it):
    """forward_onestep_CPML! (acoustic2D_VD_xPU.jl:91-137).  st: dict of numpy arrays pcur, vx, vy, fact_m0, m1x, m1y, psi_x, psi_y, xi_x,
    xi_y; cp: [(a, a_h, b, b_h)] per axis; positions 1-based ints"""
    nx, ny = st["pcur"].shape
    p, vx, vy = A1(st["pcur"]), A1(st["vx"]), A1(st["vy"])
    (ax, axh, bx, bxh), (ay, ayh, by, byh) = cp
    halo, idx, idy = st["halo"], 1.0 / st["dx"], 1.0 / st["dy"]
    for j in range(2, ny):  # update_p_CPML! over (2:nx-1, 2:ny-1)
        for i in range(2, nx):
            dvx = dtilde(vx, ax, bx, A1(st["xi_x"]), 1, (i - 1, j), idx, halo, False, 4)
            dvy = dtilde(vy, ay, by, A1(st["xi_y"]), 2, (i, j - 1), idy, halo, False, 4)
            p[(i, j)] = p[(i, j)] - st["fact_m0"][i - 1, j - 1] * (dvx + dvy)
    for s in range(possrcs.shape[0]):
        q = (int(possrcs[s, 0]), int(possrcs[s, 1]))
        p[q] = p[q] + tf[it - 1, s]
    for j in range(1, ny + 1):  # update_vx_CPML! over (1:nx-1, 1:ny)
        for i in range(1, nx):
            dp = dtilde(p, axh, bxh, A1(st["psi_x"]), 1, (i, j), idx, halo, True, 4)
            vx[(i, j)] = vx[(i, j)] - st["m1x"][i - 1, j - 1] * dp
    for j in range(1, ny):  # update_vy_CPML! over (1:nx, 1:ny-1)
        for i in range(1, nx + 1):
            dp = dtilde(p, ayh, byh, A1(st["psi_y"]), 2, (i, j), idy, halo, True, 4)
            vy[(i, j)] = vy[(i, j)] - st["m1y"][i - 1, j - 1] * dp
    for r in range(posrecs.shape[0]):
        traces[it - 1, r] = p[(int(posrecs[r, 0]), int(posrecs[r, 1]))]


# ---- acoustic2D_xPU.jl ----------------------------------------------------------------------------------------------------
def cd_forward_step(st, cp, possrcs, tf, posrecs, traces, it):
    """forward_onestep_CPML! (acoustic2D_xPU.jl:78-126); st: pold, pcur (pnew aliases pold), fact, psi_x, psi_y, xi_x, xi_y.  Rotates the handles."""
    nx, ny = st["pcur"].shape
    pold, pcur = A1(st["pold"]), A1(st["pcur"])
    (ax, axh, bx, bxh), (ay, ayh, by, byh) = cp
    halo, idx, idy = st["halo"], 1.0 / st["dx"], 1.0 / st["dy"]
    psx, psy = A1(st["psi_x"]), A1(st["psi_y"])
    for j in range(1, ny + 1):  # update_ψ_x! over (1:2halo, 1:ny)
        for i in range(1, 2 * halo + 1):
            ii = nx - halo - 1 + (i - halo) if i > halo else i
            dtilde(pcur, axh, bxh, psx, 1, (ii, j), idx, halo, True, 2)
    for j in range(1, 2 * halo + 1):  # update_ψ_y! over (1:nx, 1:2halo)
        jj = ny - halo - 1 + (j - halo) if j > halo else j
        for i in range(1, nx + 1):
            dtilde(pcur, ayh, byh, psy, 2, (i, jj), idy, halo, True, 2)
    pnew = pold
    for j in range(2, ny):
        for i in range(2, nx):
            lap = dtilde2(pcur, ax, bx, psx, A1(st["xi_x"]), 1, (i, j), idx, halo) + dtilde2(pcur, ay, by, psy, A1(st["xi_y"]), 2, (i, j), idy, halo)
            pnew[(i, j)] = 2.0 * pcur[(i, j)] - pold[(i, j)] + st["fact"][i - 1, j - 1] * lap
    for s in range(possrcs.shape[0]):
        q = (int(possrcs[s, 0]), int(possrcs[s, 1]))
        pnew[q] = pnew[q] + tf[it - 1, s]
    for r in range(posrecs.shape[0]):
        traces[it - 1, r] = pnew[(int(posrecs[r, 0]), int(posrecs[r, 1]))]
    st["pold"], st["pcur"] = st["pcur"], st["pold"]


# ---- acoustic3D_xPU.jl ----------------------------------------------------------------------------------------------------
def cd3_forward_step(st, cp, possrcs, tf, posrecs, traces, it):
    """forward_onestep_CPML! (acoustic3D_xPU.jl:96-160): st: pold, pcur (pnew aliases pold), fact, psi[3], xi[3]; rotates the handles"""
    nx, ny, nz = st["pcur"].shape
    n = (nx, ny, nz)
    pold, pcur = A1(st["pold"]), A1(st["pcur"])
    halo = st["halo"]
    inv = [1.0 / st["d"][k] for k in range(3)]
    psi = [A1(a) for a in st["psi"]]
    xi = [A1(a) for a in st["xi"]]
    for ax in range(3):  # update_ψ_x! / _y! / _z! over the 2 halo compact indices of their axis
        (a, ah, b, bh) = cp[ax]
        rng = [range(1, n[q] + 1) for q in range(3)]
        rng[ax] = range(1, 2 * halo + 1)
        for k in rng[2]:
            for j in rng[1]:
                for i in rng[0]:
                    I = [i, j, k]
                    c = I[ax]
                    I[ax] = n[ax] - halo - 1 + (c - halo) if c > halo else c
                    dtilde(pcur, ah, bh, psi[ax], ax + 1, tuple(I), inv[ax], halo, True, 2)
    pnew = pold
    for k in range(2, nz):
        for j in range(2, ny):
            for i in range(2, nx):
                I = (i, j, k)
                lap = dtilde2(pcur, cp[0][0], cp[0][2], psi[0], xi[0], 1, I, inv[0], halo)
                lap = lap + dtilde2(pcur, cp[1][0], cp[1][2], psi[1], xi[1], 2, I, inv[1], halo)
                lap = lap + dtilde2(pcur, cp[2][0], cp[2][2], psi[2], xi[2], 3, I, inv[2], halo)
                pnew[I] = 2.0 * pcur[I] - pold[I] + st["fact"][i - 1, j - 1, k - 1] * lap
    for s in range(possrcs.shape[0]):
        q = tuple(int(v) for v in possrcs[s, :])
        pnew[q] = pnew[q] + tf[it - 1, s]
    for r in range(posrecs.shape[0]):
        traces[it - 1, r] = pnew[tuple(int(v) for v in posrecs[r, :])]
    st["pold"], st["pcur"] = st["pcur"], st["pold"]
