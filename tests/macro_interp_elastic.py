"""Independent restatement of the reference's 2D elastic P-SV step (test infrastructure, Float64, tiny grids): plain Python loops that follow
update_σxx_σzz!, update_σxz!, update_ux!, update_uz! and the injection / recording kernels of elastic2D_iso_xPU.jl:1-118,120-242, the derivative
wrappers of freesurface_derivatives_4th_mirror.jl:72-244 (zero-padded 4-point stencils, mirrored stresses / displacements under a free
surface, Hooke's-law ∂uz∂z on the surface row) and ∂̃4th of fdgenerated.jl:178-195 -- written from those files, not from oracle/swref_elastic.h,
so that the two can be compared (tests/test_oracle_hardening.py)."""
from __future__ import annotations

from macro_interp import A1


def inner(f1, f2, f3, f4, inv):
    """∂x4th_inner / ∂y4th_inner"""
    return (1 / 24 * f1 - 27 / 24 * f2 + 27 / 24 * f3 - 1 / 24 * f4) * inv


def d_sxx_dx(s, i, j, inv, nx):
    if i == 1:
        return inner(0, s[(i, j)], s[(i + 1, j)], s[(i + 2, j)], inv)
    if i == nx - 1:
        return inner(s[(i - 1, j)], s[(i, j)], s[(i + 1, j)], 0, inv)
    return inner(s[(i - 1, j)], s[(i, j)], s[(i + 1, j)], s[(i + 2, j)], inv)


def d_szz_dz(s, i, j, inv, nz, free):
    if j == 1:
        if free:
            return inner(-s[(i, j + 1)], s[(i, j)], s[(i, j + 1)], s[(i, j + 2)], inv)
        return inner(0, s[(i, j)], s[(i, j + 1)], s[(i, j + 2)], inv)
    if j == nz - 1:
        return inner(s[(i, j - 1)], s[(i, j)], s[(i, j + 1)], 0, inv)
    return inner(s[(i, j - 1)], s[(i, j)], s[(i, j + 1)], s[(i, j + 2)], inv)


def d_sxz_dx(s, i, j, inv, nx):
    if i == 1:
        return inner(0, 0, s[(i, j)], s[(i + 1, j)], inv)
    if i == 2:
        return inner(0, s[(i - 1, j)], s[(i, j)], s[(i + 1, j)], inv)
    if i == nx - 1:
        return inner(s[(i - 2, j)], s[(i - 1, j)], s[(i, j)], 0, inv)
    if i == nx:
        return inner(s[(i - 2, j)], s[(i - 1, j)], 0, 0, inv)
    return inner(s[(i - 2, j)], s[(i - 1, j)], s[(i, j)], s[(i + 1, j)], inv)


def d_sxz_dz(s, i, j, inv, nz, free):
    if j == 1:
        if free:
            return inner(-s[(i, j + 1)], -s[(i, j)], s[(i, j)], s[(i, j + 1)], inv)
        return inner(0, 0, s[(i, j)], s[(i, j + 1)], inv)
    if j == 2:
        if free:
            return inner(-s[(i, j - 1)], s[(i, j - 1)], s[(i, j)], s[(i, j + 1)], inv)
        return inner(0, s[(i, j - 1)], s[(i, j)], s[(i, j + 1)], inv)
    if j == nz - 1:
        return inner(s[(i, j - 2)], s[(i, j - 1)], s[(i, j)], 0, inv)
    if j == nz:
        return inner(s[(i, j - 2)], s[(i, j - 1)], 0, 0, inv)
    return inner(s[(i, j - 2)], s[(i, j - 1)], s[(i, j)], s[(i, j + 1)], inv)


def d_ux_dx(u, i, j, inv, nx):
    if i == 1:
        return inner(0, 0, u[(i, j)], u[(i + 1, j)], inv)
    if i == 2:
        return inner(0, u[(i - 1, j)], u[(i, j)], u[(i + 1, j)], inv)
    if i == nx - 1:
        return inner(u[(i - 2, j)], u[(i - 1, j)], u[(i, j)], 0, inv)
    if i == nx:
        return inner(u[(i - 2, j)], u[(i - 1, j)], 0, 0, inv)
    return inner(u[(i - 2, j)], u[(i - 1, j)], u[(i, j)], u[(i + 1, j)], inv)


def d_uz_dz(ux, uz, lam, mu, i, j, invx, invz, nx, nz, free):
    if j == 1:
        if free:
            dux = d_ux_dx(ux, i, j, invx, nx)
            return (-lam[(i, j)] / (lam[(i, j)] + 2 * mu[(i, j)])) * dux
        return inner(0, 0, uz[(i, j)], uz[(i, j + 1)], invz)
    if j == 2:
        if free:
            return inner(uz[(i, j - 1)], uz[(i, j - 1)], uz[(i, j)], uz[(i, j + 1)], invz)
        return inner(0, uz[(i, j - 1)], uz[(i, j)], uz[(i, j + 1)], invz)
    if j == nz - 1:
        return inner(uz[(i, j - 2)], uz[(i, j - 1)], uz[(i, j)], 0, invz)
    if j == nz:
        return inner(uz[(i, j - 2)], uz[(i, j - 1)], 0, 0, invz)
    return inner(uz[(i, j - 2)], uz[(i, j - 1)], uz[(i, j)], uz[(i, j + 1)], invz)


def d_ux_dz(u, i, j, inv, nz, free):
    if j == 1:
        if free:
            return inner(u[(i, j + 1)], u[(i, j)], u[(i, j + 1)], u[(i, j + 2)], inv)
        return inner(0, u[(i, j)], u[(i, j + 1)], u[(i, j + 2)], inv)
    if j == nz - 1:
        return inner(u[(i, j - 1)], u[(i, j)], u[(i, j + 1)], 0, inv)
    return inner(u[(i, j - 1)], u[(i, j)], u[(i, j + 1)], u[(i, j + 2)], inv)


def d_uz_dx(u, i, j, inv, nx):
    if i == 1:
        return inner(0, u[(i, j)], u[(i + 1, j)], u[(i + 2, j)], inv)
    if i == nx - 1:
        return inner(u[(i - 1, j)], u[(i, j)], u[(i + 1, j)], 0, inv)
    return inner(u[(i - 1, j)], u[(i, j)], u[(i + 1, j)], u[(i + 2, j)], inv)


def dtilde4(x: A1, dx, a, b, psi: A1, I, direction, halo, half):
    """∂̃4th (fdgenerated.jl:178-195): a, b 0-based numpy vectors, psi / x 1-based views"""
    ndim = x.size(direction)
    plusone = 1 if half else 0
    idim = I[direction - 1] + plusone
    iidim = I[direction - 1] - (ndim - halo) + 1 + (halo + plusone)
    at = lambda k: tuple(k if q == direction - 1 else I[q] for q in range(2))
    if idim <= halo + plusone:
        psi[at(idim)] = b[idim - 1] * psi[at(idim)] + a[idim - 1] * dx
        return dx + psi[at(idim)]
    if idim >= ndim - halo:
        psi[at(iidim)] = b[iidim - 1] * psi[at(iidim)] + a[iidim - 1] * dx
        return dx + psi[at(iidim)]
    return dx


def forward_step(st, cp, lists, tf, momtens, traces, it, kind):
    """forward_onestep_CPML! (elastic2D_iso_xPU.jl:120-357, both source kinds).  st: numpy arrays sxx, szz, sxz, uxo, uzo, uxc, uzc (unew aliases
    uold), lam, mu, mu_hh, rho_ih, rho_jh and the eight psi arrays; cp = [(a, a_h, b, b_h)] for x, z; lists = CSR (off, ij, coef) x 4"""
    nx, nz = st["sxx"].shape
    h, free, invx, invz, dt = st["halo"], st["freetop"], 1.0 / st["dx"], 1.0 / st["dz"], st["dt"]
    (ax, axh, bx, bxh), (az, azh, bz, bzh) = cp
    sxx, szz, sxz = A1(st["sxx"]), A1(st["szz"]), A1(st["sxz"])
    ux, uz = A1(st["uxc"]), A1(st["uzc"])
    lam, mu, muhh = A1(st["lam"]), A1(st["mu"]), A1(st["mu_hh"])
    # update_σxx_σzz! over (2:nx-1, freetop ? 1:nz-1 : 2:nz-1)
    for j in range(1 if free else 2, nz):
        for i in range(2, nx):
            dux = d_ux_dx(ux, i, j, invx, nx)
            duz = d_uz_dz(ux, uz, lam, mu, i, j, invx, invz, nx, nz, free)
            dux_c = dtilde4(ux, dux, ax, bx, A1(st["psi_duxdx"]), (i - 1, j), 1, h, True)
            duz_c = dtilde4(uz, duz, az, bz, A1(st["psi_duzdz"]), (i, j - 1), 2, h, True)
            sxx[(i, j)] = (lam[(i, j)] + 2 * mu[(i, j)]) * dux_c + lam[(i, j)] * duz_c
            szz[(i, j)] = 0.0 if j == 1 else lam[(i, j)] * dux_c + (lam[(i, j)] + 2 * mu[(i, j)]) * duz_c
    # update_σxz! over (1:nx-1, 1:nz-1)
    for j in range(1, nz):
        for i in range(1, nx):
            duzx = d_uz_dx(uz, i, j, invx, nx)
            duxz = d_ux_dz(ux, i, j, invz, nz, free)
            duzx_c = dtilde4(uz, duzx, axh, bxh, A1(st["psi_duzdx"]), (i, j), 1, h, False)
            duxz_c = dtilde4(ux, duxz, azh, bzh, A1(st["psi_duxdz"]), (i, j), 2, h, False)
            sxz[(i, j)] = muhh[(i, j)] * (duzx_c + duxz_c)
    s_a, s_b, r_a, r_b = lists
    if kind == "momten":
        for s in range(len(s_a[0]) - 1):
            for p in range(s_a[0][s], s_a[0][s + 1]):
                q = (int(s_a[1][p, 0]), int(s_a[1][p, 1]))
                sxx[q] = sxx[q] + momtens[0][s] * s_a[2][p] * tf[it - 1, s]
                szz[q] = szz[q] + momtens[1][s] * s_a[2][p] * tf[it - 1, s]
            for p in range(s_b[0][s], s_b[0][s + 1]):
                q = (int(s_b[1][p, 0]), int(s_b[1][p, 1]))
                sxz[q] = sxz[q] + momtens[2][s] * s_b[2][p] * tf[it - 1, s]
    uxo, uzo = A1(st["uxo"]), A1(st["uzo"])
    uxn, uzn = uxo, uzo  # unew aliases uold from the first rotation on; the update is point-wise
    rih, rjh = A1(st["rho_ih"]), A1(st["rho_jh"])
    for j in range(1, nz + 1):  # update_ux! over (1:nx-1, 1:nz)
        for i in range(1, nx):
            d1 = d_sxx_dx(sxx, i, j, invx, nx)
            d2 = d_sxz_dz(sxz, i, j, invz, nz, free)
            d1c = dtilde4(sxx, d1, axh, bxh, A1(st["psi_dsxxdx"]), (i, j), 1, h, False)
            d2c = dtilde4(sxz, d2, az, bz, A1(st["psi_dsxzdz"]), (i, j - 1), 2, h, True)
            uxn[(i, j)] = 2 * ux[(i, j)] - uxo[(i, j)] + dt**2 / rih[(i, j)] * (d1c + d2c)
    for j in range(1, nz):  # update_uz! over (1:nx, 1:nz-1)
        for i in range(1, nx + 1):
            d1 = d_sxz_dx(sxz, i, j, invx, nx)
            d2 = d_szz_dz(szz, i, j, invz, nz, free)
            d1c = dtilde4(sxz, d1, ax, bx, A1(st["psi_dsxzdx"]), (i - 1, j), 1, h, True)
            d2c = dtilde4(szz, d2, azh, bzh, A1(st["psi_dszzdz"]), (i, j), 2, h, False)
            uzn[(i, j)] = 2 * uz[(i, j)] - uzo[(i, j)] + dt**2 / rjh[(i, j)] * (d1c + d2c)
    if kind == "extforce":
        for s in range(len(s_a[0]) - 1):
            for p in range(s_a[0][s], s_a[0][s + 1]):
                q = (int(s_a[1][p, 0]), int(s_a[1][p, 1]))
                uxn[q] = uxn[q] + s_a[2][p] * tf[it - 1, 0, s] / rih[q] * dt**2
            for p in range(s_b[0][s], s_b[0][s + 1]):
                q = (int(s_b[1][p, 0]), int(s_b[1][p, 1]))
                uzn[q] = uzn[q] + s_b[2][p] * tf[it - 1, 1, s] / rjh[q] * dt**2
    for r in range(len(r_a[0]) - 1):
        acc = 0.0
        for p in range(r_a[0][r], r_a[0][r + 1]):
            acc = acc + r_a[2][p] * uxn[(int(r_a[1][p, 0]), int(r_a[1][p, 1]))]
        traces[it - 1, 0, r] = acc
        acc = 0.0
        for p in range(r_b[0][r], r_b[0][r + 1]):
            acc = acc + r_b[2][p] * uzn[(int(r_b[1][p, 0]), int(r_b[1][p, 1]))]
        traces[it - 1, 1, r] = acc
    st["uxo"], st["uxc"] = st["uxc"], st["uxo"]
    st["uzo"], st["uzc"] = st["uzc"], st["uzo"]
