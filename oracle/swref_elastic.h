/*
 * ORACLE (test infrastructure, NOT product code) -- 2D elastic isotropic P-SV kernels.
 *
 * CPU restatement of /root/reference/src/models/elastic/backends/shared/elastic2D_iso_xPU.jl (kernels
 * update_σxx_σzz!, update_σxz!, update_ux!, update_uz!, inject_*!, record_*!), of the derivative wrappers in
 * /root/reference/src/models/elastic/backends/shared/freesurface_derivatives_4th_mirror.jl:1-234, of the C-PML
 * wrapper ∂̃4th in /root/reference/src/utils/fdgenerated.jl:178-195 and of the correlation kernels in
 * /root/reference/src/models/elastic/backends/shared/correlate_gradient_xPU.jl:1-83.
 * Included twice by swref.c (REAL = float / double).  One function per reference kernel, same sweep structure.
 *
 * Precision rule: 1/24 and 27/24 are Float64 literals (freesurface_derivatives_4th_mirror.jl:2-7), so every
 * derivative is a double even for REAL = float; integer literals (2*μ, 2*ucur, /2) stay in REAL.
 * Indices are 1-based like the reference.  Array extents (column-major):
 *   σxx, σzz, λ, μ : (nx, nz)      ux, ρ_ihalf : (nx-1, nz)      uz, ρ_jhalf : (nx, nz-1)      σxz, μ_ihalf_jhalf : (nx-1, nz-1)
 * Parity status: the reference has no golden vectors and cannot run here; pinned through its own known-answer
 * criteria only (tests/test_oracle_elastic.py).
 */

/* column-major, 1-based */
#define IX2(i, j, n1) ((size_t)((j) - 1) * (size_t)(n1) + (size_t)((i) - 1))

typedef struct {
    long nx, nz, halo;
    int freetop;
    REAL inv_dx, inv_dz, dt;
    const REAL *a_x, *a_xh, *b_x, *b_xh; /* axis 1: a,b length 2(halo+1); a_h,b_h length 2halo */
    const REAL *a_z, *a_zh, *b_z, *b_zh; /* axis 2 */
} FN(ela_geom);

/* ∂x4th_inner / ∂y4th_inner -- freesurface_derivatives_4th_mirror.jl:2-7 */
static inline double FN(e_inner)(REAL f1, REAL f2, REAL f3, REAL f4, REAL inv)
{
    return (1.0 / 24.0 * f1 - 27.0 / 24.0 * f2 + 27.0 / 24.0 * f3 - 1.0 / 24.0 * f4) * inv;
}

#define E_SXX(i, j) sxx[IX2(i, j, nx)]
#define E_SZZ(i, j) szz[IX2(i, j, nx)]
#define E_SXZ(i, j) sxz[IX2(i, j, nx - 1)]
#define E_UX(i, j) ux[IX2(i, j, nx - 1)]
#define E_UZ(i, j) uz[IX2(i, j, nx)]

/* ∂σxx∂x_4th -- :72-83 */
static inline double FN(e_dsxx_dx)(const REAL *sxx, long i, long j, REAL inv, long nx)
{
    if (i == 1)
        return FN(e_inner)(0, E_SXX(i, j), E_SXX(i + 1, j), E_SXX(i + 2, j), inv);
    else if (i == nx - 1)
        return FN(e_inner)(E_SXX(i - 1, j), E_SXX(i, j), E_SXX(i + 1, j), 0, inv);
    return FN(e_inner)(E_SXX(i - 1, j), E_SXX(i, j), E_SXX(i + 1, j), E_SXX(i + 2, j), inv);
}

/* ∂σzz∂z_4th -- :85-102 */
static inline double FN(e_dszz_dz)(const REAL *szz, long i, long j, REAL inv, long nx, long nz, int freetop)
{
    if (j == 1) {
        if (freetop)
            return FN(e_inner)(-E_SZZ(i, j + 1), E_SZZ(i, j), E_SZZ(i, j + 1), E_SZZ(i, j + 2), inv);
        return FN(e_inner)(0, E_SZZ(i, j), E_SZZ(i, j + 1), E_SZZ(i, j + 2), inv);
    } else if (j == nz - 1)
        return FN(e_inner)(E_SZZ(i, j - 1), E_SZZ(i, j), E_SZZ(i, j + 1), 0, inv);
    return FN(e_inner)(E_SZZ(i, j - 1), E_SZZ(i, j), E_SZZ(i, j + 1), E_SZZ(i, j + 2), inv);
}

/* ∂σxz∂x_4th -- :104-123 */
static inline double FN(e_dsxz_dx)(const REAL *sxz, long i, long j, REAL inv, long nx)
{
    if (i == 1)
        return FN(e_inner)(0, 0, E_SXZ(i, j), E_SXZ(i + 1, j), inv);
    else if (i == 2)
        return FN(e_inner)(0, E_SXZ(i - 1, j), E_SXZ(i, j), E_SXZ(i + 1, j), inv);
    else if (i == nx - 1)
        return FN(e_inner)(E_SXZ(i - 2, j), E_SXZ(i - 1, j), E_SXZ(i, j), 0, inv);
    else if (i == nx)
        return FN(e_inner)(E_SXZ(i - 2, j), E_SXZ(i - 1, j), 0, 0, inv);
    return FN(e_inner)(E_SXZ(i - 2, j), E_SXZ(i - 1, j), E_SXZ(i, j), E_SXZ(i + 1, j), inv);
}

/* ∂σxz∂z_4th -- :125-156 */
static inline double FN(e_dsxz_dz)(const REAL *sxz, long i, long j, REAL inv, long nx, long nz, int freetop)
{
    if (j == 1) {
        if (freetop)
            return FN(e_inner)(-E_SXZ(i, j + 1), -E_SXZ(i, j), E_SXZ(i, j), E_SXZ(i, j + 1), inv);
        return FN(e_inner)(0, 0, E_SXZ(i, j), E_SXZ(i, j + 1), inv);
    } else if (j == 2) {
        if (freetop)
            return FN(e_inner)(-E_SXZ(i, j - 1), E_SXZ(i, j - 1), E_SXZ(i, j), E_SXZ(i, j + 1), inv);
        return FN(e_inner)(0, E_SXZ(i, j - 1), E_SXZ(i, j), E_SXZ(i, j + 1), inv);
    } else if (j == nz - 1)
        return FN(e_inner)(E_SXZ(i, j - 2), E_SXZ(i, j - 1), E_SXZ(i, j), 0, inv);
    else if (j == nz)
        return FN(e_inner)(E_SXZ(i, j - 2), E_SXZ(i, j - 1), 0, 0, inv);
    return FN(e_inner)(E_SXZ(i, j - 2), E_SXZ(i, j - 1), E_SXZ(i, j), E_SXZ(i, j + 1), inv);
}

/* ∂ux∂x_4th -- :158-177 */
static inline double FN(e_dux_dx)(const REAL *ux, long i, long j, REAL inv, long nx)
{
    if (i == 1)
        return FN(e_inner)(0, 0, E_UX(i, j), E_UX(i + 1, j), inv);
    else if (i == 2)
        return FN(e_inner)(0, E_UX(i - 1, j), E_UX(i, j), E_UX(i + 1, j), inv);
    else if (i == nx - 1)
        return FN(e_inner)(E_UX(i - 2, j), E_UX(i - 1, j), E_UX(i, j), 0, inv);
    else if (i == nx)
        return FN(e_inner)(E_UX(i - 2, j), E_UX(i - 1, j), 0, 0, inv);
    return FN(e_inner)(E_UX(i - 2, j), E_UX(i - 1, j), E_UX(i, j), E_UX(i + 1, j), inv);
}

/* ∂uz∂z_4th -- :179-212 */
static inline double FN(e_duz_dz)(const REAL *ux, const REAL *uz, const REAL *lam, const REAL *mu, long i, long j, REAL inv_dx, REAL inv_dz, long nx,
                                  long nz, int freetop)
{
    if (j == 1) {
        if (freetop) {
            double dudx = FN(e_dux_dx)(ux, i, j, inv_dx, nx);
            REAL l = lam[IX2(i, j, nx)], m = mu[IX2(i, j, nx)];
            REAL fac = -l / (l + 2 * m);
            return fac * dudx;
        }
        return FN(e_inner)(0, 0, E_UZ(i, j), E_UZ(i, j + 1), inv_dz);
    } else if (j == 2) {
        if (freetop)
            return FN(e_inner)(E_UZ(i, j - 1), E_UZ(i, j - 1), E_UZ(i, j), E_UZ(i, j + 1), inv_dz);
        return FN(e_inner)(0, E_UZ(i, j - 1), E_UZ(i, j), E_UZ(i, j + 1), inv_dz);
    } else if (j == nz - 1)
        return FN(e_inner)(E_UZ(i, j - 2), E_UZ(i, j - 1), E_UZ(i, j), 0, inv_dz);
    else if (j == nz)
        return FN(e_inner)(E_UZ(i, j - 2), E_UZ(i, j - 1), 0, 0, inv_dz);
    return FN(e_inner)(E_UZ(i, j - 2), E_UZ(i, j - 1), E_UZ(i, j), E_UZ(i, j + 1), inv_dz);
}

/* ∂ux∂z_4th -- :214-231 */
static inline double FN(e_dux_dz)(const REAL *ux, long i, long j, REAL inv, long nx, long nz, int freetop)
{
    if (j == 1) {
        if (freetop)
            return FN(e_inner)(E_UX(i, j + 1), E_UX(i, j), E_UX(i, j + 1), E_UX(i, j + 2), inv);
        return FN(e_inner)(0, E_UX(i, j), E_UX(i, j + 1), E_UX(i, j + 2), inv);
    } else if (j == nz - 1)
        return FN(e_inner)(E_UX(i, j - 1), E_UX(i, j), E_UX(i, j + 1), 0, inv);
    return FN(e_inner)(E_UX(i, j - 1), E_UX(i, j), E_UX(i, j + 1), E_UX(i, j + 2), inv);
}

/* ∂uz∂x_4th -- :233-244 */
static inline double FN(e_duz_dx)(const REAL *uz, long i, long j, REAL inv, long nx)
{
    if (i == 1)
        return FN(e_inner)(0, E_UZ(i, j), E_UZ(i + 1, j), E_UZ(i + 2, j), inv);
    else if (i == nx - 1)
        return FN(e_inner)(E_UZ(i - 1, j), E_UZ(i, j), E_UZ(i + 1, j), 0, inv);
    return FN(e_inner)(E_UZ(i - 1, j), E_UZ(i, j), E_UZ(i + 1, j), E_UZ(i + 2, j), inv);
}

/* ∂̃4th -- fdgenerated.jl:178-195.  ndim = size of the differentiated array along dir, I = index passed by the kernel.
 * psi points at compact index 1 along dir with the other index applied; pstride = element stride along dir. */
static inline double FN(e_cpml)(double D, long I, long ndim, long halo, int half, const REAL *a, const REAL *b, REAL *psi, long pstride)
{
    long plusone = half ? 1 : 0;
    long idim = I + plusone;
    long iidim = I - (ndim - halo) + 1 + (halo + plusone);
    if (idim <= halo + plusone) {
        REAL *ps = psi + (size_t)(idim - 1) * (size_t)pstride;
        *ps = (REAL)((double)(REAL)(b[idim - 1] * *ps) + a[idim - 1] * D);
        return D + *ps;
    } else if (idim >= ndim - halo) {
        REAL *ps = psi + (size_t)(iidim - 1) * (size_t)pstride;
        *ps = (REAL)((double)(REAL)(b[iidim - 1] * *ps) + a[iidim - 1] * D);
        return D + *ps;
    }
    return D;
}

/* update_σxx_σzz! -- elastic2D_iso_xPU.jl:39-60, range (2:nx-1, freetop ? 1:nz-1 : 2:nz-1) (:182-183)
 * psi_duxdx (2(halo+1), nz), psi_duzdz (nx, 2(halo+1)) */
void FN(ela_update_sxx_szz)(const FN(ela_geom) * g, REAL *sxx, REAL *szz, const REAL *ux, const REAL *uz, const REAL *lam, const REAL *mu,
                            REAL *psi_duxdx, REAL *psi_duzdz)
{
    const long nx = g->nx, nz = g->nz, h = g->halo;
    const long j0 = g->freetop ? 1 : 2;
#pragma omp parallel for schedule(static)
    for (long j = j0; j <= nz - 1; ++j)
        for (long i = 2; i <= nx - 1; ++i) {
            double dudx = FN(e_dux_dx)(ux, i, j, g->inv_dx, nx);
            double dwdz = FN(e_duz_dz)(ux, uz, lam, mu, i, j, g->inv_dx, g->inv_dz, nx, nz, g->freetop);
            double dudx_c = FN(e_cpml)(dudx, i - 1, nx - 1, h, 1, g->a_x, g->b_x, psi_duxdx + (size_t)(j - 1) * (size_t)(2 * (h + 1)), 1);
            double dwdz_c = FN(e_cpml)(dwdz, j - 1, nz - 1, h, 1, g->a_z, g->b_z, psi_duzdz + (size_t)(i - 1), nx);
            REAL l = lam[IX2(i, j, nx)], m = mu[IX2(i, j, nx)];
            REAL l2m = l + 2 * m;
            sxx[IX2(i, j, nx)] = (REAL)(l2m * dudx_c + l * dwdz_c);
            if (j == 1)
                szz[IX2(i, j, nx)] = 0;
            else
                szz[IX2(i, j, nx)] = (REAL)(l * dudx_c + l2m * dwdz_c);
        }
}

/* update_σxz! -- :62-79, range (1:nx-1, 1:nz-1); psi_duzdx (2halo, nz-1), psi_duxdz (nx-1, 2halo) */
void FN(ela_update_sxz)(const FN(ela_geom) * g, REAL *sxz, const REAL *ux, const REAL *uz, const REAL *mu_hh, REAL *psi_duxdz, REAL *psi_duzdx)
{
    const long nx = g->nx, nz = g->nz, h = g->halo;
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= nz - 1; ++j)
        for (long i = 1; i <= nx - 1; ++i) {
            double dwdx = FN(e_duz_dx)(uz, i, j, g->inv_dx, nx);
            double dudz = FN(e_dux_dz)(ux, i, j, g->inv_dz, nx, nz, g->freetop);
            double dwdx_c = FN(e_cpml)(dwdx, i, nx, h, 0, g->a_xh, g->b_xh, psi_duzdx + (size_t)(j - 1) * (size_t)(2 * h), 1);
            double dudz_c = FN(e_cpml)(dudz, j, nz, h, 0, g->a_zh, g->b_zh, psi_duxdz + (size_t)(i - 1), nx - 1);
            sxz[IX2(i, j, nx - 1)] = (REAL)(mu_hh[IX2(i, j, nx - 1)] * (dwdx_c + dudz_c));
        }
}

/* update_ux! -- :1-18, range (1:nx-1, 1:nz); psi_dsxxdx (2halo, nz), psi_dsxzdz (nx-1, 2(halo+1)) */
void FN(ela_update_ux)(const FN(ela_geom) * g, REAL *uxnew, const REAL *uxcur, const REAL *uxold, const REAL *sxx, const REAL *sxz, const REAL *rho_ih,
                       REAL *psi_dsxxdx, REAL *psi_dsxzdz)
{
    const long nx = g->nx, nz = g->nz, h = g->halo;
    const REAL dt2 = g->dt * g->dt;
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= nz; ++j)
        for (long i = 1; i <= nx - 1; ++i) {
            double d1 = FN(e_dsxx_dx)(sxx, i, j, g->inv_dx, nx);
            double d2 = FN(e_dsxz_dz)(sxz, i, j, g->inv_dz, nx, nz, g->freetop);
            double c1 = FN(e_cpml)(d1, i, nx, h, 0, g->a_xh, g->b_xh, psi_dsxxdx + (size_t)(j - 1) * (size_t)(2 * h), 1);
            double c2 = FN(e_cpml)(d2, j - 1, nz - 1, h, 1, g->a_z, g->b_z, psi_dsxzdz + (size_t)(i - 1), nx - 1);
            const size_t q = IX2(i, j, nx - 1);
            REAL t = 2 * uxcur[q] - uxold[q];
            REAL f = dt2 / rho_ih[q];
            uxnew[q] = (REAL)(t + f * (c1 + c2));
        }
}

/* update_uz! -- :20-37, range (1:nx, 1:nz-1); psi_dsxzdx (2(halo+1), nz-1), psi_dszzdz (nx, 2halo) */
void FN(ela_update_uz)(const FN(ela_geom) * g, REAL *uznew, const REAL *uzcur, const REAL *uzold, const REAL *sxz, const REAL *szz, const REAL *rho_jh,
                       REAL *psi_dsxzdx, REAL *psi_dszzdz)
{
    const long nx = g->nx, nz = g->nz, h = g->halo;
    const REAL dt2 = g->dt * g->dt;
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= nz - 1; ++j)
        for (long i = 1; i <= nx; ++i) {
            double d1 = FN(e_dsxz_dx)(sxz, i, j, g->inv_dx, nx);
            double d2 = FN(e_dszz_dz)(szz, i, j, g->inv_dz, nx, nz, g->freetop);
            double c1 = FN(e_cpml)(d1, i - 1, nx - 1, h, 1, g->a_x, g->b_x, psi_dsxzdx + (size_t)(j - 1) * (size_t)(2 * (h + 1)), 1);
            double c2 = FN(e_cpml)(d2, j, nz, h, 0, g->a_zh, g->b_zh, psi_dszzdz + (size_t)(i - 1), nx);
            const size_t q = IX2(i, j, nx);
            REAL t = 2 * uzcur[q] - uzold[q];
            REAL f = dt2 / rho_jh[q];
            uznew[q] = (REAL)(t + f * (c1 + c2));
        }
}

/* inject_momten_sources2D_σxx_σzz! / σxz! -- :81-94.  Point lists in CSR form: points of source s are [off[s], off[s+1]);
 * ij (npts, 2) column-major 1-based; tf (nt, nsrc). */
void FN(ela_inject_momten)(const FN(ela_geom) * g, REAL *sxx, REAL *szz, REAL *sxz, long nsrc, const long *off_xx, const int *ij_xx, const REAL *coef_xx,
                           const long *off_xz, const int *ij_xz, const REAL *coef_xz, const REAL *Mxx, const REAL *Mzz, const REAL *Mxz, const REAL *tf,
                           long nt, long it)
{
    const long nx = g->nx;
    const long npx = off_xx[nsrc], npz = off_xz[nsrc];
    for (long s = 0; s < nsrc; ++s) {
        const REAL w = tf[(size_t)s * (size_t)nt + (size_t)(it - 1)];
        for (long p = off_xx[s]; p < off_xx[s + 1]; ++p) {
            const long i = ij_xx[p], j = ij_xx[p + npx];
            sxx[IX2(i, j, nx)] += Mxx[s] * coef_xx[p] * w;
            szz[IX2(i, j, nx)] += Mzz[s] * coef_xx[p] * w;
        }
        for (long p = off_xz[s]; p < off_xz[s + 1]; ++p) {
            const long i = ij_xz[p], j = ij_xz[p + npz];
            sxz[IX2(i, j, nx - 1)] += Mxz[s] * coef_xz[p] * w;
        }
    }
}

/* inject_external_sources2D_ux!/uz! -- :96-106; tf (nt, 2, nsrc): u[ij] += coef * tf[it, c, s] / rho[ij] * dt^2 */
void FN(ela_inject_extforce)(const FN(ela_geom) * g, REAL *ux, REAL *uz, const REAL *rho_ih, const REAL *rho_jh, long nsrc, const long *off_ux,
                             const int *ij_ux, const REAL *coef_ux, const long *off_uz, const int *ij_uz, const REAL *coef_uz, const REAL *tf, long nt,
                             long it)
{
    const long nx = g->nx;
    const long npx = off_ux[nsrc], npz = off_uz[nsrc];
    const REAL dt2 = g->dt * g->dt;
    for (long s = 0; s < nsrc; ++s) {
        const REAL wx = tf[((size_t)s * 2 + 0) * (size_t)nt + (size_t)(it - 1)];
        const REAL wz = tf[((size_t)s * 2 + 1) * (size_t)nt + (size_t)(it - 1)];
        for (long p = off_ux[s]; p < off_ux[s + 1]; ++p) {
            const size_t q = IX2(ij_ux[p], ij_ux[p + npx], nx - 1);
            ux[q] += coef_ux[p] * wx / rho_ih[q] * dt2;
        }
        for (long p = off_uz[s]; p < off_uz[s + 1]; ++p) {
            const size_t q = IX2(ij_uz[p], ij_uz[p + npz], nx);
            uz[q] += coef_uz[p] * wz / rho_jh[q] * dt2;
        }
    }
}

/* record_receivers2D_ux!/uz! + the per-receiver sum (elastic2D_iso_xPU.jl:108-118,219-230); traces (nt, 2, nrec).
 * The reference reduces with Base.mapreducedim!(+) whose @simd loop may re-associate; this restatement sums the
 * points of a receiver left to right starting from 0 (rounding-level difference only). */
void FN(ela_record)(const FN(ela_geom) * g, const REAL *ux, const REAL *uz, long nrec, const long *off_ux, const int *ij_ux, const REAL *coef_ux,
                    const long *off_uz, const int *ij_uz, const REAL *coef_uz, REAL *traces, long nt, long it)
{
    const long nx = g->nx;
    const long npx = off_ux[nrec], npz = off_uz[nrec];
    for (long r = 0; r < nrec; ++r) {
        REAL sx = 0, sz = 0;
        for (long p = off_ux[r]; p < off_ux[r + 1]; ++p)
            sx = sx + coef_ux[p] * ux[IX2(ij_ux[p], ij_ux[p + npx], nx - 1)];
        for (long p = off_uz[r]; p < off_uz[r + 1]; ++p)
            sz = sz + coef_uz[p] * uz[IX2(ij_uz[p], ij_uz[p + npz], nx)];
        traces[((size_t)r * 2 + 0) * (size_t)nt + (size_t)(it - 1)] = sx;
        traces[((size_t)r * 2 + 1) * (size_t)nt + (size_t)(it - 1)] = sz;
    }
}

/* correlate_gradients! -- elastic/backends/shared/correlate_gradient_xPU.jl:1-83.
 * uold/ucur/unew = u^{it-2}, u^{it-1}, u^{it}; adj = adjucur. */
void FN(ela_correlate)(const FN(ela_geom) * g, const REAL *adjux, const REAL *adjuz, const REAL *uxo, const REAL *uzo, const REAL *uxc, const REAL *uzc,
                       const REAL *uxn, const REAL *uzn, const REAL *lam, const REAL *mu, REAL *g_rho_ih, REAL *g_rho_jh, REAL *g_lam, REAL *g_mu,
                       REAL *g_mu_hh)
{
    const long nx = g->nx, nz = g->nz;
    const int freetop = g->freetop;
    const REAL _dt2 = 1 / (g->dt * g->dt);
    /* grad_ρ_ihalf, range (1:nx-1, 1:nz), half weight on the free-surface row */
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= nz; ++j)
        for (long i = 1; i <= nx - 1; ++i) {
            const size_t q = IX2(i, j, nx - 1);
            REAL v = adjux[q] * (uxo[q] - 2 * uxc[q] + uxn[q]) * _dt2;
            if (j == 1 && freetop)
                g_rho_ih[q] += v / 2;
            else
                g_rho_ih[q] += v;
        }
    /* grad_ρ_jhalf, range (1:nx, 1:nz-1), freeboundtop = false */
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= nz - 1; ++j)
        for (long i = 1; i <= nx; ++i) {
            const size_t q = IX2(i, j, nx);
            g_rho_jh[q] += adjuz[q] * (uzo[q] - 2 * uzc[q] + uzn[q]) * _dt2;
        }
    /* grad_λ, grad_μ, range (2:nx-1, idxσxx) with the strains of ucur_corr */
    const long j0 = freetop ? 1 : 2;
#pragma omp parallel for schedule(static)
    for (long j = j0; j <= nz - 1; ++j)
        for (long i = 2; i <= nx - 1; ++i) {
            double exx = FN(e_dux_dx)(uxc, i, j, g->inv_dx, nx);
            double exx_a = FN(e_dux_dx)(adjux, i, j, g->inv_dx, nx);
            double ezz = FN(e_duz_dz)(uxc, uzc, lam, mu, i, j, g->inv_dx, g->inv_dz, nx, nz, freetop);
            double ezz_a = FN(e_duz_dz)(adjux, adjuz, lam, mu, i, j, g->inv_dx, g->inv_dz, nx, nz, freetop);
            double div_u = exx + ezz, div_a = exx_a + ezz_a;
            const size_t q = IX2(i, j, nx);
            if (j == 1 && freetop) {
                g_lam[q] = (REAL)(g_lam[q] + div_u * div_a / 2);
                g_mu[q] = (REAL)(g_mu[q] + (exx * exx_a + ezz * ezz_a));
            } else {
                g_lam[q] = (REAL)(g_lam[q] + div_u * div_a);
                g_mu[q] = (REAL)(g_mu[q] + 2 * (exx * exx_a + ezz * ezz_a));
            }
        }
    /* grad_μ_ihalf_jhalf, range (1:nx-1, 1:nz-1) */
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= nz - 1; ++j)
        for (long i = 1; i <= nx - 1; ++i) {
            double exz = (FN(e_duz_dx)(uzc, i, j, g->inv_dx, nx) + FN(e_dux_dz)(uxc, i, j, g->inv_dz, nx, nz, freetop)) / 2;
            double exz_a = (FN(e_duz_dx)(adjuz, i, j, g->inv_dx, nx) + FN(e_dux_dz)(adjux, i, j, g->inv_dz, nx, nz, freetop)) / 2;
            const size_t q = IX2(i, j, nx - 1);
            g_mu_hh[q] = (REAL)(g_mu_hh[q] + 2 * (exz * exz_a + exz * exz_a));
        }
}

#undef IX2
#undef E_SXX
#undef E_SZZ
#undef E_SXZ
#undef E_UX
#undef E_UZ
