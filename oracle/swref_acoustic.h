/*
 * ORACLE (test infrastructure, NOT product code) -- acoustic kernels.
 *
 * CPU restatement of the reference's CPU backend arithmetic, one function per
 * reference kernel, keeping the reference's sweep structure (separate psi / field /
 * inject / record / correlate passes).  This file is a "template": it is included
 * twice by swref.c with REAL = float and REAL = double and FN(x) = x##_f32 / x##_f64.
 *
 * Precision rule (/root/reference/src/utils/fdgen.jl:61-63,131):
 * stencil weights and the literal 2.0 are Float64, so for REAL=float every expression
 * that contains them evaluates in double and rounds once on the store.  Products of
 * two REAL values (b*psi, b*xi) stay in REAL.  In C this falls out of the usual
 * arithmetic conversions as long as the weights are `double` and nothing is hoisted.
 *
 * All indices below are 1-based like the reference; IX() converts.
 * Parity status: pinned only through the reference's own known-answer tests
 * (analytic Green's functions, checkpoint == no-checkpoint, exact interpolation values);
 * the reference ships no golden vectors and Julia is not available here.
 */

/* column-major, 1-based */
#define IX2(i, j, n1) ((size_t)((j) - 1) * (size_t)(n1) + (size_t)((i) - 1))
#define IX3(i, j, k, n1, n2) (((size_t)((k) - 1) * (size_t)(n2) + (size_t)((j) - 1)) * (size_t)(n1) + (size_t)((i) - 1))

/* ------------------------------------------------------------------------------------
 * Plain finite-difference stencil with the reference's boundary rule.
 * /root/reference/src/utils/fdgen.jl:65-135 (function `∂ⁿ_`, bdcheck=true, no mirror):
 *   - all points inside         -> full left-associated sum, times inv
 *   - points missing on ONE side -> sum of the remaining points (same order), times inv
 *   - points missing on both sides (or nothing left) -> 0.0
 * `A` points at the element with index 1 along the differentiated axis (other indices
 * already applied), `stride` is the element stride of that axis, `n` its extent,
 * `I` the base index, off[] the offsets (fdidxs, fdgen.jl:49-59), c[] the Fornberg
 * weights (fdcoeffs, fdgen.jl:44-47).
 * ---------------------------------------------------------------------------------- */
static inline double FN(fd_bd)(const REAL *A, long stride, long n, long I, const double *c, const int *off, int w, REAL inv)
{
    long lo = I + off[0], hi = I + off[w - 1];
    if (lo < 1 && hi > n)
        return 0.0;
    double acc = 0.0;
    int first = 1;
    for (int k = 0; k < w; ++k) {
        long idx = I + off[k];
        if (idx < 1 || idx > n)
            continue;
        double term = c[k] * A[(size_t)(idx - 1) * (size_t)stride];
        if (first) {
            acc = term;
            first = 0;
        } else
            acc = acc + term;
    }
    if (first)
        return 0.0;
    return acc * inv;
}

/* ------------------------------------------------------------------------------------
 * C-PML first derivative, /root/reference/src/utils/fdgen.jl:137-161 (`∂̃_`).
 *   plusone = halfgrid ? 0 : 1; idim = I + plusone; ndim = n + plusone
 *   left  if idim <= halo + plusone        -> psi index idim
 *   right if idim >= ndim - halo           -> psi index idim - (ndim-halo) + 1 + (halo+plusone)
 * psi points at compact index 1 along the axis (other indices applied), stride pstride.
 * psi is updated in place and re-read after the store (rounding to REAL).
 * ---------------------------------------------------------------------------------- */
static inline double FN(cpml_d1)(double D, long I, long n, long halo, int halfgrid,
                                 const REAL *a, const REAL *b, REAL *psi, long pstride)
{
    long plusone = halfgrid ? 0 : 1;
    long idim = I + plusone;
    long ndim = n + plusone;
    if (idim <= halo + plusone) {
        REAL *ps = psi + (size_t)(idim - 1) * (size_t)pstride;
        *ps = (REAL)((double)(REAL)(b[idim - 1] * *ps) + a[idim - 1] * D);
        return D + *ps;
    } else if (idim >= ndim - halo) {
        long ii = idim - (ndim - halo) + 1 + (halo + plusone);
        REAL *ps = psi + (size_t)(ii - 1) * (size_t)pstride;
        *ps = (REAL)((double)(REAL)(b[ii - 1] * *ps) + a[ii - 1] * D);
        return D + *ps;
    }
    return D;
}

/* ------------------------------------------------------------------------------------
 * C-PML second derivative, /root/reference/src/utils/fdgen.jl:163-193 (`∂̃²_`), order 2.
 *   D2 = (1*A[i-1] + -2*A[i] + 1*A[i+1]) * inv^2      (bdcheck as above)
 *   left  if i <= halo           : dpsi = (-1*psi[i-1] + 1*psi[i]) * inv
 *   right if i >= n - halo + 1   : ii = i-(n-halo)+1+halo ; dpsi from psi[ii-1], psi[ii]
 *   xi = b*xi + a*(D2+dpsi) ; return D2 + dpsi + xi
 * ---------------------------------------------------------------------------------- */
static inline double FN(cpml_d2)(const REAL *A, long stride, long n, long i, REAL inv, long halo,
                                 const REAL *a, const REAL *b, const REAL *psi, long pstride,
                                 REAL *xi, long xstride, const double *c2, const int *off2, const double *c1)
{
    REAL inv2 = inv * inv; /* `_Δ ^ 2` evaluated in REAL (literal_pow -> x*x) */
    double D2 = FN(fd_bd)(A, stride, n, i, c2, off2, 3, inv2);
    long ii;
    if (i <= halo)
        ii = i;
    else if (i >= n - halo + 1)
        ii = i - (n - halo) + 1 + halo;
    else
        return D2;
    /* bdcheck=false stencil on psi, base index ii-1, offsets [0,+1] */
    double dpsi = (c1[0] * psi[(size_t)(ii - 2) * (size_t)pstride] + c1[1] * psi[(size_t)(ii - 1) * (size_t)pstride]) * inv;
    REAL *x = xi + (size_t)(ii - 1) * (size_t)xstride;
    *x = (REAL)((double)(REAL)(b[ii - 1] * *x) + a[ii - 1] * (D2 + dpsi));
    return D2 + dpsi + *x;
}

/* ====================================================================================
 * Acoustic constant density, N = 1, 2, 3 (n[d]=1 and unused for d >= ndim).
 * Reference: src/models/acoustic/backends/shared/acoustic{1D,2D,3D}_xPU.jl and the plain
 * loops of src/models/acoustic/backends/Acoustic{1,2,3}D_CD_CPML_Serial.jl.
 * ==================================================================================== */

typedef struct {
    int ndim;
    long n[3];
    REAL inv_d[3]; /* 1/spacing, computed in REAL (acoustic2D_xPU.jl:99-101) */
    long halo;
    /* per-axis CPML coefficients: a,b length 2(halo+1); a_h,b_h length 2halo */
    const REAL *a[3], *b[3], *a_h[3], *b_h[3];
    /* Fornberg weights computed by the caller (oracle.py: fdcoeffs) */
    double c_d1o2[2]; /* deriv 1 order 2 */
    double c_d2o2[3]; /* deriv 2 order 2 */
} FN(acou_cd_geom);

/* update_ψ_{x,y,z}! -- acoustic2D_xPU.jl:1-25, acoustic3D_xPU.jl:1-38.
 * psi[ax] has extent 2halo along axis ax and the grid extent elsewhere. */
void FN(acou_cd_update_psi)(const FN(acou_cd_geom) * g, const REAL *pcur, REAL *const psi[3])
{
    static const int off1[2] = {0, 1};
    const long nx = g->n[0], ny = g->n[1], nz = g->n[2];
    const long halo = g->halo;
    for (int ax = 0; ax < g->ndim; ++ax) {
        long m[3] = {nx, ny, nz};
        m[ax] = 2 * halo; /* psi extents */
        const long pst[3] = {1, nx, nx * ny};
        const long sst[3] = {1, m[0], m[0] * m[1]};
        REAL *ps = psi[ax];
#pragma omp parallel for collapse(2) schedule(static)
        for (long k = 1; k <= m[2]; ++k)
            for (long j = 1; j <= m[1]; ++j)
                for (long i = 1; i <= m[0]; ++i) {
                    long c[3] = {i, j, k};   /* compact (psi) index */
                    long gi[3] = {i, j, k};  /* grid index */
                    long cc = c[ax];
                    gi[ax] = cc > halo ? g->n[ax] - halo - 1 + (cc - halo) : cc;
                    /* base pointers with the differentiated axis at index 1 */
                    size_t pbase = 0, sbase = 0;
                    for (int d = 0; d < 3; ++d) {
                        if (d == ax)
                            continue;
                        pbase += (size_t)(gi[d] - 1) * (size_t)pst[d];
                        sbase += (size_t)(c[d] - 1) * (size_t)sst[d];
                    }
                    double D = FN(fd_bd)(pcur + pbase, pst[ax], g->n[ax], gi[ax], g->c_d1o2, off1, 2, g->inv_d[ax]);
                    (void)FN(cpml_d1)(D, gi[ax], g->n[ax], halo, 1, g->a_h[ax], g->b_h[ax], ps + sbase, sst[ax]);
                }
    }
}

/* update_p_CPML! -- acoustic2D_xPU.jl:27-46, acoustic3D_xPU.jl:40-61.
 * xi[ax] has extent 2(halo+1) along ax.  pnew may alias pold. */
void FN(acou_cd_update_p)(const FN(acou_cd_geom) * g, const REAL *pold, const REAL *pcur, REAL *pnew, const REAL *fact,
                          REAL *const psi[3], REAL *const xi[3])
{
    static const int off2[3] = {-1, 0, 1};
    const long nx = g->n[0], ny = g->n[1], nz = g->n[2];
    const long halo = g->halo;
    const int nd = g->ndim;
    const long pst[3] = {1, nx, nx * ny};
    const long k0 = nd >= 3 ? 2 : 1, k1 = nd >= 3 ? nz - 1 : 1;
    const long j0 = nd >= 2 ? 2 : 1, j1 = nd >= 2 ? ny - 1 : 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (long k = k0; k <= k1; ++k)
        for (long j = j0; j <= j1; ++j)
            for (long i = 2; i <= nx - 1; ++i) {
                const long gi[3] = {i, j, k};
                double lap = 0.0;
                for (int ax = 0; ax < nd; ++ax) {
                    long mp[3] = {nx, ny, nz}, mx[3] = {nx, ny, nz};
                    mp[ax] = 2 * halo;
                    mx[ax] = 2 * (halo + 1);
                    const long sst[3] = {1, mp[0], mp[0] * mp[1]};
                    const long xst[3] = {1, mx[0], mx[0] * mx[1]};
                    size_t pbase = 0, sbase = 0, xbase = 0;
                    for (int d = 0; d < 3; ++d) {
                        if (d == ax)
                            continue;
                        pbase += (size_t)(gi[d] - 1) * (size_t)pst[d];
                        sbase += (size_t)(gi[d] - 1) * (size_t)sst[d];
                        xbase += (size_t)(gi[d] - 1) * (size_t)xst[d];
                    }
                    double t = FN(cpml_d2)(pcur + pbase, pst[ax], g->n[ax], gi[ax], g->inv_d[ax], halo, g->a[ax], g->b[ax],
                                           psi[ax] + sbase, sst[ax], xi[ax] + xbase, xst[ax], g->c_d2o2, off2, g->c_d1o2);
                    lap = (ax == 0) ? t : lap + t; /* +(x, y, z) is left-associated */
                }
                size_t q = IX3(i, j, k, nx, ny);
                pnew[q] = (REAL)(2.0 * pcur[q] - pold[q] + fact[q] * lap);
            }
}

/* inject_sources! / record_receivers! / prescale_residuals! -- acoustic2D_xPU.jl:48-76.
 * pos: (npos, ndim) column-major int64 1-based grid indices; tf/traces: (nt, npos). */
void FN(acou_inject)(REAL *p, const long *n, int ndim, const long *pos, long npos, const REAL *tf, long nt, long it)
{
    for (long s = 0; s < npos; ++s) {
        long i = pos[s], j = ndim >= 2 ? pos[s + npos] : 1, k = ndim >= 3 ? pos[s + 2 * npos] : 1;
        p[IX3(i, j, k, n[0], n[1])] += tf[(size_t)s * (size_t)nt + (size_t)(it - 1)];
    }
}

void FN(acou_record)(const REAL *p, const long *n, int ndim, const long *pos, long npos, REAL *traces, long nt, long it)
{
    for (long r = 0; r < npos; ++r) {
        long i = pos[r], j = ndim >= 2 ? pos[r + npos] : 1, k = ndim >= 3 ? pos[r + 2 * npos] : 1;
        traces[(size_t)r * (size_t)nt + (size_t)(it - 1)] = p[IX3(i, j, k, n[0], n[1])];
    }
}

void FN(acou_prescale)(REAL *res, long nt, const long *n, int ndim, const long *pos, long npos, const REAL *fact)
{
    for (long r = 0; r < npos; ++r) {
        long i = pos[r], j = ndim >= 2 ? pos[r + npos] : 1, k = ndim >= 3 ? pos[r + 2 * npos] : 1;
        REAL f = fact[IX3(i, j, k, n[0], n[1])];
        for (long it = 0; it < nt; ++it)
            res[(size_t)r * (size_t)nt + (size_t)it] *= f;
    }
}

/* correlate_gradient! -- acoustic/backends/shared/correlate_gradient_xPU.jl:1-10:
 *   curgrad = curgrad + (adjcur * (pcur - 2.0*pold + pveryold) * _dt2),  _dt2 = 1/dt^2 in REAL */
void FN(acou_cd_correlate)(REAL *grad, const REAL *adjcur, const REAL *pcur, const REAL *pold, const REAL *pveryold, REAL dt, size_t ncell)
{
    const REAL _dt2 = (REAL)1 / (REAL)(dt * dt);
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < ncell; ++q)
        grad[q] = (REAL)(grad[q] + (adjcur[q] * (pcur[q] - 2.0 * pold[q] + pveryold[q]) * _dt2));
}

/* ====================================================================================
 * Acoustic variable density, staggered p / v, 4th order, N = 1, 2.
 * Reference: src/models/acoustic/backends/shared/acoustic2D_VD_xPU.jl (and the 1D twin).
 * ==================================================================================== */

typedef struct {
    int ndim;
    long n[2];
    REAL inv_d[2];
    long halo;
    const REAL *a[2], *b[2], *a_h[2], *b_h[2];
    double c_d1o4[4]; /* deriv 1 order 4 */
} FN(acou_vd_geom);

/* update_p_CPML! -- acoustic2D_VD_xPU.jl:17-37.
 *  dvx = ∂̃x(vx; I=(i-1,j), order 4, halfgrid=false, a_x,b_x, ξ_x), dvy likewise;
 *  pcur[i,j] -= fact_m0[i,j]*(dvx+dvy) on 2..nx-1 x 2..ny-1.
 *  vx is (nx-1, ny), vy is (nx, ny-1); xi[ax] has extent 2(halo+1) along ax. */
void FN(acou_vd_update_p)(const FN(acou_vd_geom) * g, REAL *p, const REAL *vx, const REAL *vy, const REAL *fact_m0, REAL *const xi[2])
{
    static const int off4[4] = {-1, 0, 1, 2};
    const long nx = g->n[0], ny = g->n[1], halo = g->halo;
    const int nd = g->ndim;
    const long j0 = nd >= 2 ? 2 : 1, j1 = nd >= 2 ? ny - 1 : 1;
#pragma omp parallel for schedule(static)
    for (long j = j0; j <= j1; ++j)
        for (long i = 2; i <= nx - 1; ++i) {
            /* x: array vx (nx-1 x ny), base index i-1 along axis 1 */
            double Dx = FN(fd_bd)(vx + (size_t)(j - 1) * (size_t)(nx - 1), 1, nx - 1, i - 1, g->c_d1o4, off4, 4, g->inv_d[0]);
            double dv = FN(cpml_d1)(Dx, i - 1, nx - 1, halo, 0, g->a[0], g->b[0], xi[0] + (size_t)(j - 1) * (size_t)(2 * (halo + 1)), 1);
            if (nd >= 2) {
                double Dy = FN(fd_bd)(vy + (size_t)(i - 1), nx, ny - 1, j - 1, g->c_d1o4, off4, 4, g->inv_d[1]);
                double dvy = FN(cpml_d1)(Dy, j - 1, ny - 1, halo, 0, g->a[1], g->b[1], xi[1] + (size_t)(i - 1), nx);
                dv = dv + dvy;
            }
            size_t q = IX2(i, j, nx);
            p[q] = (REAL)(p[q] - fact_m0[q] * dv);
        }
}

/* update_vx_CPML! / update_vy_CPML! -- acoustic2D_VD_xPU.jl:39-75.
 *  dp = ∂̃x(p; I=(i,j), order 4, halfgrid=true, a_xh,b_xh, ψ_x); vx[i,j] -= fact_m1_x[i,j]*dp on 1..nx-1 x 1..ny */
void FN(acou_vd_update_v)(const FN(acou_vd_geom) * g, const REAL *p, REAL *vx, REAL *vy, const REAL *fact_m1_x, const REAL *fact_m1_y, REAL *const psi[2])
{
    static const int off4[4] = {-1, 0, 1, 2};
    const long nx = g->n[0], ny = g->n[1], halo = g->halo;
    const int nd = g->ndim;
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= ny; ++j)
        for (long i = 1; i <= nx - 1; ++i) {
            double D = FN(fd_bd)(p + (size_t)(j - 1) * (size_t)nx, 1, nx, i, g->c_d1o4, off4, 4, g->inv_d[0]);
            double dp = FN(cpml_d1)(D, i, nx, halo, 1, g->a_h[0], g->b_h[0], psi[0] + (size_t)(j - 1) * (size_t)(2 * halo), 1);
            size_t q = (size_t)(j - 1) * (size_t)(nx - 1) + (size_t)(i - 1);
            vx[q] = (REAL)(vx[q] - fact_m1_x[q] * dp);
        }
    if (nd < 2)
        return;
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= ny - 1; ++j)
        for (long i = 1; i <= nx; ++i) {
            double D = FN(fd_bd)(p + (size_t)(i - 1), nx, ny, j, g->c_d1o4, off4, 4, g->inv_d[1]);
            double dp = FN(cpml_d1)(D, j, ny, halo, 1, g->a_h[1], g->b_h[1], psi[1] + (size_t)(i - 1), nx);
            size_t q = IX2(i, j, nx);
            vy[q] = (REAL)(vy[q] - fact_m1_y[q] * dp);
        }
}

/* correlate_gradient_m0! -- acoustic/backends/shared/correlate_gradient_xPU.jl:12-21:
 *   grad_m0 = grad_m0 - (adjp * (p_it - p_itm1) * _dt),  _dt = 1/dt in REAL (all-REAL arithmetic) */
void FN(acou_vd_correlate_m0)(REAL *grad_m0, const REAL *adjp, const REAL *p_it, const REAL *p_itm1, REAL dt, size_t ncell)
{
    const REAL _dt = (REAL)1 / dt;
#pragma omp parallel for schedule(static)
    for (size_t q = 0; q < ncell; ++q)
        grad_m0[q] = grad_m0[q] - (REAL)((REAL)(adjp[q] * (REAL)(p_it[q] - p_itm1[q])) * _dt);
}

/* correlate_gradient_m1! -- acoustic2D_VD_xPU.jl:180-199: plain 4-pt @∂x / @∂y (bdcheck, no CPML):
 *   grad_m1_x[i,j] = grad_m1_x[i,j] + adjvx[i,j] * ∂x p,  (1..nx-1, 1..ny); y likewise */
void FN(acou_vd_correlate_m1)(const FN(acou_vd_geom) * g, REAL *gx, REAL *gy, const REAL *adjvx, const REAL *adjvy, const REAL *p)
{
    static const int off4[4] = {-1, 0, 1, 2};
    const long nx = g->n[0], ny = g->n[1];
    /* the 1D method runs over (2:nx-2) only (acoustic1D_VD_xPU.jl:127-131) */
    const long ilo = g->ndim == 1 ? 2 : 1, ihi = g->ndim == 1 ? nx - 2 : nx - 1;
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= ny; ++j)
        for (long i = ilo; i <= ihi; ++i) {
            double D = FN(fd_bd)(p + (size_t)(j - 1) * (size_t)nx, 1, nx, i, g->c_d1o4, off4, 4, g->inv_d[0]);
            size_t q = (size_t)(j - 1) * (size_t)(nx - 1) + (size_t)(i - 1);
            gx[q] = (REAL)(gx[q] + adjvx[q] * D);
        }
    if (g->ndim < 2)
        return;
#pragma omp parallel for schedule(static)
    for (long j = 1; j <= ny - 1; ++j)
        for (long i = 1; i <= nx; ++i) {
            double D = FN(fd_bd)(p + (size_t)(i - 1), nx, ny, j, g->c_d1o4, off4, 4, g->inv_d[1]);
            size_t q = IX2(i, j, nx);
            gy[q] = (REAL)(gy[q] + adjvy[q] * D);
        }
}

#undef IX2
#undef IX3
