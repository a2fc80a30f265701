// ela_iso.cu -- 2D elastic isotropic P-SV displacement-stress update with C-PML and the zero-lag correlations.
//
// Reference semantics: src/models/elastic/backends/shared/elastic2D_iso_xPU.jl:1-448 (update_σxx_σzz!, update_σxz!,
// update_ux!, update_uz!, inject_*!, record_*!, the three orchestrators), the derivative wrappers of
// src/models/elastic/backends/shared/freesurface_derivatives_4th_mirror.jl:1-244 (zero padding at the grid edges, odd /
// even mirroring and the Hooke's-law row at the free surface), the C-PML wrapper ∂̃4th of src/utils/fdgenerated.jl:178-195
// and src/models/elastic/backends/shared/correlate_gradient_xPU.jl:1-83.
//
// Launch structure of one step (the reference issues 4 stencil launches plus 2 tiny launches and a host-side
// mapreduce per source and per receiver):
//   ela_sigma_kernel   : σxx, σzz and σxz from ucur            (one launch, thread per cell)
//   ela_inject_kernel  : moment-tensor injection into σ        (one launch, sources in index order)
//   ela_u_kernel       : uxnew and uznew                       (one launch)
//   ela_inject_kernel  : external-force / adjoint-source injection into unew
//   ela_record_kernel  : Σ_p coef * unew at every receiver     (one launch, thread per receiver and component)
// T = storage type, CT = type of the Float64-literal expressions (1/24, 27/24 are Float64 in the reference).
#include "common.cuh"
#include "kernels.h"

namespace swb {

namespace {

template <class T>
struct ElaParams {
    long long nx, nz;
    int halo, freetop;
    T inv_dx, inv_dz, dt;
    const T *uxo, *uzo, *uxc, *uzc;
    T *uxn, *uzn;
    T *sxx, *szz, *sxz;
    const T *lam, *mu, *rho_ih, *rho_jh, *mu_hh;
    T *psi_dsxxdx, *psi_dsxzdx, *psi_dszzdz, *psi_dsxzdz, *psi_duxdx, *psi_duzdx, *psi_duxdz, *psi_duzdz;
    const T *a_x, *a_xh, *b_x, *b_xh, *a_z, *a_zh, *b_z, *b_zh;
};

// column-major, 1-based
#define EIX(i, j, n1) ((size_t)((j) - 1) * (size_t)(n1) + (size_t)((i) - 1))

// ∂x4th_inner / ∂y4th_inner (freesurface_derivatives_4th_mirror.jl:2-7)
template <class T, class CT>
__device__ __forceinline__ CT e_inner(T f1, T f2, T f3, T f4, T inv)
{
    const CT c1 = (CT)(1.0 / 24.0), c2 = (CT)(27.0 / 24.0);
    return (((c1 * (CT)f1 - c2 * (CT)f2) + c2 * (CT)f3) - c1 * (CT)f4) * (CT)inv;
}

// derivative of an array A(n1 x *) along x at (i, j): value at index i+o is A[i+o, j] or 0 outside 1..n1
template <class T>
__device__ __forceinline__ T at_x(const T *A, long long n1, long long i, long long j)
{
    return (i >= 1 && i <= n1) ? A[EIX(i, j, n1)] : (T)0;
}
template <class T>
__device__ __forceinline__ T at_z(const T *A, long long n1, long long n2, long long i, long long j)
{
    return (j >= 1 && j <= n2) ? A[EIX(i, j, n1)] : (T)0;
}

// ∂σxx∂x_4th (:72-83): points i-1 .. i+2 of σxx (nx columns), missing ones are 0
template <class T, class CT>
__device__ __forceinline__ CT d_sxx_dx(const T *s, long long i, long long j, T inv, long long nx)
{
    return e_inner<T, CT>(at_x(s, nx, i - 1, j), s[EIX(i, j, nx)], s[EIX(i + 1, j, nx)], at_x(s, nx, i + 2, j), inv);
}
// ∂σzz∂z_4th (:85-102)
template <class T, class CT>
__device__ __forceinline__ CT d_szz_dz(const T *s, long long i, long long j, T inv, long long nx, long long nz, int freetop)
{
    if (j == 1) {
        const T f1 = freetop ? -s[EIX(i, j + 1, nx)] : (T)0;
        return e_inner<T, CT>(f1, s[EIX(i, j, nx)], s[EIX(i, j + 1, nx)], s[EIX(i, j + 2, nx)], inv);
    }
    return e_inner<T, CT>(s[EIX(i, j - 1, nx)], s[EIX(i, j, nx)], s[EIX(i, j + 1, nx)], at_z(s, nx, nz, i, j + 2), inv);
}
// ∂σxz∂x_4th / ∂ux∂x_4th (:104-123, :158-177): array with nx-1 columns, points i-2 .. i+1
template <class T, class CT>
__device__ __forceinline__ CT d_stag_dx(const T *s, long long i, long long j, T inv, long long nx)
{
    return e_inner<T, CT>(at_x(s, nx - 1, i - 2, j), at_x(s, nx - 1, i - 1, j), at_x(s, nx - 1, i, j), at_x(s, nx - 1, i + 1, j), inv);
}
// ∂σxz∂z_4th (:125-156): σxz has nz-1 rows, points j-2 .. j+1, odd mirror at the free surface
template <class T, class CT>
__device__ __forceinline__ CT d_sxz_dz(const T *s, long long i, long long j, T inv, long long nx, long long nz, int freetop)
{
    const long long n1 = nx - 1;
    if (j == 1) {
        if (freetop)
            return e_inner<T, CT>(-s[EIX(i, j + 1, n1)], -s[EIX(i, j, n1)], s[EIX(i, j, n1)], s[EIX(i, j + 1, n1)], inv);
        return e_inner<T, CT>((T)0, (T)0, s[EIX(i, j, n1)], s[EIX(i, j + 1, n1)], inv);
    }
    if (j == 2) {
        const T f1 = freetop ? -s[EIX(i, j - 1, n1)] : (T)0;
        return e_inner<T, CT>(f1, s[EIX(i, j - 1, n1)], s[EIX(i, j, n1)], s[EIX(i, j + 1, n1)], inv);
    }
    return e_inner<T, CT>(s[EIX(i, j - 2, n1)], s[EIX(i, j - 1, n1)], at_z(s, n1, nz - 1, i, j), at_z(s, n1, nz - 1, i, j + 1), inv);
}
// ∂uz∂z_4th (:179-212): uz has nz-1 rows; Hooke's-law row at the free surface, even mirror below it
template <class T, class CT>
__device__ __forceinline__ CT d_uz_dz(const T *ux, const T *uz, const T *lam, const T *mu, long long i, long long j, T inv_dx, T inv_dz, long long nx,
                                      long long nz, int freetop)
{
    if (j == 1) {
        if (freetop) {
            const CT dudx = d_stag_dx<T, CT>(ux, i, j, inv_dx, nx);
            const T l = lam[EIX(i, j, nx)], m = mu[EIX(i, j, nx)];
            const T fac = -l / (l + (T)2 * m);
            return (CT)fac * dudx;
        }
        return e_inner<T, CT>((T)0, (T)0, uz[EIX(i, j, nx)], uz[EIX(i, j + 1, nx)], inv_dz);
    }
    if (j == 2) {
        const T f1 = freetop ? uz[EIX(i, j - 1, nx)] : (T)0;
        return e_inner<T, CT>(f1, uz[EIX(i, j - 1, nx)], uz[EIX(i, j, nx)], uz[EIX(i, j + 1, nx)], inv_dz);
    }
    return e_inner<T, CT>(uz[EIX(i, j - 2, nx)], uz[EIX(i, j - 1, nx)], at_z(uz, nx, nz - 1, i, j), at_z(uz, nx, nz - 1, i, j + 1), inv_dz);
}
// ∂ux∂z_4th (:214-231): ux has nz rows, points j-1 .. j+2, even mirror at the free surface
template <class T, class CT>
__device__ __forceinline__ CT d_ux_dz(const T *ux, long long i, long long j, T inv, long long nx, long long nz, int freetop)
{
    const long long n1 = nx - 1;
    if (j == 1) {
        const T f1 = freetop ? ux[EIX(i, j + 1, n1)] : (T)0;
        return e_inner<T, CT>(f1, ux[EIX(i, j, n1)], ux[EIX(i, j + 1, n1)], ux[EIX(i, j + 2, n1)], inv);
    }
    return e_inner<T, CT>(ux[EIX(i, j - 1, n1)], ux[EIX(i, j, n1)], ux[EIX(i, j + 1, n1)], at_z(ux, n1, nz, i, j + 2), inv);
}
// ∂uz∂x_4th (:233-244): uz has nx columns, points i-1 .. i+2
template <class T, class CT>
__device__ __forceinline__ CT d_uz_dx(const T *uz, long long i, long long j, T inv, long long nx)
{
    return e_inner<T, CT>(at_x(uz, nx, i - 1, j), uz[EIX(i, j, nx)], uz[EIX(i + 1, j, nx)], at_x(uz, nx, i + 2, j), inv);
}

// ∂̃4th (fdgenerated.jl:178-195): ndim = extent of the differentiated array along the axis, I = index the kernel passes
template <class T, class CT>
__device__ __forceinline__ CT e_cpml(CT D, long long I, long long ndim, int halo, bool half, const T *a, const T *b, T *psi, long long pstride)
{
    const long long p1 = half ? 1 : 0;
    const long long idim = I + p1;
    long long k;
    if (idim <= halo + p1)
        k = idim;
    else if (idim >= ndim - halo)
        k = I - (ndim - halo) + 1 + (halo + p1);
    else
        return D;
    T *ps = psi + (size_t)(k - 1) * (size_t)pstride;
    T pn;
    const CT r = cpml_apply<T, CT>(D, a[k - 1], b[k - 1], *ps, pn);
    *ps = pn;
    return r;
}

// update_σxx_σzz! (:39-60) on (2:nx-1, j0:nz-1) and update_σxz! (:62-79) on (1:nx-1, 1:nz-1)
template <class T, class CT>
__global__ void __launch_bounds__(256) ela_sigma_kernel(ElaParams<T> P)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
    const long long j = (long long)blockIdx.y * blockDim.y + threadIdx.y + 1;
    const long long nx = P.nx, nz = P.nz;
    const int h = P.halo;
    if (i > nx || j > nz)
        return;
    const long long j0 = P.freetop ? 1 : 2;
    if (i >= 2 && i <= nx - 1 && j >= j0 && j <= nz - 1) {
        const CT dudx = d_stag_dx<T, CT>(P.uxc, i, j, P.inv_dx, nx);
        const CT dwdz = d_uz_dz<T, CT>(P.uxc, P.uzc, P.lam, P.mu, i, j, P.inv_dx, P.inv_dz, nx, nz, P.freetop);
        const CT dudx_c = e_cpml<T, CT>(dudx, i - 1, nx - 1, h, true, P.a_x, P.b_x, P.psi_duxdx + (size_t)(j - 1) * (size_t)(2 * (h + 1)), 1);
        const CT dwdz_c = e_cpml<T, CT>(dwdz, j - 1, nz - 1, h, true, P.a_z, P.b_z, P.psi_duzdz + (size_t)(i - 1), nx);
        const T l = P.lam[EIX(i, j, nx)], m = P.mu[EIX(i, j, nx)];
        const T l2m = l + (T)2 * m;
        P.sxx[EIX(i, j, nx)] = (T)((CT)l2m * dudx_c + (CT)l * dwdz_c);
        P.szz[EIX(i, j, nx)] = j == 1 ? (T)0 : (T)((CT)l * dudx_c + (CT)l2m * dwdz_c);
    }
    if (i <= nx - 1 && j <= nz - 1) {
        const CT dwdx = d_uz_dx<T, CT>(P.uzc, i, j, P.inv_dx, nx);
        const CT dudz = d_ux_dz<T, CT>(P.uxc, i, j, P.inv_dz, nx, nz, P.freetop);
        const CT dwdx_c = e_cpml<T, CT>(dwdx, i, nx, h, false, P.a_xh, P.b_xh, P.psi_duzdx + (size_t)(j - 1) * (size_t)(2 * h), 1);
        const CT dudz_c = e_cpml<T, CT>(dudz, j, nz, h, false, P.a_zh, P.b_zh, P.psi_duxdz + (size_t)(i - 1), nx - 1);
        P.sxz[EIX(i, j, nx - 1)] = (T)((CT)P.mu_hh[EIX(i, j, nx - 1)] * (dwdx_c + dudz_c));
    }
}

// update_ux! (:1-18) on (1:nx-1, 1:nz) and update_uz! (:20-37) on (1:nx, 1:nz-1)
template <class T, class CT>
__global__ void __launch_bounds__(256) ela_u_kernel(ElaParams<T> P)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
    const long long j = (long long)blockIdx.y * blockDim.y + threadIdx.y + 1;
    const long long nx = P.nx, nz = P.nz;
    const int h = P.halo;
    if (i > nx || j > nz)
        return;
    const T dt2 = P.dt * P.dt;
    if (i <= nx - 1) {
        const CT d1 = d_sxx_dx<T, CT>(P.sxx, i, j, P.inv_dx, nx);
        const CT d2 = d_sxz_dz<T, CT>(P.sxz, i, j, P.inv_dz, nx, nz, P.freetop);
        const CT c1 = e_cpml<T, CT>(d1, i, nx, h, false, P.a_xh, P.b_xh, P.psi_dsxxdx + (size_t)(j - 1) * (size_t)(2 * h), 1);
        const CT c2 = e_cpml<T, CT>(d2, j - 1, nz - 1, h, true, P.a_z, P.b_z, P.psi_dsxzdz + (size_t)(i - 1), nx - 1);
        const size_t q = EIX(i, j, nx - 1);
        const T t = (T)2 * P.uxc[q] - P.uxo[q];
        const T f = dt2 / P.rho_ih[q];
        P.uxn[q] = (T)((CT)t + (CT)f * (c1 + c2));
    }
    if (j <= nz - 1) {
        const CT d1 = d_stag_dx<T, CT>(P.sxz, i, j, P.inv_dx, nx);
        const CT d2 = d_szz_dz<T, CT>(P.szz, i, j, P.inv_dz, nx, nz, P.freetop);
        const CT c1 = e_cpml<T, CT>(d1, i - 1, nx - 1, h, true, P.a_x, P.b_x, P.psi_dsxzdx + (size_t)(j - 1) * (size_t)(2 * (h + 1)), 1);
        const CT c2 = e_cpml<T, CT>(d2, j, nz, h, false, P.a_zh, P.b_zh, P.psi_dszzdz + (size_t)(i - 1), nx);
        const size_t q = EIX(i, j, nx);
        const T t = (T)2 * P.uzc[q] - P.uzo[q];
        const T f = dt2 / P.rho_jh[q];
        P.uzn[q] = (T)((CT)t + (CT)f * (c1 + c2));
    }
}

// Source injection, one CTA: sources are processed in index order (they may share cells), the points of one source in
// parallel (distinct cells after spread_positions' merge of equal indices).
//   kind 1: σxx += Mxx[s]*coef*tf[it,s], σzz += Mzz[s]*coef*tf[it,s] (list 0);  σxz += Mxz[s]*coef*tf[it,s] (list 1)   (:81-94)
//   kind 2: ux += coef*tf[it,1,s]/ρ_ihalf*dt^2 (list 0);  uz += coef*tf[it,2,s]/ρ_jhalf*dt^2 (list 1)                  (:96-106)
template <class T>
__global__ void __launch_bounds__(256) ela_inject_kernel(int kind, long long nx, T *f0a, T *f0b, T *f1, const T *rho_ih, const T *rho_jh, swb_sinc_points l0,
                                                         swb_sinc_points l1, const T *tf, long long nt, long long it, const T *Mxx, const T *Mzz, const T *Mxz,
                                                         T dt)
{
    const long long np0 = l0.off[l0.n], np1 = l1.off[l1.n];
    const T *c0 = (const T *)l0.coef, *c1 = (const T *)l1.coef;
    const T dt2 = dt * dt;
    for (long long s = 0; s < l0.n; ++s) {
        if (kind == 1) {
            const T w = tf[(size_t)s * nt + (it - 1)];
            for (long long p = l0.off[s] + threadIdx.x; p < l0.off[s + 1]; p += blockDim.x) {
                const size_t q = EIX(l0.ij[p], l0.ij[p + np0], nx);
                f0a[q] = f0a[q] + (Mxx[s] * c0[p]) * w;
                f0b[q] = f0b[q] + (Mzz[s] * c0[p]) * w;
            }
            for (long long p = l1.off[s] + threadIdx.x; p < l1.off[s + 1]; p += blockDim.x) {
                const size_t q = EIX(l1.ij[p], l1.ij[p + np1], nx - 1);
                f1[q] = f1[q] + (Mxz[s] * c1[p]) * w;
            }
        } else {
            const T wx = tf[((size_t)s * 2 + 0) * nt + (it - 1)], wz = tf[((size_t)s * 2 + 1) * nt + (it - 1)];
            for (long long p = l0.off[s] + threadIdx.x; p < l0.off[s + 1]; p += blockDim.x) {
                const size_t q = EIX(l0.ij[p], l0.ij[p + np0], nx - 1);
                f0a[q] = f0a[q] + ((c0[p] * wx) / rho_ih[q]) * dt2;
            }
            for (long long p = l1.off[s] + threadIdx.x; p < l1.off[s + 1]; p += blockDim.x) {
                const size_t q = EIX(l1.ij[p], l1.ij[p + np1], nx);
                f1[q] = f1[q] + ((c1[p] * wz) / rho_jh[q]) * dt2;
            }
        }
        __syncthreads();
    }
}

// record_receivers2D_ux!/uz! + the per-receiver sum (:108-118, :219-230): traces[it, c, r] = Σ_p coef[p] * u[ij[p]], summed
// left to right (the reference's mapreducedim! is an @simd reduction whose association is unspecified)
template <class T>
__global__ void ela_record_kernel(long long nx, const T *ux, const T *uz, swb_sinc_points lx, swb_sinc_points lz, T *traces, long long nt, long long it)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * lx.n)
        return;
    const long long r = t >> 1;
    const int comp = (int)(t & 1);
    const swb_sinc_points &l = comp == 0 ? lx : lz;
    const long long np = l.off[l.n], n1 = comp == 0 ? nx - 1 : nx;
    const T *u = comp == 0 ? ux : uz, *c = (const T *)l.coef;
    T acc = (T)0;
    for (long long p = l.off[r]; p < l.off[r + 1]; ++p)
        acc = acc + c[p] * u[EIX(l.ij[p], l.ij[p + np], n1)];
    traces[((size_t)r * 2 + comp) * nt + (it - 1)] = acc;
}

// correlate_gradients! (elastic/backends/shared/correlate_gradient_xPU.jl:1-83), all five accumulators in one launch
template <class T, class CT>
__global__ void __launch_bounds__(256)
    ela_correlate_kernel(long long nx, long long nz, int freetop, T inv_dx, T inv_dz, T _dt2, const T *aux, const T *auz, const T *uxo, const T *uzo, const T *uxc,
                         const T *uzc, const T *uxn, const T *uzn, const T *lam, const T *mu, T *g_ri, T *g_rj, T *g_l, T *g_m, T *g_mh)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
    const long long j = (long long)blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > nx || j > nz)
        return;
    if (i <= nx - 1) { // grad_ρ_ihalf on (1:nx-1, 1:nz), all in T
        const size_t q = EIX(i, j, nx - 1);
        const T v = (aux[q] * ((uxo[q] - (T)2 * uxc[q]) + uxn[q])) * _dt2;
        g_ri[q] = g_ri[q] + ((j == 1 && freetop) ? v / (T)2 : v);
    }
    if (j <= nz - 1) { // grad_ρ_jhalf on (1:nx, 1:nz-1)
        const size_t q = EIX(i, j, nx);
        g_rj[q] = g_rj[q] + (auz[q] * ((uzo[q] - (T)2 * uzc[q]) + uzn[q])) * _dt2;
    }
    const long long j0 = freetop ? 1 : 2;
    if (i >= 2 && i <= nx - 1 && j >= j0 && j <= nz - 1) { // grad_λ, grad_μ
        const CT exx = d_stag_dx<T, CT>(uxc, i, j, inv_dx, nx), exx_a = d_stag_dx<T, CT>(aux, i, j, inv_dx, nx);
        const CT ezz = d_uz_dz<T, CT>(uxc, uzc, lam, mu, i, j, inv_dx, inv_dz, nx, nz, freetop);
        const CT ezz_a = d_uz_dz<T, CT>(aux, auz, lam, mu, i, j, inv_dx, inv_dz, nx, nz, freetop);
        const CT div_u = exx + ezz, div_a = exx_a + ezz_a;
        const size_t q = EIX(i, j, nx);
        if (j == 1 && freetop) {
            g_l[q] = (T)((CT)g_l[q] + (div_u * div_a) / (CT)2);
            g_m[q] = (T)((CT)g_m[q] + (exx * exx_a + ezz * ezz_a));
        } else {
            g_l[q] = (T)((CT)g_l[q] + div_u * div_a);
            g_m[q] = (T)((CT)g_m[q] + (CT)2 * (exx * exx_a + ezz * ezz_a));
        }
    }
    if (i <= nx - 1 && j <= nz - 1) { // grad_μ_ihalf_jhalf
        const CT exz = (d_uz_dx<T, CT>(uzc, i, j, inv_dx, nx) + d_ux_dz<T, CT>(uxc, i, j, inv_dz, nx, nz, freetop)) / (CT)2;
        const CT exz_a = (d_uz_dx<T, CT>(auz, i, j, inv_dx, nx) + d_ux_dz<T, CT>(aux, i, j, inv_dz, nx, nz, freetop)) / (CT)2;
        const size_t q = EIX(i, j, nx - 1);
        g_mh[q] = (T)((CT)g_mh[q] + (CT)2 * (exz * exz_a + exz * exz_a));
    }
}

template <class T>
ElaParams<T> make_params(const swb_ela_step_args &a)
{
    ElaParams<T> P{};
    SWB_REQUIRE(a.halo >= 0, "CPML halo size must be non-negative!");
    SWB_REQUIRE(a.n[0] >= 6 && a.n[1] >= 6, "elastic grid must have at least 6 points per axis");
    P.nx = a.n[0];
    P.nz = a.n[1];
    P.halo = a.halo;
    P.freetop = a.freetop;
    P.inv_dx = (T)1 / (T)a.spacing[0];
    P.inv_dz = (T)1 / (T)a.spacing[1];
    P.dt = (T)a.dt;
    P.uxo = (const T *)a.uold[0], P.uzo = (const T *)a.uold[1];
    P.uxc = (const T *)a.ucur[0], P.uzc = (const T *)a.ucur[1];
    P.uxn = (T *)a.unew[0], P.uzn = (T *)a.unew[1];
    P.sxx = (T *)a.sigma[0], P.szz = (T *)a.sigma[1], P.sxz = (T *)a.sigma[2];
    P.lam = (const T *)a.lambda, P.mu = (const T *)a.mu;
    P.rho_ih = (const T *)a.rho_ihalf, P.rho_jh = (const T *)a.rho_jhalf, P.mu_hh = (const T *)a.mu_ihalf_jhalf;
    P.psi_dsxxdx = (T *)a.psi_dsdx[0], P.psi_dsxzdx = (T *)a.psi_dsdx[1];
    P.psi_dszzdz = (T *)a.psi_dsdz[0], P.psi_dsxzdz = (T *)a.psi_dsdz[1];
    P.psi_duxdx = (T *)a.psi_dudx[0], P.psi_duzdx = (T *)a.psi_dudx[1];
    P.psi_duxdz = (T *)a.psi_dudz[0], P.psi_duzdz = (T *)a.psi_dudz[1];
    P.a_x = (const T *)a.cpml[0].a, P.a_xh = (const T *)a.cpml[0].a_h, P.b_x = (const T *)a.cpml[0].b, P.b_xh = (const T *)a.cpml[0].b_h;
    P.a_z = (const T *)a.cpml[1].a, P.a_zh = (const T *)a.cpml[1].a_h, P.b_z = (const T *)a.cpml[1].b, P.b_zh = (const T *)a.cpml[1].b_h;
    return P;
}

template <class T, class CT>
void ela_step_impl(const swb_ela_step_args &a, bool adjoint)
{
    ElaParams<T> P = make_params<T>(a);
    cudaStream_t st = (cudaStream_t)a.stream;
    const dim3 blk(32, 8, 1), grd(cdiv(P.nx, 32), cdiv(P.nz, 8), 1);
    ela_sigma_kernel<T, CT><<<grd, blk, 0, st>>>(P);
    check_launch("ela_sigma");
    count_launch();
    const int kind = adjoint ? 2 : a.src_kind;
    if (kind != 0)
        SWB_REQUIRE(a.it >= 1 && a.it <= a.nt_tf, "time index out of range of the source time function");
    if (kind == 1 && a.src_pts[0].n > 0) {
        ela_inject_kernel<T><<<1, 256, 0, st>>>(1, P.nx, P.sxx, P.szz, P.sxz, P.rho_ih, P.rho_jh, a.src_pts[0], a.src_pts[1], (const T *)a.srctf, a.nt_tf, a.it,
                                                (const T *)a.Mxx, (const T *)a.Mzz, (const T *)a.Mxz, P.dt);
        check_launch("ela_inject_momten");
        count_launch();
    }
    ela_u_kernel<T, CT><<<grd, blk, 0, st>>>(P);
    check_launch("ela_u");
    count_launch();
    if (kind == 2 && a.src_pts[0].n > 0) {
        ela_inject_kernel<T><<<1, 256, 0, st>>>(2, P.nx, P.uxn, nullptr, P.uzn, P.rho_ih, P.rho_jh, a.src_pts[0], a.src_pts[1], (const T *)a.srctf, a.nt_tf, a.it,
                                                nullptr, nullptr, nullptr, P.dt);
        check_launch("ela_inject_extforce");
        count_launch();
    }
    if (!adjoint && a.traces != nullptr && a.rec_pts[0].n > 0) {
        SWB_REQUIRE(a.it >= 1 && a.it <= a.nt_tr, "time index out of range of the trace buffer");
        ela_record_kernel<T><<<cdiv(2 * a.rec_pts[0].n, 128), 128, 0, st>>>(P.nx, P.uxn, P.uzn, a.rec_pts[0], a.rec_pts[1], (T *)a.traces, a.nt_tr, a.it);
        check_launch("ela_record");
        count_launch();
    }
}

template <class T, class CT>
void ela_correlate_impl(const swb_ela_correlate_args &a)
{
    const long long nx = a.n[0], nz = a.n[1];
    const dim3 blk(32, 8, 1), grd(cdiv(nx, 32), cdiv(nz, 8), 1);
    const T dt = (T)a.dt;
    ela_correlate_kernel<T, CT><<<grd, blk, 0, (cudaStream_t)a.stream>>>(
        nx, nz, a.freetop, (T)1 / (T)a.spacing[0], (T)1 / (T)a.spacing[1], (T)1 / (dt * dt), (const T *)a.adjucur[0], (const T *)a.adjucur[1], (const T *)a.u_itm2[0],
        (const T *)a.u_itm2[1], (const T *)a.u_itm1[0], (const T *)a.u_itm1[1], (const T *)a.u_it[0], (const T *)a.u_it[1], (const T *)a.lambda, (const T *)a.mu,
        (T *)a.grad_rho_ihalf, (T *)a.grad_rho_jhalf, (T *)a.grad_lambda, (T *)a.grad_mu, (T *)a.grad_mu_ihalf_jhalf);
    check_launch("ela_correlate");
    count_launch();
}

} // namespace

// adjoint = true: the residuals in a.srctf (nt, 2, nrec) are injected as external forces through a.src_pts (the receivers'
// sinc lists), nothing is recorded (elastic2D_iso_xPU.jl:359-448)
void ela_step(const swb_ela_step_args &a, bool adjoint)
{
    if (a.dtype == SWB_F64)
        ela_step_impl<double, double>(a, adjoint);
    else if (a.dtype == SWB_F32) {
        if (a.flags & SWB_FLAG_FAST_F32)
            ela_step_impl<float, float>(a, adjoint);
        else
            ela_step_impl<float, double>(a, adjoint);
    } else
        throw Error(SWB_ERR_ARG, "dtype must be SWB_F32 or SWB_F64");
}

void ela_correlate(const swb_ela_correlate_args &a)
{
    if (a.dtype == SWB_F64)
        ela_correlate_impl<double, double>(a);
    else if (a.dtype == SWB_F32) {
        if (a.flags & SWB_FLAG_FAST_F32)
            ela_correlate_impl<float, float>(a);
        else
            ela_correlate_impl<float, double>(a);
    } else
        throw Error(SWB_ERR_ARG, "dtype must be SWB_F32 or SWB_F64");
}

} // namespace swb
