// acou_vd.cu -- 2D (and 1D) acoustic variable-density staggered pressure/velocity update with C-PML.
//
// A 1D grid (acoustic1D_VD_xPU.jl:1-140) is the ny = 1 case of the same kernels: no y derivative, no vy, the pressure update runs
// over (2:nx-1) of the single row and correlate_gradient_m1! over (2:nx-2) (acoustic1D_VD_xPU.jl:127-131).
//
// Reference semantics: src/models/acoustic/backends/shared/acoustic2D_VD_xPU.jl:1-199,
// src/models/acoustic/backends/shared/correlate_gradient_xPU.jl:12-21, stencils from
// src/utils/fdgen.jl:65-161 (4th-order staggered first derivative, zero padding outside the array).
// This file holds the one-launch-per-reference-kernel path (fine-grained C ABI, parity debugging).
// The fused engine kernels live in acou_vd_fused.cu.
#include "common.cuh"
#include "kernels.h"

namespace swb {

template <class T>
struct VdParams {
    int halo;
    long long nx, ny;
    T inv_dx, inv_dy;
    T *p, *vx, *vy;
    const T *m0, *m1x, *m1y;
    T *psi_x, *psi_y, *xi_x, *xi_y;
    const T *a_x, *b_x, *a_xh, *b_xh, *a_y, *b_y, *a_yh, *b_yh;
    double c4[4];
};

// 4-point staggered first derivative with the reference's boundary rule (fdgen.jl:96-131):
// A points at index 1 along the axis, base index I (1-based), offsets {-1,0,+1,+2}.
template <class T, class CT>
__device__ __forceinline__ CT fd4_bd(const T *A, long long stride, long long n, long long I, const double *c, T inv)
{
    const long long lo = I - 1, hi = I + 2;
    const T *q = A + (I - 1) * stride; // element at index I
    if (lo >= 1 && hi <= n) {
        CT acc = (((CT)c[0] * (CT)q[-stride] + (CT)c[1] * (CT)q[0]) + (CT)c[2] * (CT)q[stride]) + (CT)c[3] * (CT)q[2 * stride];
        return acc * (CT)inv;
    }
    if (lo < 1 && hi > n)
        return (CT)0;
    CT acc = (CT)0;
    bool first = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        long long idx = I - 1 + k;
        if (idx < 1 || idx > n)
            continue;
        CT term = (CT)c[k] * (CT)q[(k - 1) * stride];
        acc = first ? term : acc + term;
        first = false;
    }
    return first ? (CT)0 : acc * (CT)inv;
}

// update_p_CPML! (acoustic2D_VD_xPU.jl:17-37): thread per interior cell
template <class T, class CT>
__global__ void __launch_bounds__(256) vd_update_p_kernel(VdParams<T> P)
{
    const long long nx = P.nx, ny = P.ny;
    const bool one_d = ny == 1;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 2; // 1-based
    const long long j = one_d ? 1 : (long long)blockIdx.y * blockDim.y + threadIdx.y + 2;
    const int h = P.halo;
    if (i > nx - 1 || (one_d ? threadIdx.y != 0 || blockIdx.y != 0 : j > ny - 1))
        return;
    // dvx/dx: array vx (nx-1, ny), I = i-1, halfgrid=false -> idim = i, ndim = nx
    CT Dx = fd4_bd<T, CT>(P.vx + (size_t)(j - 1) * (nx - 1), 1, nx - 1, i - 1, P.c4, P.inv_dx);
    if (i <= h + 1 || i >= nx - h) {
        long long ii = i <= h + 1 ? i : i - nx + 2 * h + 2;
        T *x = P.xi_x + (size_t)(j - 1) * (2 * (h + 1)) + (ii - 1);
        T xn;
        Dx = cpml_apply<T, CT>(Dx, P.a_x[ii - 1], P.b_x[ii - 1], *x, xn);
        *x = xn;
    }
    const size_t q = (size_t)(j - 1) * nx + (i - 1);
    if (one_d) { // acoustic1D_VD_xPU.jl:28: pcur[i] -= fact_m0[i] * ∂vx∂x
        P.p[q] = (T)((CT)P.p[q] - (CT)P.m0[q] * Dx);
        return;
    }
    CT Dy = fd4_bd<T, CT>(P.vy + (i - 1), nx, ny - 1, j - 1, P.c4, P.inv_dy);
    if (j <= h + 1 || j >= ny - h) {
        long long jj = j <= h + 1 ? j : j - ny + 2 * h + 2;
        T *x = P.xi_y + (size_t)(jj - 1) * nx + (i - 1);
        T xn;
        Dy = cpml_apply<T, CT>(Dy, P.a_y[jj - 1], P.b_y[jj - 1], *x, xn);
        *x = xn;
    }
    P.p[q] = (T)((CT)P.p[q] - (CT)P.m0[q] * (Dx + Dy));
}

// update_vx_CPML! + update_vy_CPML! (acoustic2D_VD_xPU.jl:39-75) in one launch
template <class T, class CT>
__global__ void __launch_bounds__(256) vd_update_v_kernel(VdParams<T> P)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1; // 1-based
    const long long j = (long long)blockIdx.y * blockDim.y + threadIdx.y + 1;
    const long long nx = P.nx, ny = P.ny;
    const int h = P.halo;
    if (i > nx || j > ny)
        return;
    if (i <= nx - 1) {
        CT D = fd4_bd<T, CT>(P.p + (size_t)(j - 1) * nx, 1, nx, i, P.c4, P.inv_dx);
        if (i <= h || i >= nx - h) {
            long long ii = i <= h ? i : i - nx + 2 * h + 1;
            T *s = P.psi_x + (size_t)(j - 1) * (2 * h) + (ii - 1);
            T sn;
            D = cpml_apply<T, CT>(D, P.a_xh[ii - 1], P.b_xh[ii - 1], *s, sn);
            *s = sn;
        }
        const size_t q = (size_t)(j - 1) * (nx - 1) + (i - 1);
        P.vx[q] = (T)((CT)P.vx[q] - (CT)P.m1x[q] * D);
    }
    if (j <= ny - 1) {
        CT D = fd4_bd<T, CT>(P.p + (i - 1), nx, ny, j, P.c4, P.inv_dy);
        if (j <= h || j >= ny - h) {
            long long jj = j <= h ? j : j - ny + 2 * h + 1;
            T *s = P.psi_y + (size_t)(jj - 1) * nx + (i - 1);
            T sn;
            D = cpml_apply<T, CT>(D, P.a_yh[jj - 1], P.b_yh[jj - 1], *s, sn);
            *s = sn;
        }
        const size_t q = (size_t)(j - 1) * nx + (i - 1);
        P.vy[q] = (T)((CT)P.vy[q] - (CT)P.m1y[q] * D);
    }
}

// correlate_gradient_m0!: all-T arithmetic (no Float64 literal in the reference expression)
template <class T>
__global__ void __launch_bounds__(256) vd_correlate_m0_kernel(T *g, const T *adjp, const T *p_it, const T *p_itm1, T _dt, size_t n)
{
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; q < n; q += stride) {
        T d = p_it[q] - p_itm1[q];
        T t = adjp[q] * d;
        g[q] = g[q] - t * _dt;
    }
}

// correlate_gradient_m1_kernel_x!/y! (acoustic2D_VD_xPU.jl:187-199): plain 4-point @∂x/@∂y, no CPML
template <class T, class CT>
__global__ void __launch_bounds__(256) vd_correlate_m1_kernel(T *gx, T *gy, const T *avx, const T *avy, const T *p, long long nx, long long ny,
                                                              T inv_dx, T inv_dy, double c0, double c1, double c2, double c3)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1;
    const long long j = (long long)blockIdx.y * blockDim.y + threadIdx.y + 1;
    if (i > nx || j > ny)
        return;
    const double c[4] = {c0, c1, c2, c3};
    if (ny == 1 ? (i >= 2 && i <= nx - 2) : i <= nx - 1) { // 1D: (2:nx-2), acoustic1D_VD_xPU.jl:130
        CT D = fd4_bd<T, CT>(p + (size_t)(j - 1) * nx, 1, nx, i, c, inv_dx);
        const size_t q = (size_t)(j - 1) * (nx - 1) + (i - 1);
        gx[q] = (T)((CT)gx[q] + (CT)avx[q] * D);
    }
    if (j <= ny - 1) {
        CT D = fd4_bd<T, CT>(p + (i - 1), nx, ny, j, c, inv_dy);
        const size_t q = (size_t)(j - 1) * nx + (i - 1);
        gy[q] = (T)((CT)gy[q] + (CT)avy[q] * D);
    }
}

template <class T>
static VdParams<T> make_params(const swb_acou_vd_step_args &a)
{
    VdParams<T> P{};
    SWB_REQUIRE(a.halo >= 0, "CPML halo size must be non-negative!");
    for (int d = 0; d < 2; ++d) // n[1] = 1: a 1D grid
        SWB_REQUIRE(a.n[d] >= 2 * (int64_t)a.halo + 3 || (d == 1 && a.n[1] == 1), "Number grid points in the dimensions with C-PML boundaries must be at least 2*halo+3!");
    P.halo = a.halo;
    P.nx = a.n[0];
    P.ny = a.n[1];
    P.inv_dx = (T)1 / (T)a.spacing[0];
    P.inv_dy = a.n[1] == 1 ? (T)0 : (T)1 / (T)a.spacing[1];
    P.p = (T *)a.pcur;
    P.vx = (T *)a.vcur[0];
    P.vy = (T *)a.vcur[1];
    P.m0 = (const T *)a.fact_m0;
    P.m1x = (const T *)a.fact_m1_stag[0];
    P.m1y = (const T *)a.fact_m1_stag[1];
    P.psi_x = (T *)a.psi[0];
    P.psi_y = (T *)a.psi[1];
    P.xi_x = (T *)a.xi[0];
    P.xi_y = (T *)a.xi[1];
    P.a_x = (const T *)a.cpml[0].a;
    P.b_x = (const T *)a.cpml[0].b;
    P.a_xh = (const T *)a.cpml[0].a_h;
    P.b_xh = (const T *)a.cpml[0].b_h;
    P.a_y = (const T *)a.cpml[1].a;
    P.b_y = (const T *)a.cpml[1].b;
    P.a_yh = (const T *)a.cpml[1].a_h;
    P.b_yh = (const T *)a.cpml[1].b_h;
    const FdWeights &w = fd_weights();
    for (int k = 0; k < 4; ++k)
        P.c4[k] = w.d1o4[k];
    return P;
}

template <class T, class CT>
static void launch_p(const VdParams<T> &P, cudaStream_t st)
{
    dim3 blk(32, 8, 1), grd(cdiv(P.nx - 2, 32), P.ny == 1 ? 1 : cdiv(P.ny - 2, 8), 1);
    if (P.ny == 1)
        blk = dim3(256, 1, 1), grd = dim3(cdiv(P.nx - 2, 256), 1, 1);
    vd_update_p_kernel<T, CT><<<grd, blk, 0, st>>>(P);
    check_launch("vd_update_p");
    count_launch();
}

template <class T, class CT>
static void launch_v(const VdParams<T> &P, cudaStream_t st)
{
    dim3 blk(32, 8, 1), grd(cdiv(P.nx, 32), cdiv(P.ny, 8), 1);
    vd_update_v_kernel<T, CT><<<grd, blk, 0, st>>>(P);
    check_launch("vd_update_v");
    count_launch();
}

template <class T, class CT>
static void vd_step_impl(const swb_acou_vd_step_args &a, bool adjoint)
{
    VdParams<T> P = make_params<T>(a);
    cudaStream_t st = (cudaStream_t)a.stream;
    if (!adjoint) { // p, inject, v, record (acoustic2D_VD_xPU.jl:117-137)
        const int nd = a.n[1] == 1 ? 1 : 2; // positions are (npos, ndim)
        launch_p<T, CT>(P, st);
        launch_inject<T>(P.p, nd, a.n, a.src, a.it, st);
        launch_v<T, CT>(P, st);
        launch_record<T>(P.p, nd, a.n, a.rec, a.it, st);
    } else { // v, p, inject (acoustic2D_VD_xPU.jl:163-178)
        const int nd = a.n[1] == 1 ? 1 : 2;
        launch_v<T, CT>(P, st);
        launch_p<T, CT>(P, st);
        launch_inject<T>(P.p, nd, a.n, a.src, a.it, st);
    }
}

void vd_step(const swb_acou_vd_step_args &a, bool adjoint)
{
    if (a.dtype == SWB_F64)
        vd_step_impl<double, double>(a, adjoint);
    else if (a.dtype == SWB_F32) {
        if (a.flags & SWB_FLAG_FAST_F32)
            vd_step_impl<float, float>(a, adjoint);
        else
            vd_step_impl<float, double>(a, adjoint);
    } else
        throw Error(SWB_ERR_ARG, "dtype must be SWB_F32 or SWB_F64");
}

void vd_correlate_m0(int dtype, size_t ncells, void *g, const void *adjp, const void *p_it, const void *p_itm1, double dt, cudaStream_t st)
{
    unsigned blocks = (unsigned)std::min<size_t>((ncells + 255) / 256, 148u * 16u);
    if (blocks == 0)
        return;
    if (dtype == SWB_F64)
        vd_correlate_m0_kernel<double><<<blocks, 256, 0, st>>>((double *)g, (const double *)adjp, (const double *)p_it, (const double *)p_itm1, 1.0 / dt, ncells);
    else
        vd_correlate_m0_kernel<float><<<blocks, 256, 0, st>>>((float *)g, (const float *)adjp, (const float *)p_it, (const float *)p_itm1, 1.0f / (float)dt, ncells);
    check_launch("vd_correlate_m0");
    count_launch();
}

template <class T, class CT>
static void corr_m1_impl(const int64_t *n, const double *spacing, void *const g[2], const void *const av[2], const void *p, cudaStream_t st)
{
    const FdWeights &w = fd_weights();
    dim3 blk(32, 8, 1), grd(cdiv(n[0], 32), cdiv(n[1], 8), 1);
    vd_correlate_m1_kernel<T, CT><<<grd, blk, 0, st>>>((T *)g[0], (T *)g[1], (const T *)av[0], (const T *)av[1], (const T *)p, n[0], n[1],
                                                       (T)1 / (T)spacing[0], (T)1 / (T)spacing[1], w.d1o4[0], w.d1o4[1], w.d1o4[2], w.d1o4[3]);
    check_launch("vd_correlate_m1");
    count_launch();
}

void vd_correlate_m1(int dtype, int flags, const int64_t *n, const double *spacing, void *const g[2], const void *const av[2], const void *p, cudaStream_t st)
{
    if (dtype == SWB_F64)
        corr_m1_impl<double, double>(n, spacing, g, av, p, st);
    else if (flags & SWB_FLAG_FAST_F32)
        corr_m1_impl<float, float>(n, spacing, g, av, p, st);
    else
        corr_m1_impl<float, double>(n, spacing, g, av, p, st);
}

} // namespace swb
