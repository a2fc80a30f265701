// acou_cd.cu -- acoustic constant-density pressure update, 2D and 3D, C-PML.
//
// Reference semantics: src/models/acoustic/backends/shared/acoustic2D_xPU.jl:1-169,
// acoustic3D_xPU.jl:1-219, correlate_gradient_xPU.jl:1-10 and the stencil generators
// src/utils/fdgen.jl:65-215.  This file holds the "one launch per reference kernel" path
// (psi sweep, p/xi sweep, inject, record) used by the fine-grained C ABI, plus the fused
// single-launch step used by the engine (psi is double-buffered there so that the psi
// update, the xi update and the pressure update happen in one pass over memory).
#include "common.cuh"
#include "kernels.h"

namespace swb {

template <class T>
struct CdParams {
    int ndim, halo;
    long long n[3];
    T inv_d[3];
    const T *pold, *pcur, *fact;
    T *pnew;
    T *psi[3];        // updated in place (L1 path) or read (fused path: psi_in)
    T *psi_out[3];    // fused path only
    T *xi[3];
    const T *a[3], *b[3], *a_h[3], *b_h[3];
    double c1[2], c2[3];
    // fused path: sources / receivers
    long long nsrc, nrec, nt_tf, nt_tr, it;
    const long long *possrc, *posrec;
    const T *srctf;
    T *traces;
};

// ------------------------------------------------------------------------------------------------
// psi sweep: one thread per compact psi element of every axis (update_ψ_x!/y!/z!)
// ------------------------------------------------------------------------------------------------
template <class T, class CT>
__global__ void __launch_bounds__(256) cd_update_psi_kernel(CdParams<T> P, long long cnt0, long long cnt1, long long total)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total)
        return;
    int ax = 0;
    if (t >= cnt0) {
        t -= cnt0;
        ax = 1;
        if (t >= cnt1) {
            t -= cnt1;
            ax = 2;
        }
    }
    long long m[3] = {P.n[0], P.n[1], P.n[2]};
    m[ax] = 2 * P.halo;
    long long c[3];
    c[0] = t % m[0];
    long long r = t / m[0];
    c[1] = r % m[1];
    c[2] = r / m[1];
    long long g[3] = {c[0], c[1], c[2]}; // 0-based grid index
    long long cc = c[ax] + 1;            // 1-based compact index
    long long gi = cc > P.halo ? P.n[ax] - P.halo - 1 + (cc - P.halo) : cc; // 1-based grid index along ax
    g[ax] = gi - 1;
    const long long pst[3] = {1, P.n[0], P.n[0] * P.n[1]};
    size_t q = lin3(g[0], g[1], g[2], P.n[0], P.n[1]);
    // (c1[0]*p[gi] + c1[1]*p[gi+1]) * _d   (fdgen.jl:126, offsets {0,+1}; gi+1 <= n always holds here)
    CT D = ((CT)P.c1[0] * (CT)P.pcur[q] + (CT)P.c1[1] * (CT)P.pcur[q + pst[ax]]) * (CT)P.inv_d[ax];
    T *ps = P.psi[ax] + t;
    T pn;
    (void)cpml_apply<T, CT>(D, P.a_h[ax][cc - 1], P.b_h[ax][cc - 1], *ps, pn);
    *ps = pn;
}

// one axis of @∇̃² (fdgen.jl:163-193) for an interior cell; psi already holds this step's values
template <class T, class CT>
__device__ __forceinline__ CT cd_d2_axis(const CdParams<T> &P, int ax, const T *pc, long long stride, long long i1 /*1-based*/,
                                         const T *psi_base, long long pstride, T *xi_base, long long xstride)
{
    const long long n = P.n[ax];
    const int halo = P.halo;
    T inv = P.inv_d[ax];
    T inv2 = inv * inv;
    CT D2 = (((CT)P.c2[0] * (CT)pc[-stride] + (CT)P.c2[1] * (CT)pc[0]) + (CT)P.c2[2] * (CT)pc[stride]) * (CT)inv2;
    long long ii;
    if (i1 <= halo)
        ii = i1;
    else if (i1 >= n - halo + 1)
        ii = i1 - (n - halo) + 1 + halo;
    else
        return D2;
    CT dpsi = ((CT)P.c1[0] * (CT)psi_base[(ii - 2) * pstride] + (CT)P.c1[1] * (CT)psi_base[(ii - 1) * pstride]) * (CT)inv;
    T *x = xi_base + (ii - 1) * xstride;
    T bx = P.b[ax][ii - 1] * *x;
    T xn = (T)((CT)bx + (CT)P.a[ax][ii - 1] * (D2 + dpsi));
    *x = xn;
    return (D2 + dpsi) + (CT)xn;
}

// p / xi sweep (update_p_CPML!): one thread per interior cell
template <class T, class CT, int NDIM>
__global__ void __launch_bounds__(256) cd_update_p_kernel(CdParams<T> P)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x + 1; // 0-based, interior starts at 1
    const long long j = NDIM >= 2 ? (long long)blockIdx.y * blockDim.y + threadIdx.y + 1 : 0;
    const long long k = NDIM == 3 ? (long long)blockIdx.z + 1 : 0;
    const long long nx = P.n[0], ny = P.n[1];
    if (i > nx - 2 || (NDIM >= 2 && j > ny - 2))
        return;
    const size_t q = lin3(i, j, k, nx, ny);
    const T *pc = P.pcur + q;
    const int h = P.halo;
    // psi_x (2h, ny, nz), xi_x (2(h+1), ny, nz): base = element with compact index 1 at (j,k)
    CT lap = cd_d2_axis<T, CT>(P, 0, pc, 1, i + 1, P.psi[0] + (size_t)(2 * h) * ((size_t)k * ny + j), 1,
                               P.xi[0] + (size_t)(2 * (h + 1)) * ((size_t)k * ny + j), 1);
    // psi_y (nx, 2h, nz), xi_y (nx, 2(h+1), nz)
    if (NDIM >= 2)
        lap = lap + cd_d2_axis<T, CT>(P, 1, pc, nx, j + 1, P.psi[1] + (size_t)k * nx * (2 * h) + i, nx,
                                      P.xi[1] + (size_t)k * nx * (2 * (h + 1)) + i, nx);
    if (NDIM == 3) // psi_z (nx, ny, 2h), xi_z (nx, ny, 2(h+1))
        lap = lap + cd_d2_axis<T, CT>(P, 2, pc, nx * ny, k + 1, P.psi[2] + (size_t)j * nx + i, nx * ny,
                                      P.xi[2] + (size_t)j * nx + i, nx * ny);
    // pnew = 2.0*pcur - pold + fact*lap   (acoustic2D_xPU.jl:43)
    P.pnew[q] = (T)((((CT)2.0 * (CT)pc[0]) - (CT)P.pold[q]) + (CT)P.fact[q] * lap);
}

// inject_sources!: p[pos_s] += tf[it, s].  Sources that share a cell are summed by the first of
// them in index order, which reproduces the CPU loop order deterministically.
template <class T>
__global__ void inject_kernel(T *p, int ndim, long long n0, long long n1, const long long *pos, long long npos, const T *tf, long long nt, long long it)
{
    long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= npos)
        return;
    long long i = pos[s] - 1, j = ndim >= 2 ? pos[s + npos] - 1 : 0, k = ndim >= 3 ? pos[s + 2 * npos] - 1 : 0;
    size_t q = lin3(i, j, k, n0, n1);
    for (long long r = 0; r < s; ++r) {
        long long ri = pos[r] - 1, rj = ndim >= 2 ? pos[r + npos] - 1 : 0, rk = ndim >= 3 ? pos[r + 2 * npos] - 1 : 0;
        if (ri == i && rj == j && rk == k)
            return; // an earlier source owns this cell
    }
    T v = p[q];
    for (long long r = s; r < npos; ++r) {
        long long ri = pos[r] - 1, rj = ndim >= 2 ? pos[r + npos] - 1 : 0, rk = ndim >= 3 ? pos[r + 2 * npos] - 1 : 0;
        if (ri == i && rj == j && rk == k)
            v = v + tf[(size_t)r * nt + (it - 1)];
    }
    p[q] = v;
}

template <class T>
__global__ void record_kernel(const T *p, int ndim, long long n0, long long n1, const long long *pos, long long npos, T *traces, long long nt, long long it)
{
    long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= npos)
        return;
    long long i = pos[r] - 1, j = ndim >= 2 ? pos[r + npos] - 1 : 0, k = ndim >= 3 ? pos[r + 2 * npos] - 1 : 0;
    traces[(size_t)r * nt + (it - 1)] = p[lin3(i, j, k, n0, n1)];
}

template <class T>
__global__ void prescale_kernel(T *res, long long nt, int ndim, long long n0, long long n1, const long long *pos, long long npos, const T *fact)
{
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long r = blockIdx.y;
    if (t >= nt || r >= npos)
        return;
    long long i = pos[r] - 1, j = ndim >= 2 ? pos[r + npos] - 1 : 0, k = ndim >= 3 ? pos[r + 2 * npos] - 1 : 0;
    res[(size_t)r * nt + t] *= fact[lin3(i, j, k, n0, n1)];
}

// correlate_gradient!: grad = grad + adj*(p_itm2 - 2.0*p_itm1 + p_it)*_dt2
template <class T, class CT>
__global__ void __launch_bounds__(256) cd_correlate_kernel(T *grad, const T *adj, const T *pm2, const T *pm1, const T *p0, T _dt2, size_t n)
{
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; q < n; q += stride) {
        CT lapt = ((CT)pm2[q] - (CT)2.0 * (CT)pm1[q]) + (CT)p0[q];
        grad[q] = (T)((CT)grad[q] + ((CT)adj[q] * lapt) * (CT)_dt2);
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
template <class T>
static CdParams<T> make_params(const swb_acou_cd_step_args &a)
{
    CdParams<T> P{};
    SWB_REQUIRE(a.ndim >= 1 && a.ndim <= 3, "acoustic CD: ndim must be 1, 2 or 3");
    SWB_REQUIRE(a.halo >= 0, "CPML halo size must be non-negative!");
    P.ndim = a.ndim;
    P.halo = a.halo;
    for (int d = 0; d < 3; ++d) {
        P.n[d] = d < a.ndim ? a.n[d] : 1;
        if (d < a.ndim) {
            SWB_REQUIRE(a.n[d] >= 2 * (int64_t)a.halo + 3, "Number grid points in the dimensions with C-PML boundaries must be at least 2*halo+3!");
            P.inv_d[d] = (T)1 / (T)a.spacing[d];
        }
        P.psi[d] = (T *)a.psi[d];
        P.xi[d] = (T *)a.xi[d];
        P.a[d] = (const T *)a.cpml[d].a;
        P.b[d] = (const T *)a.cpml[d].b;
        P.a_h[d] = (const T *)a.cpml[d].a_h;
        P.b_h[d] = (const T *)a.cpml[d].b_h;
    }
    P.pold = (const T *)a.pold;
    P.pcur = (const T *)a.pcur;
    P.pnew = (T *)a.pnew;
    P.fact = (const T *)a.fact;
    const FdWeights &w = fd_weights();
    P.c1[0] = w.d1o2[0];
    P.c1[1] = w.d1o2[1];
    for (int k = 0; k < 3; ++k)
        P.c2[k] = w.d2o2[k];
    return P;
}

template <class T>
void launch_inject(T *p, int ndim, const int64_t *n, const swb_points &pts, int64_t it, cudaStream_t st)
{
    if (pts.n <= 0)
        return;
    SWB_REQUIRE(it >= 1 && it <= pts.nt, "time index out of range of the source time function");
    inject_kernel<T><<<cdiv(pts.n, 128), 128, 0, st>>>(p, ndim, n[0], ndim >= 2 ? n[1] : 1, (const long long *)pts.pos, pts.n, (const T *)pts.tf, pts.nt, it);
    check_launch("inject");
    count_launch();
}

template <class T>
void launch_record(const T *p, int ndim, const int64_t *n, const swb_points &pts, int64_t it, cudaStream_t st)
{
    if (pts.n <= 0 || pts.tf == nullptr)
        return;
    SWB_REQUIRE(it >= 1 && it <= pts.nt, "time index out of range of the trace buffer");
    record_kernel<T><<<cdiv(pts.n, 128), 128, 0, st>>>(p, ndim, n[0], ndim >= 2 ? n[1] : 1, (const long long *)pts.pos, pts.n, (T *)pts.tf, pts.nt, it);
    check_launch("record");
    count_launch();
}

template void launch_inject<float>(float *, int, const int64_t *, const swb_points &, int64_t, cudaStream_t);
template void launch_inject<double>(double *, int, const int64_t *, const swb_points &, int64_t, cudaStream_t);
template void launch_record<float>(const float *, int, const int64_t *, const swb_points &, int64_t, cudaStream_t);
template void launch_record<double>(const double *, int, const int64_t *, const swb_points &, int64_t, cudaStream_t);

template <class T, class CT>
static void cd_step_impl(const swb_acou_cd_step_args &a, bool record)
{
    CdParams<T> P = make_params<T>(a);
    cudaStream_t st = (cudaStream_t)a.stream;
    const int h = a.halo;
    if (h > 0) {
        long long cnt[3] = {0, 0, 0};
        for (int ax = 0; ax < a.ndim; ++ax) {
            long long c = 2 * h;
            for (int d = 0; d < a.ndim; ++d)
                if (d != ax)
                    c *= P.n[d];
            cnt[ax] = c;
        }
        long long total = cnt[0] + cnt[1] + cnt[2];
        cd_update_psi_kernel<T, CT><<<cdiv(total, 256), 256, 0, st>>>(P, cnt[0], cnt[1], total);
        check_launch("cd_update_psi");
        count_launch();
    }
    dim3 blk(32, 8, 1);
    if (a.ndim == 1) { // acoustic1D_xPU.jl:78-95
        cd_update_p_kernel<T, CT, 1><<<cdiv(P.n[0] - 2, 256), 256, 0, st>>>(P);
    } else if (a.ndim == 2) {
        dim3 grd(cdiv(P.n[0] - 2, 32), cdiv(P.n[1] - 2, 8), 1);
        cd_update_p_kernel<T, CT, 2><<<grd, blk, 0, st>>>(P);
    } else {
        SWB_REQUIRE(P.n[2] - 2 <= 65535, "3D grid too deep for this launch geometry");
        dim3 grd(cdiv(P.n[0] - 2, 32), cdiv(P.n[1] - 2, 8), (unsigned)(P.n[2] - 2));
        cd_update_p_kernel<T, CT, 3><<<grd, blk, 0, st>>>(P);
    }
    check_launch("cd_update_p");
    count_launch();
    launch_inject<T>((T *)a.pnew, a.ndim, a.n, a.src, a.it, st);
    if (record)
        launch_record<T>((const T *)a.pnew, a.ndim, a.n, a.rec, a.it, st);
}

void cd_step(const swb_acou_cd_step_args &a, bool record)
{
    if (a.dtype == SWB_F64)
        cd_step_impl<double, double>(a, record);
    else if (a.dtype == SWB_F32) {
        if (a.flags & SWB_FLAG_FAST_F32)
            cd_step_impl<float, float>(a, record);
        else
            cd_step_impl<float, double>(a, record);
    } else
        throw Error(SWB_ERR_ARG, "dtype must be SWB_F32 or SWB_F64");
}

void cd_correlate(int dtype, int flags, size_t ncells, void *grad, const void *adj, const void *pm2, const void *pm1, const void *p0, double dt, cudaStream_t st)
{
    unsigned blocks = (unsigned)std::min<size_t>((ncells + 255) / 256, 148u * 16u);
    if (blocks == 0)
        return;
    if (dtype == SWB_F64) {
        double _dt2 = 1.0 / (dt * dt);
        cd_correlate_kernel<double, double><<<blocks, 256, 0, st>>>((double *)grad, (const double *)adj, (const double *)pm2, (const double *)pm1, (const double *)p0, _dt2, ncells);
    } else {
        float dtf = (float)dt;
        float _dt2 = 1.0f / (dtf * dtf);
        if (flags & SWB_FLAG_FAST_F32)
            cd_correlate_kernel<float, float><<<blocks, 256, 0, st>>>((float *)grad, (const float *)adj, (const float *)pm2, (const float *)pm1, (const float *)p0, _dt2, ncells);
        else
            cd_correlate_kernel<float, double><<<blocks, 256, 0, st>>>((float *)grad, (const float *)adj, (const float *)pm2, (const float *)pm1, (const float *)p0, _dt2, ncells);
    }
    check_launch("cd_correlate");
    count_launch();
}

void prescale_residuals(int dtype, int ndim, const int64_t *n, void *res, int64_t nt, int64_t nrec, const int64_t *pos, const void *fact, cudaStream_t st)
{
    if (nrec <= 0 || nt <= 0)
        return;
    SWB_REQUIRE(nrec <= 65535, "too many receivers for this launch geometry");
    dim3 grd(cdiv(nt, 128), (unsigned)nrec, 1);
    if (dtype == SWB_F64)
        prescale_kernel<double><<<grd, 128, 0, st>>>((double *)res, nt, ndim, n[0], ndim >= 2 ? n[1] : 1, (const long long *)pos, nrec, (const double *)fact);
    else
        prescale_kernel<float><<<grd, 128, 0, st>>>((float *)res, nt, ndim, n[0], ndim >= 2 ? n[1] : 1, (const long long *)pos, nrec, (const float *)fact);
    check_launch("prescale_residuals");
    count_launch();
}

} // namespace swb
