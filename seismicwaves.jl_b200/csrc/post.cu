// post.cu -- device-side versions of the O(N) host steps that sit either side of the time loops:
// material factors (precompute_fact!), gradient post-processing (back_interp, mute, chain rules),
// L2 adjoint source + misfit, gradient accumulation.  Each kernel reproduces the reference's
// arithmetic (types and association order) so that results match the CPU path.
#include "common.cuh"
#include <map>
#include <mutex>
#include <vector>
#include <algorithm>
#include "kernels.h"
#include <vector>
#include <algorithm>
#include <cmath>

namespace swb {

static inline unsigned grid1d(size_t n) { return (unsigned)std::min<size_t>((n + 255) / 256, 148u * 32u); }

#define GRID_STRIDE(q, n) for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < (n); q += (size_t)gridDim.x * blockDim.x)

// precompute_fact! (acou_models.jl:60): fact = (dt^2) .* (vp .^ 2), all in T
template <class T>
__global__ void cd_fact_kernel(const T *vp, T dt2, T *fact, size_t n)
{
    GRID_STRIDE(q, n) { fact[q] = dt2 * (vp[q] * vp[q]); }
}

void post_cd_fact(int dtype, size_t n, const void *vp, double dt, void *fact, cudaStream_t st)
{
    if (dtype == SWB_F64)
        cd_fact_kernel<double><<<grid1d(n), 256, 0, st>>>((const double *)vp, dt * dt, (double *)fact, n);
    else {
        float d = (float)dt;
        cd_fact_kernel<float><<<grid1d(n), 256, 0, st>>>((const float *)vp, d * d, (float *)fact, n);
    }
    check_launch("cd_fact");
    count_launch();
}

// precompute_fact! (acou_models.jl:292-300):
//   fact_m0 = vp.^2 .* rho .* dt                               (T)
//   fact_m1_stag[d] = interp(1 ./ rho, d) .* dt                (Float64 interpolation weights, rounded to T on copyto!)
// interp: 0 = arithmetic mean of the two neighbours, 1 = harmonic (1 ./ itp(1 ./ m)) -- utils/interpolations.jl:35-43
template <class T>
__global__ void vd_facts_kernel(const T *vp, const T *rho, T dt, int interp, T *m0, T *m1x, T *m1y, long long nx, long long ny)
{
    const size_t n = (size_t)nx * ny;
    GRID_STRIDE(q, n)
    {
        const long long i = q % nx, j = q / nx;
        const T v = vp[q], r = rho[q];
        m0[q] = ((v * v) * r) * dt;
        const T ir = (T)1 / r;
        if (i < nx - 1) {
            const T ir2 = (T)1 / rho[q + 1];
            double m;
            if (interp == 0)
                m = 0.5 * (double)ir + 0.5 * (double)ir2;
            else
                m = 1.0 / (0.5 * (double)((T)1 / ir) + 0.5 * (double)((T)1 / ir2));
            m1x[(size_t)j * (nx - 1) + i] = (T)(m * (double)dt);
        }
        if (j < ny - 1) {
            const T ir2 = (T)1 / rho[q + nx];
            double m;
            if (interp == 0)
                m = 0.5 * (double)ir + 0.5 * (double)ir2;
            else
                m = 1.0 / (0.5 * (double)((T)1 / ir) + 0.5 * (double)((T)1 / ir2));
            m1y[q] = (T)(m * (double)dt);
        }
    }
}

void post_vd_facts(int dtype, const int64_t *n, const void *vp, const void *rho, double dt, int interp, void *m0, void *m1x, void *m1y, cudaStream_t st)
{
    const size_t nn = (size_t)n[0] * n[1];
    if (dtype == SWB_F64)
        vd_facts_kernel<double><<<grid1d(nn), 256, 0, st>>>((const double *)vp, (const double *)rho, dt, interp, (double *)m0, (double *)m1x, (double *)m1y, n[0], n[1]);
    else
        vd_facts_kernel<float><<<grid1d(nn), 256, 0, st>>>((const float *)vp, (const float *)rho, (float)dt, interp, (float *)m0, (float *)m1x, (float *)m1y, n[0], n[1]);
    check_launch("vd_facts");
    count_launch();
}

// mutearoundmultiplepoints! (utils/mute_grad.jl:3-80).  One thread per cell of the bounding box of all
// windows; each thread walks the points in order, so overlapping windows multiply in the reference's order.
// win: (npos, ndim) int64 centre indices ijkpt (1-based), pos: (npos, ndim) of T in metres.
template <class T>
__global__ void mute_kernel(T *arr, int ndim, long long n0, long long n1, long long n2, T d0, T d1, T d2, const long long *win, const T *pos,
                            long long npos, int radius, T rmax, long long lo0, long long lo1, long long lo2, long long e0, long long e1, long long e2)
{
    const size_t tot = (size_t)e0 * e1 * e2;
    GRID_STRIDE(t, tot)
    {
        const long long i = lo0 + (long long)(t % e0), j = lo1 + (long long)((t / e0) % e1), k = lo2 + (long long)(t / ((size_t)e0 * e1)); // 1-based
        const size_t q = lin3(i - 1, j - 1, k - 1, n0, n1);
        T val = arr[q];
        bool touched = false;
        // xyzcur[d] = (idx-1)*spacing[d]: T arithmetic, stored in Float64
        const double c0 = (double)((T)(i - 1) * d0), c1 = (double)((T)(j - 1) * d1), c2 = (double)((T)(k - 1) * d2);
        for (long long s = 0; s < npos; ++s) {
            const long long w0 = win[s], w1 = ndim >= 2 ? win[s + npos] : 1, w2 = ndim >= 3 ? win[s + 2 * npos] : 1;
            if (i < w0 - radius || i > w0 + radius)
                continue;
            if (ndim >= 2 && (j < w1 - radius || j > w1 + radius))
                continue;
            if (ndim >= 3 && (k < w2 - radius || k > w2 + radius))
                continue;
            double x = (double)pos[s] - c0;
            double r2 = x * x;
            if (ndim >= 2) {
                double y = (double)pos[s + npos] - c1;
                r2 = r2 + y * y;
            }
            if (ndim >= 3) {
                double z = (double)pos[s + 2 * npos] - c2;
                r2 = r2 + z * z;
            }
            const double r = sqrt(r2);
            if (r <= (double)rmax) {
                const double att = r / (double)rmax;
                val = (T)((double)val * att);
                touched = true;
            }
        }
        if (touched)
            arr[q] = val;
    }
}

// grow-only per-device scratch for the window centres: a stream-ordered cudaMallocAsync / cudaFreeAsync pair here made
// the memory pool trim itself at the following synchronisation, which was measured at 0.1 - 1 s per shot
static void *mute_scratch(size_t bytes)
{
    static std::mutex mtx;
    static std::map<int, std::pair<void *, size_t>> cache;
    std::lock_guard<std::mutex> lk(mtx);
    int dev = 0;
    SWB_CUDA(cudaGetDevice(&dev));
    auto &e = cache[dev];
    if (e.second < bytes) {
        if (e.first)
            cudaFree(e.first);
        e.first = nullptr;
        e.second = 0;
        const size_t want = std::max<size_t>(bytes, 1 << 16);
        SWB_CUDA(cudaMalloc(&e.first, want));
        e.second = want;
    }
    return e.first;
}

template <class T>
static void mute_impl(int ndim, const int64_t *n, const double *spacing, void *arr, int64_t npos, const void *dev_positions, int radius, cudaStream_t st)
{
    if (radius == 0 || npos == 0)
        return;
    SWB_REQUIRE(radius > 0, "mutearoundpoint!(): The smoothing radius must be positive.");
    // window centres on the host: ijkpt = floor(div(x, spacing)) + 1 with Base.div = round((x - rem(x, y)) / y)
    std::vector<T> hpos((size_t)npos * ndim);
    SWB_CUDA(cudaMemcpyAsync(hpos.data(), dev_positions, hpos.size() * sizeof(T), cudaMemcpyDeviceToHost, st));
    SWB_CUDA(cudaStreamSynchronize(st));
    std::vector<long long> win((size_t)npos * ndim);
    long long lo[3] = {1, 1, 1}, hi[3] = {1, 1, 1};
    T maxsp = (T)spacing[0];
    for (int d = 1; d < ndim; ++d)
        maxsp = std::max(maxsp, (T)spacing[d]);
    const T rmax = (T)radius * maxsp;
    for (int d = 0; d < ndim; ++d) {
        lo[d] = n[d] + 1;
        hi[d] = 0;
        const T sp = (T)spacing[d];
        const T extent = sp * (T)(n[d] - 1);
        for (int64_t s = 0; s < npos; ++s) {
            const T x = hpos[(size_t)d * npos + s];
            if (!((T)0 <= x && x <= extent))
                throw Error(SWB_ERR_ARG, "mutearoundpoint!(): The point lies outside the grid on dimension " + std::to_string(d + 1) + ".");
            const T res = std::nearbyint((x - std::fmod(x, sp)) / sp);
            const long long c = (long long)std::floor(res) + 1;
            win[(size_t)d * npos + s] = c;
            lo[d] = std::min<long long>(lo[d], c - radius);
            hi[d] = std::max<long long>(hi[d], c + radius);
        }
        lo[d] = std::max<long long>(lo[d], 1);
        hi[d] = std::min<long long>(hi[d], n[d]);
        if (hi[d] < lo[d])
            return;
    }
    long long *dwin = (long long *)mute_scratch(win.size() * sizeof(long long));
    SWB_CUDA(cudaMemcpyAsync(dwin, win.data(), win.size() * sizeof(long long), cudaMemcpyHostToDevice, st));
    const long long e0 = hi[0] - lo[0] + 1, e1 = ndim >= 2 ? hi[1] - lo[1] + 1 : 1, e2 = ndim >= 3 ? hi[2] - lo[2] + 1 : 1;
    mute_kernel<T><<<grid1d((size_t)e0 * e1 * e2), 256, 0, st>>>((T *)arr, ndim, n[0], ndim >= 2 ? n[1] : 1, ndim >= 3 ? n[2] : 1, (T)spacing[0],
                                                                 ndim >= 2 ? (T)spacing[1] : (T)0, ndim >= 3 ? (T)spacing[2] : (T)0, dwin,
                                                                 (const T *)dev_positions, npos, radius, rmax, lo[0], lo[1], lo[2], e0, e1, e2);
    check_launch("mute");
    count_launch();
    SWB_CUDA(cudaStreamSynchronize(st)); // `win` (host) was the async copy source
}

void post_mute(int dtype, int ndim, const int64_t *n, const double *spacing, void *arr, int64_t npos, const void *dev_positions, int radius, cudaStream_t st)
{
    if (dtype == SWB_F64)
        mute_impl<double>(ndim, n, spacing, arr, npos, dev_positions, radius, st);
    else
        mute_impl<float>(ndim, n, spacing, arr, npos, dev_positions, radius, st);
}

// acou_gradient.jl:93 + accumulate_gradient! (acou_models.jl:64):
//   gradient = (convert(T, 2.0) ./ (vp .^ 3)) .* gradient ; totgrad .+= gradient
template <class T>
__global__ void cd_chain_kernel(const T *g, const T *vp, T *tot, size_t n)
{
    GRID_STRIDE(q, n)
    {
        const T v = vp[q];
        const T v3 = (v * v) * v;
        const T cur = ((T)2 / v3) * g[q];
        tot[q] = tot[q] + cur;
    }
}

void post_cd_chain_accumulate(int dtype, size_t n, const void *g, const void *vp, void *total, cudaStream_t st)
{
    if (dtype == SWB_F64)
        cd_chain_kernel<double><<<grid1d(n), 256, 0, st>>>((const double *)g, (const double *)vp, (double *)total, n);
    else
        cd_chain_kernel<float><<<grid1d(n), 256, 0, st>>>((const float *)g, (const float *)vp, (float *)total, n);
    check_launch("cd_chain");
    count_launch();
}

// back_interp of the staggered m1 gradients onto the grid (acou_gradient.jl:182-185, utils/interpolations.jl:14-28,45-47)
//   g1 = (0 + back_x) + back_y ;  back_d[i] = (0 + g_d[i]*w_lo) + g_d[i-1]*w_hi
// arithmetic: w = T(0.5); harmonic: w = (itp^2 / m[idx]^2) / 2 in Float64 with m = 1 ./ rho
template <class T>
__device__ __forceinline__ T backinterp_pair(T g_here, bool has_here, T g_prev, bool has_prev, int interp, T m_self, T m_next, T m_prev)
{
    T res = (T)0;
    if (interp == 0) {
        if (has_here)
            res = res + g_here * (T)0.5;
        if (has_prev)
            res = res + g_prev * (T)0.5;
        return res;
    }
    if (has_here) { // staggered point between self and next; derivative w.r.t. the lower node (perm 0)
        const double itp = 1.0 / (0.5 * (double)((T)1 / m_self) + 0.5 * (double)((T)1 / m_next));
        const double w = ((itp * itp) / (double)(m_self * m_self)) / 2.0;
        res = (T)((double)res + (double)g_here * w);
    }
    if (has_prev) { // staggered point between prev and self; derivative w.r.t. the upper node (perm 1)
        const double itp = 1.0 / (0.5 * (double)((T)1 / m_prev) + 0.5 * (double)((T)1 / m_self));
        const double w = ((itp * itp) / (double)(m_self * m_self)) / 2.0;
        res = (T)((double)res + (double)g_prev * w);
    }
    return res;
}

template <class T>
__global__ void vd_backinterp_kernel(const T *rho, int interp, const T *gx, const T *gy, T *g1, long long nx, long long ny)
{
    const size_t n = (size_t)nx * ny;
    GRID_STRIDE(q, n)
    {
        const long long i = q % nx, j = q / nx;
        const T m = (T)1 / rho[q];
        const T mxn = i < nx - 1 ? (T)1 / rho[q + 1] : (T)0, mxp = i > 0 ? (T)1 / rho[q - 1] : (T)0;
        const T myn = j < ny - 1 ? (T)1 / rho[q + nx] : (T)0, myp = j > 0 ? (T)1 / rho[q - nx] : (T)0;
        const T bx = backinterp_pair<T>(i < nx - 1 ? gx[(size_t)j * (nx - 1) + i] : (T)0, i < nx - 1, i > 0 ? gx[(size_t)j * (nx - 1) + i - 1] : (T)0, i > 0,
                                        interp, m, mxn, mxp);
        const T by = backinterp_pair<T>(j < ny - 1 ? gy[q] : (T)0, j < ny - 1, j > 0 ? gy[q - nx] : (T)0, j > 0, interp, m, myn, myp);
        g1[q] = ((T)0 + bx) + by;
    }
}

void post_vd_backinterp(int dtype, const int64_t *n, const void *rho, int interp, const void *g1x, const void *g1y, void *g1, cudaStream_t st)
{
    const size_t nn = (size_t)n[0] * n[1];
    if (dtype == SWB_F64)
        vd_backinterp_kernel<double><<<grid1d(nn), 256, 0, st>>>((const double *)rho, interp, (const double *)g1x, (const double *)g1y, (double *)g1, n[0], n[1]);
    else
        vd_backinterp_kernel<float><<<grid1d(nn), 256, 0, st>>>((const float *)rho, interp, (const float *)g1x, (const float *)g1y, (float *)g1, n[0], n[1]);
    check_launch("vd_backinterp");
    count_launch();
}

// acou_gradient.jl:199-202 + accumulate_gradient! (acou_models.jl:302-305):
//   vp : -2 .* g0 ./ (vp.^3 .* rho) ;  rho : -g0 ./ (vp.^2 .* rho.^2) .- g1 ./ rho
template <class T>
__global__ void vd_chain_kernel(const T *g0, const T *g1, const T *vp, const T *rho, T *tvp, T *trho, size_t n)
{
    GRID_STRIDE(q, n)
    {
        const T v = vp[q], r = rho[q];
        const T v2 = v * v, v3 = v2 * v, r2 = r * r;
        const T gvp = ((T)(-2) * g0[q]) / (v3 * r);
        const T grho = ((-g0[q]) / (v2 * r2)) - (g1[q] / r);
        tvp[q] = tvp[q] + gvp;
        trho[q] = trho[q] + grho;
    }
}

void post_vd_chain_accumulate(int dtype, size_t n, const void *g0, const void *g1, const void *vp, const void *rho, void *tvp, void *trho, cudaStream_t st)
{
    if (dtype == SWB_F64)
        vd_chain_kernel<double><<<grid1d(n), 256, 0, st>>>((const double *)g0, (const double *)g1, (const double *)vp, (const double *)rho, (double *)tvp, (double *)trho, n);
    else
        vd_chain_kernel<float><<<grid1d(n), 256, 0, st>>>((const float *)g0, (const float *)g1, (const float *)vp, (const float *)rho, (float *)tvp, (float *)trho, n);
    check_launch("vd_chain");
    count_launch();
}

// ---- elastic: staggered material averages (precomp_elaprop!, ela_models.jl:156-173; utils/interpolations.jl:35-43) ----
//   ρ_ihalf = interp(ρ, 1), ρ_jhalf = interp(ρ, 2), μ_ihalf_jhalf = interp(μ, [1, 2]); BSpline(Linear()) at half indices
//   = nested two-point means with Float64 weights 0.5 (first axis innermost); harmonic = 1 ./ itp(1 ./ m)
template <class T>
__device__ __forceinline__ double mean2(T a, T b, int interp)
{
    if (interp == 0)
        return 0.5 * (double)a + 0.5 * (double)b;
    return 0.5 * (double)((T)1 / a) + 0.5 * (double)((T)1 / b); // caller inverts
}
template <class T>
__global__ void ela_props_kernel(const T *rho, const T *mu, int irho, int imu, T *rho_ih, T *rho_jh, T *mu_hh, long long nx, long long nz)
{
    const size_t n = (size_t)nx * nz;
    GRID_STRIDE(q, n)
    {
        const long long i = q % nx, j = q / nx;
        if (i < nx - 1) {
            const double m = mean2<T>(rho[q], rho[q + 1], irho);
            rho_ih[(size_t)j * (nx - 1) + i] = (T)(irho == 0 ? m : 1.0 / m);
        }
        if (j < nz - 1) {
            const double m = mean2<T>(rho[q], rho[q + nx], irho);
            rho_jh[q] = (T)(irho == 0 ? m : 1.0 / m);
        }
        if (i < nx - 1 && j < nz - 1) {
            const double a0 = mean2<T>(mu[q], mu[q + 1], imu), a1 = mean2<T>(mu[q + nx], mu[q + nx + 1], imu);
            const double m = 0.5 * a0 + 0.5 * a1;
            mu_hh[(size_t)j * (nx - 1) + i] = (T)(imu == 0 ? m : 1.0 / m);
        }
    }
}

void post_ela_props(int dtype, const int64_t *n, const void *rho, const void *mu, int interp_rho, int interp_mu, void *rho_ih, void *rho_jh, void *mu_hh,
                    cudaStream_t st)
{
    const size_t nn = (size_t)n[0] * n[1];
    if (dtype == SWB_F64)
        ela_props_kernel<double><<<grid1d(nn), 256, 0, st>>>((const double *)rho, (const double *)mu, interp_rho, interp_mu, (double *)rho_ih, (double *)rho_jh,
                                                            (double *)mu_hh, n[0], n[1]);
    else
        ela_props_kernel<float><<<grid1d(nn), 256, 0, st>>>((const float *)rho, (const float *)mu, interp_rho, interp_mu, (float *)rho_ih, (float *)rho_jh,
                                                           (float *)mu_hh, n[0], n[1]);
    check_launch("ela_props");
    count_launch();
}

// ---- elastic: back_interp of the staggered gradients (ela_gradient.jl:163-165, utils/interpolations.jl:14-28,45-47) ----
//   gradient_ρ = 0 + (back_interp(g_ρ_ihalf, 1) + back_interp(g_ρ_jhalf, 2));  gradient_μ = grad_μ + back_interp(g_μ_ihalf_jhalf, [1, 2])
// 2-D back_interp visits the permutations (0,0), (0,1), (1,0), (1,1) in that order: cell (I,J) receives
// g[I,J], g[I,J-1], g[I-1,J], g[I-1,J-1] (0-based staggered indices), each times ∂f∂m.
template <class T>
__device__ __forceinline__ double harm4_itp(const T *mu, long long nx, long long i, long long j)
{ // harmonic interpolant at staggered point (i, j): 1 / (nested means of 1/μ)
    const size_t q = (size_t)j * nx + i;
    const double a0 = 0.5 * (double)((T)1 / mu[q]) + 0.5 * (double)((T)1 / mu[q + 1]);
    const double a1 = 0.5 * (double)((T)1 / mu[q + nx]) + 0.5 * (double)((T)1 / mu[q + nx + 1]);
    return 1.0 / (0.5 * a0 + 0.5 * a1);
}
template <class T>
__global__ void ela_backinterp_kernel(const T *rho, const T *mu, int irho, int imu, const T *g_ri, const T *g_rj, const T *g_mh, const T *g_mu, T *out_rho,
                                      T *out_mu, long long nx, long long nz)
{
    const size_t n = (size_t)nx * nz;
    GRID_STRIDE(q, n)
    {
        const long long i = q % nx, j = q / nx;
        const T m = rho[q];
        const T mxn = i < nx - 1 ? rho[q + 1] : (T)0, mxp = i > 0 ? rho[q - 1] : (T)0;
        const T mzn = j < nz - 1 ? rho[q + nx] : (T)0, mzp = j > 0 ? rho[q - nx] : (T)0;
        const T bx = backinterp_pair<T>(i < nx - 1 ? g_ri[(size_t)j * (nx - 1) + i] : (T)0, i < nx - 1, i > 0 ? g_ri[(size_t)j * (nx - 1) + i - 1] : (T)0, i > 0,
                                        irho, m, mxn, mxp);
        const T bz = backinterp_pair<T>(j < nz - 1 ? g_rj[q] : (T)0, j < nz - 1, j > 0 ? g_rj[q - nx] : (T)0, j > 0, irho, m, mzn, mzp);
        out_rho[q] = (T)0 + (bx + bz);
        T res = (T)0;
        const long long di[4] = {0, 0, -1, -1}, dj[4] = {0, -1, 0, -1};
        const T mm = mu[q];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const long long si = i + di[k], sj = j + dj[k];
            if (si < 0 || sj < 0 || si > nx - 2 || sj > nz - 2)
                continue;
            const T g = g_mh[(size_t)sj * (nx - 1) + si];
            if (imu == 0)
                res = res + g * (T)0.25;
            else {
                const double itp = harm4_itp<T>(mu, nx, si, sj);
                const double w = ((itp * itp) / (double)(mm * mm)) / 4.0;
                res = (T)((double)res + (double)g * w);
            }
        }
        out_mu[q] = g_mu[q] + res;
    }
}

void post_ela_backinterp(int dtype, const int64_t *n, const void *rho, const void *mu, int interp_rho, int interp_mu, const void *g_ri, const void *g_rj,
                         const void *g_mh, const void *g_mu, void *out_rho, void *out_mu, cudaStream_t st)
{
    const size_t nn = (size_t)n[0] * n[1];
    if (dtype == SWB_F64)
        ela_backinterp_kernel<double><<<grid1d(nn), 256, 0, st>>>((const double *)rho, (const double *)mu, interp_rho, interp_mu, (const double *)g_ri,
                                                                 (const double *)g_rj, (const double *)g_mh, (const double *)g_mu, (double *)out_rho, (double *)out_mu,
                                                                 n[0], n[1]);
    else
        ela_backinterp_kernel<float><<<grid1d(nn), 256, 0, st>>>((const float *)rho, (const float *)mu, interp_rho, interp_mu, (const float *)g_ri, (const float *)g_rj,
                                                                (const float *)g_mh, (const float *)g_mu, (float *)out_rho, (float *)out_mu, n[0], n[1]);
    check_launch("ela_backinterp");
    count_launch();
}

// L2 misfit on the device (L2Misfit.jl:24-95) for an identity or Diagonal inverse covariance and optional windows:
//   r = syn - obs;  r = mask .* r (if windows);  adjsrc = -(w .* r) (= -invcov * r);  misfit += dot(r, w .* r)/2 (accumulated in double)
// w, mask: one entry per time sample (the fastest index of syn); NULL = ones.
template <class T>
__global__ void l2_adjsrc_kernel(const T *syn, const T *obs, const T *w, const T *mask, T *adj, double *acc, size_t n, size_t nt)
{
    double local = 0.0;
    GRID_STRIDE(q, n)
    {
        T r = obs ? syn[q] - obs[q] : syn[q];
        const size_t t = q % nt;
        if (mask)
            r = mask[t] * r;
        const T wr = w ? w[t] * r : r;
        adj[q] = -wr;
        local += (double)r * (double)wr;
    }
    for (int o = 16; o > 0; o >>= 1)
        local += __shfl_down_sync(0xffffffffu, local, o);
    __shared__ double sm[8];
    const int wp = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0)
        sm[wp] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k)
            s += sm[k];
        atomicAdd(acc, 0.5 * s);
    }
}

void post_l2_adjsrc(int dtype, size_t n, size_t nt, const void *syn, const void *obs, const void *w, const void *mask, void *adj, double *acc, cudaStream_t st)
{
    if (n == 0)
        return;
    if (dtype == SWB_F64)
        l2_adjsrc_kernel<double><<<grid1d(n), 256, 0, st>>>((const double *)syn, (const double *)obs, (const double *)w, (const double *)mask, (double *)adj, acc, n, nt);
    else
        l2_adjsrc_kernel<float><<<grid1d(n), 256, 0, st>>>((const float *)syn, (const float *)obs, (const float *)w, (const float *)mask, (float *)adj, acc, n, nt);
    check_launch("l2_adjsrc");
    count_launch();
}

template <class T>
__global__ void axpy_kernel(const T *x, T *y, size_t n)
{
    GRID_STRIDE(q, n) { y[q] = y[q] + x[q]; }
}

void post_axpy(int dtype, size_t n, const void *x, void *y, cudaStream_t st)
{
    if (n == 0)
        return;
    if (dtype == SWB_F64)
        axpy_kernel<double><<<grid1d(n), 256, 0, st>>>((const double *)x, (double *)y, n);
    else
        axpy_kernel<float><<<grid1d(n), 256, 0, st>>>((const float *)x, (float *)y, n);
    check_launch("axpy");
    count_launch();
}

} // namespace swb
