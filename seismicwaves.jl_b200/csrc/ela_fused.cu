// ela_fused.cu -- single-launch 2D elastic isotropic P-SV time step for the per-shot engine (stresses kept on chip).
//
// The reference advances one step with four stencil sweeps that round-trip the stresses through memory
// (src/models/elastic/backends/shared/elastic2D_iso_xPU.jl:1-79,120-239: update_σxx_σzz!, update_σxz!, update_ux!,
// update_uz!; 19 array passes per cell) plus two tiny launches per source and per receiver, and one more full sweep for the
// zero-lag correlations (elastic/backends/shared/correlate_gradient_xPU.jl:47-83).  The stresses are a pure function of
// the current displacements (σ = C : ε(ucur), recomputed from scratch every step), so here a CTA
//   1. has one thread request the whole shared-memory working set of its TX x TZ tile as TMA boxes (cp.async.bulk.tensor,
//      one mbarrier): ux, uz plus a 4-cell halo; λ, μ, μ_ihalf_jhalf plus a 2-cell halo straight into the arrays that will
//      hold σxx, σzz, σxz (each cell's factors are consumed by the thread that then overwrites them with that cell's
//      stresses); for an adjoint launch that also correlates, the forward displacements u[it-1] plus a 2-cell halo.  uold and
//      dt²/ρ of the owned cells go to registers meanwhile,
//   2. computes σxx, σzz, σxz on the tile plus a 2-cell halo -- the halo is recomputed, not exchanged; the C-PML memory
//      variables of the displacement derivatives are read from the `in` copy and written (owned cells only) to the `out`
//      copy -- and applies the moment-tensor injection to its on-chip stresses; when correlating, the λ, μ, μ_ihalf_jhalf
//      gradients are accumulated here, where the adjoint strain components are in registers anyway,
//   3. computes uxnew, uznew of the tile (with the C-PML memory variables of the stress derivatives) and stores them;
//      when correlating, accumulates the two ρ gradients.
// A thread owns 16 bytes of consecutive cells (4 Float32 / 2 Float64) of one row: it reads whole vectors from shared memory
// (x neighbours from the two adjacent vectors of the row, z neighbours from the same vector of the rows above / below) and, in
// interior tiles (no C-PML strip, no grid edge, no free surface -- > 90 % of a production grid), moves its owned cells as one
// 128-bit global access per array.  Tiles that touch a strip, an edge or the free surface run the same body with the
// reference's range tests, mirrored free-surface rows and ∂̃ memory variables applied per cell.
// HBM traffic per cell-update: ux, uz read, uold read / unew written in place (4), λ, μ, μ_ihalf_jhalf, dt²/ρ_ihalf,
// dt²/ρ_jhalf read = 11 values; the correlation adds u[it-2], u[it-1], u[it] (6) and five accumulators read +
// written (10).  The arithmetic is the reference's, operation for operation: the derivative wrappers of
// freesurface_derivatives_4th_mirror.jl:1-244 (zero padding outside every array -- here literal zeros of the padded
// planes --, odd / even mirroring and the Hooke's-law row at the free surface), ∂̃4th of src/utils/fdgenerated.jl:178-195,
// stresses rounded to T before the displacement update reads them.  (SWB_FLAG_FAST_F32 evaluates the 4-point derivative as
// 27/24 (f3 - f2) + 1/24 (f1 - f4) and contracts a*b+c into FMAs; every other mode keeps the reference's operation order.)
// External-force / adjoint-source injection and the receiver sums stay separate small launches (elf_inject_force,
// elf_record); elf_correlate is the stand-alone correlation (last adjoint step, step-by-step path).
#include "ela_fused.h"
#include "tma.cuh"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>
#include <type_traits>

namespace swb {

int elf_tz(int dtype, bool adjoint, long long nx, long long nz)
{
    // candidates, tallest first: the tallest tile that keeps two CTAs per SM is fastest on large grids (measured at 4096 x 2048,
    // profiles/: Float32 forward 24 rows 75 us vs 16 rows 81 us, adjoint + correlation 16 rows 171 us vs 24 rows 195 us; Float64
    // forward 12 rows 137 us vs 8 rows 153 us); a small grid takes the tallest height that still yields three CTAs per SM, because a
    // step of a grid that is one partial wave of CTAs is the latency of one tile (load -> stresses -> displacements -> store)
    static const int env32 = [] {
        const char *e = std::getenv("SWB_ELF_TZ");
        const int v = e ? std::atoi(e) : 0;
        return (v == 8 || v == 16 || v == 24) ? v : 0;
    }();
    static const int env64 = [] {
        const char *e = std::getenv("SWB_ELF_TZ64");
        const int v = e ? std::atoi(e) : 0;
        return (v == 4 || v == 8 || v == 12) ? v : 0;
    }();
    if (dtype == SWB_F64 ? env64 : env32)
        return dtype == SWB_F64 ? env64 : env32;
    const int c32f[3] = {24, 16, 8}, c32a[3] = {16, 8, 8}, c64f[3] = {12, 8, 4}, c64a[3] = {8, 4, 4};
    const int *c = dtype == SWB_F64 ? (adjoint ? c64a : c64f) : (adjoint ? c32a : c32f);
    const long long ntx = (nx + ELF_TX - 1) / ELF_TX;
    for (int k = 0; k < 3; ++k)
        if (k == 2 || ntx * ((nz + c[k] - 1) / c[k]) >= 3 * 148) // measured (Float32 forward): 1024^2 23.1 / 20.3 / 21.4 us at 24 / 16 / 8 rows; 640^2 19.3 / 18.5 / 17.2
            return c[k];
    return c[2];
}

// 2D tensor map of a whole padded plane (all guard rows included), box = ELF_SW columns x box_rows rows
void elf_make_tmap(CUtensorMap *out, int dtype, const void *plane_base, long long ld, long long rows, int box_rows)
{
    make_tmap_2d(out, dtype, plane_base, ld, rows, ELF_SW, box_rows);
}

namespace {

constexpr int TX = ELF_TX, NTHR = 256;
constexpr int W = ELF_SW; // row pitch of every staged array: columns -4 .. TX+3

// 16-byte vectors of consecutive cells
__device__ __forceinline__ void ldv(const float *p, float (&o)[4])
{
    const float4 v = *reinterpret_cast<const float4 *>(p);
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
}
__device__ __forceinline__ void ldv(const double *p, double (&o)[2])
{
    const double2 v = *reinterpret_cast<const double2 *>(p);
    o[0] = v.x, o[1] = v.y;
}
__device__ __forceinline__ void ldv_ro(const float *p, float (&o)[4])
{
    const float4 v = __ldg(reinterpret_cast<const float4 *>(p));
    o[0] = v.x, o[1] = v.y, o[2] = v.z, o[3] = v.w;
}
__device__ __forceinline__ void ldv_ro(const double *p, double (&o)[2])
{
    const double2 v = __ldg(reinterpret_cast<const double2 *>(p));
    o[0] = v.x, o[1] = v.y;
}
__device__ __forceinline__ void stv(float *p, const float (&o)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[3]); }
__device__ __forceinline__ void stv(double *p, const double (&o)[2]) { *reinterpret_cast<double2 *>(p) = make_double2(o[0], o[1]); }

// ∂x4th_inner / ∂y4th_inner (freesurface_derivatives_4th_mirror.jl:2-7)
template <class T, class CT>
__device__ __forceinline__ CT inner4(T f1, T f2, T f3, T f4, T inv)
{
    const CT c1 = (CT)(1.0 / 24.0), c2 = (CT)(27.0 / 24.0);
    return (((c1 * (CT)f1 - c2 * (CT)f2) + c2 * (CT)f3) - c1 * (CT)f4) * (CT)inv;
}
template <>
__device__ __forceinline__ float inner4<float, float>(float f1, float f2, float f3, float f4, float inv) // SWB_FLAG_FAST_F32
{
    return __fmaf_rn(1.125f * inv, f3 - f2, ((1.0f / 24.0f) * inv) * (f1 - f4));
}
// a * b + c: two roundings like the reference (the library is compiled with -fmad=false), one FMA under SWB_FLAG_FAST_F32
template <class CT>
__device__ __forceinline__ CT mad2(CT a, CT b, CT c, bool)
{
    return a * b + c;
}
__device__ __forceinline__ float mad2(float a, float b, float c, int) { return __fmaf_rn(a, b, c); }
template <class T, class CT>
struct FastSel {
    typedef bool type;
};
template <>
struct FastSel<float, float> {
    typedef int type;
};
#define MAD(a, b, c) mad2((CT)(a), (CT)(b), (CT)(c), typename FastSel<T, CT>::type(0))

// ∂̃4th (fdgenerated.jl:178-195) with separate in / out memory-variable arrays; `own`: this CTA owns the cell and stores psi
template <class T, class CT>
__device__ __forceinline__ CT cpml4(CT D, int I, int ndim, int halo, bool half, const T *__restrict__ a, const T *__restrict__ b, const T *__restrict__ psi_in,
                                    T *__restrict__ psi_out, long long base, long long stride, bool own)
{
    const int p1 = half ? 1 : 0;
    const int idim = I + p1;
    int k;
    if (idim <= halo + p1)
        k = idim;
    else if (idim >= ndim - halo)
        k = I - (ndim - halo) + 1 + (halo + p1);
    else
        return D;
    const long long q = base + (long long)(k - 1) * stride;
    T pn;
    const CT r = cpml_apply<T, CT>(D, a[k - 1], b[k - 1], psi_in[q], pn);
    if (own)
        psi_out[q] = pn;
    return r;
}

// ∂̃4th for a whole vector of cells, split into an index step, a fetch step and a finish step so that the loads of every cell and
// of both derivative kinds of a group are in flight together (the per-cell form above serialises one dependent global load per
// cell and kind).  kk = 1-based index into the strip's memory variable / coefficient profiles, 0 = outside the strips (∂̃ = ∂).
__device__ __forceinline__ int cpml_k(int I, int ndim, int halo, int p1)
{
    const int idim = I + p1;
    if (idim <= halo + p1)
        return idim;
    if (idim >= ndim - halo)
        return I - (ndim - halo) + 1 + (halo + p1);
    return 0;
}
template <class T, int V>
struct CpmlVec {
    T a[V], b[V], psi[V];
};
template <class T, int V>
__device__ __forceinline__ void cpml_fetch(CpmlVec<T, V> &c, const int (&kk)[V], const int (&off)[V], const T *__restrict__ a, const T *__restrict__ b,
                                           const T *__restrict__ psi_in)
{
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const bool in = kk[k] > 0;
        c.a[k] = in ? a[kk[k] - 1] : (T)0;
        c.b[k] = in ? b[kk[k] - 1] : (T)0;
        c.psi[k] = in ? psi_in[off[k]] : (T)0;
    }
}
template <class T, class CT, int V>
__device__ __forceinline__ void cpml_finish(CT (&D)[V], const CpmlVec<T, V> &c, const int (&kk)[V], const int (&off)[V], const bool (&st)[V], T *__restrict__ psi_out)
{
#pragma unroll
    for (int k = 0; k < V; ++k)
        if (kk[k] > 0) {
            T pn;
            D[k] = cpml_apply<T, CT>(D[k], c.a[k], c.b[k], c.psi[k], pn);
            if (st[k])
                psi_out[off[k]] = pn;
        }
}

template <class T, int TZ, bool ADJ>
struct ElaSmem {
    static constexpr int UH = TZ + 8, SH = TZ + 4;
    T ux[UH * W], uz[UH * W];                     // displacements, rows -4 .. TZ+3, columns -4 .. TX+3 (TMA destinations: 128-byte aligned)
    T sxx[SH * W], szz[SH * W], sxz[SH * W];      // λ, μ, μ_ihalf_jhalf, then the stresses: rows -2 .. TZ+1
    T fx[ADJ ? SH * W : 8], fz[ADJ ? SH * W : 8]; // forward u[it-1] for the correlation, rows -2 .. TZ+1
    unsigned long long bar, bar2; // mbarriers of the two row halves of the staged working set
    static_assert(TZ % 4 == 0, "rows x 544 bytes must be a multiple of 128");
};

// moment-tensor injection into the on-chip stresses (inject_momten_sources2D_σxx_σzz! / _σxz! :81-94)
template <class T, int TZ, bool ADJ>
__device__ __forceinline__ void inject_mt(const ElaFusedParams<T> &P, ElaSmem<T, TZ, ADJ> &S, int tile, int tid)
{
    if (P.mt_it <= 0)
        return;
    int e = P.mt_off[tile];
    const int e1 = P.mt_off[tile + 1];
    while (e < e1) { // sources in index order (they may share cells); the points of one source in parallel (distinct cells)
        const int s = P.mt_src[e];
        int en = e + 1;
        while (en < e1 && P.mt_src[en] == s)
            ++en;
        const T w = P.srctf[(long long)s * P.nt + (P.mt_it - 1)];
        for (int k = e + tid; k < en; k += NTHR) {
            const int cell = P.mt_cell[k];
            const T cf = P.mt_coef[k];
            if (cell < ELF_MT_FIELD) {
                S.sxx[cell] = S.sxx[cell] + (P.Mxx[s] * cf) * w;
                S.szz[cell] = S.szz[cell] + (P.Mzz[s] * cf) * w;
            } else
                S.sxz[cell - ELF_MT_FIELD] = S.sxz[cell - ELF_MT_FIELD] + (P.Mxz[s] * cf) * w;
        }
        __syncthreads();
        e = en;
    }
}

// ---- the tile body -------------------------------------------------------------------------------------------------------------
// EDGE = false: every cell of the tile's stress region lies inside all update ranges, outside every C-PML strip and below the
// free-surface rows, so the reference's expressions reduce to the plain 4-point derivatives.  EDGE = true: the reference's
// range tests, free-surface rows and ∂̃ memory variables per cell.  The two outermost vectors of a stress row (columns
// -4 .. -1 and TX .. TX+3) hold two needed and two junk cells each; junk is computed from in-bounds shared memory and never read.
// everything a vector task of the tile body needs
template <class T, int TZ, bool ADJ>
struct TileCtx {
    const ElaFusedParams<T> &P;
    ElaSmem<T, TZ, ADJ> &S;
    int x0, z0, nx, nz, h, j0;
    long long ld;
    bool ft, xs, zs;
    T idx_, idz_;
};

// phase 2, one vector of the stress region (row rr, vector column vv).  SP = false: the interior expressions; SP = true: range
// tests, free-surface rows and ∂̃ per cell
template <class T, class CT, int TZ, bool ADJ, bool SP>
__device__ __forceinline__ void stress_task(const TileCtx<T, TZ, ADJ> &C, const int rr, const int vv)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int NV = W / V, NVT = TX / V;
    const ElaFusedParams<T> &P = C.P;
    ElaSmem<T, TZ, ADJ> &S = C.S;
    const int x0 = C.x0, z0 = C.z0, nx = C.nx, nz = C.nz, h = C.h, j0 = C.j0;
    const long long ld = C.ld;
    const bool ft = C.ft, xs = C.xs, zs = C.zs;
    const T idx_ = C.idx_, idz_ = C.idz_;
    (void)NV, (void)NVT, (void)j0, (void)ld, (void)xs, (void)zs, (void)nx, (void)nz, (void)h, (void)ft;
    const int r = rr - 2, c0 = vv * V - 4; // first cell of the vector
    const int ou = (rr + 2) * W + vv * V;  // ... in ux / uz
    const int os = rr * W + vv * V;        // ... in the stress / factor / forward-field arrays
    const int J = z0 + r + 1;              // 1-based reference row
    {   // rows whose inputs (ux, uz rows rr .. rr+4, factors row rr, forward fields rows rr-2 .. rr+2) reach into the second half
        constexpr int UH1 = elf_half_rows(TZ + 8, sizeof(T)), SH1 = elf_half_rows(TZ + 4, sizeof(T));
        constexpr int RR1 = (UH1 - 4) < (ADJ ? SH1 - 2 : SH1) ? (UH1 - 4) : (ADJ ? SH1 - 2 : SH1);
        if (rr >= RR1)
            mbar_wait(&S.bar2, 0);
    }
    if (SP && (J < 1 || J > nz - 1)) {     // outside every update range: zero stresses
        T zero_[V];
#pragma unroll
        for (int k = 0; k < V; ++k)
            zero_[k] = (T)0;
        stv(S.sxx + os, zero_);
        stv(S.szz + os, zero_);
        stv(S.sxz + os, zero_);
        return;
    }
    T X[3][V], Z[3][V], xm1[V], xp1[V], xp2[V], zm2[V], zm1[V], zp1[V], l[V], m[V], mh[V];
    ldv(S.ux + ou - V, X[0]);
    ldv(S.ux + ou, X[1]);
    ldv(S.ux + ou + V, X[2]);
    ldv(S.uz + ou - V, Z[0]);
    ldv(S.uz + ou, Z[1]);
    ldv(S.uz + ou + V, Z[2]);
    ldv(S.ux + ou - W, xm1);
    ldv(S.ux + ou + W, xp1);
    ldv(S.ux + ou + 2 * W, xp2);
    ldv(S.uz + ou - 2 * W, zm2);
    ldv(S.uz + ou - W, zm1);
    ldv(S.uz + ou + W, zp1);
    ldv(S.sxx + os, l);
    ldv(S.szz + os, m);
    ldv(S.sxz + os, mh);
    const T *Xf = &X[0][0], *Zf = &Z[0][0]; // columns c0 - V .. c0 + 2 V - 1 of row r
    T oxx[V], ozz[V], oxz[V];
    CT dudx[V], dwdz[V], dwdx[V], dudz[V]; // before ∂̃: the adjoint strains of the correlation
    bool v1[V], v2[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = c0 + k, I = x0 + c + 1;
        v1[k] = v2[k] = true;
        if (SP) {
            const bool inreg = c >= -2 && c <= TX + 1;
            v1[k] = inreg && I >= 2 && I <= nx - 1 && J >= j0 && J <= nz - 1;
            v2[k] = inreg && I >= 1 && I <= nx - 1 && J >= 1 && J <= nz - 1;
        }
        dudx[k] = inner4<T, CT>(Xf[V + k - 2], Xf[V + k - 1], Xf[V + k], Xf[V + k + 1], idx_);
        if (SP && ft && J == 1) { // Hooke's law on the free-surface row (:179-195)
            const T fac = -l[k] / (l[k] + (T)2 * m[k]);
            dwdz[k] = (CT)fac * dudx[k];
        } else
            dwdz[k] = inner4<T, CT>((SP && ft && J == 2) ? zm1[k] : zm2[k], zm1[k], Zf[V + k], zp1[k], idz_);
        dwdx[k] = inner4<T, CT>(Zf[V + k - 1], Zf[V + k], Zf[V + k + 1], Zf[V + k + 2], idx_);
        // (even mirror of ux at the free surface, :214-222)
        dudz[k] = inner4<T, CT>((SP && ft && J == 1) ? xp1[k] : xm1[k], Xf[V + k], xp1[k], xp2[k], idz_);
    }
    CT dudx_c[V], dwdz_c[V], dwdx_c[V], dudz_c[V];
#pragma unroll
    for (int k = 0; k < V; ++k)
        dudx_c[k] = dudx[k], dwdz_c[k] = dwdz[k], dwdx_c[k] = dwdx[k], dudz_c[k] = dudz[k];
    if (SP) { // ∂̃ by direction (a tile usually reaches the strips of one direction only): x: ψ_∂ux∂x (4) for σxx, σzz and ψ_∂uz∂x (5)
              // for σxz; z: ψ_∂uz∂z (7) and ψ_∂ux∂z (6).  Both kinds of a direction are fetched before either is used.
        int ka[V], kb[V], oa[V], ob[V];
        bool st[V];
        CpmlVec<T, V> ca, cb;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const int c = c0 + k;
            st[k] = r >= 0 && r < TZ && c >= 0 && c < TX;
        }
        if (xs) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const int I = x0 + c0 + k + 1;
                ka[k] = v1[k] ? cpml_k(I - 1, nx - 1, h, 1) : 0;
                oa[k] = (J - 1) * (2 * (h + 1)) + (ka[k] - 1);
                kb[k] = v2[k] ? cpml_k(I, nx, h, 0) : 0;
                ob[k] = (J - 1) * (2 * h) + (kb[k] - 1);
            }
            cpml_fetch<T, V>(ca, ka, oa, P.a_x, P.b_x, P.psi_in[4]);
            cpml_fetch<T, V>(cb, kb, ob, P.a_xh, P.b_xh, P.psi_in[5]);
            cpml_finish<T, CT, V>(dudx_c, ca, ka, oa, st, P.psi_out[4]);
            cpml_finish<T, CT, V>(dwdx_c, cb, kb, ob, st, P.psi_out[5]);
        }
        const int kz7 = zs ? cpml_k(J - 1, nz - 1, h, 1) : 0, kz6 = zs ? cpml_k(J, nz, h, 0) : 0;
        if (kz7 > 0 || kz6 > 0) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const int I = x0 + c0 + k + 1;
                ka[k] = v1[k] ? kz7 : 0;
                oa[k] = (I - 1) + (kz7 - 1) * nx;
                kb[k] = v2[k] ? kz6 : 0;
                ob[k] = (I - 1) + (kz6 - 1) * (nx - 1);
            }
            cpml_fetch<T, V>(ca, ka, oa, P.a_z, P.b_z, P.psi_in[7]);
            cpml_fetch<T, V>(cb, kb, ob, P.a_zh, P.b_zh, P.psi_in[6]);
            cpml_finish<T, CT, V>(dwdz_c, ca, ka, oa, st, P.psi_out[7]);
            cpml_finish<T, CT, V>(dudz_c, cb, kb, ob, st, P.psi_out[6]);
        }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const T l2m = l[k] + (T)2 * m[k];
        oxx[k] = (T)MAD(l2m, dudx_c[k], (CT)l[k] * dwdz_c[k]);
        ozz[k] = (SP && J == 1) ? (T)0 : (T)MAD(l[k], dudx_c[k], (CT)l2m * dwdz_c[k]);
        oxz[k] = (T)((CT)mh[k] * (dwdx_c[k] + dudz_c[k]));
        if (SP && !v1[k])
            oxx[k] = ozz[k] = (T)0;
        if (SP && !v2[k])
            oxz[k] = (T)0;
    }
    stv(S.sxx + os, oxx);
    stv(S.szz + os, ozz);
    stv(S.sxz + os, oxz);
    if (ADJ) { // grad_λ, grad_μ, grad_μ_ihalf_jhalf of the owned cells (correlate_gradient_xPU.jl:62-82)
        if (r >= 0 && r < TZ && c0 >= 0 && c0 < TX) {
            const long long q = (long long)(z0 + r) * ld + (x0 + c0);
            T gl[V], gm[V], gh[V], F[3][V], G[3][V], fm1[V], fp1[V], fp2[V], gm2[V], gm1[V], gp1[V];
            if (!SP) {
                ldv(P.g_l + q, gl);
                ldv(P.g_m + q, gm);
                ldv(P.g_mh + q, gh);
            }
            ldv(S.fx + os - V, F[0]);
            ldv(S.fx + os, F[1]);
            ldv(S.fx + os + V, F[2]);
            ldv(S.fz + os - V, G[0]);
            ldv(S.fz + os, G[1]);
            ldv(S.fz + os + V, G[2]);
            ldv(S.fx + os - W, fm1);
            ldv(S.fx + os + W, fp1);
            ldv(S.fx + os + 2 * W, fp2);
            ldv(S.fz + os - 2 * W, gm2);
            ldv(S.fz + os - W, gm1);
            ldv(S.fz + os + W, gp1);
            const T *Ff = &F[0][0], *Gf = &G[0][0];
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const int I = x0 + c0 + k + 1;
                const bool surf = SP && ft && J == 1;
                const bool w1 = !SP || (I >= 2 && I <= nx - 1 && J >= j0 && J <= nz - 1);
                const bool w2 = !SP || (I <= nx - 1 && J <= nz - 1);
                if (w1) {
                    if (SP)
                        gl[k] = P.g_l[q + k], gm[k] = P.g_m[q + k];
                    const CT exx = inner4<T, CT>(Ff[V + k - 2], Ff[V + k - 1], Ff[V + k], Ff[V + k + 1], idx_);
                    CT ezz;
                    if (surf) {
                        const T fac = -l[k] / (l[k] + (T)2 * m[k]);
                        ezz = (CT)fac * exx;
                    } else
                        ezz = inner4<T, CT>((SP && ft && J == 2) ? gm1[k] : gm2[k], gm1[k], Gf[V + k], gp1[k], idz_);
                    const CT exx_a = dudx[k], ezz_a = dwdz[k];
                    const CT div_u = exx + ezz, div_a = exx_a + ezz_a;
                    if (surf) {
                        gl[k] = (T)((CT)gl[k] + (div_u * div_a) / (CT)2);
                        gm[k] = (T)((CT)gm[k] + (exx * exx_a + ezz * ezz_a));
                    } else {
                        gl[k] = (T)((CT)gl[k] + div_u * div_a);
                        gm[k] = (T)((CT)gm[k] + (CT)2 * (exx * exx_a + ezz * ezz_a));
                    }
                    if (SP)
                        P.g_l[q + k] = gl[k], P.g_m[q + k] = gm[k];
                }
                if (w2) {
                    if (SP)
                        gh[k] = P.g_mh[q + k];
                    const CT fdwdx = inner4<T, CT>(Gf[V + k - 1], Gf[V + k], Gf[V + k + 1], Gf[V + k + 2], idx_);
                    const CT fdudz = inner4<T, CT>(surf ? fp1[k] : fm1[k], Ff[V + k], fp1[k], fp2[k], idz_);
                    const CT exz = (fdwdx + fdudz) / (CT)2, exz_a = (dwdx[k] + dudz[k]) / (CT)2;
                    gh[k] = (T)((CT)gh[k] + (CT)2 * (exz * exz_a + exz * exz_a));
                    if (SP)
                        P.g_mh[q + k] = gh[k];
                }
            }
            if (!SP) {
                stv(P.g_l + q, gl);
                stv(P.g_m + q, gm);
                stv(P.g_mh + q, gh);
            }
        }
    }
}

// phase 3, one vector of the tile (row r, vector column v)
template <class T, class CT, int TZ, bool ADJ, bool SP>
__device__ __forceinline__ void disp_task(const TileCtx<T, TZ, ADJ> &C, const int r, const int v, const T (&uxo_)[16 / sizeof(T)], const T (&uzo_)[16 / sizeof(T)],
                                          const T (&fi_)[16 / sizeof(T)], const T (&fj_)[16 / sizeof(T)])
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int NV = W / V, NVT = TX / V;
    const ElaFusedParams<T> &P = C.P;
    ElaSmem<T, TZ, ADJ> &S = C.S;
    const int x0 = C.x0, z0 = C.z0, nx = C.nx, nz = C.nz, h = C.h, j0 = C.j0;
    const long long ld = C.ld;
    const bool ft = C.ft, xs = C.xs, zs = C.zs;
    const T idx_ = C.idx_, idz_ = C.idz_;
    (void)NV, (void)NVT, (void)j0, (void)ld, (void)xs, (void)zs, (void)nx, (void)nz, (void)h, (void)ft;
    const int c0 = v * V;
    const int J = z0 + r + 1;
    if (SP && J > nz)
        return; // below the grid
    const long long q = (long long)(z0 + r) * ld + (x0 + c0);
    const int os = (r + 2) * W + c0 + 4;
    T A[3][V], B[3][V], bm2[V], bm1[V], bp1[V], zm1[V], z0v[V], zp1[V], zp2[V], ucx[V], ucz[V];
    ldv(S.sxx + os - V, A[0]);
    ldv(S.sxx + os, A[1]);
    ldv(S.sxx + os + V, A[2]);
    ldv(S.sxz + os - V, B[0]);
    ldv(S.sxz + os, B[1]);
    ldv(S.sxz + os + V, B[2]);
    ldv(S.sxz + os - 2 * W, bm2);
    ldv(S.sxz + os - W, bm1);
    ldv(S.sxz + os + W, bp1);
    ldv(S.szz + os - W, zm1);
    ldv(S.szz + os, z0v);
    ldv(S.szz + os + W, zp1);
    ldv(S.szz + os + 2 * W, zp2);
    ldv(S.ux + os + 2 * W, ucx);
    ldv(S.uz + os + 2 * W, ucz);
    const T *Af = &A[0][0], *Bf = &B[0][0];
    T nx_[V], nz_[V];
    bool vx[V], vz[V];
    CT a1[V], a2[V], b1[V], b2[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int I = x0 + c0 + k + 1;
        vx[k] = !SP || (I <= nx - 1 && J <= nz);
        vz[k] = !SP || (I <= nx && J <= nz - 1);
        a1[k] = inner4<T, CT>(Af[V + k - 1], Af[V + k], Af[V + k + 1], Af[V + k + 2], idx_);
        if (SP && ft && J == 1) // odd mirror of σxz at the free surface (:125-140)
            a2[k] = inner4<T, CT>(-bp1[k], -Bf[V + k], Bf[V + k], bp1[k], idz_);
        else if (SP && ft && J == 2)
            a2[k] = inner4<T, CT>(-bm1[k], bm1[k], Bf[V + k], bp1[k], idz_);
        else
            a2[k] = inner4<T, CT>(bm2[k], bm1[k], Bf[V + k], bp1[k], idz_);
        b1[k] = inner4<T, CT>(Bf[V + k - 2], Bf[V + k - 1], Bf[V + k], Bf[V + k + 1], idx_);
        if (SP && ft && J == 1) // odd mirror of σzz at the free surface (:85-93)
            b2[k] = inner4<T, CT>(-zp1[k], z0v[k], zp1[k], zp2[k], idz_);
        else
            b2[k] = inner4<T, CT>(zm1[k], z0v[k], zp1[k], zp2[k], idz_);
    }
    if (SP) { // ∂̃ by direction: x: ψ_∂σxx∂x (0) for ux, ψ_∂σxz∂x (1) for uz; z: ψ_∂σxz∂z (3) for ux, ψ_∂σzz∂z (2) for uz
        int ka[V], kb[V], oa[V], ob[V];
        CpmlVec<T, V> ca, cb;
        if (xs) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const int I = x0 + c0 + k + 1;
                ka[k] = vx[k] ? cpml_k(I, nx, h, 0) : 0;
                oa[k] = (J - 1) * (2 * h) + (ka[k] - 1);
                kb[k] = vz[k] ? cpml_k(I - 1, nx - 1, h, 1) : 0;
                ob[k] = (J - 1) * (2 * (h + 1)) + (kb[k] - 1);
            }
            cpml_fetch<T, V>(ca, ka, oa, P.a_xh, P.b_xh, P.psi_in[0]);
            cpml_fetch<T, V>(cb, kb, ob, P.a_x, P.b_x, P.psi_in[1]);
            cpml_finish<T, CT, V>(a1, ca, ka, oa, vx, P.psi_out[0]);
            cpml_finish<T, CT, V>(b1, cb, kb, ob, vz, P.psi_out[1]);
        }
        const int kz3 = zs ? cpml_k(J - 1, nz - 1, h, 1) : 0, kz2 = zs ? cpml_k(J, nz, h, 0) : 0;
        if (kz3 > 0 || kz2 > 0) {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                const int I = x0 + c0 + k + 1;
                ka[k] = vx[k] ? kz3 : 0;
                oa[k] = (I - 1) + (kz3 - 1) * (nx - 1);
                kb[k] = vz[k] ? kz2 : 0;
                ob[k] = (I - 1) + (kz2 - 1) * nx;
            }
            cpml_fetch<T, V>(ca, ka, oa, P.a_z, P.b_z, P.psi_in[3]);
            cpml_fetch<T, V>(cb, kb, ob, P.a_zh, P.b_zh, P.psi_in[2]);
            cpml_finish<T, CT, V>(a2, ca, ka, oa, vx, P.psi_out[3]);
            cpml_finish<T, CT, V>(b2, cb, kb, ob, vz, P.psi_out[2]);
        }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const T tx = fma((T)2, ucx[k], -uxo_[k]); // = 2 ucur - uold rounded once (2 ucur is exact)
        nx_[k] = (T)MAD(fi_[k], a1[k] + a2[k], tx);
        const T tz = fma((T)2, ucz[k], -uzo_[k]);
        nz_[k] = (T)MAD(fj_[k], b1[k] + b2[k], tz);
    }
    if (!SP) {
        stv(P.uxn + q, nx_);
        stv(P.uzn + q, nz_);
    } else {
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (vx[k])
                P.uxn[q + k] = nx_[k];
            if (vz[k])
                P.uzn[q + k] = nz_[k];
        }
    }
    if (ADJ) { // grad_ρ_ihalf, grad_ρ_jhalf (correlate_gradient_xPU.jl:50-60), all in T
        T fxo[V], fzo[V], fxn[V], fzn[V], gi[V], gj[V], fcx[V], fcz[V];
        ldv_ro(P.fxo + q, fxo);
        ldv_ro(P.fzo + q, fzo);
        ldv_ro(P.fxn + q, fxn);
        ldv_ro(P.fzn + q, fzn);
        ldv(P.g_ri + q, gi);
        ldv(P.g_rj + q, gj);
        ldv(S.fx + os, fcx);
        ldv(S.fz + os, fcz);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            const T vi = (ucx[k] * ((fxo[k] - (T)2 * fcx[k]) + fxn[k])) * P.inv_dt2;
            gi[k] = gi[k] + ((SP && ft && J == 1) ? vi / (T)2 : vi);
            gj[k] = gj[k] + (ucz[k] * ((fzo[k] - (T)2 * fcz[k]) + fzn[k])) * P.inv_dt2;
        }
        if (!SP) {
            stv(P.g_ri + q, gi);
            stv(P.g_rj + q, gj);
        } else {
#pragma unroll
            for (int k = 0; k < V; ++k) {
                if (vx[k])
                    P.g_ri[q + k] = gi[k];
                if (vz[k])
                    P.g_rj[q + k] = gj[k];
            }
        }
    }
}

// Tile rows are issued bottom strip rows first, then from the top: the rows that touch the bottom C-PML strip run the slow
// per-cell body and would otherwise form a tail of long CTAs at the end of the grid.
template <int TZ>
__device__ __forceinline__ int tile_row(int halo, int nty, int brow, int rev)
{
    const int nedge = min(nty, (halo + TZ + 3 + TZ - 1) / TZ);
    if (brow < nedge)
        return nty - nedge + brow;
    return rev ? nty - 1 - brow : brow - nedge; // serpentine: every other launch walks the remaining rows upwards
}

template <class T, class CT, int TZ, bool ADJ, bool EDGE>
__device__ __forceinline__ void ela_tile(const ElaFusedParams<T> &P, ElaSmem<T, TZ, ADJ> &S, const int bx, const int trow)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int NV = W / V;   // vectors per staged row
    constexpr int NVT = TX / V; // vectors per tile row
    constexpr int SH = TZ + 4, UH = TZ + 8;
    constexpr int NP3 = TZ * NVT / NTHR; // vectors per thread in the displacement update
    static_assert(TZ * NVT % NTHR == 0, "tile rows must split evenly over the CTA");
    const int tid = threadIdx.x;
    const int x0 = bx * TX, z0 = trow * TZ;
    const int nx = P.nx, nz = P.nz, h = P.halo;
    const long long ld = P.ld;
    const bool ft = P.freetop != 0;
    const int j0 = ft ? 1 : 2;
    const int tile = trow * P.ntx + bx;
    const T idx_ = P.inv_dx, idz_ = P.inv_dz;
    // does the tile's stress region reach an x / a z strip (or edge)?  ∂̃ is the identity elsewhere (CTA-uniform tests)
    const int mm = max(h + 1, 2);
    const bool xs = EDGE && !(x0 - 1 > mm && x0 + TX + 2 < nx - 1 - h);
    const bool zs = EDGE && !((P.top_inactive || z0 - 1 > mm) && z0 + TZ + 2 < nz - 1 - h);

    // ---- phase 1: one thread requests the shared-memory working set (TMA), all threads their owned uold and dt²/ρ ----------
    //      The working set arrives in two row halves on two mbarriers: the stress rows that need only the first half start while the
    //      second half is still in flight.
    if (tid == 0) {
        constexpr int UH1 = elf_half_rows(UH, sizeof(T)), SH1 = elf_half_rows(SH, sizeof(T));
        static_assert(UH - UH1 == UH1, "the halves of ux / uz share one descriptor");
        mbar_init(&S.bar, 1);
        mbar_init(&S.bar2, 1);
        mbar_expect_tx(&S.bar, (unsigned)((2 * UH1 + (ADJ ? 5 : 3) * SH1) * W * sizeof(T)));
        mbar_expect_tx(&S.bar2, (unsigned)((2 * (UH - UH1) + (ADJ ? 5 : 3) * (SH - SH1)) * W * sizeof(T)));
        const int yu = ELF_GB + z0 - 4, ys = ELF_GB + z0 - 2;
        tma_load_2d(S.ux, &P.tm[0], x0 - 4, yu, &S.bar);
        tma_load_2d(S.uz, &P.tm[1], x0 - 4, yu, &S.bar);
        tma_load_2d(S.sxx, &P.tm[2], x0 - 4, ys, &S.bar);
        tma_load_2d(S.szz, &P.tm[3], x0 - 4, ys, &S.bar);
        tma_load_2d(S.sxz, &P.tm[4], x0 - 4, ys, &S.bar);
        if constexpr (ADJ) {
            tma_load_2d(S.fx, &P.tm[5], x0 - 4, ys, &S.bar);
            tma_load_2d(S.fz, &P.tm[6], x0 - 4, ys, &S.bar);
        }
        tma_load_2d(S.ux + UH1 * W, &P.tm[0], x0 - 4, yu + UH1, &S.bar2);
        tma_load_2d(S.uz + UH1 * W, &P.tm[1], x0 - 4, yu + UH1, &S.bar2);
        tma_load_2d(S.sxx + SH1 * W, &P.tm[7], x0 - 4, ys + SH1, &S.bar2);
        tma_load_2d(S.szz + SH1 * W, &P.tm[8], x0 - 4, ys + SH1, &S.bar2);
        tma_load_2d(S.sxz + SH1 * W, &P.tm[9], x0 - 4, ys + SH1, &S.bar2);
        if constexpr (ADJ) {
            tma_load_2d(S.fx + SH1 * W, &P.tm[10], x0 - 4, ys + SH1, &S.bar2);
            tma_load_2d(S.fz + SH1 * W, &P.tm[11], x0 - 4, ys + SH1, &S.bar2);
        }
    }
    T r_uxo[NP3][V], r_uzo[NP3][V], r_fi[NP3][V], r_fj[NP3][V];
#pragma unroll
    for (int n = 0; n < NP3; ++n) { // all inside the padded plane (junk beyond the arrays is masked at the point of use)
        const int t = tid + NTHR * n;
        const int r = t / NVT, v = t - r * NVT;
        const long long q = (long long)(z0 + r) * ld + (x0 + v * V);
        ldv(P.uxo + q, r_uxo[n]);
        ldv(P.uzo + q, r_uzo[n]);
        ldv_ro(P.fac_ih + q, r_fi[n]);
        ldv_ro(P.fac_jh + q, r_fj[n]);
    }
    __syncthreads(); // (the mbarrier's initialisation becomes visible to the waiting threads)
    mbar_wait(&S.bar, 0);

    // Edge tiles split their vectors into plain ones (the interior expressions are exact: no strip, no edge, no free-surface row
    // within the vector's needed cells) and special ones, enumerated compactly in a second pass, so that a C-PML column strip of
    // a few vectors does not drag whole warps through the per-cell code.  plain cells: I in [Ilo, Ihi], J in [Jlo, Jhi].
    const int Ilo = max(h + 2, 2), Ihi = nx - h - 2, Jlo = P.top_inactive ? 3 : max(h + 2, 3), Jhi = nz - h - 2;
    const TileCtx<T, TZ, ADJ> C{P, S, x0, z0, nx, nz, h, j0, ld, ft, xs, zs, idx_, idz_};

    // ---- phase 2: stresses on the tile + 2-cell halo (update_σxx_σzz! :39-60, update_σxz! :62-79) ------------------------
    if (!EDGE) {
        for (int t = tid; t < SH * NV; t += NTHR) {
            const int rr = t / NV;
            stress_task<T, CT, TZ, ADJ, false>(C, rr, t - rr * NV);
        }
    } else {
        // special rows [0, za) and [zb, SH), special vector columns [0, xa) and [xb, NV) of the stress region
        const int za = min(max(Jlo - (z0 - 1), 0), SH), zb = min(max(Jhi - (z0 - 1) + 1, za), SH);
        int xa = 0, xb = NV;
        for (int vv = 0; vv < NV; ++vv) {
            const int lo_c = max(V * vv - 4, -2), hi_c = min(V * vv - 4 + V - 1, TX + 1);
            if (x0 + lo_c + 1 < Ilo)
                xa = vv + 1;
            if (x0 + hi_c + 1 > Ihi && xb == NV)
                xb = vv;
        }
        xb = max(xb, xa);
        for (int t = tid; t < SH * NV; t += NTHR) {
            const int rr = t / NV, vv = t - rr * NV;
            if (rr >= za && rr < zb && vv >= xa && vv < xb)
                stress_task<T, CT, TZ, ADJ, false>(C, rr, vv);
        }
        const int nzs = za + (SH - zb), nxs = xa + (NV - xb);
        const int n1 = nzs * NV, n2 = (SH - nzs) * nxs;
        for (int t = tid; t < n1 + n2; t += NTHR) {
            int rr, vv;
            if (t < n1) {
                const int i = t / NV;
                vv = t - i * NV;
                rr = i < za ? i : zb + (i - za);
            } else {
                const int t2 = t - n1, i = t2 / nxs, jx = t2 - i * nxs;
                rr = za + i;
                vv = jx < xa ? jx : xb + (jx - xa);
            }
            stress_task<T, CT, TZ, ADJ, true>(C, rr, vv);
        }
    }
    mbar_wait(&S.bar2, 0); // (every thread observes the second half itself before phase 3 reads ux, uz)
    __syncthreads();
    inject_mt<T, TZ, ADJ>(P, S, tile, tid);

    // ---- phase 3: displacements of the tile (update_ux! :1-18, update_uz! :20-37) -----------------------------------------
    if (!EDGE) {
#pragma unroll
        for (int n = 0; n < NP3; ++n) {
            const int t = tid + NTHR * n;
            const int r = t / NVT;
            disp_task<T, CT, TZ, ADJ, false>(C, r, t - r * NVT, r_uxo[n], r_uzo[n], r_fi[n], r_fj[n]);
        }
    } else {
        const int za = min(max(Jlo - (z0 + 1), 0), TZ), zb = min(max(Jhi - (z0 + 1) + 1, za), TZ);
        int xa = 0, xb = NVT;
        for (int v = 0; v < NVT; ++v) {
            if (x0 + V * v + 1 < Ilo)
                xa = v + 1;
            if (x0 + V * v + V > Ihi && xb == NVT)
                xb = v;
        }
        xb = max(xb, xa);
#pragma unroll
        for (int n = 0; n < NP3; ++n) {
            const int t = tid + NTHR * n;
            const int r = t / NVT, v = t - r * NVT;
            if (r >= za && r < zb && v >= xa && v < xb)
                disp_task<T, CT, TZ, ADJ, false>(C, r, v, r_uxo[n], r_uzo[n], r_fi[n], r_fj[n]);
        }
        const int nzs = za + (TZ - zb), nxs = xa + (NVT - xb);
        const int n1 = nzs * NVT, n2 = (TZ - nzs) * nxs;
        for (int t = tid; t < n1 + n2; t += NTHR) {
            int r, v;
            if (t < n1) {
                const int i = t / NVT;
                v = t - i * NVT;
                r = i < za ? i : zb + (i - za);
            } else {
                const int t2 = t - n1, i = t2 / nxs, jx = t2 - i * nxs;
                r = za + i;
                v = jx < xa ? jx : xb + (jx - xa);
            }
            const long long q = (long long)(z0 + r) * ld + (x0 + v * V);
            T uxo_[V], uzo_[V], fi_[V], fj_[V]; // (the owner of these cells has not written them yet: they are written here)
            ldv(P.uxo + q, uxo_);
            ldv(P.uzo + q, uzo_);
            ldv_ro(P.fac_ih + q, fi_);
            ldv_ro(P.fac_jh + q, fj_);
            disp_task<T, CT, TZ, ADJ, true>(C, r, v, uxo_, uzo_, fi_, fj_);
        }
    }

    // ---- external forces / adjoint sources on the owned cells: one parallel pass per round (see ElaFusedParams::fi_off) --------
    if (P.fi_it > 0) {
        const int g0 = P.fi_off[tile], g1 = P.fi_off[tile + 1];
        if (g0 < g1) {
            __syncthreads(); // the tile's unew is in memory
            const T dt2 = P.dt * P.dt;
            for (int g = g0; g < g1; ++g) {
                const int e0 = P.fi_roff[g], e1 = P.fi_roff[g + 1];
                for (int k = e0 + tid; k < e1; k += NTHR) {
                    const int cell = P.fi_cell[k], s_ = P.fi_src[k];
                    const int comp = cell >= ELF_MT_FIELD ? 1 : 0;
                    const int o = cell - comp * ELF_MT_FIELD;
                    const long long q = (long long)(z0 + o / TX) * ld + (x0 + o % TX);
                    const T cf = P.fi_coef[k];
                    const T w = P.fi_tf[((long long)s_ * 2 + comp) * P.nt + (P.fi_it - 1)];
                    if (comp == 0)
                        P.uxn[q] = P.uxn[q] + ((cf * w) / P.rho_ih[q]) * dt2;
                    else
                        P.uzn[q] = P.uzn[q] + ((cf * w) / P.rho_jh[q]) * dt2;
                }
                if (g + 1 < g1)
                    __syncthreads();
            }
        }
    }
}
#undef MAD

// PART 0: every tile in one launch; 1: interior tiles only (edge tiles exit at once); 2: edge tiles only, blockIdx.x enumerates
// them compactly: the full rows outside [era, erb) first, then the columns outside [eca, ecb) of the rows inside.
template <class T, class CT, int TZ, bool ADJ, int PART>
__global__ void __launch_bounds__(NTHR, (PART == 1 && !ADJ && sizeof(T) == 4 && TZ == 16) ? 3 : 2) ela_fused_kernel(const __grid_constant__ ElaFusedParams<T> P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ElaSmem<T, TZ, ADJ> &S = *reinterpret_cast<ElaSmem<T, TZ, ADJ> *>(smem_raw);
    int bx, trow;
    if (PART == 2) {
        const int t = (int)blockIdx.x;
        const int nrs = P.era + (P.ntz - P.erb), n1 = nrs * P.ntx;
        if (t < n1) {
            const int i = t / P.ntx;
            bx = t - i * P.ntx;
            trow = i < (P.ntz - P.erb) ? P.erb + i : i - (P.ntz - P.erb); // bottom strip rows first
        } else {
            const int ncs = P.eca + (P.ntx - P.ecb), t2 = t - n1, i = t2 / ncs, j = t2 - i * ncs;
            trow = P.era + i;
            bx = j < P.eca ? j : P.ecb + (j - P.eca);
        }
    } else {
        bx = (int)blockIdx.x;
        trow = tile_row<TZ>(P.halo, P.ntz, (int)blockIdx.y, P.rev);
    }
    // interior tile: every cell of the tile's stress region (1-based indices x0-1 .. x0+TX+2, z0-1 .. z0+TZ+2) lies inside all
    // update ranges, outside every C-PML strip and below the free-surface rows (host: ela_interior_range)
    const bool interior = (trow >= P.era && trow < P.erb && bx >= P.eca && bx < P.ecb) || P.dbg_all_interior;
    if (PART == 1) {
        if (interior)
            ela_tile<T, CT, TZ, ADJ, false>(P, S, bx, trow);
    } else if (PART == 2) {
        ela_tile<T, CT, TZ, ADJ, true>(P, S, bx, trow);
    } else {
        if (interior)
            ela_tile<T, CT, TZ, ADJ, false>(P, S, bx, trow);
        else
            ela_tile<T, CT, TZ, ADJ, true>(P, S, bx, trow);
    }
}

// ---- small kernels on the padded layout ---------------------------------------------------------------------------------
#define PIX(i, j) ((long long)((j) - 1) * ld + ((i) - 1))

template <class T>
__global__ void __launch_bounds__(256) elf_inject_force_kernel(long long ld, T *ux, T *uz, const T *rho_ih, const T *rho_jh, swb_sinc_points l0, swb_sinc_points l1,
                                                               const T *tf, long long nt, long long it, T dt)
{
    const long long np0 = l0.off[l0.n], np1 = l1.off[l1.n];
    const T *c0 = (const T *)l0.coef, *c1 = (const T *)l1.coef;
    const T dt2 = dt * dt;
    for (long long s = 0; s < l0.n; ++s) { // sources in index order (they may share cells), the points of one source in parallel
        const T wx = tf[((size_t)s * 2 + 0) * nt + (it - 1)], wz = tf[((size_t)s * 2 + 1) * nt + (it - 1)];
        for (long long p = l0.off[s] + threadIdx.x; p < l0.off[s + 1]; p += blockDim.x) {
            const long long q = PIX(l0.ij[p], l0.ij[p + np0]);
            ux[q] = ux[q] + ((c0[p] * wx) / rho_ih[q]) * dt2;
        }
        for (long long p = l1.off[s] + threadIdx.x; p < l1.off[s + 1]; p += blockDim.x) {
            const long long q = PIX(l1.ij[p], l1.ij[p + np1]);
            uz[q] = uz[q] + ((c1[p] * wz) / rho_jh[q]) * dt2;
        }
        __syncthreads();
    }
}

template <class T>
__global__ void elf_record_kernel(long long ld, const T *ux, const T *uz, swb_sinc_points lx, swb_sinc_points lz, T *traces, long long nt, long long it)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 2 * lx.n)
        return;
    const long long r = t >> 1;
    const int comp = (int)(t & 1);
    const swb_sinc_points &l = comp == 0 ? lx : lz;
    const long long np = l.off[l.n];
    const T *u = comp == 0 ? ux : uz, *c = (const T *)l.coef;
    T acc = (T)0;
    for (long long p = l.off[r]; p < l.off[r + 1]; ++p) // left to right, like the oracle
        acc = acc + c[p] * u[PIX(l.ij[p], l.ij[p + np])];
    traces[((size_t)r * 2 + comp) * nt + (it - 1)] = acc;
}

// correlate_gradients! on padded planes: the zero padding outside every array replaces the wrappers' range tests
template <class T, class CT>
__global__ void __launch_bounds__(256) elf_correlate_kernel(int nx, int nz, int freetop, long long ld, T inv_dx, T inv_dz, T _dt2, const T *aux, const T *auz,
                                                            const T *uxo, const T *uzo, const T *uxc, const T *uzc, const T *uxn, const T *uzn, const T *lam,
                                                            const T *mu, T *g_ri, T *g_rj, T *g_l, T *g_m, T *g_mh)
{
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x) + 1;
    const int j = (int)(blockIdx.y * blockDim.y + threadIdx.y) + 1;
    if (i > nx || j > nz)
        return;
    const long long q = PIX(i, j);
    const bool ft = freetop != 0;
    if (i <= nx - 1) { // grad_ρ_ihalf on (1:nx-1, 1:nz), all in T
        const T v = (aux[q] * ((uxo[q] - (T)2 * uxc[q]) + uxn[q])) * _dt2;
        g_ri[q] = g_ri[q] + ((j == 1 && ft) ? v / (T)2 : v);
    }
    if (j <= nz - 1) // grad_ρ_jhalf on (1:nx, 1:nz-1)
        g_rj[q] = g_rj[q] + (auz[q] * ((uzo[q] - (T)2 * uzc[q]) + uzn[q])) * _dt2;
    const int j0 = ft ? 1 : 2;
    if (i >= 2 && i <= nx - 1 && j >= j0 && j <= nz - 1) { // grad_λ, grad_μ
        const CT exx = inner4<T, CT>(uxc[q - 2], uxc[q - 1], uxc[q], uxc[q + 1], inv_dx);
        const CT exx_a = inner4<T, CT>(aux[q - 2], aux[q - 1], aux[q], aux[q + 1], inv_dx);
        CT ezz, ezz_a;
        if (ft && j == 1) {
            const T l = lam[q], m = mu[q];
            const T fac = -l / (l + (T)2 * m);
            ezz = (CT)fac * exx;
            ezz_a = (CT)fac * exx_a;
        } else if (ft && j == 2) {
            ezz = inner4<T, CT>(uzc[q - ld], uzc[q - ld], uzc[q], uzc[q + ld], inv_dz);
            ezz_a = inner4<T, CT>(auz[q - ld], auz[q - ld], auz[q], auz[q + ld], inv_dz);
        } else {
            ezz = inner4<T, CT>(uzc[q - 2 * ld], uzc[q - ld], uzc[q], uzc[q + ld], inv_dz);
            ezz_a = inner4<T, CT>(auz[q - 2 * ld], auz[q - ld], auz[q], auz[q + ld], inv_dz);
        }
        const CT div_u = exx + ezz, div_a = exx_a + ezz_a;
        if (j == 1 && ft) {
            g_l[q] = (T)((CT)g_l[q] + (div_u * div_a) / (CT)2);
            g_m[q] = (T)((CT)g_m[q] + (exx * exx_a + ezz * ezz_a));
        } else {
            g_l[q] = (T)((CT)g_l[q] + div_u * div_a);
            g_m[q] = (T)((CT)g_m[q] + (CT)2 * (exx * exx_a + ezz * ezz_a));
        }
    }
    if (i <= nx - 1 && j <= nz - 1) { // grad_μ_ihalf_jhalf
        const CT dwdx = inner4<T, CT>(uzc[q - 1], uzc[q], uzc[q + 1], uzc[q + 2], inv_dx);
        const CT dwdx_a = inner4<T, CT>(auz[q - 1], auz[q], auz[q + 1], auz[q + 2], inv_dx);
        CT dudz, dudz_a;
        if (ft && j == 1) {
            dudz = inner4<T, CT>(uxc[q + ld], uxc[q], uxc[q + ld], uxc[q + 2 * ld], inv_dz);
            dudz_a = inner4<T, CT>(aux[q + ld], aux[q], aux[q + ld], aux[q + 2 * ld], inv_dz);
        } else {
            dudz = inner4<T, CT>(uxc[q - ld], uxc[q], uxc[q + ld], uxc[q + 2 * ld], inv_dz);
            dudz_a = inner4<T, CT>(aux[q - ld], aux[q], aux[q + ld], aux[q + 2 * ld], inv_dz);
        }
        const CT exz = (dwdx + dudz) / (CT)2, exz_a = (dwdx_a + dudz_a) / (CT)2;
        g_mh[q] = (T)((CT)g_mh[q] + (CT)2 * (exz * exz_a + exz * exz_a));
    }
}
#undef PIX

template <class T>
__global__ void __launch_bounds__(256) elf_dt2_over_rho_kernel(long long ld, long long w, long long hgt, const T *rho, T *fac, T dt)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (i >= w || j >= hgt)
        return;
    const T dt2 = dt * dt;
    fac[j * ld + i] = dt2 / rho[j * ld + i];
}

} // namespace

template <class T, class CT, int TZ, bool ADJ, int PART>
static void ela_fused_launch_p(const ElaFusedParams<T> &P, dim3 grd, cudaStream_t st)
{
    const size_t smem = sizeof(ElaSmem<T, TZ, ADJ>);
    int dev = 0;
    SWB_CUDA(cudaGetDevice(&dev));
    static bool done[64] = {}; // opt-in to > 48 KB of dynamic shared memory once per device and instantiation
    if (dev < 64 && !done[dev]) {
        SWB_CUDA(cudaFuncSetAttribute(ela_fused_kernel<T, CT, TZ, ADJ, PART>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done[dev] = true;
    }
    ela_fused_kernel<T, CT, TZ, ADJ, PART><<<grd, NTHR, smem, st>>>(P);
    check_launch("ela_fused_kernel");
    count_launch();
}

template <class T, class CT, int TZ, bool ADJ>
static void ela_fused_launch_v(const ElaFusedParams<T> &P0, cudaStream_t st, cudaStream_t st_edge)
{
    static const int dbg = [] { const char *e = std::getenv("SWB_ELF_DEBUG_ALL_INTERIOR"); return e ? std::atoi(e) : 0; }();
    ElaFusedParams<T> P = P0;
    P.dbg_all_interior = dbg;
    P.ntx = (int)cdiv(P.nx, TX);
    P.ntz = (int)cdiv(P.nz, TZ);
    // interior tile rows / columns: x0 - 1 > m, x0 + TX + 2 < nx - 1 - halo (likewise z, with m = 2 above an inactive top strip)
    const int m = std::max(P.halo + 1, 2), mz = P.top_inactive ? 2 : m;
    auto range = [](int n_tiles, int t, int lo_excl, int hi_excl, int &a, int &b) { // tiles i with i*t - 1 > lo_excl and i*t + t + 2 < hi_excl
        a = n_tiles, b = n_tiles;
        for (int i = n_tiles - 1; i >= 0; --i) {
            const bool in = i * t - 1 > lo_excl && i * t + t + 2 < hi_excl;
            if (in)
                a = i;
            else if (a == n_tiles)
                b = i;
        }
        if (a == n_tiles)
            a = b = 0;
    };
    range(P.ntx, TX, m, P.nx - 1 - P.halo, P.eca, P.ecb);
    range(P.ntz, TZ, mz, P.nz - 1 - P.halo, P.era, P.erb);
    const int n_int = (P.erb - P.era) * (P.ecb - P.eca), n_edge = P.ntx * P.ntz - n_int;
    if (st_edge != nullptr && n_int > 0 && n_edge > 0 && !dbg) {
        ela_fused_launch_p<T, CT, TZ, ADJ, 1>(P, dim3(P.ntx, P.ntz, 1), st);
        ela_fused_launch_p<T, CT, TZ, ADJ, 2>(P, dim3(n_edge, 1, 1), st_edge);
    } else
        ela_fused_launch_p<T, CT, TZ, ADJ, 0>(P, dim3(P.ntx, P.ntz, 1), st);
}

template <class T, class CT, int TZ>
static void ela_fused_launch_a(const ElaFusedParams<T> &P, cudaStream_t st, cudaStream_t st_edge)
{
    if (P.corr)
        ela_fused_launch_v<T, CT, TZ, true>(P, st, st_edge);
    else
        ela_fused_launch_v<T, CT, TZ, false>(P, st, st_edge);
}

template <>
void ela_fused_launch<float>(const ElaFusedParams<float> &P, bool fast, cudaStream_t st, cudaStream_t st_edge)
{
    SWB_REQUIRE(P.tz == 8 || P.tz == 16 || P.tz == 24, "fused elastic step: unsupported tile height");
    if (fast) {
        if (P.tz == 8)
            ela_fused_launch_a<float, float, 8>(P, st, st_edge);
        else if (P.tz == 16)
            ela_fused_launch_a<float, float, 16>(P, st, st_edge);
        else
            ela_fused_launch_a<float, float, 24>(P, st, st_edge);
    } else {
        if (P.tz == 8)
            ela_fused_launch_a<float, double, 8>(P, st, st_edge);
        else if (P.tz == 16)
            ela_fused_launch_a<float, double, 16>(P, st, st_edge);
        else
            ela_fused_launch_a<float, double, 24>(P, st, st_edge);
    }
}
template <>
void ela_fused_launch<double>(const ElaFusedParams<double> &P, bool, cudaStream_t st, cudaStream_t st_edge)
{
    SWB_REQUIRE(P.tz == 4 || P.tz == 8 || P.tz == 12, "fused elastic step: unsupported tile height");
    if (P.tz == 4)
        ela_fused_launch_a<double, double, 4>(P, st, st_edge);
    else if (P.tz == 8)
        ela_fused_launch_a<double, double, 8>(P, st, st_edge);
    else
        ela_fused_launch_a<double, double, 12>(P, st, st_edge);
}

void elf_dt2_over_rho(int dtype, long long ld, long long w, long long hgt, const void *rho, void *fac, double dt, cudaStream_t st)
{
    const dim3 grd(cdiv(w, 256), (unsigned)hgt, 1);
    if (dtype == SWB_F64)
        elf_dt2_over_rho_kernel<double><<<grd, 256, 0, st>>>(ld, w, hgt, (const double *)rho, (double *)fac, dt);
    else
        elf_dt2_over_rho_kernel<float><<<grd, 256, 0, st>>>(ld, w, hgt, (const float *)rho, (float *)fac, (float)dt);
    check_launch("elf_dt2_over_rho");
    count_launch();
}

void elf_inject_force(int dtype, long long ld, void *ux, void *uz, const void *rho_ih, const void *rho_jh, const swb_sinc_points &l0, const swb_sinc_points &l1,
                      const void *tf, long long nt, long long it, double dt, cudaStream_t st)
{
    if (l0.n <= 0)
        return;
    if (dtype == SWB_F64)
        elf_inject_force_kernel<double><<<1, 256, 0, st>>>(ld, (double *)ux, (double *)uz, (const double *)rho_ih, (const double *)rho_jh, l0, l1, (const double *)tf,
                                                          nt, it, dt);
    else
        elf_inject_force_kernel<float><<<1, 256, 0, st>>>(ld, (float *)ux, (float *)uz, (const float *)rho_ih, (const float *)rho_jh, l0, l1, (const float *)tf, nt, it,
                                                         (float)dt);
    check_launch("elf_inject_force");
    count_launch();
}

void elf_record(int dtype, long long ld, const void *ux, const void *uz, const swb_sinc_points &lx, const swb_sinc_points &lz, void *traces, long long nt,
                long long it, cudaStream_t st)
{
    if (lx.n <= 0)
        return;
    if (dtype == SWB_F64)
        elf_record_kernel<double><<<cdiv(2 * lx.n, 128), 128, 0, st>>>(ld, (const double *)ux, (const double *)uz, lx, lz, (double *)traces, nt, it);
    else
        elf_record_kernel<float><<<cdiv(2 * lx.n, 128), 128, 0, st>>>(ld, (const float *)ux, (const float *)uz, lx, lz, (float *)traces, nt, it);
    check_launch("elf_record");
    count_launch();
}

template <class T, class CT>
static void elf_correlate_t(const ElaCorrPadded &a, cudaStream_t st)
{
    const dim3 blk(32, 8, 1), grd(cdiv(a.nx, 32), cdiv(a.nz, 8), 1);
    const T dt = (T)a.dt;
    elf_correlate_kernel<T, CT><<<grd, blk, 0, st>>>(a.nx, a.nz, a.freetop, a.ld, (T)1 / (T)a.dx, (T)1 / (T)a.dz, (T)1 / (dt * dt), (const T *)a.aux, (const T *)a.auz,
                                                     (const T *)a.uxo, (const T *)a.uzo, (const T *)a.uxc, (const T *)a.uzc, (const T *)a.uxn, (const T *)a.uzn,
                                                     (const T *)a.lam, (const T *)a.mu, (T *)a.g_ri, (T *)a.g_rj, (T *)a.g_l, (T *)a.g_m, (T *)a.g_mh);
    check_launch("elf_correlate");
    count_launch();
}

void elf_correlate(const ElaCorrPadded &a, cudaStream_t st)
{
    if (a.dtype == SWB_F64)
        elf_correlate_t<double, double>(a, st);
    else if (a.flags & SWB_FLAG_FAST_F32)
        elf_correlate_t<float, float>(a, st);
    else
        elf_correlate_t<float, double>(a, st);
}

} // namespace swb
