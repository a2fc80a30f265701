// ela_fused.cu -- single-launch 2D elastic isotropic P-SV time step for the per-shot engine (stresses kept on chip).
//
// The reference advances one step with four stencil sweeps that round-trip the stresses through memory
// (src/models/elastic/backends/shared/elastic2D_iso_xPU.jl:1-79,120-239: update_σxx_σzz!, update_σxz!, update_ux!,
// update_uz!; 19 array passes per cell) plus two tiny launches per source and per receiver.  The stresses are a pure
// function of the current displacements (σ = C : ε(ucur), recomputed from scratch every step), so here a CTA
//   1. requests every input of its TX x TZ tile at once: ux, uz plus a 4-cell halo and the stress-update factors into shared
//      memory (cp.async), uold and the densities of the owned cells into registers,
//   2. computes σxx, σzz, σxz on the tile plus a 2-cell halo into shared memory -- the halo is recomputed, not exchanged;
//      the C-PML memory variables of the displacement derivatives are read from the `in` copy and written (owned cells
//      only) to the `out` copy -- and applies the moment-tensor injection to its on-chip stresses,
//   3. computes uxnew, uznew of the tile (with the C-PML memory variables of the stress derivatives) and stores them.
// HBM traffic per cell-update: ux, uz read, uold read / unew written in place (4), λ, μ, μ_ihalf_jhalf, ρ_ihalf, ρ_jhalf
// read = 11 values.  The arithmetic is the reference's, operation for operation: the derivative wrappers of
// freesurface_derivatives_4th_mirror.jl:1-244 (zero padding outside every array -- here literal zeros of the padded
// planes --, odd / even mirroring and the Hooke's-law row at the free surface), ∂̃4th of src/utils/fdgenerated.jl:178-195,
// stresses rounded to T before the displacement update reads them.
// External-force / adjoint-source injection and the receiver sums stay separate small launches (elf_inject_force,
// elf_record); the zero-lag correlations run in elf_correlate on the same padded planes.
#include "ela_fused.h"
#include "kernels.h"

namespace swb {

namespace {

constexpr int TX = ELF_TX, TZ = ELF_TZ, NTHR = 256;
constexpr int UW = TX + 8;  // staged displacement rows: columns -4 .. TX+3
constexpr int UH = TZ + 8;  // rows -4 .. TZ+3
constexpr int SW = ELF_SW;  // stress rows: columns -2 .. TX+1
constexpr int SH = TZ + 4;  // rows -2 .. TZ+1

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
template <class T>
__device__ __forceinline__ void cp_async_chunk(T *smem, const T *gmem) // 4 elements
{
    cp_async16(smem, gmem);
    if (sizeof(T) == 8)
        cp_async16((char *)smem + 16, (const char *)gmem + 16);
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// ∂x4th_inner / ∂y4th_inner (freesurface_derivatives_4th_mirror.jl:2-7)
template <class T, class CT>
__device__ __forceinline__ CT inner4(T f1, T f2, T f3, T f4, T inv)
{
    const CT c1 = (CT)(1.0 / 24.0), c2 = (CT)(27.0 / 24.0);
    return (((c1 * (CT)f1 - c2 * (CT)f2) + c2 * (CT)f3) - c1 * (CT)f4) * (CT)inv;
}

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

template <class T>
struct ElaSmem {
    T ux[UH * UW], uz[UH * UW];                  // displacements, rows -4 .. TZ+3, columns -4 .. TX+3
    T lam[SH * UW], mu[SH * UW], muhh[SH * UW];  // stress-update factors, rows -2 .. TZ+1, columns -4 .. TX+3
    T sxx[SH * SW], szz[SH * SW], sxz[SH * SW];  // stresses, rows -2 .. TZ+1, columns -2 .. TX+1
};
constexpr int NP3 = TZ * TX / NTHR; // cells per thread in the displacement update (fixed mapping: column tid % TX, rows tid / TX + 2 n)

template <class T, class CT, bool EDGE>
__device__ __forceinline__ void ela_tile(const ElaFusedParams<T> &P, ElaSmem<T> &S)
{
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX, z0 = blockIdx.y * TZ;
    const int nx = P.nx, nz = P.nz, h = P.halo;
    const long long ld = P.ld;
    const bool ft = P.freetop != 0;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;

    // ---- phase 1: every input of the tile is requested before anything is computed, so that a CTA has its whole working set
    //      in flight at once: ux, uz (+4-cell halo) and λ, μ, μ_ihalf_jhalf (+2 halo rows) -> shared memory with cp.async;
    //      uold, ρ_ihalf, ρ_jhalf of the owned cells -> registers (used once, in phase 3)
    for (int idx = tid; idx < UH * (UW / 4); idx += NTHR) {
        const int r = idx / (UW / 4), ch = idx - r * (UW / 4);
        const int gx = x0 - 4 + 4 * ch;
        const long long q = (long long)(z0 - 4 + r) * ld + gx;
        T *dx = S.ux + r * UW + 4 * ch, *dz = S.uz + r * UW + 4 * ch;
        if (gx >= ld) { // beyond the row pitch (last tile of a row): zeros, like everything outside the arrays
            dx[0] = dx[1] = dx[2] = dx[3] = (T)0;
            dz[0] = dz[1] = dz[2] = dz[3] = (T)0;
        } else {
            cp_async_chunk(dx, P.uxc + q);
            cp_async_chunk(dz, P.uzc + q);
        }
    }
    for (int idx = tid; idx < SH * (UW / 4); idx += NTHR) {
        const int r = idx / (UW / 4), ch = idx - r * (UW / 4);
        const int gx = x0 - 4 + 4 * ch;
        const long long q = (long long)(z0 - 2 + r) * ld + gx;
        if (gx < ld) { // (factors beyond the pitch only feed cells that are masked out)
            cp_async_chunk(S.lam + r * UW + 4 * ch, P.lam + q);
            cp_async_chunk(S.mu + r * UW + 4 * ch, P.mu + q);
            cp_async_chunk(S.muhh + r * UW + 4 * ch, P.mu_hh + q);
        }
    }
    T r_uxo[NP3], r_uzo[NP3], r_ri[NP3], r_rj[NP3];
    {
        const int c = tid % TX, rb = tid / TX;
#pragma unroll
        for (int n = 0; n < NP3; ++n) { // all inside the padded plane (junk beyond the arrays is masked at the point of use)
            const long long q = (long long)(z0 + rb + (NTHR / TX) * n) * ld + (x0 + c);
            r_uxo[n] = P.uxo[q];
            r_uzo[n] = P.uzo[q];
            r_ri[n] = P.rho_ih[q];
            r_rj[n] = P.rho_jh[q];
        }
    }
    cp_async_wait_all();
    __syncthreads();

#define UX(r, c) S.ux[((r) + 4) * UW + (c) + 4]
#define UZ(r, c) S.uz[((r) + 4) * UW + (c) + 4]
#define SXX(r, c) S.sxx[((r) + 2) * SW + (c) + 2]
#define SZZ(r, c) S.szz[((r) + 2) * SW + (c) + 2]
#define SXZ(r, c) S.sxz[((r) + 2) * SW + (c) + 2]

    // ---- phase 2: stresses on the tile + 2-cell halo (update_σxx_σzz! :39-60, update_σxz! :62-79) ------------------------
    for (int idx = tid; idx < SH * SW; idx += NTHR) {
        const int rr = idx / SW, cc = idx - rr * SW;
        const int r = rr - 2, c = cc - 2;
        const int I = x0 + c + 1, J = z0 + r + 1; // 1-based reference indices
        const bool own = r >= 0 && r < TZ && c >= 0 && c < TX;
        const int qm = rr * UW + c + 4; // this cell in the staged factor arrays
        T sxx = (T)0, szz = (T)0, sxz = (T)0;
        const int j0 = ft ? 1 : 2;
        if (!EDGE || (I >= 2 && I <= nx - 1 && J >= j0 && J <= nz - 1)) {
            const CT dudx = inner4<T, CT>(UX(r, c - 2), UX(r, c - 1), UX(r, c), UX(r, c + 1), P.inv_dx);
            const T l = S.lam[qm], m = S.mu[qm];
            CT dwdz;
            if (EDGE && ft && J == 1) { // Hooke's law on the free-surface row (:179-195)
                const T fac = -l / (l + (T)2 * m);
                dwdz = (CT)fac * dudx;
            } else if (EDGE && ft && J == 2)
                dwdz = inner4<T, CT>(UZ(r - 1, c), UZ(r - 1, c), UZ(r, c), UZ(r + 1, c), P.inv_dz);
            else
                dwdz = inner4<T, CT>(UZ(r - 2, c), UZ(r - 1, c), UZ(r, c), UZ(r + 1, c), P.inv_dz);
            CT dudx_c = dudx, dwdz_c = dwdz;
            if (EDGE) {
                dudx_c = cpml4<T, CT>(dudx, I - 1, nx - 1, h, true, P.a_x, P.b_x, P.psi_in[4], P.psi_out[4], (long long)(J - 1) * (2 * (h + 1)), 1, own);
                dwdz_c = cpml4<T, CT>(dwdz, J - 1, nz - 1, h, true, P.a_z, P.b_z, P.psi_in[7], P.psi_out[7], (long long)(I - 1), nx, own);
            }
            const T l2m = l + (T)2 * m;
            sxx = (T)((CT)l2m * dudx_c + (CT)l * dwdz_c);
            szz = (EDGE && J == 1) ? (T)0 : (T)((CT)l * dudx_c + (CT)l2m * dwdz_c);
        }
        if (!EDGE || (I >= 1 && I <= nx - 1 && J >= 1 && J <= nz - 1)) {
            const CT dwdx = inner4<T, CT>(UZ(r, c - 1), UZ(r, c), UZ(r, c + 1), UZ(r, c + 2), P.inv_dx);
            CT dudz;
            if (EDGE && ft && J == 1) // even mirror of ux at the free surface (:214-222)
                dudz = inner4<T, CT>(UX(r + 1, c), UX(r, c), UX(r + 1, c), UX(r + 2, c), P.inv_dz);
            else
                dudz = inner4<T, CT>(UX(r - 1, c), UX(r, c), UX(r + 1, c), UX(r + 2, c), P.inv_dz);
            CT dwdx_c = dwdx, dudz_c = dudz;
            if (EDGE) {
                dwdx_c = cpml4<T, CT>(dwdx, I, nx, h, false, P.a_xh, P.b_xh, P.psi_in[5], P.psi_out[5], (long long)(J - 1) * (2 * h), 1, own);
                dudz_c = cpml4<T, CT>(dudz, J, nz, h, false, P.a_zh, P.b_zh, P.psi_in[6], P.psi_out[6], (long long)(I - 1), nx - 1, own);
            }
            sxz = (T)((CT)S.muhh[qm] * (dwdx_c + dudz_c));
        }
        S.sxx[idx] = sxx;
        S.szz[idx] = szz;
        S.sxz[idx] = sxz;
    }
    __syncthreads();

    // ---- moment-tensor injection into the on-chip stresses (inject_momten_sources2D_σxx_σzz! / _σxz! :81-94) --------------
    if (P.mt_it > 0) {
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
                if (cell < ELF_SREGION) {
                    S.sxx[cell] = S.sxx[cell] + (P.Mxx[s] * cf) * w;
                    S.szz[cell] = S.szz[cell] + (P.Mzz[s] * cf) * w;
                } else
                    S.sxz[cell - ELF_SREGION] = S.sxz[cell - ELF_SREGION] + (P.Mxz[s] * cf) * w;
            }
            __syncthreads();
            e = en;
        }
    }

    // ---- phase 3: displacements of the tile (update_ux! :1-18, update_uz! :20-37) -----------------------------------------
    const T dt2 = P.dt * P.dt;
#pragma unroll
    for (int n = 0; n < NP3; ++n) {
        const int r = tid / TX + (NTHR / TX) * n, c = tid % TX;
        const int I = x0 + c + 1, J = z0 + r + 1;
        if (EDGE && (I > nx || J > nz))
            continue;
        const long long q = (long long)(z0 + r) * ld + (x0 + c);
        if (!EDGE || I <= nx - 1) {
            const CT d1 = inner4<T, CT>(SXX(r, c - 1), SXX(r, c), SXX(r, c + 1), SXX(r, c + 2), P.inv_dx);
            CT d2;
            if (EDGE && ft && J == 1) // odd mirror of σxz at the free surface (:125-140)
                d2 = inner4<T, CT>(-SXZ(r + 1, c), -SXZ(r, c), SXZ(r, c), SXZ(r + 1, c), P.inv_dz);
            else if (EDGE && ft && J == 2)
                d2 = inner4<T, CT>(-SXZ(r - 1, c), SXZ(r - 1, c), SXZ(r, c), SXZ(r + 1, c), P.inv_dz);
            else
                d2 = inner4<T, CT>(SXZ(r - 2, c), SXZ(r - 1, c), SXZ(r, c), SXZ(r + 1, c), P.inv_dz);
            CT c1 = d1, c2 = d2;
            if (EDGE) {
                c1 = cpml4<T, CT>(d1, I, nx, h, false, P.a_xh, P.b_xh, P.psi_in[0], P.psi_out[0], (long long)(J - 1) * (2 * h), 1, true);
                c2 = cpml4<T, CT>(d2, J - 1, nz - 1, h, true, P.a_z, P.b_z, P.psi_in[3], P.psi_out[3], (long long)(I - 1), nx - 1, true);
            }
            const T t = (T)2 * UX(r, c) - r_uxo[n];
            const T f = dt2 / r_ri[n];
            P.uxn[q] = (T)((CT)t + (CT)f * (c1 + c2));
        }
        if (!EDGE || J <= nz - 1) {
            const CT d1 = inner4<T, CT>(SXZ(r, c - 2), SXZ(r, c - 1), SXZ(r, c), SXZ(r, c + 1), P.inv_dx);
            CT d2;
            if (EDGE && ft && J == 1) // odd mirror of σzz at the free surface (:85-93)
                d2 = inner4<T, CT>(-SZZ(r + 1, c), SZZ(r, c), SZZ(r + 1, c), SZZ(r + 2, c), P.inv_dz);
            else
                d2 = inner4<T, CT>(SZZ(r - 1, c), SZZ(r, c), SZZ(r + 1, c), SZZ(r + 2, c), P.inv_dz);
            CT c1 = d1, c2 = d2;
            if (EDGE) {
                c1 = cpml4<T, CT>(d1, I - 1, nx - 1, h, true, P.a_x, P.b_x, P.psi_in[1], P.psi_out[1], (long long)(J - 1) * (2 * (h + 1)), 1, true);
                c2 = cpml4<T, CT>(d2, J, nz, h, false, P.a_zh, P.b_zh, P.psi_in[2], P.psi_out[2], (long long)(I - 1), nx, true);
            }
            const T t = (T)2 * UZ(r, c) - r_uzo[n];
            const T f = dt2 / r_rj[n];
            P.uzn[q] = (T)((CT)t + (CT)f * (c1 + c2));
        }
    }
#undef UX
#undef UZ
#undef SXX
#undef SZZ
#undef SXZ
}

template <class T, class CT>
__global__ void __launch_bounds__(NTHR) ela_fused_kernel(const __grid_constant__ ElaFusedParams<T> P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ElaSmem<T> &S = *reinterpret_cast<ElaSmem<T> *>(smem_raw);
    const int x0 = blockIdx.x * TX, z0 = blockIdx.y * TZ;
    // interior tile: every cell of the tile's stress region (1-based indices x0-1 .. x0+TX+2, z0-1 .. z0+TZ+2) lies inside all
    // update ranges, outside every C-PML strip and below the free-surface rows
    const int m = max(P.halo + 1, 2);
    const bool interior = x0 - 1 > m && x0 + TX + 2 < P.nx - 1 - P.halo && z0 - 1 > m && z0 + TZ + 2 < P.nz - 1 - P.halo;
    if (interior)
        ela_tile<T, CT, false>(P, S);
    else
        ela_tile<T, CT, true>(P, S);
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

} // namespace

template <class T>
void ela_fused_launch(const ElaFusedParams<T> &P, bool fast, cudaStream_t st)
{
    const dim3 grd(cdiv(P.nx, TX), cdiv(P.nz, TZ), 1);
    const size_t smem = sizeof(ElaSmem<T>);
    int dev = 0;
    SWB_CUDA(cudaGetDevice(&dev));
    static bool done[2][64] = {}; // opt-in to > 48 KB of dynamic shared memory once per device and instantiation
    const int v = (sizeof(T) == 4 && fast) ? 0 : 1;
    if (dev < 64 && !done[v][dev]) {
        if (v == 0)
            SWB_CUDA(cudaFuncSetAttribute(ela_fused_kernel<T, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else
            SWB_CUDA(cudaFuncSetAttribute(ela_fused_kernel<T, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done[v][dev] = true;
    }
    if (v == 0)
        ela_fused_kernel<T, T><<<grd, NTHR, smem, st>>>(P);
    else
        ela_fused_kernel<T, double><<<grd, NTHR, smem, st>>>(P);
    check_launch("ela_fused_kernel");
    count_launch();
}
template void ela_fused_launch<float>(const ElaFusedParams<float> &, bool, cudaStream_t);
template void ela_fused_launch<double>(const ElaFusedParams<double> &, bool, cudaStream_t);

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
