// ela_fused.h -- parameter blocks and layout constants of the fused 2D elastic P-SV step (ela_fused.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace swb {

constexpr int ELF_TX = 128;    // tile width in cells
constexpr int ELF_TZ_MAX = 32; // largest tile height any instantiation uses
constexpr int ELF_GB = 5;      // zero guard rows in front of a padded plane (4 halo rows + the left halo of row -4)
constexpr int ELF_GA = ELF_TZ_MAX + 8; // zero guard rows behind it
// tile height in cells, per storage type and launch kind (forward / adjoint): two CTAs fit the 227 KB of shared memory of an SM
int elf_tz(int dtype, bool adjoint, long long nx, long long nz);

// Padded plane: row pitch ld >= nx + 8 (multiple of 32 elements): at least 4 zero columns after the last cell of a row
// and 4 before the first cell of the next one, so the 4-point stencils of a tile's halo read zeros outside every
// array's index range -- the reference's derivative wrappers pad with zeros there
// (src/models/elastic/backends/shared/freesurface_derivatives_4th_mirror.jl:72-244).
inline long long elf_ld(long long nx) { return ((nx + 8 + 31) / 32) * 32; }
inline size_t elf_plane_elems(long long nx, long long nz) { return (size_t)elf_ld(nx) * (size_t)(nz + ELF_GB + ELF_GA); }
inline size_t elf_origin(long long nx) { return (size_t)elf_ld(nx) * ELF_GB; }

template <class T>
struct ElaFusedParams {
    // TMA descriptors of the padded planes staged through shared memory: ux, uz (current), λ, μ, μ_ihalf_jhalf, forward ux, uz [it-1]
    // (each array arrives as two row halves on two mbarriers: [0,1] ux, uz; [2..4] factors and [5,6] forward fields, first half;
    //  [7..9] factors and [10,11] forward fields, second half -- the halves of ux / uz have equal heights and share a descriptor)
    alignas(64) CUtensorMap tm[12];
    int nx, nz, halo, freetop, tz;
    int ntx, ntz;           // tile grid
    int era, erb, eca, ecb; // interior tiles (plain expressions everywhere): tile rows [era, erb) x tile columns [eca, ecb)
    int top_inactive; // the top C-PML strip has a = 0 (free surface): its memory variables stay 0 and ∂̃ = ∂ there
    long long ld;
    T inv_dx, inv_dz, dt;
    // padded planes, pointers to cell (1,1) of each array (ux: (nx-1, nz), uz: (nx, nz-1), ...)
    const T *uxc, *uzc, *uxo, *uzo;
    T *uxn, *uzn; // may alias uxo / uzo
    // λ, μ, μ_ihalf_jhalf and the displacement-update factors dt^2 / ρ_ihalf, dt^2 / ρ_jhalf (zero outside their arrays)
    const T *lam, *mu, *mu_hh, *fac_ih, *fac_jh;
    // C-PML memory variables, dense reference layouts (ela_models.jl:275-299), double-buffered because a tile recomputes
    // the stresses of its halo: 0 ψ_∂σxx∂x (2h, nz)  1 ψ_∂σxz∂x (2(h+1), nz-1)  2 ψ_∂σzz∂z (nx, 2h)  3 ψ_∂σxz∂z (nx-1, 2(h+1))
    //                           4 ψ_∂ux∂x (2(h+1), nz)  5 ψ_∂uz∂x (2h, nz-1)  6 ψ_∂ux∂z (nx-1, 2h)  7 ψ_∂uz∂z (nx, 2(h+1))
    const T *psi_in[8];
    T *psi_out[8];
    const T *a_x, *a_xh, *b_x, *b_xh, *a_z, *a_zh, *b_z, *b_zh;
    // moment-tensor injection into the on-chip stresses: per-tile lists over the tile's stress region (tile + 2 halo cells),
    // entries grouped by source in index order.  mt_it = 0 disables.
    const int *mt_off, *mt_cell, *mt_src; // cell = field * ELF_MT_FIELD + offset inside the staged region; field 0: σxx and σzz, 1: σxz
    const T *mt_coef;
    const T *srctf, *Mxx, *Mzz, *Mxz;
    long long nt;
    int mt_it;
    // external-force / adjoint-source injection into the freshly stored unew (inject_external_sources2D! and the residual
    // injection of adjoint_onestep_CPML!, elastic2D_iso_xPU.jl:96-106,433-440): per-tile lists over the owned cells, split
    // into rounds (round r = the r-th contribution of every cell, in source order; a round is one parallel pass): fi_off maps a tile
    // to its rounds, fi_roff a round to its entries; cell = component * ELF_MT_FIELD + row * ELF_TX + column.  fi_it = 0 disables.
    const int *fi_off, *fi_roff, *fi_cell, *fi_src;
    const T *fi_coef, *fi_tf, *rho_ih, *rho_jh;
    int fi_it;
    // zero-lag correlation fused into an adjoint launch (correlate_gradients!, elastic/backends/shared/
    // correlate_gradient_xPU.jl:47-83): this launch's *current* adjoint field (the one the previous launch produced) against
    // the forward displacements of three consecutive steps.  corr = 0 disables.
    int corr;
    const T *fxo, *fzo, *fxc, *fzc, *fxn, *fzn; // forward u[it-2], u[it-1], u[it]
    T *g_ri, *g_rj, *g_l, *g_m, *g_mh;
    T inv_dt2;
    int dbg_all_interior; // timing experiment only (SWB_ELF_DEBUG_ALL_INTERIOR=1): wrong results in the strips
    int rev;              // 1: tile rows after the edge rows in descending order (serpentine sweep: a launch starts on the rows the previous one left in L2)
};

// st_edge != nullptr: the interior tiles and the edge tiles (C-PML strips, grid edges, free surface) run as two kernels, the edge
// one on st_edge beside the interior one (disjoint tiles, both read only the previous time levels; the caller forks / joins the
// streams): the interior kernel then carries none of the per-cell code and fits one more CTA per SM.
template <class T>
void ela_fused_launch(const ElaFusedParams<T> &P, bool fast, cudaStream_t st, cudaStream_t st_edge);
template <>
void ela_fused_launch<float>(const ElaFusedParams<float> &P, bool fast, cudaStream_t st, cudaStream_t st_edge);
template <>
void ela_fused_launch<double>(const ElaFusedParams<double> &P, bool fast, cudaStream_t st, cudaStream_t st_edge);

// rows of the first of the two halves a staged array of `rows` rows arrives in (a multiple of 4 Float32 / 2 Float64 rows keeps the
// second half's shared-memory destination 128-byte aligned)
__host__ __device__ inline constexpr int elf_half_rows(int rows, size_t esize)
{
    return esize == 4 ? ((rows / 2 + 3) / 4) * 4 : ((rows / 2 + 1) / 2) * 2;
}

// tensor map of a whole padded plane (plane_base = first guard row), box = ELF_SW columns x box_rows rows
void elf_make_tmap(CUtensorMap *out, int dtype, const void *plane_base, long long ld, long long rows, int box_rows);

// staged region of a tile in shared memory: rows -2 .. TZ+1 (stresses; displacements -4 .. TZ+3), columns -4 .. TX+3
constexpr int ELF_SW = ELF_TX + 8;
constexpr int ELF_MT_FIELD = 1 << 24;
inline int elf_mt_cell(int field, int r, int c) { return field * ELF_MT_FIELD + (r + 2) * ELF_SW + (c + 4); }

// small kernels on the padded layout (sinc lists as uploaded by the engine: 1-based (i, j) into the array they address)
// external-force / adjoint-source injection into unew (elastic2D_iso_xPU.jl:96-106), sources in index order
void elf_inject_force(int dtype, long long ld, void *ux, void *uz, const void *rho_ih, const void *rho_jh, const swb_sinc_points &l0, const swb_sinc_points &l1,
                      const void *tf, long long nt, long long it, double dt, cudaStream_t st);
// traces[it, c, r] = sum_p coef[p] * u[ij[p]] (elastic2D_iso_xPU.jl:108-118, 219-230)
void elf_record(int dtype, long long ld, const void *ux, const void *uz, const swb_sinc_points &lx, const swb_sinc_points &lz, void *traces, long long nt,
                long long it, cudaStream_t st);
// fac[q] = dt^2 / rho[q] on the (w, hgt) cells of a padded plane (the factor update_ux! / update_uz! compute per cell and
// step, elastic2D_iso_xPU.jl:1-37), everything else of the plane stays zero
void elf_dt2_over_rho(int dtype, long long ld, long long w, long long hgt, const void *rho, void *fac, double dt, cudaStream_t st);
// correlate_gradients! (elastic/backends/shared/correlate_gradient_xPU.jl:1-83) on padded planes
struct ElaCorrPadded {
    int dtype, flags, nx, nz, freetop;
    long long ld;
    double dx, dz, dt;
    const void *aux, *auz, *uxo, *uzo, *uxc, *uzc, *uxn, *uzn, *lam, *mu;
    void *g_ri, *g_rj, *g_l, *g_m, *g_mh;
};
void elf_correlate(const ElaCorrPadded &a, cudaStream_t st);

} // namespace swb
