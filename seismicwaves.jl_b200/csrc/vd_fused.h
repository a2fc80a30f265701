// vd_fused.h -- parameter block and layout constants of the fused 2D variable-density step (acou_vd_fused.cu).
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace swb {

constexpr int VDF_TX = 128;          // tile width in cells
constexpr int VDF_TY = 16;           // tile height in cells (large grids)
constexpr int VDF_TY_SMALL = 8;      // ... of grids with fewer than two 16-row tiles per SM (a step is the latency of one tile there)
inline int vdf_ty(long long nx, long long ny) { return ((nx + VDF_TX - 1) / VDF_TX) * ((ny + VDF_TY - 1) / VDF_TY) >= 2 * 148 ? VDF_TY : VDF_TY_SMALL; } // measured: 640^2 (200 tiles) 11.2 us at 16 rows, 9.5 at 8; 1024^2 (512 tiles) 13.7 vs 14.4
constexpr int PAD_GUARD_BEFORE = 4;  // zero rows in front of a padded plane
constexpr int PAD_GUARD_AFTER = 36;  // zero rows behind it (>= tile height + 4)

// pitch (in elements) of a padded plane: at least 4 zero columns after the last cell, multiple of 32 elements
inline long long padded_ld(long long nx) { return ((nx + 4 + 31) / 32) * 32; }
inline size_t padded_plane_elems(long long nx, long long ny) { return (size_t)padded_ld(nx) * (size_t)(ny + PAD_GUARD_BEFORE + PAD_GUARD_AFTER); }
inline size_t padded_origin(long long nx) { return (size_t)padded_ld(nx) * PAD_GUARD_BEFORE; }

constexpr int VDF_SW = VDF_TX + 16; // staged row width of p, vx, m1x, p_it: columns -8 .. TX+7

template <class T>
struct VdFusedParams {
    // TMA descriptors of the padded planes staged through shared memory (whole plane incl. guard rows; boxes: p_in VDF_SW x (TY+6),
    // vx_in / m1x VDF_SW x TY, vy_in / m1y VDF_TX x (TY+3), pc_it VDF_SW x (TY+3))
    alignas(64) CUtensorMap tm[6];
    int nx, ny, halo, ty; // ty: tile height (VDF_TY or VDF_TY_SMALL)
    long long ld;
    int do_v, do_p, adj;
    T inv_dx, inv_dy, inv_dt;
    // fields (pointers to cell (0,0) of padded planes)
    const T *p_in, *vx_in, *vy_in;
    T *p_out, *vx_out, *vy_out;
    const T *m0, *m1x, *m1y;
    // C-PML memory variables, dense reference layout: psi_x (2h, ny), psi_y (nx, 2h), xi_x (2(h+1), ny), xi_y (nx, 2(h+1))
    const T *psi_x_in, *psi_y_in, *xi_x_in, *xi_y_in;
    T *psi_x_out, *psi_y_out, *xi_x_out, *xi_y_out;
    const T *a_x, *b_x, *a_xh, *b_xh, *a_y, *b_y, *a_yh, *b_yh;
    double c4[4];
    // injection (per-tile CSR lists): p_out[cell] += inj_tf[inj_it, idx]; inj_it = 0 disables
    const int *inj_off, *inj_cell, *inj_idx;
    const T *inj_tf;
    long long inj_nt;
    int inj_it;
    // recording: traces[rec_it, idx] = p_in[cell]; rec_it = 0 disables
    const int *rec_off, *rec_cell, *rec_idx;
    T *traces;
    long long rec_nt;
    int rec_it;
    // adjoint-mode correlation
    const T *pc_it, *pc_itm1;
    T *g0, *g1x, *g1y;
    int dbg_all_interior; // timing experiment only (SWB_VD_DEBUG_ALL_INTERIOR=1): wrong results in the strips
    int rev;              // 1: the tile rows after the edge rows are issued in descending order (serpentine sweep, see vd_tile_row)
};

template <class T>
void vd_fused_launch(const VdFusedParams<T> &P, bool fast, cudaStream_t st);

} // namespace swb
