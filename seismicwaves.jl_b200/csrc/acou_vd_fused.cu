// acou_vd_fused.cu -- single-launch 2D acoustic variable-density step for the per-shot engine.
//
// One launch performs, for every tile of the grid:
//     v  <- v - m1 * grad~(p)            update_vx_CPML! / update_vy_CPML!   (acoustic2D_VD_xPU.jl:39-75)
//     p  <- p - m0 * div~(v)             update_p_CPML!                     (acoustic2D_VD_xPU.jl:17-37)
//     p[src] += tf[it, s]                inject_sources!                    (acoustic2D_VD_xPU.jl:1-7)
//     traces[it, r] = p_in[rec]          record_receivers!                  (acoustic2D_VD_xPU.jl:9-15)
// and in adjoint mode additionally the zero-lag correlations
//     grad_m0 -= adjp * (p_it - p_itm1) / dt            (correlate_gradient_xPU.jl:12-21)
//     grad_m1_x/y += adjv_x/y * d_x/y p_it              (acoustic2D_VD_xPU.jl:180-199)
// "v then p then inject" is exactly the reference's adjoint step (acoustic2D_VD_xPU.jl:163-178).  The forward
// step is "p, inject, v, record"; the engine runs it with the same kernel shifted by half a step: launch k
// computes v^k from (v^{k-1}, p^k), records p^k and then p^{k+1} from (p^k, v^k) plus the injection of step k+1.
// The values are those of the reference sequence, operation for operation.
//
// Data layout (engine-owned, not the dense Julia layout): every full-size field lives in a padded plane with
// row pitch ld >= nx + 4 (multiple of 32 elements), GUARD_BEFORE zero rows in front and GUARD_AFTER behind.
// Everything outside the field's index range is kept at zero, so the 4-point stencils see the reference's
// "missing neighbours contribute nothing" rule (src/utils/fdgen.jl:96-131) without a bounds test, and the
// tail of row j doubles as the left halo of row j+1.  Fields are ping-ponged (in -> out) because a tile
// recomputes the v values of its halo from the neighbours' old state.
//
// Work decomposition: a CTA owns a TX x TY tile of cells (TX = 128 = one warp-row of 16-byte chunks).
// Phase 1 stages p (+3 halo) -- and in adjoint mode the stored forward field p_it -- in shared memory with
// cp.async; phase 2 computes v_new on the tile plus the halo the p update needs (3 columns / 3 rows,
// recomputed instead of exchanged) into shared memory and writes the owned part; phase 3 computes p_new,
// applies the tile's injection list, correlates, and writes.  HBM traffic per cell: p, vx, vy read + written,
// m0, m1x, m1y read = 9 values; adjoint + correlation: 17.
#include "common.cuh"
#include "kernels.h"
#include "vd_fused.h"
#include "tma.cuh"
#include <cstdlib>

namespace swb {

namespace {

constexpr int TX = VDF_TX;        // tile width (cells)
constexpr int NCH = TX / 4;       // 16-byte (float) / 32-byte (double) chunks per tile row
constexpr int NTHR = 256;
constexpr int SW = VDF_SW;        // shared row width: columns -8 .. TX+7

template <class T>
struct alignas(16) Chunk {
    T v[4];
};

template <class T>
__device__ __forceinline__ Chunk<T> ldg_chunk(const T *p)
{
    return *reinterpret_cast<const Chunk<T> *>(p);
}
template <class T>
__device__ __forceinline__ void st_chunk(T *p, const Chunk<T> &c)
{
    *reinterpret_cast<Chunk<T> *>(p) = c;
}

// 4-point staggered first derivative, offsets {-1, 0, +1, +2} around q (fdgen.jl:65-135): left-associated sum
// of Float64-literal weights times the samples, then times 1/spacing.
template <class T, class CT>
__device__ __forceinline__ CT fd4(const VdFusedParams<T> &P, T fm1, T f0, T f1, T f2, T inv)
{
    CT acc = (((CT)P.c4[0] * (CT)fm1 + (CT)P.c4[1] * (CT)f0) + (CT)P.c4[2] * (CT)f1) + (CT)P.c4[3] * (CT)f2;
    return acc * (CT)inv;
}

// smem layout (elements of T); every input of the tile is staged by TMA so that a CTA has its whole
// working set (~50 KB in Float32) in flight at once -- the SM keeps HBM busy with 3-4 resident CTAs instead of
// depending on per-thread load/use latency.
template <class T, int TY, bool ADJ>
struct VdSmem {
    static constexpr int P_OFF = 0;                              // p_in   rows -3 .. TY+2, cols -8 .. TX+7
    static constexpr int VX_OFF = P_OFF + (TY + 6) * SW;         // vx     rows  0 .. TY-1, cols -8 .. TX+7 (updated in place)
    static constexpr int M1X_OFF = VX_OFF + TY * SW;             // m1x    same shape
    static constexpr int VY_OFF = M1X_OFF + TY * SW;             // vy     rows -2 .. TY,   cols 0 .. TX-1   (updated in place)
    static constexpr int M1Y_OFF = VY_OFF + (TY + 3) * TX;       // m1y    same shape
    static constexpr int PI_OFF = M1Y_OFF + (TY + 3) * TX;       // p_it   rows -1 .. TY+1, cols -8 .. TX+7 (adjoint only)
    static constexpr int TOTAL = PI_OFF + (ADJ ? (TY + 3) * SW : 0);
    static constexpr int BAR_OFF = TOTAL; // mbarrier of the TMA staging (8 bytes)
    static constexpr size_t BYTES = sizeof(T) * (size_t)TOTAL + 16;
};

// phase 1 of both tile bodies: the elected thread arms the mbarrier and requests every staged array of the tile as one TMA box
// (zeros outside the padded planes); the caller waits on the barrier after its own register loads.  Two parts around the wait for
// the previous step's grid (programmatic dependent launch): the material factors never change, the fields are that step's output.
template <class T, int TY, bool ADJ>
__device__ __forceinline__ void vd_stage_tma_material(const VdFusedParams<T> &P, T *sm, int x0, int y0)
{
    typedef VdSmem<T, TY, ADJ> L;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + L::BAR_OFF);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        unsigned bytes = (unsigned)(((TY + 6) * SW + TY * SW + (TY + 3) * TX) * sizeof(T));
        if (P.do_v)
            bytes += (unsigned)((TY * SW + (TY + 3) * TX) * sizeof(T));
        if (ADJ)
            bytes += (unsigned)((TY + 3) * SW * sizeof(T));
        mbar_expect_tx(bar, bytes);
        const int gb = PAD_GUARD_BEFORE;
        if (P.do_v) {
            tma_load_2d(sm + L::M1X_OFF, &P.tm[2], x0 - 8, gb + y0, bar);
            tma_load_2d(sm + L::M1Y_OFF, &P.tm[4], x0, gb + y0 - 2, bar);
        }
    }
}
template <class T, int TY, bool ADJ>
__device__ __forceinline__ void vd_stage_tma_fields(const VdFusedParams<T> &P, T *sm, int x0, int y0)
{
    typedef VdSmem<T, TY, ADJ> L;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(sm + L::BAR_OFF);
    if (threadIdx.x == 0) {
        const int gb = PAD_GUARD_BEFORE;
        tma_load_2d(sm + L::P_OFF, &P.tm[0], x0 - 8, gb + y0 - 3, bar);
        tma_load_2d(sm + L::VX_OFF, &P.tm[1], x0 - 8, gb + y0, bar);
        tma_load_2d(sm + L::VY_OFF, &P.tm[3], x0, gb + y0 - 2, bar);
        if (ADJ)
            tma_load_2d(sm + L::PI_OFF, &P.tm[5], x0 - 8, gb + y0 - 1, bar);
    }
}
template <class T, int TY, bool ADJ>
__device__ __forceinline__ void vd_stage_wait(T *sm)
{
    __syncthreads(); // (the mbarrier's initialisation becomes visible to the waiting threads)
    mbar_wait(reinterpret_cast<unsigned long long *>(sm + VdSmem<T, TY, ADJ>::BAR_OFF), 0);
}

// ---- chunk bodies of the edge tiles.  SPC = true: the reference's range tests and C-PML per cell; SPC = false: the chunk lies
//      inside every update range and outside every strip (plain expressions).  vd_tile runs the plain chunks in place and
//      enumerates the special ones compactly in a second pass, so that a strip a few chunks wide does not drag whole warps
//      through the per-cell code.
template <class T>
struct VdCtx {
    const VdFusedParams<T> &P;
    T *sp, *svx, *sm1x, *svy, *sm1y, *spi;
    int x0, y0, nx, ny, h;
    long long ld;
};

// vx_new of chunk column c (-1 .. NCH) of tile row r
template <class T, class CT, bool ADJ, int TY, bool SPC>
__device__ __forceinline__ void vd_edge_vx(const VdCtx<T> &C, const int r, const int c)
{
    const VdFusedParams<T> &P = C.P;
    T *const sp = C.sp, *const svx = C.svx, *const sm1x = C.sm1x, *const svy = C.svy, *const sm1y = C.sm1y, *const spi = C.spi;
    const int x0 = C.x0, y0 = C.y0, nx = C.nx, ny = C.ny, h = C.h;
    const long long ld = C.ld;
    (void)sp, (void)svx, (void)sm1x, (void)svy, (void)sm1y, (void)spi, (void)nx, (void)ny, (void)h;
    const int gx = x0 + 4 * c, gy = y0 + r; // 0-based global column of the chunk's first cell / row
    if (gx >= ld)
        return;
    const long long q = (long long)gy * ld + gx;
    T *vs = svx + r * SW + 4 * c;
    const Chunk<T> vin = ldg_chunk(vs);
    const Chunk<T> m1 = ldg_chunk(sm1x + r * SW + 4 * c);
    const T *ps = sp + r * SW + 4 * c; // p at column 4c
    const Chunk<T> pa = ldg_chunk(ps - 4), pb = ldg_chunk(ps), pc = ldg_chunk(ps + 4);
    const T w[7] = {pa.v[3], pb.v[0], pb.v[1], pb.v[2], pb.v[3], pc.v[0], pc.v[1]};
    const bool owned = c >= 0 && c < NCH && gy < ny;
    Chunk<T> out = vin, g1 = {};
    T wi[7];
    if (ADJ && owned) {
        g1 = ldg_chunk(P.g1x + q);
        const T *pis = spi + r * SW + 4 * c;
        const Chunk<T> ia = ldg_chunk(pis - 4), ib = ldg_chunk(pis), ic = ldg_chunk(pis + 4);
        wi[0] = ia.v[3], wi[1] = ib.v[0], wi[2] = ib.v[1], wi[3] = ib.v[2], wi[4] = ib.v[3], wi[5] = ic.v[0], wi[6] = ic.v[1];
    }
    // C-PML values of the chunk's cells are fetched together before use (one dependent global load per cell otherwise)
    T ca[4], cb[4], cs[4];
    int cq[4];
    #pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int I = gx + e + 1, J = gy + 1;
        const bool in = SPC && (I >= 1 && I <= nx - 1 && J <= ny) && (I <= h || I >= nx - h);
        const int ii = I <= h ? I : I - nx + 2 * h + 1;
        cq[e] = in ? (J - 1) * (2 * h) + (ii - 1) : -1;
        ca[e] = in ? P.a_xh[ii - 1] : (T)0;
        cb[e] = in ? P.b_xh[ii - 1] : (T)0;
        cs[e] = in ? P.psi_x_in[cq[e]] : (T)0;
    }
    #pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int I = gx + e + 1, J = gy + 1; // 1-based reference indices
        if (SPC && !(I >= 1 && I <= nx - 1 && J <= ny))
            continue; // outside update_vx_CPML!'s range: stays as it is (zero)
        CT D = fd4<T, CT>(P, w[e], w[e + 1], w[e + 2], w[e + 3], P.inv_dx);
        if (SPC && cq[e] >= 0) {
            T sn;
            D = cpml_apply<T, CT>(D, ca[e], cb[e], cs[e], sn);
            if (owned)
                P.psi_x_out[cq[e]] = sn;
        }
        out.v[e] = (T)((CT)vin.v[e] - (CT)m1.v[e] * D);
        if (ADJ && owned) {
            const CT Dc = fd4<T, CT>(P, wi[e], wi[e + 1], wi[e + 2], wi[e + 3], P.inv_dx);
            g1.v[e] = (T)((CT)g1.v[e] + (CT)out.v[e] * Dc);
        }
    }
    st_chunk(vs, out);
    if (owned) {
        st_chunk(P.vx_out + q, out);
        if (ADJ)
            st_chunk(P.g1x + q, g1);
    }
}

// vy_new of chunk column c of staged row rr (tile row rr - 2)
template <class T, class CT, bool ADJ, int TY, bool SPC>
__device__ __forceinline__ void vd_edge_vy(const VdCtx<T> &C, const int rr, const int c)
{
    const VdFusedParams<T> &P = C.P;
    T *const sp = C.sp, *const svx = C.svx, *const sm1x = C.sm1x, *const svy = C.svy, *const sm1y = C.sm1y, *const spi = C.spi;
    const int x0 = C.x0, y0 = C.y0, nx = C.nx, ny = C.ny, h = C.h;
    const long long ld = C.ld;
    (void)sp, (void)svx, (void)sm1x, (void)svy, (void)sm1y, (void)spi, (void)nx, (void)ny, (void)h;
    const int r = rr - 2;
    const int gx = x0 + 4 * c, gy = y0 + r;
    if (gx >= ld)
        return;
    const long long q = (long long)gy * ld + gx;
    T *vs = svy + rr * TX + 4 * c;
    const Chunk<T> vin = ldg_chunk(vs);
    const Chunk<T> m1 = ldg_chunk(sm1y + rr * TX + 4 * c);
    const T *ps = sp + r * SW + 4 * c;
    const Chunk<T> pa = ldg_chunk(ps - SW), pb = ldg_chunk(ps), pc = ldg_chunk(ps + SW), pd = ldg_chunk(ps + 2 * SW);
    const bool owned = r >= 0 && r < TY && gy < ny;
    Chunk<T> out = vin, g1 = {}, ia = {}, ib = {}, ic = {}, id = {};
    if (ADJ && owned) {
        g1 = ldg_chunk(P.g1y + q);
        const T *pis = spi + r * SW + 4 * c;
        ia = ldg_chunk(pis - SW), ib = ldg_chunk(pis), ic = ldg_chunk(pis + SW), id = ldg_chunk(pis + 2 * SW);
    }
    T ca = (T)0, cb = (T)0, cs[4];
    int cq[4];
    {
        const int J = gy + 1;
        const bool rowin = SPC && J >= 1 && J <= ny - 1 && (J <= h || J >= ny - h);
        const int jj = J <= h ? J : J - ny + 2 * h + 1;
        if (rowin) {
            ca = P.a_yh[jj - 1];
            cb = P.b_yh[jj - 1];
        }
    #pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int I = gx + e + 1;
            const bool in = rowin && I <= nx;
            cq[e] = in ? (jj - 1) * nx + (I - 1) : -1;
            cs[e] = in ? P.psi_y_in[cq[e]] : (T)0;
        }
    }
    #pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int I = gx + e + 1, J = gy + 1;
        if (SPC && !(I <= nx && J >= 1 && J <= ny - 1))
            continue;
        CT D = fd4<T, CT>(P, pa.v[e], pb.v[e], pc.v[e], pd.v[e], P.inv_dy);
        if (SPC && cq[e] >= 0) {
            T sn;
            D = cpml_apply<T, CT>(D, ca, cb, cs[e], sn);
            if (owned)
                P.psi_y_out[cq[e]] = sn;
        }
        out.v[e] = (T)((CT)vin.v[e] - (CT)m1.v[e] * D);
        if (ADJ && owned) {
            const CT Dc = fd4<T, CT>(P, ia.v[e], ib.v[e], ic.v[e], id.v[e], P.inv_dy);
            g1.v[e] = (T)((CT)g1.v[e] + (CT)out.v[e] * Dc);
        }
    }
    st_chunk(vs, out);
    if (owned) {
        st_chunk(P.vy_out + q, out);
        if (ADJ)
            st_chunk(P.g1y + q, g1);
    }
}

// p_new, injection and m0 correlation of chunk column c of tile row r
template <class T, class CT, bool ADJ, int TY, bool SPC>
__device__ __forceinline__ void vd_edge_p(const VdCtx<T> &C, const int r, const int c, const int ie0, const int ie1)
{
    const VdFusedParams<T> &P = C.P;
    T *const sp = C.sp, *const svx = C.svx, *const sm1x = C.sm1x, *const svy = C.svy, *const sm1y = C.sm1y, *const spi = C.spi;
    const int x0 = C.x0, y0 = C.y0, nx = C.nx, ny = C.ny, h = C.h;
    const long long ld = C.ld;
    (void)sp, (void)svx, (void)sm1x, (void)svy, (void)sm1y, (void)spi, (void)nx, (void)ny, (void)h;
    const int idx = r * NCH + c;
    const int gx = x0 + 4 * c, gy = y0 + r;
    if (gx >= ld || gy >= ny)
        return;
    const long long q = (long long)gy * ld + gx;
    Chunk<T> g0 = {}, pm1 = {};
    if (ADJ) {
        g0 = ldg_chunk(P.g0 + q);
        pm1 = ldg_chunk(P.pc_itm1 + q);
    }
    const Chunk<T> m0 = ldg_chunk(P.m0 + q);
    const T *vxs = svx + r * SW + 4 * c;
    const Chunk<T> xa = ldg_chunk(vxs - 4), xb = ldg_chunk(vxs), xc = ldg_chunk(vxs + 4);
    const T wx[7] = {xa.v[2], xa.v[3], xb.v[0], xb.v[1], xb.v[2], xb.v[3], xc.v[0]}; // vx at columns 4c-2 .. 4c+4
    const T *vys = svy + (r + 2) * TX + 4 * c;
    const Chunk<T> ya = ldg_chunk(vys - 2 * TX), yb = ldg_chunk(vys - TX), yc = ldg_chunk(vys), yd = ldg_chunk(vys + TX);
    const Chunk<T> pin = ldg_chunk(sp + r * SW + 4 * c);
    Chunk<T> out = pin;
    T xa_[4], xb_[4], xs_[4], ya_ = (T)0, yb_ = (T)0, ys_[4];
    int xq[4], yq[4];
    {
        const int J = gy + 1;
        const bool rowok = SPC && J >= 2 && J <= ny - 1;
        const bool rowin = rowok && (J <= h + 1 || J >= ny - h);
        const int jj = J <= h + 1 ? J : J - ny + 2 * h + 2;
        if (rowin) {
            ya_ = P.a_y[jj - 1];
            yb_ = P.b_y[jj - 1];
        }
    #pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int I = gx + e + 1;
            const bool ok = rowok && I >= 2 && I <= nx - 1;
            const bool inx = ok && (I <= h + 1 || I >= nx - h);
            const int ii = I <= h + 1 ? I : I - nx + 2 * h + 2;
            xq[e] = inx ? (J - 1) * (2 * (h + 1)) + (ii - 1) : -1;
            xa_[e] = inx ? P.a_x[ii - 1] : (T)0;
            xb_[e] = inx ? P.b_x[ii - 1] : (T)0;
            xs_[e] = inx ? P.xi_x_in[xq[e]] : (T)0;
            const bool iny = ok && rowin;
            yq[e] = iny ? (jj - 1) * nx + (I - 1) : -1;
            ys_[e] = iny ? P.xi_y_in[yq[e]] : (T)0;
        }
    }
    #pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int I = gx + e + 1, J = gy + 1;
        if (SPC && !(I >= 2 && I <= nx - 1 && J >= 2 && J <= ny - 1))
            continue; // update_p_CPML! touches interior cells only
        // d vx / dx at I-1 (backward staggered): vx[I-2 .. I+1]
        CT Dx = fd4<T, CT>(P, wx[e], wx[e + 1], wx[e + 2], wx[e + 3], P.inv_dx);
        if (SPC && xq[e] >= 0) {
            T sn;
            Dx = cpml_apply<T, CT>(Dx, xa_[e], xb_[e], xs_[e], sn);
            P.xi_x_out[xq[e]] = sn;
        }
        CT Dy = fd4<T, CT>(P, ya.v[e], yb.v[e], yc.v[e], yd.v[e], P.inv_dy);
        if (SPC && yq[e] >= 0) {
            T sn;
            Dy = cpml_apply<T, CT>(Dy, ya_, yb_, ys_[e], sn);
            P.xi_y_out[yq[e]] = sn;
        }
        out.v[e] = (T)((CT)pin.v[e] - (CT)m0.v[e] * (Dx + Dy));
    }
    // inject_sources!: entries of this tile in source-index order (deterministic for coincident sources)
    for (int e = ie0; e < ie1; ++e) {
        const int cell = P.inj_cell[e];
        if ((cell >> 2) == idx)
            out.v[cell & 3] = out.v[cell & 3] + P.inj_tf[(size_t)P.inj_idx[e] * P.inj_nt + (P.inj_it - 1)];
    }
    if (ADJ) { // grad_m0 = grad_m0 - adjp * (p_it - p_itm1) * (1/dt), all in T (correlate_gradient_xPU.jl:12-21)
        const Chunk<T> pit = ldg_chunk(spi + r * SW + 4 * c);
    #pragma unroll
        for (int e = 0; e < 4; ++e) {
            const T d = pit.v[e] - pm1.v[e];
            const T t = out.v[e] * d;
            g0.v[e] = g0.v[e] - t * P.inv_dt;
        }
        st_chunk(P.g0 + q, g0);
    }
    st_chunk(P.p_out + q, out);
}

// Tile rows are issued bottom strip rows first, then from the top: the rows that touch the bottom C-PML strip run the slow
// generic body and would otherwise form a tail of long CTAs at the end of the grid.
// Serpentine sweep (rev = 1 on every other launch): the remaining rows are issued in descending order, so a launch starts where
// the previous one ended -- on the rows whose fields (written or read a moment ago) and material factors still sit in the 126 MB L2.
template <int TY>
__device__ __forceinline__ int vd_tile_row(int halo, int rev)
{
    const int nty = (int)gridDim.y;
    const int nedge = min(nty, (halo + 6 + TY + TY - 1) / TY);
    const int b = (int)blockIdx.y;
    if (b < nedge)
        return nty - nedge + b;
    return rev ? nty - 1 - b : b - nedge;
}

template <class T, class CT, bool ADJ, int TY, bool edge>
__device__ __forceinline__ void vd_tile(const VdFusedParams<T> &P, unsigned char *smem_raw)
{
    typedef VdSmem<T, TY, ADJ> L;
    T *sm = reinterpret_cast<T *>(smem_raw);
    // sp / svx / sm1x / spi point at (row 0, column 0) of the tile; svy / sm1y at (row -2, column 0)
    T *sp = sm + L::P_OFF + 3 * SW + 8, *svx = sm + L::VX_OFF + 8, *sm1x = sm + L::M1X_OFF + 8;
    T *svy = sm + L::VY_OFF, *sm1y = sm + L::M1Y_OFF, *spi = sm + L::PI_OFF + SW + 8;

    const int tid = threadIdx.x;
    const int trow = vd_tile_row<TY>(P.halo, P.rev);
    const int x0 = blockIdx.x * TX, y0 = trow * TY;
    const int nx = P.nx, ny = P.ny, h = P.halo;
    const long long ld = P.ld;
    const int tile = trow * gridDim.x + blockIdx.x;

    // ---- phase 1: stage every input of the tile in shared memory (TMA) ---------------------------------------
    vd_stage_tma_material<T, TY, ADJ>(P, sm, x0, y0);
    pdl_wait(); // the previous step's grid has completed: its fields are visible
    vd_stage_tma_fields<T, TY, ADJ>(P, sm, x0, y0);
    vd_stage_wait<T, TY, ADJ>(sm);

    // record_receivers!: traces[it, r] = p_in[rec]  (p_in is the pressure after the reference's step `rec_it`)
    if (P.rec_it > 0) {
        const int e0 = P.rec_off[tile], e1 = P.rec_off[tile + 1];
        for (int e = e0 + tid; e < e1; e += NTHR) {
            const int cell = P.rec_cell[e];
            const int r = cell / TX, c = cell - r * TX;
            P.traces[(size_t)P.rec_idx[e] * P.rec_nt + (P.rec_it - 1)] = sp[r * SW + c];
        }
    }

    if (P.do_v) {
        const VdCtx<T> C{P, sp, svx, sm1x, svy, sm1y, spi, x0, y0, nx, ny, h, ld};
        // ---- phase 2a: vx_new on columns -4 .. TX+3 (chunks -1 .. NCH), rows 0 .. TY-1 --------------------
        {
            // plain chunk columns [ca, cb): all four cells inside 1 .. nx-1 and outside the x strips; rows below the grid do nothing
            const int nr = min(TY, ny - y0);
            int ca = NCH + 1, cb = NCH + 1;
            for (int c = NCH; c >= -1; --c) {
                const int gx = x0 + 4 * c;
                const bool plain = gx >= max(h, 0) && gx + 4 <= nx - h - 1;
                if (plain)
                    ca = c;
                else if (ca == NCH + 1)
                    cb = c;
            }
            if (ca == NCH + 1)
                ca = cb = -1; // no plain column
            for (int idx = tid; idx < nr * (NCH + 2); idx += NTHR) {
                const int r = idx / (NCH + 2), c = idx - r * (NCH + 2) - 1;
                if (c >= ca && c < cb)
                    vd_edge_vx<T, CT, ADJ, TY, false>(C, r, c);
            }
            const int ncs = (ca + 1) + (NCH + 1 - cb);
            for (int idx = tid; idx < nr * ncs; idx += NTHR) {
                const int r = idx / ncs, j = idx - r * ncs;
                vd_edge_vx<T, CT, ADJ, TY, true>(C, r, j < ca + 1 ? j - 1 : cb + (j - (ca + 1)));
            }
        }

        // ---- phase 2b: vy_new on rows -2 .. TY, columns 0 .. TX-1 -----------------------------------------
        {
            // plain rows [ra, rb) of the staged rows 0 .. TY+2 (tile row rr - 2): 1-based J = y0 + rr - 1 inside (h, ny - h); plain
            // columns [0, cb): all four cells <= nx
            const int ra = min(max(h + 2 - y0, 0), TY + 3), rb = min(max(ny - h - y0 + 1, ra), TY + 3);
            const int cb = min(max((nx - x0) / 4, 0), NCH);
            for (int idx = tid; idx < (TY + 3) * NCH; idx += NTHR) {
                const int rr = idx / NCH, c = idx - rr * NCH;
                if (rr >= ra && rr < rb && c < cb)
                    vd_edge_vy<T, CT, ADJ, TY, false>(C, rr, c);
            }
            const int nrs = ra + (TY + 3 - rb), ncs = NCH - cb;
            const int n1 = nrs * NCH, n2 = (rb - ra) * ncs;
            for (int idx = tid; idx < n1 + n2; idx += NTHR) {
                int rr, c;
                if (idx < n1) {
                    const int i = idx / NCH;
                    c = idx - i * NCH;
                    rr = i < ra ? i : rb + (i - ra);
                } else {
                    const int t2 = idx - n1, i = t2 / ncs;
                    rr = ra + i;
                    c = cb + (t2 - i * ncs);
                }
                vd_edge_vy<T, CT, ADJ, TY, true>(C, rr, c);
            }
        }
        if (!P.do_p)
            return;
        __syncthreads();
    }

    // ---- phase 3: p_new on the tile, injection, m0 correlation -------------------------------------------
    const int ie0 = P.inj_it > 0 ? P.inj_off[tile] : 0, ie1 = P.inj_it > 0 ? P.inj_off[tile + 1] : 0;
    {
        const VdCtx<T> C{P, sp, svx, sm1x, svy, sm1y, spi, x0, y0, nx, ny, h, ld};
        // plain rows [ra, rb): 1-based J = y0 + r + 1 in [max(2, h+2), ny-h-1]; plain chunk columns [ca, cb): I likewise
        const int lo = max(2, h + 2);
        const int ra = min(max(lo - y0 - 1, 0), TY), rb = min(max(ny - h - 1 - y0, ra), TY);
        int ca = NCH, cb = NCH;
        for (int c = NCH - 1; c >= 0; --c) {
            const int gx = x0 + 4 * c;
            const bool plain = gx + 1 >= lo && gx + 4 <= nx - h - 1;
            if (plain)
                ca = c;
            else if (ca == NCH)
                cb = c;
        }
        if (ca == NCH)
            ca = cb = 0; // no plain column
        for (int idx = tid; idx < TY * NCH; idx += NTHR) {
            const int r = idx / NCH, c = idx - r * NCH;
            if (r >= ra && r < rb && c >= ca && c < cb)
                vd_edge_p<T, CT, ADJ, TY, false>(C, r, c, ie0, ie1);
        }
        const int nrs = ra + (TY - rb), ncs = ca + (NCH - cb);
        const int n1 = nrs * NCH, n2 = (rb - ra) * ncs;
        for (int idx = tid; idx < n1 + n2; idx += NTHR) {
            int r, c;
            if (idx < n1) {
                const int i = idx / NCH;
                c = idx - i * NCH;
                r = i < ra ? i : rb + (i - ra);
            } else {
                const int t2 = idx - n1, i = t2 / ncs, j = t2 - i * ncs;
                r = ra + i;
                c = j < ca ? j : cb + (j - ca);
            }
            vd_edge_p<T, CT, ADJ, TY, true>(C, r, c, ie0, ie1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Interior tiles (no C-PML strip, no grid edge inside the tile or its halo): fixed thread -> cell mapping.
// lane = 16-byte chunk column, warp w owns rows w, w+8, ... ; every loop below has a compile-time trip
// count, so the index arithmetic of the generic body (divisions, range tests) disappears.
// ------------------------------------------------------------------------------------------------
template <class T, class CT>
struct Fd4 {
    // reference-faithful form: left-associated sum of weight * sample in CT, then * 1/spacing
    static __device__ __forceinline__ CT d(const VdFusedParams<T> &P, T fm1, T f0, T f1, T f2, T inv) { return fd4<T, CT>(P, fm1, f0, f1, f2, inv); }
    static __device__ __forceinline__ T upd(T a, T m, CT D) { return (T)((CT)a - (CT)m * D); }
    static __device__ __forceinline__ T acc(T g, T a, CT D) { return (T)((CT)g + (CT)a * D); }
};
template <>
struct Fd4<float, float> {
    // SWB_FLAG_FAST_F32: (9/8 (f1 - f0) - 1/24 (f2 - fm1)) / spacing with fused multiply-adds
    static __device__ __forceinline__ float d(const VdFusedParams<float> &, float fm1, float f0, float f1, float f2, float inv)
    {
        return __fmaf_rn(1.125f, f1 - f0, (-1.0f / 24.0f) * (f2 - fm1)) * inv;
    }
    static __device__ __forceinline__ float upd(float a, float m, float D) { return __fmaf_rn(-m, D, a); }
    static __device__ __forceinline__ float acc(float g, float a, float D) { return __fmaf_rn(a, D, g); }
};

template <class T, class CT, bool ADJ, int TY>
__device__ __forceinline__ void vd_tile_interior(const VdFusedParams<T> &P, unsigned char *smem_raw)
{
    typedef VdSmem<T, TY, ADJ> L;
    typedef Fd4<T, CT> A;
    constexpr int NW = NTHR / 32;
    static_assert(NCH == 32, "one lane per chunk column");
    T *sm = reinterpret_cast<T *>(smem_raw);
    T *sp = sm + L::P_OFF + 3 * SW + 8, *svx = sm + L::VX_OFF + 8, *sm1x = sm + L::M1X_OFF + 8;
    T *svy = sm + L::VY_OFF, *sm1y = sm + L::M1Y_OFF, *spi = sm + L::PI_OFF + SW + 8;

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int trow = vd_tile_row<TY>(P.halo, P.rev);
    const int x0 = blockIdx.x * TX, y0 = trow * TY;
    const long long ld = P.ld;
    const int tile = trow * gridDim.x + blockIdx.x;
    const long long o = (long long)y0 * ld + x0 + 4 * lane; // this lane's chunk in row 0 of the tile

    // ---- phase 1: everything the tile needs goes into shared memory (TMA), m0 straight to registers ---------
    vd_stage_tma_material<T, TY, ADJ>(P, sm, x0, y0);
    Chunk<T> m0r[TY / NW]; // m0 is used once per cell: in flight together with the TMA traffic
    if (P.do_p) {
#pragma unroll
        for (int k = 0; k < TY / NW; ++k)
            m0r[k] = ldg_chunk(P.m0 + o + (long long)(w + NW * k) * ld);
    }
    pdl_wait(); // the previous step's grid has completed: its fields are visible (the material requests above overlap its tail)
    vd_stage_tma_fields<T, TY, ADJ>(P, sm, x0, y0);
    vd_stage_wait<T, TY, ADJ>(sm);

    if (P.rec_it > 0) { // record_receivers!
        const int e0 = P.rec_off[tile], e1 = P.rec_off[tile + 1];
        for (int e = e0 + (int)threadIdx.x; e < e1; e += NTHR) {
            const int cell = P.rec_cell[e];
            const int r = cell / TX, c = cell - r * TX;
            P.traces[(size_t)P.rec_idx[e] * P.rec_nt + (P.rec_it - 1)] = sp[r * SW + c];
        }
    }

    if (P.do_v) {
        // ---- phase 2a: vx_new, own chunks ----------------------------------------------------------------
#pragma unroll
        for (int k = 0; k < TY / NW; ++k) {
            const int r = w + NW * k;
            T *vs = svx + r * SW + 4 * lane;
            const Chunk<T> vin = ldg_chunk(vs), m1 = ldg_chunk(sm1x + r * SW + 4 * lane);
            const T *ps = sp + r * SW + 4 * lane;
            const Chunk<T> pa = ldg_chunk(ps - 4), pb = ldg_chunk(ps), pc = ldg_chunk(ps + 4);
            const long long q = o + (long long)r * ld;
            Chunk<T> out;
            const CT D0 = A::d(P, pa.v[3], pb.v[0], pb.v[1], pb.v[2], P.inv_dx), D1 = A::d(P, pb.v[0], pb.v[1], pb.v[2], pb.v[3], P.inv_dx);
            const CT D2 = A::d(P, pb.v[1], pb.v[2], pb.v[3], pc.v[0], P.inv_dx), D3 = A::d(P, pb.v[2], pb.v[3], pc.v[0], pc.v[1], P.inv_dx);
            out.v[0] = A::upd(vin.v[0], m1.v[0], D0), out.v[1] = A::upd(vin.v[1], m1.v[1], D1);
            out.v[2] = A::upd(vin.v[2], m1.v[2], D2), out.v[3] = A::upd(vin.v[3], m1.v[3], D3);
            st_chunk(vs, out);
            st_chunk(P.vx_out + q, out);
            if (ADJ) {
                Chunk<T> g1 = ldg_chunk(P.g1x + q);
                const T *is = spi + r * SW + 4 * lane;
                const Chunk<T> ia = ldg_chunk(is - 4), ib = ldg_chunk(is), ic = ldg_chunk(is + 4);
                g1.v[0] = A::acc(g1.v[0], out.v[0], A::d(P, ia.v[3], ib.v[0], ib.v[1], ib.v[2], P.inv_dx));
                g1.v[1] = A::acc(g1.v[1], out.v[1], A::d(P, ib.v[0], ib.v[1], ib.v[2], ib.v[3], P.inv_dx));
                g1.v[2] = A::acc(g1.v[2], out.v[2], A::d(P, ib.v[1], ib.v[2], ib.v[3], ic.v[0], P.inv_dx));
                g1.v[3] = A::acc(g1.v[3], out.v[3], A::d(P, ib.v[2], ib.v[3], ic.v[0], ic.v[1], P.inv_dx));
                st_chunk(P.g1x + q, g1);
            }
        }
        // halo columns -2, -1 and TX, TX+1 (recomputed, not stored): 2 columns x 2 sides x TY rows = 4 TY items
        if (threadIdx.x < 4 * TY) {
            const int r = threadIdx.x >> 2, j = threadIdx.x & 3;
            const int col = j < 2 ? j - 2 : TX + (j - 2);
            const T *ps = sp + r * SW + col;
            const CT D = A::d(P, ps[-1], ps[0], ps[1], ps[2], P.inv_dx);
            svx[r * SW + col] = A::upd(svx[r * SW + col], sm1x[r * SW + col], D);
        }
        // ---- phase 2b: vy_new on rows -2 .. TY -----------------------------------------------------------
#pragma unroll
        for (int k = 0; k < (TY + 3 + NW - 1) / NW; ++k) {
            const int rr = w + NW * k;
            if (rr < TY + 3) {
                const int r = rr - 2;
                T *vs = svy + rr * TX + 4 * lane;
                const Chunk<T> vin = ldg_chunk(vs), m1 = ldg_chunk(sm1y + rr * TX + 4 * lane);
                const T *ps = sp + r * SW + 4 * lane;
                const Chunk<T> pa = ldg_chunk(ps - SW), pb = ldg_chunk(ps), pc = ldg_chunk(ps + SW), pd = ldg_chunk(ps + 2 * SW);
                Chunk<T> out;
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    out.v[e] = A::upd(vin.v[e], m1.v[e], A::d(P, pa.v[e], pb.v[e], pc.v[e], pd.v[e], P.inv_dy));
                st_chunk(vs, out);
                if (r >= 0 && r < TY) {
                    const long long q = o + (long long)r * ld;
                    st_chunk(P.vy_out + q, out);
                    if (ADJ) {
                        Chunk<T> g1 = ldg_chunk(P.g1y + q);
                        const T *is = spi + r * SW + 4 * lane;
                        const Chunk<T> ia = ldg_chunk(is - SW), ib = ldg_chunk(is), ic = ldg_chunk(is + SW), id = ldg_chunk(is + 2 * SW);
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            g1.v[e] = A::acc(g1.v[e], out.v[e], A::d(P, ia.v[e], ib.v[e], ic.v[e], id.v[e], P.inv_dy));
                        st_chunk(P.g1y + q, g1);
                    }
                }
            }
        }
        if (!P.do_p)
            return;
        __syncthreads();
    }

    // ---- phase 3: p_new, injection, m0 correlation -------------------------------------------------------
    const int ie0 = P.inj_it > 0 ? P.inj_off[tile] : 0, ie1 = P.inj_it > 0 ? P.inj_off[tile + 1] : 0;
#pragma unroll
    for (int k = 0; k < TY / NW; ++k) {
        const int r = w + NW * k;
        const long long q = o + (long long)r * ld;
        Chunk<T> g0 = {}, pm1 = {};
        if (ADJ) {
            g0 = ldg_chunk(P.g0 + q);
            pm1 = ldg_chunk(P.pc_itm1 + q);
        }
        const Chunk<T> m0 = m0r[k];
        const T *vxs = svx + r * SW + 4 * lane;
        const Chunk<T> xa = ldg_chunk(vxs - 4), xb = ldg_chunk(vxs), xc = ldg_chunk(vxs + 4);
        const T *vys = svy + (r + 2) * TX + 4 * lane;
        const Chunk<T> ya = ldg_chunk(vys - 2 * TX), yb = ldg_chunk(vys - TX), yc = ldg_chunk(vys), yd = ldg_chunk(vys + TX);
        const Chunk<T> pin = ldg_chunk(sp + r * SW + 4 * lane);
        CT Dx[4];
        Dx[0] = A::d(P, xa.v[2], xa.v[3], xb.v[0], xb.v[1], P.inv_dx), Dx[1] = A::d(P, xa.v[3], xb.v[0], xb.v[1], xb.v[2], P.inv_dx);
        Dx[2] = A::d(P, xb.v[0], xb.v[1], xb.v[2], xb.v[3], P.inv_dx), Dx[3] = A::d(P, xb.v[1], xb.v[2], xb.v[3], xc.v[0], P.inv_dx);
        Chunk<T> out;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            out.v[e] = A::upd(pin.v[e], m0.v[e], Dx[e] + A::d(P, ya.v[e], yb.v[e], yc.v[e], yd.v[e], P.inv_dy));
        for (int e = ie0; e < ie1; ++e) { // inject_sources!, this tile's entries in source-index order
            const int cell = P.inj_cell[e];
            if ((cell >> 2) == r * NCH + lane) {
                const T add = P.inj_tf[(size_t)P.inj_idx[e] * P.inj_nt + (P.inj_it - 1)];
                const int sub = cell & 3;
                out.v[0] = sub == 0 ? out.v[0] + add : out.v[0];
                out.v[1] = sub == 1 ? out.v[1] + add : out.v[1];
                out.v[2] = sub == 2 ? out.v[2] + add : out.v[2];
                out.v[3] = sub == 3 ? out.v[3] + add : out.v[3];
            }
        }
        if (ADJ) {
            const Chunk<T> pit = ldg_chunk(spi + r * SW + 4 * lane);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const T d = pit.v[e] - pm1.v[e];
                const T t = out.v[e] * d;
                g0.v[e] = g0.v[e] - t * P.inv_dt;
            }
            st_chunk(P.g0 + q, g0);
        }
        st_chunk(P.p_out + q, out);
    }
}

template <class T, class CT, bool ADJ, int TY>
__global__ void __launch_bounds__(NTHR, sizeof(T) == 4 ? (ADJ ? 3 : 4) : 1) vd_fused_kernel(const __grid_constant__ VdFusedParams<T> P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    pdl_trigger(); // the next step's CTAs may take the SM slots this grid frees while it drains
    const int x0 = blockIdx.x * TX, y0 = vd_tile_row<TY>(P.halo, P.rev) * TY, h = P.halo;
    // block-uniform: does the tile (with the halo it recomputes) touch a C-PML strip or the grid edge?
    const bool edge = (x0 - 8 <= h + 2) || (x0 + TX + 8 >= P.nx - h - 2) || (y0 - 4 <= h + 2) || (y0 + TY + 4 >= P.ny - h - 2);
    if (edge && !P.dbg_all_interior)
        vd_tile<T, CT, ADJ, TY, true>(P, smem_raw);
    else
        vd_tile_interior<T, CT, ADJ, TY>(P, smem_raw);
}

template <class T, class CT, bool ADJ, int TY>
void launch_one(const VdFusedParams<T> &P, cudaStream_t st)
{
    const size_t smem = VdSmem<T, TY, ADJ>::BYTES;
    auto kern = vd_fused_kernel<T, CT, ADJ, TY>;
    static bool configured = false; // per instantiation
    if (!configured) {
        SWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    dim3 grd(cdiv(P.nx, TX), cdiv(P.ny, TY), 1);
    if (pdl_enabled()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grd, cfg.blockDim = dim3(NTHR, 1, 1), cfg.dynamicSmemBytes = smem, cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        SWB_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    } else
        kern<<<grd, NTHR, smem, st>>>(P);
    check_launch("vd_fused");
    count_launch();
}

} // namespace

template <class T, int TY>
static void vd_fused_launch_ty(const VdFusedParams<T> &P, bool fast, cudaStream_t st)
{
    if (sizeof(T) == 8 || !fast) {
        if (P.adj)
            launch_one<T, double, true, TY>(P, st);
        else
            launch_one<T, double, false, TY>(P, st);
    } else {
        if (P.adj)
            launch_one<T, T, true, TY>(P, st);
        else
            launch_one<T, T, false, TY>(P, st);
    }
}

template <class T>
void vd_fused_launch(const VdFusedParams<T> &P0, bool fast, cudaStream_t st)
{
    static const int dbg = [] { const char *e = std::getenv("SWB_VD_DEBUG_ALL_INTERIOR"); return e ? std::atoi(e) : 0; }();
    VdFusedParams<T> P = P0;
    P.dbg_all_interior = dbg;
    SWB_REQUIRE(P.ty == VDF_TY || P.ty == VDF_TY_SMALL, "fused VD step: unsupported tile height");
    if (P.ty == VDF_TY)
        vd_fused_launch_ty<T, VDF_TY>(P, fast, st);
    else
        vd_fused_launch_ty<T, VDF_TY_SMALL>(P, fast, st);
}

template void vd_fused_launch<float>(const VdFusedParams<float> &, bool, cudaStream_t);
template void vd_fused_launch<double>(const VdFusedParams<double> &, bool, cudaStream_t);

} // namespace swb
