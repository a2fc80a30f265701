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

namespace swb {

namespace {

constexpr int TX = VDF_TX;        // tile width (cells)
constexpr int NCH = TX / 4;       // 16-byte (float) / 32-byte (double) chunks per tile row
constexpr int NTHR = 256;
constexpr int SW = TX + 16;       // shared row width: columns -8 .. TX+7

template <class T>
struct alignas(16) Chunk {
    T v[4];
};

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
template <class T>
__device__ __forceinline__ void cp_async_chunk(T *smem, const T *gmem)
{
    cp_async16(smem, gmem);
    if (sizeof(T) == 8)
        cp_async16((char *)smem + 16, (const char *)gmem + 16);
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

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

template <class T, class CT, bool ADJ, int TY, bool edge>
__device__ __forceinline__ void vd_tile(const VdFusedParams<T> &P, unsigned char *smem_raw)
{
    T *sp = reinterpret_cast<T *>(smem_raw);          // p_in      rows -3 .. TY+2, cols -8 .. TX+7
    T *svx = sp + (TY + 6) * SW;                      // vx_new    rows  0 .. TY-1, cols -8 .. TX+7
    T *svy = svx + TY * SW;                           // vy_new    rows -2 .. TY,   cols  0 .. TX-1
    T *spi = svy + (TY + 3) * TX;                     // p_it      rows -1 .. TY+1, cols -8 .. TX+7 (adjoint only)

    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int nx = P.nx, ny = P.ny, h = P.halo;
    const long long ld = P.ld;
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;

    // ---- phase 1: stage p_in (and p_it) in shared memory -------------------------------------------------
    {
        const T *g = P.p_in + (long long)(y0 - 3) * ld + (x0 - 8);
        for (int idx = tid; idx < (TY + 6) * (NCH + 4); idx += NTHR) {
            const int r = idx / (NCH + 4), c = idx - r * (NCH + 4);
            const bool halo_row = r < 3 || r >= TY + 3;
            if (halo_row && (c < 2 || c >= NCH + 2))
                continue; // corners are never read
            if (x0 - 8 + 4 * c >= ld) { // beyond the row pitch: only feeds cells that are never stored
                Chunk<T> z = {};
                st_chunk(sp + r * SW + 4 * c, z);
                continue;
            }
            cp_async_chunk(sp + r * SW + 4 * c, g + (long long)r * ld + 4 * c);
        }
        if (ADJ) {
            const T *gi = P.pc_it + (long long)(y0 - 1) * ld + (x0 - 8);
            for (int idx = tid; idx < (TY + 3) * (NCH + 4); idx += NTHR) {
                const int r = idx / (NCH + 4), c = idx - r * (NCH + 4);
                if (c < 1 || c >= NCH + 3)
                    continue;
                if (x0 - 8 + 4 * c >= ld) {
                    Chunk<T> z = {};
                    st_chunk(spi + r * SW + 4 * c, z);
                    continue;
                }
                cp_async_chunk(spi + r * SW + 4 * c, gi + (long long)r * ld + 4 * c);
            }
        }
        cp_async_wait_all();
        __syncthreads();
    }

    // record_receivers!: traces[it, r] = p_in[rec]  (p_in is the pressure after the reference's step `rec_it`)
    if (P.rec_it > 0) {
        const int e0 = P.rec_off[tile], e1 = P.rec_off[tile + 1];
        for (int e = e0 + tid; e < e1; e += NTHR) {
            const int cell = P.rec_cell[e];
            const int r = cell / TX, c = cell - r * TX;
            P.traces[(size_t)P.rec_idx[e] * P.rec_nt + (P.rec_it - 1)] = sp[(r + 3) * SW + c + 8];
        }
    }

    // ---- phase 2a: vx_new on columns -4 .. TX+3 (chunks -1 .. NCH), rows 0 .. TY-1 ------------------------
    for (int idx = tid; idx < TY * (NCH + 2); idx += NTHR) {
        const int r = idx / (NCH + 2), c = idx - r * (NCH + 2) - 1;
        const int gx = x0 + 4 * c, gy = y0 + r; // 0-based global column of the chunk's first cell / row
        Chunk<T> out = {};
        if (gx < ld) {
            const long long q = (long long)gy * ld + gx;
            const Chunk<T> vin = ldg_chunk(P.vx_in + q);
            out = vin;
            if (P.do_v) {
                const Chunk<T> m1 = ldg_chunk(P.m1x + q);
                const T *ps = sp + (r + 3) * SW + 4 * c + 8; // p at column 4c
                const Chunk<T> pa = ldg_chunk(ps - 4), pb = ldg_chunk(ps), pc = ldg_chunk(ps + 4);
                const T w[7] = {pa.v[3], pb.v[0], pb.v[1], pb.v[2], pb.v[3], pc.v[0], pc.v[1]};
                const bool owned = c >= 0 && c < NCH && gy < ny;
                Chunk<T> g1 = {};
                T wi[7];
                if (ADJ && owned) {
                    g1 = ldg_chunk(P.g1x + q);
                    const T *pis = spi + (r + 1) * SW + 4 * c + 8;
                    const Chunk<T> ia = ldg_chunk(pis - 4), ib = ldg_chunk(pis), ic = ldg_chunk(pis + 4);
                    wi[0] = ia.v[3], wi[1] = ib.v[0], wi[2] = ib.v[1], wi[3] = ib.v[2], wi[4] = ib.v[3], wi[5] = ic.v[0], wi[6] = ic.v[1];
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int I = gx + e + 1, J = gy + 1; // 1-based reference indices
                    if (edge && !(I >= 1 && I <= nx - 1 && J <= ny))
                        continue; // outside update_vx_CPML!'s range: stays as it is (zero)
                    CT D = fd4<T, CT>(P, w[e], w[e + 1], w[e + 2], w[e + 3], P.inv_dx);
                    if (edge && (I <= h || I >= nx - h)) {
                        const int ii = I <= h ? I : I - nx + 2 * h + 1;
                        const size_t qs = (size_t)(J - 1) * (2 * h) + (ii - 1);
                        T sn;
                        D = cpml_apply<T, CT>(D, P.a_xh[ii - 1], P.b_xh[ii - 1], P.psi_x_in[qs], sn);
                        if (owned)
                            P.psi_x_out[qs] = sn;
                    }
                    out.v[e] = (T)((CT)vin.v[e] - (CT)m1.v[e] * D);
                    if (ADJ && owned) {
                        const CT Dc = fd4<T, CT>(P, wi[e], wi[e + 1], wi[e + 2], wi[e + 3], P.inv_dx);
                        g1.v[e] = (T)((CT)g1.v[e] + (CT)out.v[e] * Dc);
                    }
                }
                if (owned) {
                    st_chunk(P.vx_out + q, out);
                    if (ADJ)
                        st_chunk(P.g1x + q, g1);
                }
            }
        }
        st_chunk(svx + r * SW + 4 * c + 8, out);
    }

    // ---- phase 2b: vy_new on rows -2 .. TY, columns 0 .. TX-1 ---------------------------------------------
    for (int idx = tid; idx < (TY + 3) * NCH; idx += NTHR) {
        const int rr = idx / NCH, c = idx - rr * NCH;
        const int r = rr - 2;
        const int gx = x0 + 4 * c, gy = y0 + r;
        Chunk<T> out = {};
        if (gx < ld) {
            const long long q = (long long)gy * ld + gx;
            const Chunk<T> vin = ldg_chunk(P.vy_in + q);
            out = vin;
            if (P.do_v) {
                const Chunk<T> m1 = ldg_chunk(P.m1y + q);
                const T *ps = sp + (r + 3) * SW + 4 * c + 8;
                const Chunk<T> pa = ldg_chunk(ps - SW), pb = ldg_chunk(ps), pc = ldg_chunk(ps + SW), pd = ldg_chunk(ps + 2 * SW);
                const bool owned = r >= 0 && r < TY && gy < ny;
                Chunk<T> g1 = {}, ia = {}, ib = {}, ic = {}, id = {};
                if (ADJ && owned) {
                    g1 = ldg_chunk(P.g1y + q);
                    const T *pis = spi + (r + 1) * SW + 4 * c + 8;
                    ia = ldg_chunk(pis - SW), ib = ldg_chunk(pis), ic = ldg_chunk(pis + SW), id = ldg_chunk(pis + 2 * SW);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int I = gx + e + 1, J = gy + 1;
                    if (edge && !(I <= nx && J >= 1 && J <= ny - 1))
                        continue;
                    CT D = fd4<T, CT>(P, pa.v[e], pb.v[e], pc.v[e], pd.v[e], P.inv_dy);
                    if (edge && (J <= h || J >= ny - h)) {
                        const int jj = J <= h ? J : J - ny + 2 * h + 1;
                        const size_t qs = (size_t)(jj - 1) * nx + (I - 1);
                        T sn;
                        D = cpml_apply<T, CT>(D, P.a_yh[jj - 1], P.b_yh[jj - 1], P.psi_y_in[qs], sn);
                        if (owned)
                            P.psi_y_out[qs] = sn;
                    }
                    out.v[e] = (T)((CT)vin.v[e] - (CT)m1.v[e] * D);
                    if (ADJ && owned) {
                        const CT Dc = fd4<T, CT>(P, ia.v[e], ib.v[e], ic.v[e], id.v[e], P.inv_dy);
                        g1.v[e] = (T)((CT)g1.v[e] + (CT)out.v[e] * Dc);
                    }
                }
                if (owned) {
                    st_chunk(P.vy_out + q, out);
                    if (ADJ)
                        st_chunk(P.g1y + q, g1);
                }
            }
        }
        st_chunk(svy + rr * TX + 4 * c, out);
    }
    if (!P.do_p)
        return;
    __syncthreads();

    // ---- phase 3: p_new on the tile, injection, m0 correlation -------------------------------------------
    const int ie0 = P.inj_it > 0 ? P.inj_off[tile] : 0, ie1 = P.inj_it > 0 ? P.inj_off[tile + 1] : 0;
    for (int idx = tid; idx < TY * NCH; idx += NTHR) {
        const int r = idx / NCH, c = idx - r * NCH;
        const int gx = x0 + 4 * c, gy = y0 + r;
        if (gx >= ld || gy >= ny)
            continue;
        const long long q = (long long)gy * ld + gx;
        const Chunk<T> m0 = ldg_chunk(P.m0 + q);
        const T *vxs = svx + r * SW + 4 * c + 8;
        const Chunk<T> xa = ldg_chunk(vxs - 4), xb = ldg_chunk(vxs), xc = ldg_chunk(vxs + 4);
        const T wx[7] = {xa.v[2], xa.v[3], xb.v[0], xb.v[1], xb.v[2], xb.v[3], xc.v[0]}; // vx at columns 4c-2 .. 4c+4
        const T *vys = svy + (r + 2) * TX + 4 * c;
        const Chunk<T> ya = ldg_chunk(vys - 2 * TX), yb = ldg_chunk(vys - TX), yc = ldg_chunk(vys), yd = ldg_chunk(vys + TX);
        const Chunk<T> pin = ldg_chunk(sp + (r + 3) * SW + 4 * c + 8);
        Chunk<T> out = pin;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int I = gx + e + 1, J = gy + 1;
            if (edge && !(I >= 2 && I <= nx - 1 && J >= 2 && J <= ny - 1))
                continue; // update_p_CPML! touches interior cells only
            // d vx / dx at I-1 (backward staggered): vx[I-2 .. I+1]
            CT Dx = fd4<T, CT>(P, wx[e], wx[e + 1], wx[e + 2], wx[e + 3], P.inv_dx);
            if (edge && (I <= h + 1 || I >= nx - h)) {
                const int ii = I <= h + 1 ? I : I - nx + 2 * h + 2;
                const size_t qs = (size_t)(J - 1) * (2 * (h + 1)) + (ii - 1);
                T sn;
                Dx = cpml_apply<T, CT>(Dx, P.a_x[ii - 1], P.b_x[ii - 1], P.xi_x_in[qs], sn);
                P.xi_x_out[qs] = sn;
            }
            CT Dy = fd4<T, CT>(P, ya.v[e], yb.v[e], yc.v[e], yd.v[e], P.inv_dy);
            if (edge && (J <= h + 1 || J >= ny - h)) {
                const int jj = J <= h + 1 ? J : J - ny + 2 * h + 2;
                const size_t qs = (size_t)(jj - 1) * nx + (I - 1);
                T sn;
                Dy = cpml_apply<T, CT>(Dy, P.a_y[jj - 1], P.b_y[jj - 1], P.xi_y_in[qs], sn);
                P.xi_y_out[qs] = sn;
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
            Chunk<T> g0 = ldg_chunk(P.g0 + q);
            const Chunk<T> pit = ldg_chunk(spi + (r + 1) * SW + 4 * c + 8);
            const Chunk<T> pm1 = ldg_chunk(P.pc_itm1 + q);
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
__global__ void __launch_bounds__(NTHR) vd_fused_kernel(const VdFusedParams<T> P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY, h = P.halo;
    // block-uniform: does the tile (with the halo it recomputes) touch a C-PML strip or the grid edge?
    const bool edge = (x0 - 8 <= h + 2) || (x0 + TX + 8 >= P.nx - h - 2) || (y0 - 4 <= h + 2) || (y0 + TY + 4 >= P.ny - h - 2);
    if (edge)
        vd_tile<T, CT, ADJ, TY, true>(P, smem_raw);
    else
        vd_tile<T, CT, ADJ, TY, false>(P, smem_raw);
}

template <class T, class CT, bool ADJ, int TY>
void launch_one(const VdFusedParams<T> &P, cudaStream_t st)
{
    const size_t smem = sizeof(T) * ((size_t)(TY + 6) * SW + (size_t)TY * SW + (size_t)(TY + 3) * TX + (ADJ ? (size_t)(TY + 3) * SW : 0));
    auto kern = vd_fused_kernel<T, CT, ADJ, TY>;
    static bool configured = false; // per instantiation
    if (!configured) {
        SWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SWB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        configured = true;
    }
    dim3 grd(cdiv(P.nx, TX), cdiv(P.ny, TY), 1);
    kern<<<grd, NTHR, smem, st>>>(P);
    check_launch("vd_fused");
    count_launch();
}

} // namespace

template <class T>
void vd_fused_launch(const VdFusedParams<T> &P, bool fast, cudaStream_t st)
{
    constexpr int TY = VDF_TY;
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

template void vd_fused_launch<float>(const VdFusedParams<float> &, bool, cudaStream_t);
template void vd_fused_launch<double>(const VdFusedParams<double> &, bool, cudaStream_t);

} // namespace swb
