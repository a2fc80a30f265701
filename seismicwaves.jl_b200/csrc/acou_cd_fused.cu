// acou_cd_fused.cu -- fused acoustic constant-density time step (2D and 3D) for the per-shot engine.
//
// One step of the reference is five or six launches (src/models/acoustic/backends/shared/acoustic2D_xPU.jl:1-126,
// acoustic3D_xPU.jl:1-219):
//     psi_d <- b_h psi_d + a_h d_d p            update_ψ_x! / _y! / _z!   (C-PML strips, from pcur)
//     xi_d  <- b xi_d + a (d2_d p + d_d psi_d)  inside update_p_CPML!      (src/utils/fdgen.jl:163-193)
//     pnew  <- 2.0 pcur - pold + fact * lap~    update_p_CPML!
//     pnew[src] += tf[it, s]                    inject_sources!
//     traces[it, r] = pnew[rec]                 record_receivers!
// plus, in the gradient's adjoint loop, the zero-lag correlation of the freshly updated adjoint field with the
// stored forward fields, grad += adj * (p_itm2 - 2.0 p_itm1 + p_it) / dt^2 (correlate_gradient_xPU.jl:1-10).
// Here a step is two kernels that write disjoint cells, read only the previous time levels and therefore run
// concurrently (the engine forks the rim kernel onto a second stream inside its CUDA graph):
//
//  * cd_bulk_kernel -- every 16-byte vector of cells that lies outside the C-PML strips and off the grid faces
//    (85 % of a 768^3 grid with halo 20, 98 % of 4096^2).  2.5D register-queue march: a thread owns V = 16 bytes /
//    sizeof(T) consecutive x cells of one (x, y) column position and marches along z over `zc` planes with
//    p[k-1], p[k], p[k+1] in registers; a warp owns one 32 V cell row.  x neighbours come from the adjacent lanes
//    with warp shuffles (lanes 0 / 31 fetch the one halo column each); y neighbours are read through L1 (they are
//    the lines the neighbouring warps of the CTA brought in as their own centre values one iteration earlier), so
//    there is no shared-memory staging and no barrier and warps drift freely; z neighbours come from the queue.
//    The loads of iteration k+1 (pcur[k+2], pold[k+1], fact[k+1], halo column) are issued before the arithmetic of
//    iteration k.  HBM traffic per cell-update: pcur, pold, fact read, pnew written = 4 values; adjoint +
//    correlation adds p_itm2, p_itm1, p_it read and grad read + written = 9 values.
//  * cd_rim_kernel -- the strips and faces, enumerated compactly as up to six boxes of vectors, one thread per
//    vector, generic code: a cell in a strip recomputes the two psi values its d_d psi needs from (psi_old, pcur)
//    instead of reading what a separate psi sweep stored, so psi is double-buffered; xi has exactly one owner per
//    entry and is updated in place.  Cells on the grid faces are never updated by the reference (pnew aliases pold
//    there); here they copy pold so that pnew may live anywhere.
// The values are those of the reference sequence operation for operation (SWB_FLAG_FAST_F32 swaps in FMA forms).
// A 2D grid (nx, ny) is run as (nx, 1, ny): no y term, every warp of a bulk CTA marches on its own x range.
#include "cd_fused.h"
#include "kernels.h"

namespace swb {

namespace {

template <class T, int V>
struct alignas(16) CVec {
    T v[V];
};

template <class T, int V>
__device__ __forceinline__ CVec<T, V> ldv(const T *p)
{
    return *reinterpret_cast<const CVec<T, V> *>(p);
}
template <class T, int V>
__device__ __forceinline__ void stv(T *p, const CVec<T, V> &c)
{
    *reinterpret_cast<CVec<T, V> *>(p) = c;
}

__device__ __forceinline__ float shfl_up1(float x) { return __shfl_up_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ float shfl_dn1(float x) { return __shfl_down_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ double shfl_up1(double x) { return __shfl_up_sync(0xffffffffu, x, 1); }
__device__ __forceinline__ double shfl_dn1(double x) { return __shfl_down_sync(0xffffffffu, x, 1); }

// plain second difference (lo - 2 mid + hi) / d^2: reference order with the Fornberg weights (fdgen.jl:96-131), or FMA form
template <class CT, bool FMA>
__device__ __forceinline__ CT d2(const CT (&w)[3], CT lo, CT mid, CT hi, CT inv2)
{
    if (FMA)
        return (fma((CT)-2.0, mid, lo) + hi) * inv2;
    return ((w[0] * lo + w[1] * mid) + w[2] * hi) * inv2;
}

// pnew = 2.0 pcur - pold + fact * lap  (acoustic2D_xPU.jl:43)
template <class T, class CT, bool FMA>
__device__ __forceinline__ T leapfrog(CT pc, T po, T fc, CT lap)
{
    if (FMA)
        return (T)fma((CT)fc, lap, fma((CT)2.0, pc, -(CT)po));
    return (T)((((CT)2.0 * pc) - (CT)po) + (CT)fc * lap);
}

// grad + adj * (pm2 - 2.0 pm1 + p0) * _dt2  (correlate_gradient_xPU.jl:1-10)
template <class T, class CT, bool FMA>
__device__ __forceinline__ T correlate(T g, T adj, T pm2, T pm1, T p0, T inv_dt2)
{
    if (FMA) {
        const CT lapt = fma((CT)-2.0, (CT)pm1, (CT)pm2) + (CT)p0;
        return (T)fma((CT)adj * lapt, (CT)inv_dt2, (CT)g);
    }
    const CT lapt = ((CT)pm2 - (CT)2.0 * (CT)pm1) + (CT)p0;
    return (T)((CT)g + ((CT)adj * lapt) * (CT)inv_dt2);
}

// point sources / receivers of this CTA touching the vector whose first cell has in-CTA code `code0`
template <class T, int V>
__device__ __forceinline__ void inject_points(const CdFusedParams<T> &P, const CdPointList &L, int cta, int code0, CVec<T, V> &out)
{
    const int e0 = L.off[cta], e1 = L.off[cta + 1];
    for (int e = e0; e < e1; ++e) { // entries in source-index order: summed like the CPU loop
        const int d = L.cell[e] - code0;
        if (d >= 0 && d < V) {
            const T s = P.inj_tf[(long long)L.idx[e] * P.inj_nt + (P.inj_it - 1)];
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v == d)
                    out.v[v] = out.v[v] + s;
        }
    }
}
template <class T, int V>
__device__ __forceinline__ void record_points(const CdFusedParams<T> &P, const CdPointList &L, int cta, int code0, const CVec<T, V> &out)
{
    const int e0 = L.off[cta], e1 = L.off[cta + 1];
    for (int e = e0; e < e1; ++e) {
        const int d = L.cell[e] - code0;
        if (d >= 0 && d < V) {
            T val = (T)0;
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (v == d)
                    val = out.v[v];
            P.traces[(long long)L.idx[e] * P.rec_nt + (P.rec_it - 1)] = val;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// bulk kernel
// ---------------------------------------------------------------------------------------------------------------------
// (bx, by, bz) = block index in the (gdx, gdy, .) bulk grid; lane / ty = thread index in the (32, bdy) block
template <class T, class CT, bool HAS_Y, bool ADJ, bool FMA>
__device__ __forceinline__ void cd_bulk_body(const CdFusedParams<T> &P, const int bx, const int by, const int bz, const int gdx, const int gdy, const int bdy,
                                             const int lane, const int ty)
{
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int TX = 32 * V;
    constexpr int TY = HAS_Y ? CDF_TY : 1;
    typedef CVec<T, V> VT;

    const int xt = HAS_Y ? bx : bx * bdy + ty;
    const int i0 = xt * TX + lane * V;
    const int j = HAS_Y ? P.jlo + by * TY + ty : 0;
    if (xt * TX >= P.nx || j >= P.jhi)
        return; // warps are independent (no barriers); a warp leaves only as a whole (shuffles below)
    const int k0 = P.klo + bz * P.zc, k1 = min(k0 + P.zc, P.khi);
    const int iv = xt * 32 + lane;
    const bool ld_ok = i0 < P.nx;                   // this lane's vector exists: it feeds the neighbours' shuffles
    const bool st_ok = iv >= P.ivlo && iv < P.ivhi; // this lane's vector belongs to the bulk
    const bool hx_any = st_ok && (lane == 0 || lane == 31);
    const long long ld = P.ld, plane = P.plane;
    long long off = (long long)k0 * plane + (long long)j * ld + i0; // this thread's vector in plane k
    const long long hxo = lane == 0 ? -1 : V;
    const int cta = (bz * gdy + by) * gdx + bx;
    const bool has_inj = P.inj_it > 0 && P.inj[0].off[cta + 1] > P.inj[0].off[cta];
    const bool has_rec = P.rec_it > 0 && P.rec[0].off[cta + 1] > P.rec[0].off[cta];

    const CT w2[3] = {(CT)P.c2[0], (CT)P.c2[1], (CT)P.c2[2]};
    const CT i2x = (CT)(P.inv_d[0] * P.inv_d[0]), i2y = (CT)(P.inv_d[1] * P.inv_d[1]), i2z = (CT)(P.inv_d[2] * P.inv_d[2]);

    CT pm[V], pc[V], pp[V], hxc = (CT)0;
    VT po = {}, fc = {};
#pragma unroll
    for (int v = 0; v < V; ++v)
        pm[v] = pc[v] = pp[v] = (CT)0;
    // ---- prologue: fill the queue for plane k0 (bulk planes always have both z neighbours) ----------------------------
    if (ld_ok) {
        const VT a = ldv<T, V>(P.pcur + off - plane), b = ldv<T, V>(P.pcur + off), c = ldv<T, V>(P.pcur + off + plane);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            pm[v] = (CT)a.v[v];
            pc[v] = (CT)b.v[v];
            pp[v] = (CT)c.v[v];
        }
    }
    if (st_ok) {
        po = ldv<T, V>(P.pold + off);
        fc = ldv<T, V>(P.fact + off);
    }
    if (hx_any)
        hxc = (CT)P.pcur[off + hxo];

#pragma unroll 1
    for (int k = k0; k < k1; ++k) {
        // ---- loads: the next iteration's centre / pold / fact / halo column, this iteration's y neighbours (L1) -------
        const bool more = k + 1 < k1;
        VT pn = {}, po_n = {}, fc_n = {}, yu = {}, yd = {};
        T hx_n = (T)0;
        if (ld_ok && more)
            pn = ldv<T, V>(P.pcur + off + 2 * plane);
        if (st_ok) {
            if (HAS_Y) {
                yu = ldv<T, V>(P.pcur + off - ld);
                yd = ldv<T, V>(P.pcur + off + ld);
            }
            if (more) {
                po_n = ldv<T, V>(P.pold + off + plane);
                fc_n = ldv<T, V>(P.fact + off + plane);
            }
        }
        if (hx_any && more)
            hx_n = P.pcur[off + plane + hxo];
        VT c2 = {}, c1 = {}, c0 = {}, g = {};
        if (ADJ && st_ok) {
            c2 = ldv<T, V>(P.pm2 + off);
            c1 = ldv<T, V>(P.pm1 + off);
            c0 = ldv<T, V>(P.p0 + off);
            g = ldv<T, V>(P.grad + off);
        }
        CT xl_in = shfl_up1(pc[V - 1]);
        CT xr_in = shfl_dn1(pc[0]);
        if (lane == 0)
            xl_in = hxc;
        if (lane == 31)
            xr_in = hxc;
        if (st_ok) {
            VT out;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const CT xl = v > 0 ? pc[v > 0 ? v - 1 : 0] : xl_in;
                const CT xr = v < V - 1 ? pc[v < V - 1 ? v + 1 : 0] : xr_in;
                CT lap = d2<CT, FMA>(w2, xl, pc[v], xr, i2x);
                if (HAS_Y)
                    lap = lap + d2<CT, FMA>(w2, (CT)yu.v[v], pc[v], (CT)yd.v[v], i2y);
                lap = lap + d2<CT, FMA>(w2, pm[v], pc[v], pp[v], i2z);
                out.v[v] = leapfrog<T, CT, FMA>(pc[v], po.v[v], fc.v[v], lap);
            }
            const int code0 = ((k - k0) * bdy + ty) * TX + lane * V;
            if (has_inj)
                inject_points<T, V>(P, P.inj[0], cta, code0, out);
            stv<T, V>(P.pnew + off, out);
            if (P.peer_lo != nullptr && k == 1) // boundary planes of a z slab go straight into the neighbour's ghost plane (NVLink peer store)
                stv<T, V>(P.peer_lo + (off - plane), out);
            if (P.peer_hi != nullptr && k == P.nz - 2)
                stv<T, V>(P.peer_hi + (off - (long long)(P.nz - 2) * plane), out);
            if (has_rec)
                record_points<T, V>(P, P.rec[0], cta, code0, out);
            if (ADJ) {
#pragma unroll
                for (int v = 0; v < V; ++v)
                    g.v[v] = correlate<T, CT, FMA>(g.v[v], out.v[v], c2.v[v], c1.v[v], c0.v[v], P.inv_dt2);
                stv<T, V>(P.grad + off, g);
            }
        }
        // ---- rotate the queue ------------------------------------------------------------------------------------------
#pragma unroll
        for (int v = 0; v < V; ++v) {
            pm[v] = pc[v];
            pc[v] = pp[v];
            pp[v] = (CT)pn.v[v];
        }
        po = po_n;
        fc = fc_n;
        hxc = (CT)hx_n;
        off += plane;
    }
}

template <class T, class CT, bool HAS_Y, bool ADJ, bool FMA>
__global__ void __launch_bounds__(HAS_Y ? 32 * CDF_TY : 32 * CDF_W2D, HAS_Y ? (sizeof(CT) == 4 ? (ADJ ? 3 : 4) : 2) : (sizeof(CT) == 4 ? 8 : 4))
    cd_bulk_kernel(const CdFusedParams<T> P)
{
    const int bz = P.rev ? (int)gridDim.z - 1 - (int)blockIdx.z : (int)blockIdx.z;
    cd_bulk_body<T, CT, HAS_Y, ADJ, FMA>(P, (int)blockIdx.x, (int)blockIdx.y, bz, (int)gridDim.x, (int)gridDim.y, (int)blockDim.y, (int)threadIdx.x,
                                         (int)threadIdx.y);
}

// ---------------------------------------------------------------------------------------------------------------------
// rim kernel
// ---------------------------------------------------------------------------------------------------------------------

// One axis of @∇̃² for an interior cell (fdgen.jl:163-193 with the psi update of acoustic2D_xPU.jl:1-25 inlined).
//   lo, mid, hi : p at c-1, c, c+1 along the axis;  c1b: 1-based index of the cell along the axis;  n: axis extent
//   entry ii (1-based) of psi lives at psi[(ii-1)*stride] (base pointers already offset to this cell's line)
template <class T, class CT, bool FMA>
__device__ __forceinline__ CT cd_axis_term(const CT (&w1)[2], const CT (&w2)[3], CT lo, CT mid, CT hi, int c1b, int n, int h, T inv, const T *__restrict__ a,
                                           const T *__restrict__ b, const T *__restrict__ a_h, const T *__restrict__ b_h, const T *__restrict__ psi_in,
                                           T *__restrict__ psi_out, T *__restrict__ xi, long long stride)
{
    const CT D2 = d2<CT, FMA>(w2, lo, mid, hi, (CT)(inv * inv));
    int ii;
    if (c1b <= h)
        ii = c1b;
    else if (c1b >= n - h + 1)
        ii = c1b - (n - h) + 1 + h;
    else
        return D2;
    // psi[ii] belongs to grid cell c (difference p[c+1]-p[c]), psi[ii-1] to cell c-1
    const CT Dhi = (w1[0] * mid + w1[1] * hi) * (CT)inv;
    const CT Dlo = (w1[0] * lo + w1[1] * mid) * (CT)inv;
    T psi_hi, psi_lo;
    (void)cpml_apply<T, CT>(Dhi, a_h[ii - 1], b_h[ii - 1], psi_in[(long long)(ii - 1) * stride], psi_hi);
    (void)cpml_apply<T, CT>(Dlo, a_h[ii - 2], b_h[ii - 2], psi_in[(long long)(ii - 2) * stride], psi_lo);
    psi_out[(long long)(ii - 1) * stride] = psi_hi;
    if (c1b == 2 || c1b == n - h + 1) // the strip's first entry has no interior cell of its own
        psi_out[(long long)(ii - 2) * stride] = psi_lo;
    const CT dpsi = (w1[0] * (CT)psi_lo + w1[1] * (CT)psi_hi) * (CT)inv;
    T *x = xi + (long long)(ii - 1) * stride;
    const T bx = b[ii - 1] * *x;
    const T xn = (T)((CT)bx + (CT)a[ii - 1] * (D2 + dpsi));
    *x = xn;
    return (D2 + dpsi) + (CT)xn;
}

// compact strip index of 1-based cell index c along an axis of extent n (0: not in a strip)
__device__ __forceinline__ int strip_index(int c, int n, int h)
{
    if (c <= h)
        return c;
    if (c >= n - h + 1)
        return c - (n - h) + 1 + h;
    return 0;
}

// The same operator for the V cells of one vector along an axis whose memory-variable arrays have x fastest (y and z):
// the cells share the strip index ii, so psi / xi move as 16-byte vectors (row pitch ld keeps them aligned).
//   ok[v]: the cell is updated (interior in x); other cells keep their psi / xi
template <class T, class CT, int V, bool FMA>
__device__ __forceinline__ void cd_axis_term_vec(CT (&term)[V], const CT (&w1)[2], const CT (&w2)[3], const CT (&lo)[V], const CT (&mid)[V], const CT (&hi)[V],
                                                 const bool (&ok)[V], int c1b, int n, int h, T inv, const T *__restrict__ a, const T *__restrict__ b,
                                                 const T *__restrict__ a_h, const T *__restrict__ b_h, const T *__restrict__ psi_in, T *__restrict__ psi_out,
                                                 T *__restrict__ xi, long long stride, bool lo_on = true, bool hi_on = true)
{
    const CT inv2 = (CT)(inv * inv);
#pragma unroll
    for (int v = 0; v < V; ++v)
        term[v] = d2<CT, FMA>(w2, lo[v], mid[v], hi[v], inv2);
    const int ii = strip_index(c1b, n, h);
    if (ii == 0 || (ii <= h ? !lo_on : !hi_on)) // a slab-interior end of the axis has no strip
        return;
    typedef CVec<T, V> VT;
    const VT ph = ldv<T, V>(psi_in + (long long)(ii - 1) * stride), pl = ldv<T, V>(psi_in + (long long)(ii - 2) * stride);
    VT xo = ldv<T, V>(xi + (long long)(ii - 1) * stride);
    const T ah1 = a_h[ii - 1], bh1 = b_h[ii - 1], ah0 = a_h[ii - 2], bh0 = b_h[ii - 2], a1 = a[ii - 1], b1 = b[ii - 1];
    VT psi_hi = ph, psi_lo = pl;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if (!ok[v])
            continue;
        const CT Dhi = (w1[0] * mid[v] + w1[1] * hi[v]) * (CT)inv;
        const CT Dlo = (w1[0] * lo[v] + w1[1] * mid[v]) * (CT)inv;
        (void)cpml_apply<T, CT>(Dhi, ah1, bh1, ph.v[v], psi_hi.v[v]);
        (void)cpml_apply<T, CT>(Dlo, ah0, bh0, pl.v[v], psi_lo.v[v]);
        const CT dpsi = (w1[0] * (CT)psi_lo.v[v] + w1[1] * (CT)psi_hi.v[v]) * (CT)inv;
        const T bx = b1 * xo.v[v];
        const T xn = (T)((CT)bx + (CT)a1 * (term[v] + dpsi));
        xo.v[v] = xn;
        term[v] = (term[v] + dpsi) + (CT)xn;
    }
    stv<T, V>(psi_out + (long long)(ii - 1) * stride, psi_hi);
    if (c1b == 2 || c1b == n - h + 1) // the strip's first entry has no interior cell of its own
        stv<T, V>(psi_out + (long long)(ii - 2) * stride, psi_lo);
    stv<T, V>(xi + (long long)(ii - 1) * stride, xo);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <class T, class CT, bool HAS_Y, bool ADJ, bool FMA>
__device__ __forceinline__ void cd_rim_body(const CdFusedParams<T> &P, const int ecta, const int tid, const int nbox, const long long nrimvec, const int cta0 = 0)
{
    // ecta: CTA index in the enumeration of the first nbox boxes (nrimvec vectors); cta: its number in the point lists
    constexpr int V = 16 / (int)sizeof(T);
    typedef CVec<T, V> VT;
    const int cta = cta0 + ecta;
    const long long vid = (long long)ecta * CDF_RIM_T + tid;
    if (vid >= nrimvec)
        return;
    int bi = 0;
#pragma unroll
    for (int b = 1; b < CDF_MAX_BOX; ++b)
        if (b < nbox && vid >= P.box[b].start)
            bi = b;
    const CdBox &B = P.box[bi];
    const unsigned loc = (unsigned)(vid - B.start); // a box holds fewer than 2^31 vectors (checked on the host)
    const unsigned r = loc / (unsigned)B.nvx, kk = r / (unsigned)B.ny;
    const int iv = B.iv0 + (int)(loc - r * (unsigned)B.nvx);
    const int j = B.j0 + (int)(r - kk * (unsigned)B.ny), k = B.k0 + (int)kk;
    const int i0 = iv * V;
    const int nx = P.nx, ny = P.ny, nz = P.nz, h = P.halo;
    if ((k == 0 && P.ghost_lo) || (k == nz - 1 && P.ghost_hi))
        return; // a ghost plane of a z slab: written by the neighbour that owns it
    const long long ld = P.ld, plane = P.plane;
    const long long off = (long long)k * plane + (long long)j * ld + i0;

    // The memory variables of a strip cell are addressed by indices only, but their loads sit behind the first use of the pressure
    // values in program order (three dependent batches of DRAM latency per thread, ncu: the stalls sit on the first use of each
    // batch): ask for their lines now, so that those loads hit the L1 when the code reaches them.
    if (h > 0 && P.rim_prefetch) {
        const int sz = strip_index(k + 1, nz, h);
        if (sz >= 2 && (sz <= h ? P.zpml_lo != 0 : P.zpml_hi != 0) && sz <= 2 * h) {
            const long long o = (long long)j * ld + i0, st = ld * ny;
            prefetch_l1(P.psi_in[2] + o + (long long)(sz - 1) * st);
            prefetch_l1(P.psi_in[2] + o + (long long)(sz - 2) * st);
            prefetch_l1(P.xi[2] + o + (long long)(sz - 1) * st);
        }
        if (HAS_Y) {
            const int sy = strip_index(j + 1, ny, h);
            if (sy >= 2 && sy <= 2 * h) {
                const long long o = (long long)k * ld * (2 * h) + i0, ox = (long long)k * ld * (2 * (h + 1)) + i0;
                prefetch_l1(P.psi_in[1] + o + (long long)(sy - 1) * ld);
                prefetch_l1(P.psi_in[1] + o + (long long)(sy - 2) * ld);
                prefetch_l1(P.xi[1] + ox + (long long)(sy - 1) * ld);
            }
        }
        if (i0 + 1 <= h || i0 + V >= nx - h + 1) {
            const long long jk = (long long)k * ny + j;
            const int sx = max(2, strip_index(i0 + 1 <= h ? i0 + 1 : min(i0 + V, nx), nx, h));
            prefetch_l1(P.psi_in[0] + jk * (2 * h) + min(sx - 2, 2 * h - 1));
            prefetch_l1(P.xi[0] + jk * (2 * (h + 1)) + min(sx - 1, 2 * h + 1));
        }
    }
    VT out = ldv<T, V>(P.pold + off); // faces (and pitch padding) keep pold
    VT c2 = {}, c1 = {}, c0 = {}, g = {};
    if (ADJ) {
        c2 = ldv<T, V>(P.pm2 + off);
        c1 = ldv<T, V>(P.pm1 + off);
        c0 = ldv<T, V>(P.p0 + off);
        g = ldv<T, V>(P.grad + off);
    }
    const bool y_int = !HAS_Y || (j >= 1 && j <= ny - 2);
    const bool z_int = k >= 1 && k <= nz - 2;
    if (y_int && z_int) {
        const VT pcv = ldv<T, V>(P.pcur + off), fcv = ldv<T, V>(P.fact + off);
        const VT zmv = ldv<T, V>(P.pcur + off - plane), zpv = ldv<T, V>(P.pcur + off + plane);
        VT yuv = {}, ydv = {};
        if (HAS_Y) {
            yuv = ldv<T, V>(P.pcur + off - ld);
            ydv = ldv<T, V>(P.pcur + off + ld);
        }
        const T xl_in = i0 > 0 ? P.pcur[off - 1] : (T)0;
        const T xr_in = i0 + V < nx ? P.pcur[off + V] : (T)0;
        const CT w1[2] = {(CT)P.c1[0], (CT)P.c1[1]};
        const CT w2[3] = {(CT)P.c2[0], (CT)P.c2[1], (CT)P.c2[2]};
        CT pc[V], lo[V], hi[V], tx[V], tyz[V];
        bool ok[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            pc[v] = (CT)pcv.v[v];
            ok[v] = i0 + v >= 1 && i0 + v <= nx - 2;
        }
        // x term: plain everywhere, C-PML for the cells inside the x strips (psi_x / xi_x have the strip index fastest)
        const CT i2x = (CT)(P.inv_d[0] * P.inv_d[0]);
        const bool x_strip = h > 0 && (i0 + 1 <= h || i0 + V >= nx - h + 1);
        const long long jk = (long long)k * ny + j;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const CT xl = (CT)(v > 0 ? pcv.v[v > 0 ? v - 1 : 0] : xl_in);
            const CT xr = (CT)(v < V - 1 ? pcv.v[v < V - 1 ? v + 1 : 0] : xr_in);
            if (x_strip && ok[v])
                tx[v] = cd_axis_term<T, CT, FMA>(w1, w2, xl, pc[v], xr, i0 + v + 1, nx, h, P.inv_d[0], P.a[0], P.b[0], P.a_h[0], P.b_h[0],
                                                 P.psi_in[0] + jk * (2 * h), P.psi_out[0] + jk * (2 * h), P.xi[0] + jk * (2 * (h + 1)), 1);
            else
                tx[v] = d2<CT, FMA>(w2, xl, pc[v], xr, i2x);
        }
        if (HAS_Y) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                lo[v] = (CT)yuv.v[v];
                hi[v] = (CT)ydv.v[v];
            }
            const long long o = (long long)k * ld * (2 * h) + i0, ox = (long long)k * ld * (2 * (h + 1)) + i0;
            cd_axis_term_vec<T, CT, V, FMA>(tyz, w1, w2, lo, pc, hi, ok, j + 1, ny, h, P.inv_d[1], P.a[1], P.b[1], P.a_h[1], P.b_h[1], P.psi_in[1] + o,
                                            P.psi_out[1] + o, P.xi[1] + ox, ld);
#pragma unroll
            for (int v = 0; v < V; ++v)
                tx[v] = tx[v] + tyz[v];
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            lo[v] = (CT)zmv.v[v];
            hi[v] = (CT)zpv.v[v];
        }
        {
            const long long o = (long long)j * ld + i0;
            cd_axis_term_vec<T, CT, V, FMA>(tyz, w1, w2, lo, pc, hi, ok, k + 1, nz, h, P.inv_d[2], P.a[2], P.b[2], P.a_h[2], P.b_h[2], P.psi_in[2] + o,
                                            P.psi_out[2] + o, P.xi[2] + o, ld * ny, P.zpml_lo != 0, P.zpml_hi != 0);
        }
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (ok[v])
                out.v[v] = leapfrog<T, CT, FMA>(pc[v], out.v[v], fcv.v[v], tx[v] + tyz[v]);
    }
    const int code0 = tid * V;
    if (P.inj_it > 0)
        inject_points<T, V>(P, P.inj[1], cta, code0, out);
    stv<T, V>(P.pnew + off, out);
    if (P.peer_lo != nullptr && k == 1)
        stv<T, V>(P.peer_lo + (off - plane), out);
    if (P.peer_hi != nullptr && k == nz - 2)
        stv<T, V>(P.peer_hi + (off - (long long)(nz - 2) * plane), out);
    if (P.rec_it > 0)
        record_points<T, V>(P, P.rec[1], cta, code0, out);
    if (ADJ) {
#pragma unroll
        for (int v = 0; v < V; ++v)
            g.v[v] = correlate<T, CT, FMA>(g.v[v], out.v[v], c2.v[v], c1.v[v], c0.v[v], P.inv_dt2);
        stv<T, V>(P.grad + off, g);
    }
}

#ifndef CDF_RIM_MINB
#define CDF_RIM_MINB 8
#endif
template <class T, class CT, bool HAS_Y, bool ADJ, bool FMA>
__global__ void __launch_bounds__(CDF_RIM_T, sizeof(CT) == 4 ? CDF_RIM_MINB : CDF_RIM_MINB / 2) cd_rim_kernel(const CdFusedParams<T> P)
{
    // (3D with the march: only the leading z-plane boxes, numbered after the march's CTAs in the point lists)
    if (HAS_Y && P.nbox_v > 0)
        cd_rim_body<T, CT, HAS_Y, ADJ, FMA>(P, (int)blockIdx.x, (int)threadIdx.x, P.nbox_v, P.nrimvec_v, P.rimv_cta0);
    else
        cd_rim_body<T, CT, HAS_Y, ADJ, FMA>(P, (int)blockIdx.x, (int)threadIdx.x, P.nbox, P.nrimvec);
}

// ---------------------------------------------------------------------------------------------------------------------
// rim kernel, z-marching form (3D)
// ---------------------------------------------------------------------------------------------------------------------
// The per-vector rim kernel above re-fetches three planes of pcur per cell and spends ~650 warp instructions per vector, more than
// half of them address arithmetic (ncu, 768^3: 3.9 TB/s on its own traffic, issue slots 43 % busy at 42 % occupancy).  Here a thread
// owns one (x-vector, row) column of a rim box and marches it along z like a bulk thread: pcur[k-1], pcur[k], pcur[k+1] in a
// register queue, and everything plane k + 1 needs -- pcur[k+2], pold, fact, the y neighbours and the memory variables of the
// box's own strip axis -- is requested before the arithmetic of plane k.  Everything that does not change along the march is hoisted:
// "this row is a y face" is a property of the thread, the C-PML coefficients come from per-CTA shared-memory tables (x strips: one
// record per strip vector; z strips: one record per strip index) or registers (y strip: one strip index per row), addresses advance
// by running offsets, the two prefetch buffers swap roles in a loop unrolled by two, and the last iteration re-requests its own plane
// instead of predicating every prefetch.  The V + 1 staggered derivatives / psi values of an x-strip vector are computed once each
// (psi_lo of cell v + 1 is psi_hi of cell v: same operands), and along a z strip psi_lo of plane k + 1 is the psi_hi plane k just
// computed.  Strips of the other axes that cross a box (edges and corners of the grid) fetch their memory variables on demand.
// Values are those of cd_rim_body operation for operation.

// one cell, one axis of @∇̃² with preloaded memory variables (same operations as cd_axis_term / cd_axis_term_vec)
template <class T, class CT>
__device__ __forceinline__ CT cd_axis_cell(const CT (&w1)[2], CT D2, CT lo, CT mid, CT hi, T inv, T ah1, T bh1, T ah0, T bh0, T a1, T b1, T ph, T pl, T xo,
                                           T &psi_hi, T &psi_lo, T &xn)
{
    const CT Dhi = (w1[0] * mid + w1[1] * hi) * (CT)inv;
    (void)cpml_apply<T, CT>(Dhi, ah1, bh1, ph, psi_hi);
    const CT Dlo = (w1[0] * lo + w1[1] * mid) * (CT)inv;
    (void)cpml_apply<T, CT>(Dlo, ah0, bh0, pl, psi_lo);
    const CT dpsi = (w1[0] * (CT)psi_lo + w1[1] * (CT)psi_hi) * (CT)inv;
    const T bx = b1 * xo;
    xn = (T)((CT)bx + (CT)a1 * (D2 + dpsi));
    return (D2 + dpsi) + (CT)xn;
}

// x-strip description of the vector starting at cell i0: first strip index s0 (cell v has index s0 + v), mask of updated strip cells,
// mask of cells that also own the strip's first psi entry
template <int V>
__device__ __forceinline__ void cd_xstrip_desc(int i0, int nx, int h, int &s0, int &xm, int &xfirst)
{
    s0 = 0, xm = 0, xfirst = 0;
    if (h <= 0 || !(i0 + 1 <= h || i0 + V >= nx - h + 1))
        return;
    const bool lo_side = i0 + 1 <= h; // (the host guarantees nx >= 2 h + V: a vector never touches both strips)
    s0 = lo_side ? i0 + 1 : i0 + 1 - (nx - h) + 1 + h;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const int c1b = i0 + v + 1;
        const bool in = lo_side ? c1b <= h : c1b >= nx - h + 1;
        if (in && c1b >= 2 && c1b <= nx - 1) {
            xm |= 1 << v;
            if (c1b == 2 || c1b == nx - h + 1)
                xfirst |= 1 << v;
        }
    }
}
// psi entries s0 - 2 + e, e = 0 .. V, needed by the updated cells
template <int V>
__device__ __forceinline__ int cd_xstrip_need(int xm)
{
    return (xm | (xm << 1)) & ((1 << (V + 1)) - 1);
}
// per-CTA table in shared memory: record r (5 V values) of x-strip vector r: a_h[e], e < V | b_h[e], e < V | a[v] | b[v] | a_h[V], b_h[V]
template <class T, int V>
__device__ __forceinline__ void cd_xstrip_table(const CdFusedParams<T> &P, T *xtab, int tid, int nthreads)
{
    const int nrec = P.ivlo + ((int)(P.ld / V) - P.ivhi);
    for (int r = tid; r < nrec; r += nthreads) {
        const int iv = r < P.ivlo ? r : P.ivhi + (r - P.ivlo);
        int s0, xm, xf;
        cd_xstrip_desc<V>(iv * V, P.nx, P.halo, s0, xm, xf);
        const int need = cd_xstrip_need<V>(xm);
        T *q = xtab + r * (5 * V);
#pragma unroll
        for (int e = 0; e <= V; ++e) {
            const bool on = (need >> e) & 1;
            const T ah = on ? P.a_h[0][s0 - 2 + e] : (T)0, bh = on ? P.b_h[0][s0 - 2 + e] : (T)0;
            q[e < V ? e : 4 * V] = ah;
            q[e < V ? V + e : 4 * V + 1] = bh;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const bool on = (xm >> v) & 1;
            q[2 * V + v] = on ? P.a[0][s0 - 1 + v] : (T)0;
            q[3 * V + v] = on ? P.b[0][s0 - 1 + v] : (T)0;
        }
    }
}

// x term of @∇̃² for the V cells of a strip vector.  psv[e] = psi_in[s0 - 2 + e], xiv[v] = xi[s0 - 1 + v] (already loaded);
// ps_out -> psi_out[s0 - 2], xs -> xi[s0 - 1] of this (row, plane).  Same operations per cell as cd_axis_term.
template <class T, class CT, int V, bool FMA>
__device__ __forceinline__ void cd_xstrip_vec(CT (&lap)[V], const CT (&w1)[2], const CT (&w2)[3], const CVec<T, V> &qc, T xl_in, T xr_in, T inv, int xm, int xfirst,
                                              const T *rec, const T (&psv)[V + 1], const T (&xiv)[V], T *__restrict__ ps_out, T *__restrict__ xs)
{
    typedef CVec<T, V> VT;
    const VT ah = *reinterpret_cast<const VT *>(rec), bh = *reinterpret_cast<const VT *>(rec + V);
    const VT a = *reinterpret_cast<const VT *>(rec + 2 * V), b = *reinterpret_cast<const VT *>(rec + 3 * V);
    const T ahV = rec[4 * V], bhV = rec[4 * V + 1];
    CT p[V + 2];
    p[0] = (CT)xl_in;
#pragma unroll
    for (int v = 0; v < V; ++v)
        p[v + 1] = (CT)qc.v[v];
    p[V + 1] = (CT)xr_in;
    T psn[V + 1];
#pragma unroll
    for (int e = 0; e <= V; ++e) { // staggered derivative between cells e - 1 and e of the vector, and its memory variable
        const CT D = (w1[0] * p[e] + w1[1] * p[e + 1]) * (CT)inv;
        (void)cpml_apply<T, CT>(D, e < V ? ah.v[e < V ? e : 0] : ahV, e < V ? bh.v[e < V ? e : 0] : bhV, psv[e], psn[e]);
    }
    const CT inv2 = (CT)(inv * inv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const CT D2 = d2<CT, FMA>(w2, p[v], p[v + 1], p[v + 2], inv2);
        if ((xm >> v) & 1) {
            const CT dpsi = (w1[0] * (CT)psn[v] + w1[1] * (CT)psn[v + 1]) * (CT)inv;
            const T bx = b.v[v] * xiv[v];
            const T xn = (T)((CT)bx + (CT)a.v[v] * (D2 + dpsi));
            lap[v] = (D2 + dpsi) + (CT)xn;
            ps_out[v + 1] = psn[v + 1];
            if ((xfirst >> v) & 1) // the strip's first entry has no interior cell of its own
                ps_out[v] = psn[v];
            xs[v] = xn;
        } else
            lap[v] = D2;
    }
}

#ifndef CDF_RIMZ_MINB
// 3 CTAs of 168 registers per SM.  Measured on 768^3 (rim launch alone): 4 CTAs / 128 registers the same 0.53-0.54 ms; single-buffered loads
// behind an L2 prefetch two planes ahead at 4 / 5 / 6 CTAs per SM 0.57 / 0.66 / 0.78 ms (spills): more warps do not pay here
#define CDF_RIMZ_MINB 3
#endif
template <class T, int V>
struct RimPre { // what a march thread requests one plane ahead
    CVec<T, V> po, fc, yu, yd; // pold, fact, pcur of rows j - 1 and j + 1
    T xl, xr;                  // pcur of the cells left and right of the vector
    CVec<T, V> ph, pl, xo;     // y-strip box: psi_y[sy - 1], psi_y[sy - 2], xi_y[sy - 1]
    T psv[V + 1], xiv[V];      // own strip axis x: psi_x[s0 - 2 ..], xi_x[s0 - 1 ..]
};

// KIND 0: an x-strip box (rows and planes of the bulk), KIND 1: a y-strip box (all x: its corner columns cross the x strips)
template <class T, class CT, int KIND, bool ADJ, bool FMA>
__device__ __forceinline__ void cd_rimz_body(const CdFusedParams<T> &P, const CdBox &B, const CdRimTile &R, const int local, const int cta, const int tid,
                                             const T *xtab)
{
    constexpr int V = 16 / (int)sizeof(T);
    typedef CVec<T, V> VT;
    typedef RimPre<T, V> Pre;
    const int li = tid % R.tw, lj = tid / R.tw;
    const int tx_ = local % R.ntx, r_ = local / R.ntx, ty_ = r_ % R.nty, tz_ = r_ / R.nty;
    const int ivl = tx_ * R.tw + li, jl = ty_ * R.th + lj;
    if (lj >= R.th || ivl >= B.nvx || jl >= B.ny)
        return; // no barriers below, no shuffles: threads leave freely
    const int iv = B.iv0 + ivl, j = B.j0 + jl, i0 = iv * V;
    const int k0 = B.k0 + tz_ * R.zc, k1 = min(k0 + R.zc, B.k0 + B.nz); // interior planes only: 1 <= k <= nz - 2, no z strip
    const int nx = P.nx, ny = P.ny, nz = P.nz, h = P.halo;
    const long long ld = P.ld, plane = P.plane;
    const bool has_inj = P.inj_it > 0 && P.inj[1].off[cta + 1] > P.inj[1].off[cta];
    const bool has_rec = P.rec_it > 0 && P.rec[1].off[cta + 1] > P.rec[1].off[cta];
    int okm = 0;
#pragma unroll
    for (int v = 0; v < V; ++v)
        if (i0 + v >= 1 && i0 + v <= nx - 2)
            okm |= 1 << v;
    long long off = (long long)k0 * plane + (long long)j * ld + i0;

    auto finish = [&](int k, long long o, VT out, const VT &c2, const VT &c1, const VT &c0, VT g) {
        const int code0 = ((k - k0) * CDF_RIM_T + tid) * V;
        if (has_inj)
            inject_points<T, V>(P, P.inj[1], cta, code0, out);
        stv<T, V>(P.pnew + o, out);
        if (P.peer_lo != nullptr && k == 1) // boundary planes of a z slab go straight into the neighbour's ghost plane (NVLink peer store)
            stv<T, V>(P.peer_lo + (o - plane), out);
        if (P.peer_hi != nullptr && k == nz - 2)
            stv<T, V>(P.peer_hi + (o - (long long)(nz - 2) * plane), out);
        if (has_rec)
            record_points<T, V>(P, P.rec[1], cta, code0, out);
        if (ADJ) {
#pragma unroll
            for (int v = 0; v < V; ++v)
                g.v[v] = correlate<T, CT, FMA>(g.v[v], out.v[v], c2.v[v], c1.v[v], c0.v[v], P.inv_dt2);
            stv<T, V>(P.grad + o, g);
        }
    };

    if (KIND == 1 && !(j >= 1 && j <= ny - 2)) { // a y face row is never updated (pnew aliases pold in the reference)
#pragma unroll 1
        for (int k = k0; k < k1; ++k, off += plane) {
            VT c2 = {}, c1 = {}, c0 = {}, g = {};
            if (ADJ) {
                c2 = ldv<T, V>(P.pm2 + off);
                c1 = ldv<T, V>(P.pm1 + off);
                c0 = ldv<T, V>(P.p0 + off);
                g = ldv<T, V>(P.grad + off);
            }
            finish(k, off, ldv<T, V>(P.pold + off), c2, c1, c0, g);
        }
        return;
    }

    const CT w1[2] = {(CT)P.c1[0], (CT)P.c1[1]};
    const CT w2[3] = {(CT)P.c2[0], (CT)P.c2[1], (CT)P.c2[2]};
    const CT i2x = (CT)(P.inv_d[0] * P.inv_d[0]), i2y = (CT)(P.inv_d[1] * P.inv_d[1]), i2z = (CT)(P.inv_d[2] * P.inv_d[2]);
    // x strip of this vector (KIND 0: the box itself; KIND 1: the columns where the box crosses the x strips)
    int s0, xm, xfirst;
    cd_xstrip_desc<V>(i0, nx, h, s0, xm, xfirst);
    const int xneed = cd_xstrip_need<V>(xm);
    const T *xrec = xtab + (iv < P.ivlo ? iv : P.ivlo + (iv - P.ivhi)) * (5 * V);
    const long long xpinc = (long long)ny * (2 * h), xxinc = (long long)ny * (2 * (h + 1));
    long long xpo = ((long long)k0 * ny + j) * (2 * h) + (s0 - 2), xxo = ((long long)k0 * ny + j) * (2 * (h + 1)) + (s0 - 1);
    // y strip of this row (KIND 1)
    const int sy = (KIND == 1 && h > 0) ? strip_index(j + 1, ny, h) : 0;
    T yah1 = (T)0, ybh1 = (T)0, yah0 = (T)0, ybh0 = (T)0, ya1 = (T)0, yb1 = (T)0;
    if (KIND == 1 && sy > 0)
        yah1 = P.a_h[1][sy - 1], ybh1 = P.b_h[1][sy - 1], yah0 = P.a_h[1][sy - 2], ybh0 = P.b_h[1][sy - 2], ya1 = P.a[1][sy - 1], yb1 = P.b[1][sy - 1];
    const bool ylo_store = j + 1 == 2 || j + 1 == ny - h + 1;
    const long long ypinc = ld * (2 * h), yxinc = ld * (2 * (h + 1));
    long long ypo = (long long)k0 * ypinc + i0 + (long long)(sy - 2) * ld, yxo_ = (long long)k0 * yxinc + i0 + (long long)(sy - 1) * ld;

    // requests of one plane
    auto load_pre = [&](Pre &q, long long o, long long xp, long long xx, long long yp, long long yx) {
        q.po = ldv<T, V>(P.pold + o);
        q.fc = ldv<T, V>(P.fact + o);
        q.yu = ldv<T, V>(P.pcur + o - ld);
        q.yd = ldv<T, V>(P.pcur + o + ld);
        if (i0 > 0)
            q.xl = P.pcur[o - 1];
        if (i0 + V < nx)
            q.xr = P.pcur[o + V];
        if (KIND == 0) {
            const T *ps = P.psi_in[0] + xp, *xs = P.xi[0] + xx;
#pragma unroll
            for (int e = 0; e <= V; ++e)
                if ((xneed >> e) & 1)
                    q.psv[e] = ps[e];
#pragma unroll
            for (int v = 0; v < V; ++v)
                if ((xm >> v) & 1)
                    q.xiv[v] = xs[v];
        }
        if (KIND == 1 && sy > 0) {
            q.pl = ldv<T, V>(P.psi_in[1] + yp);
            q.ph = ldv<T, V>(P.psi_in[1] + yp + ld);
            q.xo = ldv<T, V>(P.xi[1] + yx);
        }
    };

    VT qm = ldv<T, V>(P.pcur + off - plane), qc = ldv<T, V>(P.pcur + off), qp = ldv<T, V>(P.pcur + off + plane);
    Pre A = {}, Bf = {};
    load_pre(A, off, xpo, xxo, ypo, yxo_);

    auto iter = [&](const int k, Pre &cur, Pre &nxt) {
        const bool more = k + 1 < k1; // the last plane re-requests itself (cache hits) instead of predicating every prefetch
        const long long dn = more ? plane : 0;
        // L1TEX returns a warp's loads in order: what this plane still has to fetch itself goes out before the prefetch of the next one
        VT c2 = {}, c1 = {}, c0 = {}, g = {};
        if (ADJ) {
            c2 = ldv<T, V>(P.pm2 + off);
            c1 = ldv<T, V>(P.pm1 + off);
            c0 = ldv<T, V>(P.p0 + off);
            g = ldv<T, V>(P.grad + off);
        }
        T cpsv[V + 1], cxiv[V]; // a y-strip box crossing an x strip: those memory variables on demand
        if (KIND != 0 && xm != 0) {
            const T *ps = P.psi_in[0] + xpo, *xs = P.xi[0] + xxo;
#pragma unroll
            for (int e = 0; e <= V; ++e)
                cpsv[e] = ((xneed >> e) & 1) ? ps[e] : (T)0;
#pragma unroll
            for (int v = 0; v < V; ++v)
                cxiv[v] = ((xm >> v) & 1) ? xs[v] : (T)0;
        }
        const VT qn = ldv<T, V>(P.pcur + off + dn + plane);
        load_pre(nxt, off + dn, xpo + (more ? xpinc : 0), xxo + (more ? xxinc : 0), ypo + (more ? ypinc : 0), yxo_ + (more ? yxinc : 0));
        const T xl_in = cur.xl, xr_in = cur.xr;
        CT pc[V], lap[V];
#pragma unroll
        for (int v = 0; v < V; ++v)
            pc[v] = (CT)qc.v[v];
        // ---- x term ----------------------------------------------------------------------------------------------------
        if (KIND == 0)
            cd_xstrip_vec<T, CT, V, FMA>(lap, w1, w2, qc, xl_in, xr_in, P.inv_d[0], xm, xfirst, xrec, cur.psv, cur.xiv, P.psi_out[0] + xpo, P.xi[0] + xxo);
        else if (xm != 0)
            cd_xstrip_vec<T, CT, V, FMA>(lap, w1, w2, qc, xl_in, xr_in, P.inv_d[0], xm, xfirst, xrec, cpsv, cxiv, P.psi_out[0] + xpo, P.xi[0] + xxo);
        else {
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const CT xl = (CT)(v > 0 ? qc.v[v > 0 ? v - 1 : 0] : xl_in);
                const CT xr = (CT)(v < V - 1 ? qc.v[v < V - 1 ? v + 1 : 0] : xr_in);
                lap[v] = d2<CT, FMA>(w2, xl, pc[v], xr, i2x);
            }
        }
        // ---- y term ----------------------------------------------------------------------------------------------------
        if (KIND != 0 && sy > 0) {
            const VT ph = cur.ph, pl = cur.pl, xo = cur.xo;
            VT psi_hi = ph, psi_lo = pl, xn = xo;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                CT t = d2<CT, FMA>(w2, (CT)cur.yu.v[v], pc[v], (CT)cur.yd.v[v], i2y);
                if ((okm >> v) & 1)
                    t = cd_axis_cell<T, CT>(w1, t, (CT)cur.yu.v[v], pc[v], (CT)cur.yd.v[v], P.inv_d[1], yah1, ybh1, yah0, ybh0, ya1, yb1, ph.v[v], pl.v[v], xo.v[v],
                                            psi_hi.v[v], psi_lo.v[v], xn.v[v]);
                lap[v] = lap[v] + t;
            }
            stv<T, V>(P.psi_out[1] + ypo + ld, psi_hi);
            if (ylo_store)
                stv<T, V>(P.psi_out[1] + ypo, psi_lo);
            stv<T, V>(P.xi[1] + yxo_, xn);
        } else {
#pragma unroll
            for (int v = 0; v < V; ++v)
                lap[v] = lap[v] + d2<CT, FMA>(w2, (CT)cur.yu.v[v], pc[v], (CT)cur.yd.v[v], i2y);
        }
        // ---- z term (plain on these planes), leapfrog ----------------------------------------------------------------------
        VT out = cur.po;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const CT t = d2<CT, FMA>(w2, (CT)qm.v[v], pc[v], (CT)qp.v[v], i2z);
            if ((okm >> v) & 1)
                out.v[v] = leapfrog<T, CT, FMA>(pc[v], cur.po.v[v], cur.fc.v[v], lap[v] + t);
        }
        finish(k, off, out, c2, c1, c0, g);
        qm = qc;
        qc = qp;
        qp = qn;
        off += plane;
        xpo += xpinc, xxo += xxinc;
        ypo += ypinc, yxo_ += yxinc;
    };
    int k = k0;
#pragma unroll 1
    for (; k + 1 < k1; k += 2) {
        iter(k, A, Bf);
        iter(k + 1, Bf, A);
    }
    if (k < k1)
        iter(k, A, Bf);
}

template <class T, class CT, bool ADJ, bool FMA>
__global__ void __launch_bounds__(CDF_RIM_T, CDF_RIMZ_MINB) cd_rimz_kernel(const CdFusedParams<T> P)
{
    constexpr int V = 16 / (int)sizeof(T);
    const int cta = (int)blockIdx.x;
    int bi = P.nbox_v; // (the z-plane boxes in front belong to the per-vector kernel)
#pragma unroll
    for (int b = 1; b < CDF_MAX_BOX; ++b)
        if (b > P.nbox_v && b < P.nbox && cta >= P.rt[b].cta0)
            bi = b;
    const CdBox &B = P.box[bi];
    const CdRimTile &R = P.rt[bi];
    const int local = cta - R.cta0;
    if ((P.rim_skip >> R.kind) & 1)
        return;
    extern __shared__ __align__(16) unsigned char cdf_smem[];
    T *xtab = reinterpret_cast<T *>(cdf_smem);
    cd_xstrip_table<T, V>(P, xtab, (int)threadIdx.x, CDF_RIM_T);
    __syncthreads();
    if (R.kind == 0)
        cd_rimz_body<T, CT, 0, ADJ, FMA>(P, B, R, local, cta, (int)threadIdx.x, xtab);
    else
        cd_rimz_body<T, CT, 1, ADJ, FMA>(P, B, R, local, cta, (int)threadIdx.x, xtab);
}

// Small 2D grids: both CTA kinds in one launch (the 2D bulk CTA and the rim CTA have 128 threads each).  A step of a grid that
// fills the GPU for a few microseconds is bound by launch and fork / join latency, not by bandwidth: one kernel node per step
// instead of two parallel ones plus the join.
static_assert(32 * CDF_W2D == CDF_RIM_T, "the merged 2D launch needs equal CTA sizes");
template <class T, class CT, bool ADJ, bool FMA>
__global__ void __launch_bounds__(CDF_RIM_T, sizeof(CT) == 4 ? 8 : 4) cd_merged2d_kernel(const CdFusedParams<T> P, const int nbulk, const int gdx)
{
    const int b = (int)blockIdx.x;
    if (b < nbulk)
        cd_bulk_body<T, CT, false, ADJ, FMA>(P, b % gdx, 0, P.rev ? nbulk / gdx - 1 - b / gdx : b / gdx, gdx, 1, CDF_W2D, (int)threadIdx.x & 31, (int)threadIdx.x >> 5);
    else
        cd_rim_body<T, CT, false, ADJ, FMA>(P, b - nbulk, (int)threadIdx.x, P.nbox, P.nrimvec);
}

} // namespace

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
CdFusedGeom cd_fused_geom(size_t esize, int nx, int ny, int nz, int halo, bool has_y, int zc, bool zpml_lo, bool zpml_hi, int rim_zc)
{
    CdFusedGeom g{};
    g.v = cdf_vec(esize);
    g.tx = cdf_tx(esize);
    g.has_y = has_y;
    g.ty = has_y ? CDF_TY : CDF_W2D;
    g.nx = nx, g.ny = ny, g.nz = nz;
    g.hs = std::max(halo, 1);
    g.zc = zc;
    const int nvec = (int)(cdf_ld(nx, esize) / g.v);
    g.ivlo = (g.hs + g.v - 1) / g.v;
    g.ivhi = std::max(g.ivlo, (nx - g.hs) / g.v);
    g.jlo = has_y ? g.hs : 0;
    g.jhi = has_y ? std::max(g.jlo, ny - g.hs) : 1;
    g.zpml_lo = zpml_lo, g.zpml_hi = zpml_hi;
    g.klo = zpml_lo ? g.hs : 1;
    g.khi = std::max(g.klo, nz - (zpml_hi ? g.hs : 1));
    g.ntx = (nx + g.tx - 1) / g.tx;
    g.nty = has_y ? (g.jhi - g.jlo + CDF_TY - 1) / CDF_TY : 1;
    g.ntz = (g.khi - g.klo + zc - 1) / zc;
    g.gx = has_y ? (unsigned)g.ntx : (unsigned)((g.ntx + CDF_W2D - 1) / CDF_W2D);
    g.gy = (unsigned)g.nty;
    g.gz = (unsigned)g.ntz;
    if (g.ivhi <= g.ivlo || g.nty == 0 || g.ntz == 0)
        g.gx = g.gy = g.gz = 0;
    // rim boxes: z slabs, then y slabs between them, then x slabs inside both
    long long start = 0;
    g.rim_zc = has_y && nx >= 2 * halo + g.v ? std::max(rim_zc, 0) : 0; // (a vector of the march never touches both x strips)
    long long cta0 = 0;
    auto add = [&](int iv0, int nvx, int j0, int nyb, int k0, int nzb, int kind) {
        if (nvx <= 0 || nyb <= 0 || nzb <= 0)
            return;
        SWB_REQUIRE((long long)nvx * nyb * nzb < (1ll << 31), "grid too large for the fused CD rim enumeration");
        CdRimTile &t = g.rt[g.nbox];
        CdBox &b = g.box[g.nbox++];
        b.iv0 = iv0, b.j0 = j0, b.k0 = k0, b.nvx = nvx, b.ny = nyb, b.nz = nzb, b.start = start;
        start += (long long)nvx * nyb * nzb;
        if (g.rim_zc > 0 && kind == 2) { // stays with the per-vector kernel (added first: its enumeration is a prefix of the box list)
            g.nbox_v = g.nbox;
            g.nrimvec_v = start;
            t = CdRimTile{};
            t.kind = 2;
            t.cta0 = 0x7fffffff;
        } else if (g.rim_zc > 0) { // march enumeration: tiles of columns x chunks of planes
            t.tw = std::min(nvx, 32);
            t.th = std::min(CDF_RIM_T / t.tw, nyb);
            t.ntx = (nvx + t.tw - 1) / t.tw;
            t.nty = (nyb + t.th - 1) / t.th;
            t.zc = kind == 2 ? std::min(nzb, 32) : std::min(nzb, g.rim_zc);
            t.ntz = (nzb + t.zc - 1) / t.zc;
            t.kind = kind;
            SWB_REQUIRE(cta0 + (long long)t.ntx * t.nty * t.ntz < (1ll << 31), "grid too large for the fused CD rim march");
            t.cta0 = (int)cta0;
            cta0 += (long long)t.ntx * t.nty * t.ntz;
        }
    };
    const int zlo = std::min(g.klo, nz), zhi = std::max(g.khi, zlo); // planes [0, zlo) and [zhi, nz) are rim
    add(0, nvec, 0, ny, 0, zlo, 2);
    add(0, nvec, 0, ny, zhi, nz - zhi, 2);
    if (has_y) {
        const int ylo = std::min(g.jlo, ny), yhi = std::max(g.jhi, ylo);
        add(0, nvec, 0, ylo, zlo, zhi - zlo, 1);
        add(0, nvec, yhi, ny - yhi, zlo, zhi - zlo, 1);
    }
    add(0, g.ivlo, g.jlo, g.jhi - g.jlo, zlo, zhi - zlo, 0);
    add(g.ivhi, nvec - g.ivhi, g.jlo, g.jhi - g.jlo, zlo, zhi - zlo, 0);
    g.nrimvec = start;
    g.nrimz_cta = (int)cta0;
    return g;
}

int cd_fused_locate(const CdFusedGeom &g, int i, int j, int k, int *cta, int *code)
{
    const int iv = i / g.v;
    const bool bulk = g.gx > 0 && iv >= g.ivlo && iv < g.ivhi && j >= g.jlo && j < g.jhi && k >= g.klo && k < g.khi;
    if (bulk) {
        const int xt = i / g.tx, xl = i % g.tx;
        const int bz = (k - g.klo) / g.zc, kl = (k - g.klo) % g.zc;
        int bx, by, ty;
        if (g.has_y) {
            bx = xt;
            by = (j - g.jlo) / CDF_TY;
            ty = (j - g.jlo) % CDF_TY;
        } else {
            bx = xt / CDF_W2D;
            by = 0;
            ty = xt % CDF_W2D;
        }
        *cta = (bz * (int)g.gy + by) * (int)g.gx + bx;
        *code = (kl * g.ty + ty) * g.tx + xl;
        return 0;
    }
    for (int b = 0; b < g.nbox; ++b) {
        const CdBox &B = g.box[b];
        if (iv >= B.iv0 && iv < B.iv0 + B.nvx && j >= B.j0 && j < B.j0 + B.ny && k >= B.k0 && k < B.k0 + B.nz) {
            if (g.rim_zc > 0 && b < g.nbox_v) { // z-plane box: per-vector kernel, CTAs numbered after the march's
                const long long lin = B.start + ((long long)(k - B.k0) * B.ny + (j - B.j0)) * B.nvx + (iv - B.iv0);
                *cta = g.nrimz_cta + (int)(lin / CDF_RIM_T);
                *code = (int)(lin % CDF_RIM_T) * g.v + i % g.v;
                return 1;
            }
            if (g.rim_zc > 0) {
                const CdRimTile &R = g.rt[b];
                const int ivl = iv - B.iv0, jl = j - B.j0, kl = k - B.k0;
                *cta = R.cta0 + ((kl / R.zc) * R.nty + jl / R.th) * R.ntx + ivl / R.tw;
                *code = ((kl % R.zc) * CDF_RIM_T + (jl % R.th) * R.tw + ivl % R.tw) * g.v + i % g.v;
                return 1;
            }
            const long long lin = B.start + ((long long)(k - B.k0) * B.ny + (j - B.j0)) * B.nvx + (iv - B.iv0);
            *cta = (int)(lin / CDF_RIM_T);
            *code = (int)(lin % CDF_RIM_T) * g.v + i % g.v;
            return 1;
        }
    }
    throw Error(SWB_ERR_STATE, "fused CD geometry: cell belongs to neither the bulk nor the rim");
}

template <class T>
void cd_fused_fill_geom(CdFusedParams<T> &P, const CdFusedGeom &g)
{
    P.zpml_lo = g.zpml_lo ? 1 : 0, P.zpml_hi = g.zpml_hi ? 1 : 0;
    P.jlo = g.jlo, P.jhi = g.jhi, P.klo = g.klo, P.khi = g.khi, P.ivlo = g.ivlo, P.ivhi = g.ivhi, P.zc = g.zc;
    P.nbox = g.nbox;
    for (int b = 0; b < g.nbox; ++b)
        P.box[b] = g.box[b];
    P.nrimvec = g.nrimvec;
    for (int b = 0; b < g.nbox; ++b)
        P.rt[b] = g.rt[b];
    P.nbox_v = g.rim_zc > 0 ? g.nbox_v : 0, P.nrimvec_v = g.nrimvec_v, P.rimv_cta0 = g.nrimz_cta;
}

template <class T>
void cd_fused_launch(const CdFusedParams<T> &P, const CdFusedGeom &g, bool adj, bool fast, cudaStream_t st, cudaStream_t st_rim, bool merged)
{
    SWB_REQUIRE(g.gy <= 65535 && g.gz <= 65535, "grid too large for the fused CD launch geometry");
    const dim3 grd(g.gx, g.gy, g.gz), blk(32, g.ty, 1);
    const unsigned nrim = (unsigned)g.ncta_rim();
    const bool f32fast = sizeof(T) == 4 && fast, has_y = g.has_y;
    // x-strip coefficient table of the marching rim kernel: one record of 5 V values per strip vector
    const size_t rimz_smem = (size_t)(g.ivlo + ((int)(cdf_ld(g.nx, sizeof(T)) / g.v) - g.ivhi)) * 5 * g.v * sizeof(T);
    if (merged && !has_y && g.gx > 0 && nrim > 0) {
        const int nbulk = g.gx * g.gz;
#define SWB_CDF_M(CT, AD, FM)                                                                            \
    cd_merged2d_kernel<T, CT, AD, FM><<<(unsigned)nbulk + nrim, CDF_RIM_T, 0, st>>>(P, nbulk, g.gx)
        if (f32fast) {
            if (adj)
                SWB_CDF_M(T, true, true);
            else
                SWB_CDF_M(T, false, true);
        } else {
            if (adj)
                SWB_CDF_M(double, true, false);
            else
                SWB_CDF_M(double, false, false);
        }
#undef SWB_CDF_M
        check_launch("cd_merged2d_kernel");
        count_launch();
        return;
    }
#define SWB_CDF_GO(CT, HY, AD, FM)                                                      \
    do {                                                                                \
        if (g.gx > 0) {                                                                 \
            cd_bulk_kernel<T, CT, HY, AD, FM><<<grd, blk, 0, st>>>(P);                  \
            check_launch("cd_bulk_kernel");                                             \
            count_launch();                                                             \
        }                                                                               \
        if (nrim > 0 && HY && g.rim_zc > 0) {                                           \
            if (g.ncta_rimv() > 0) {                                                    \
                cd_rim_kernel<T, CT, HY, AD, FM><<<g.ncta_rimv(), CDF_RIM_T, 0, st_rim>>>(P); \
                check_launch("cd_rim_kernel");                                          \
                count_launch();                                                         \
            }                                                                           \
            if (g.nrimz_cta > 0) {                                                      \
                cd_rimz_kernel<T, CT, AD, FM><<<g.nrimz_cta, CDF_RIM_T, rimz_smem, st_rim>>>(P); \
                check_launch("cd_rimz_kernel");                                         \
                count_launch();                                                         \
            }                                                                           \
        } else if (nrim > 0) {                                                          \
            cd_rim_kernel<T, CT, HY, AD, FM><<<nrim, CDF_RIM_T, 0, st_rim>>>(P);        \
            check_launch("cd_rim_kernel");                                              \
            count_launch();                                                             \
        }                                                                               \
    } while (0)
#define SWB_CDF_CT(CT, FM)                    \
    do {                                      \
        if (has_y) {                          \
            if (adj)                          \
                SWB_CDF_GO(CT, true, true, FM);   \
            else                              \
                SWB_CDF_GO(CT, true, false, FM);  \
        } else {                              \
            if (adj)                          \
                SWB_CDF_GO(CT, false, true, FM);  \
            else                              \
                SWB_CDF_GO(CT, false, false, FM); \
        }                                     \
    } while (0)
    if (f32fast)
        SWB_CDF_CT(T, true);
    else
        SWB_CDF_CT(double, false);
#undef SWB_CDF_CT
#undef SWB_CDF_GO
}

template void cd_fused_fill_geom<float>(CdFusedParams<float> &, const CdFusedGeom &);
template void cd_fused_fill_geom<double>(CdFusedParams<double> &, const CdFusedGeom &);
template void cd_fused_launch<float>(const CdFusedParams<float> &, const CdFusedGeom &, bool, bool, cudaStream_t, cudaStream_t, bool);
template void cd_fused_launch<double>(const CdFusedParams<double> &, const CdFusedGeom &, bool, bool, cudaStream_t, cudaStream_t, bool);

} // namespace swb
