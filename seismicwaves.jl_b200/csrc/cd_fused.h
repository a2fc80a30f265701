// cd_fused.h -- parameter block and layout constants of the fused acoustic constant-density step (acou_cd_fused.cu).
#pragma once
#include "common.cuh"

namespace swb {

constexpr int CDF_TY = 8;      // rows of a 3D bulk tile (one warp per row)
constexpr int CDF_W2D = 4;     // warps per CTA in the 2D bulk variant (each warp owns its own x range)
constexpr int CDF_RIM_T = 128; // threads (= vectors) per CTA of the rim kernel
constexpr int CDF_MAX_BOX = 6;
constexpr int CDF_RIM_ZC = 16; // planes per chunk of the 3D rim march (engine default; SWB_CDF_RIM_ZC overrides, 0 = per-vector rim kernel only)

// elements per 16-byte vector and cells per warp-row
inline int cdf_vec(size_t esize) { return (int)(16 / esize); }
inline int cdf_tx(size_t esize) { return 32 * cdf_vec(esize); }
// row pitch (elements) of an engine-owned CD field: rows start 16-byte aligned
inline long long cdf_ld(long long nx, size_t esize) { const long long v = cdf_vec(esize); return (nx + v - 1) / v * v; }

// The step is split over two kernels that write disjoint cells and may run concurrently:
//   bulk: every 16-byte vector of cells that contains no C-PML strip cell and no face cell, on rows / planes that
//         are neither -- plain Laplacian, register-queue march along z;
//   rim : all other vectors (strips and faces), enumerated compactly as up to six boxes of vectors.
// Both see a 3D grid (nx, ny, nz).  A 2D simulation (nx, ny2d) is presented as (nx, 1, ny2d): the y term of the
// Laplacian is compiled out (HAS_Y = false) and the reference's y axis rides on z.
struct CdBox {
    int iv0, j0, k0;    // origin: vector index along x, row, plane
    int nvx, ny, nz;    // extent in vectors / rows / planes
    long long start;    // index of the box's first vector in the rim enumeration
};

// March enumeration of a rim box (3D): a CTA owns a tile of tw x th (x-vector, row) columns and marches them over zc planes.
// kind = the strip axis whose memory variables the march keeps one plane ahead in registers (0: x strips, 1: y strips, 2: the z
// planes); strips of the other axes that cross the box (corners) are handled by the generic on-demand code.
struct CdRimTile {
    int tw, th;        // tile width (vectors) and height (rows), tw * th <= CDF_RIM_T
    int ntx, nty, ntz; // tiles along x and y, chunks along z
    int zc;            // planes per chunk
    int cta0;          // first CTA of the box
    int kind;
};

struct CdPointList {
    const int *off, *cell, *idx; // per-CTA CSR: entries [off[cta], off[cta+1]) = (in-CTA cell code, point index)
};

template <class T>
struct CdFusedParams {
    int nx, ny, nz, halo;
    int zpml_lo, zpml_hi;  // 0: that z end is an interior face of a z-slab decomposition (one ghost plane, no C-PML strip)
    // z-slab decomposition with the halo exchange fused into the step (peer memory over NVLink): the freshly computed first / last
    // owned plane (k = 1 / k = nz - 2) is also stored into the neighbour's ghost plane -- peer_lo / peer_hi point at that plane of
    // the neighbour's pnew (null: no neighbour, or the exchange is done by NCCL after the step).  ghost_lo / ghost_hi: plane 0 /
    // nz - 1 is a ghost plane the neighbour owns; this slab's kernels leave it alone.
    T *peer_lo, *peer_hi;
    int ghost_lo, ghost_hi;
    int rim_skip;     // timing experiments only: bit k set = the marching rim kernel skips boxes of kind k (wrong results)
    int rim_prefetch; // 1: rim threads prefetch their C-PML memory variables before the first dependent use
    int rev; // 1: the bulk z chunks are issued from the last to the first (serpentine sweep: a step starts on the planes the previous one left in L2)
    long long ld, plane;   // row pitch and plane pitch (elements) of pcur / pold / pnew / fact / grad / stored fields
    T inv_d[3];            // 1 / spacing along kernel axes x, y, z
    const T *pcur, *pold, *fact;
    T *pnew;               // may alias pold
    // C-PML memory variables: psi_x (2h, ny, nz) as in the reference; psi_y (ld, 2h, nz) and psi_z (ld, ny, 2h) with the
    // fields' row pitch so that a thread's V cells move as one aligned vector; xi_* likewise with 2(h+1).
    // psi is double-buffered (a cell needs the new psi of its lower neighbour, which that
    // neighbour's thread computes too); xi is updated in place (one owner per entry).
    const T *psi_in[3];
    T *psi_out[3];
    T *xi[3];
    const T *a[3], *b[3], *a_h[3], *b_h[3];
    double c1[2], c2[3];
    // injection: pnew[cell] += inj_tf[inj_it, idx]; inj_it = 0 disables.  recording: traces[rec_it, idx] = pnew[cell].
    // One CSR per kernel (index 0: bulk, 1: rim).
    CdPointList inj[2], rec[2];
    const T *inj_tf;
    T *traces;
    long long inj_nt, rec_nt;
    int inj_it, rec_it;
    // adjoint-mode zero-lag correlation: grad += pnew * (pm2 - 2 pm1 + p0) * inv_dt2
    const T *pm2, *pm1, *p0;
    T *grad;
    T inv_dt2;
    // bulk geometry: rows [jlo, jhi), planes [klo, khi) in chunks of zc, x-vectors [ivlo, ivhi)
    int jlo, jhi, klo, khi, ivlo, ivhi, zc;
    // rim geometry
    int nbox;
    CdBox box[CDF_MAX_BOX];
    long long nrimvec;
    CdRimTile rt[CDF_MAX_BOX]; // march enumeration of the x / y strip boxes (cd_rimz_kernel)
    // 3D with the march: the z-plane boxes (the first nbox_v boxes, nrimvec_v vectors) stay with the one-thread-per-vector kernel -- whole
    // planes have all the parallelism it wants, and a march over the 20 planes of a strip is latency-bound (768^3: 0.15 ms against 0.10);
    // its CTAs are numbered after the march's in the point lists (rimv_cta0)
    int nbox_v, rimv_cta0;
    long long nrimvec_v;
};

// host-side geometry shared by the launcher and the code that builds the per-CTA point lists
struct CdFusedGeom {
    int v, tx, ty, ntx, nty, ntz, zc;
    bool has_y;
    int nx, ny, nz, hs;                 // hs = max(halo, 1): thickness of the rim along each axis
    bool zpml_lo, zpml_hi;              // false: slab-interior z end (rim = the ghost plane only)
    int jlo, jhi, klo, khi, ivlo, ivhi; // bulk ranges
    unsigned gx, gy, gz;                // bulk launch grid (0 blocks if the bulk is empty)
    int nbox;
    CdBox box[CDF_MAX_BOX];
    long long nrimvec;
    int rim_zc = 0;                     // > 0: the rim is marched along z in chunks of rim_zc planes (cd_rimz_kernel), 0: one thread per vector
    CdRimTile rt[CDF_MAX_BOX];
    int nrimz_cta = 0;                  // CTAs of the march (x / y strip boxes)
    int nbox_v = 0;                     // with the march: the leading z-plane boxes, run by the per-vector kernel
    long long nrimvec_v = 0;
    int ncta_bulk() const { return (int)(gx * gy * gz); }
    int ncta_rimv() const { return (int)(((rim_zc > 0 ? nrimvec_v : nrimvec) + CDF_RIM_T - 1) / CDF_RIM_T); }
    int ncta_rim() const { return rim_zc > 0 ? nrimz_cta + ncta_rimv() : ncta_rimv(); }
};
CdFusedGeom cd_fused_geom(size_t esize, int nx, int ny, int nz, int halo, bool has_y, int zc, bool zpml_lo = true, bool zpml_hi = true, int rim_zc = 0);
// which kernel owns the 0-based cell (i, j, k): returns 0 (bulk) or 1 (rim), the CTA index in launch order and the packed in-CTA cell code
int cd_fused_locate(const CdFusedGeom &g, int i, int j, int k, int *cta, int *code);
template <class T>
void cd_fused_fill_geom(CdFusedParams<T> &P, const CdFusedGeom &g);

// launches the bulk kernel on `st` and the rim kernel on `st_rim` (may equal st)
template <class T>
void cd_fused_launch(const CdFusedParams<T> &P, const CdFusedGeom &g, bool adj, bool fast, cudaStream_t st, cudaStream_t st_rim, bool merged = false);

} // namespace swb
