// capi.cu -- the extern "C" surface declared in include/swb200.h.
#include "engine.h"
#include "cd_fused.h"
#include <unordered_set>
#include <dlfcn.h>
#include <nccl.h>
#include <cstring>

namespace swb {
const char *last_error_cstr();
}
using namespace swb;

template <class T>
__global__ void fill_kernel(T *p, T v, size_t n)
{
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x)
        p[q] = v;
}


extern "C" {

const char *swb_last_error(void) { return last_error_cstr(); }
int32_t swb_abi_version(void) { return SWB_ABI_VERSION; }

int32_t swb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10)
            ++ok;
    }
    return ok;
}

int64_t swb_launch_count(void) { return (int64_t)g_launches.load(); }

int32_t swb_diag_cd_partition(int32_t esize, int32_t nx, int32_t ny, int32_t nz, int32_t halo, int32_t zc, int32_t zpml_lo, int32_t zpml_hi,
                              int32_t rim_zc, int64_t *counts)
{
    SWB_API_BEGIN
    SWB_REQUIRE((esize == 4 || esize == 8) && nx >= 3 && ny >= 1 && nz >= 3 && halo >= 0 && zc >= 1 && counts != nullptr, "bad arguments");
    const bool has_y = ny > 1;
    const CdFusedGeom g = cd_fused_geom((size_t)esize, nx, ny, nz, halo, has_y, zc, zpml_lo != 0, zpml_hi != 0, rim_zc < 0 ? CDF_RIM_ZC : rim_zc);
    counts[0] = counts[1] = counts[2] = 0;
    std::unordered_set<unsigned long long> seen;
    seen.reserve((size_t)nx * ny * nz * 2);
    const long long bulk_slots = (long long)g.zc * g.ty * g.tx, vec_slots = (long long)CDF_RIM_T * g.v;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                int cta = -1, code = -1;
                const int which = cd_fused_locate(g, i, j, k, &cta, &code);
                long long slots;
                int kind;
                if (which == 0) {
                    SWB_REQUIRE(cta >= 0 && cta < g.ncta_bulk(), "bulk CTA out of range");
                    slots = bulk_slots, kind = 0;
                } else {
                    SWB_REQUIRE(cta >= 0 && cta < g.ncta_rim(), "rim CTA out of range");
                    const bool marched = g.rim_zc > 0 && cta < g.nrimz_cta;
                    kind = marched ? 1 : 2;
                    slots = vec_slots;
                    if (marched) { // the chunk length of the box this CTA belongs to
                        int b = g.nbox_v;
                        for (int q = g.nbox_v + 1; q < g.nbox; ++q)
                            if (cta >= g.rt[q].cta0)
                                b = q;
                        slots = (long long)g.rt[b].zc * vec_slots;
                    }
                }
                SWB_REQUIRE(code >= 0 && code < slots, "in-CTA slot out of range");
                const unsigned long long key = ((unsigned long long)which << 62) | ((unsigned long long)cta << 26) | (unsigned long long)code;
                SWB_REQUIRE(code < (1 << 26), "slot does not fit the check's key");
                SWB_REQUIRE(seen.insert(key).second, "two cells share one (kernel, CTA, slot)");
                counts[kind]++;
            }
    SWB_API_END
}

// ---- 1. device buffers --------------------------------------------------------------------------
int32_t swb_set_device(int32_t device)
{
    SWB_API_BEGIN
    SWB_CUDA(cudaSetDevice(device));
    SWB_API_END
}

int32_t swb_malloc(void **dev_ptr, size_t nbytes)
{
    SWB_API_BEGIN
    SWB_REQUIRE(dev_ptr != nullptr, "null output pointer");
    *dev_ptr = nullptr;
    if (nbytes == 0)
        return SWB_OK;
    cudaError_t e = cudaMalloc(dev_ptr, nbytes);
    if (e != cudaSuccess)
        throw Error(e == cudaErrorMemoryAllocation ? SWB_ERR_NOMEM : SWB_ERR_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    SWB_CUDA(cudaMemset(*dev_ptr, 0, nbytes));
    SWB_API_END
}

int32_t swb_free(void *dev_ptr)
{
    SWB_API_BEGIN
    if (dev_ptr)
        SWB_CUDA(cudaFree(dev_ptr));
    SWB_API_END
}

int32_t swb_memcpy_h2d(void *dst, const void *src, size_t nbytes)
{
    SWB_API_BEGIN
    if (nbytes)
        SWB_CUDA(cudaMemcpy(dst, src, nbytes, cudaMemcpyHostToDevice));
    SWB_API_END
}

int32_t swb_memcpy_d2h(void *dst, const void *src, size_t nbytes)
{
    SWB_API_BEGIN
    if (nbytes)
        SWB_CUDA(cudaMemcpy(dst, src, nbytes, cudaMemcpyDeviceToHost));
    SWB_API_END
}

int32_t swb_memcpy_d2d(void *dst, const void *src, size_t nbytes)
{
    SWB_API_BEGIN
    if (nbytes)
        SWB_CUDA(cudaMemcpy(dst, src, nbytes, cudaMemcpyDeviceToDevice));
    SWB_API_END
}

int32_t swb_fill(void *dev_ptr, int32_t dtype, double value, size_t first, size_t count)
{
    SWB_API_BEGIN
    SWB_REQUIRE(dtype == SWB_F32 || dtype == SWB_F64, "dtype must be SWB_F32 or SWB_F64");
    if (count) {
        unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, 148u * 32u);
        if (dtype == SWB_F64)
            fill_kernel<double><<<blocks, 256>>>((double *)dev_ptr + first, value, count);
        else
            fill_kernel<float><<<blocks, 256>>>((float *)dev_ptr + first, (float)value, count);
        check_launch("fill");
        count_launch();
    }
    SWB_API_END
}

int32_t swb_synchronize(void)
{
    SWB_API_BEGIN
    SWB_CUDA(cudaDeviceSynchronize());
    SWB_API_END
}

// ---- 2. backend-module functions ------------------------------------------------------------------
int32_t swb_acou_cd_forward_onestep(const swb_acou_cd_step_args *a)
{
    SWB_API_BEGIN
    SWB_REQUIRE(a != nullptr, "null argument struct");
    cd_step(*a, a->rec.n > 0 && a->rec.tf != nullptr);
    SWB_API_END
}

int32_t swb_acou_cd_adjoint_onestep(const swb_acou_cd_step_args *a)
{
    SWB_API_BEGIN
    SWB_REQUIRE(a != nullptr, "null argument struct");
    cd_step(*a, false);
    SWB_API_END
}

int32_t swb_acou_cd_correlate_gradient(int32_t dtype, int32_t flags, size_t ncells, void *grad, const void *adjcur, const void *p_itm2,
                                       const void *p_itm1, const void *p_it, double dt, void *stream)
{
    SWB_API_BEGIN
    SWB_REQUIRE(dtype == SWB_F32 || dtype == SWB_F64, "dtype must be SWB_F32 or SWB_F64");
    cd_correlate(dtype, flags, ncells, grad, adjcur, p_itm2, p_itm1, p_it, dt, (cudaStream_t)stream);
    SWB_API_END
}

int32_t swb_prescale_residuals(int32_t dtype, int32_t ndim, const int64_t *n, void *residuals, int64_t nt, int64_t nrec, const int64_t *posrecs,
                               const void *fact, void *stream)
{
    SWB_API_BEGIN
    SWB_REQUIRE(dtype == SWB_F32 || dtype == SWB_F64, "dtype must be SWB_F32 or SWB_F64");
    SWB_REQUIRE(ndim >= 1 && ndim <= 3, "ndim must be 1..3");
    prescale_residuals(dtype, ndim, n, residuals, nt, nrec, posrecs, fact, (cudaStream_t)stream);
    SWB_API_END
}

int32_t swb_acou_vd_forward_onestep(const swb_acou_vd_step_args *a)
{
    SWB_API_BEGIN
    SWB_REQUIRE(a != nullptr, "null argument struct");
    vd_step(*a, false);
    SWB_API_END
}

int32_t swb_acou_vd_adjoint_onestep(const swb_acou_vd_step_args *a)
{
    SWB_API_BEGIN
    SWB_REQUIRE(a != nullptr, "null argument struct");
    vd_step(*a, true);
    SWB_API_END
}

int32_t swb_acou_vd_correlate_gradient_m0(int32_t dtype, int32_t, size_t ncells, void *grad_m0, const void *adjp, const void *p_it,
                                          const void *p_itm1, double dt, void *stream)
{
    SWB_API_BEGIN
    SWB_REQUIRE(dtype == SWB_F32 || dtype == SWB_F64, "dtype must be SWB_F32 or SWB_F64");
    vd_correlate_m0(dtype, ncells, grad_m0, adjp, p_it, p_itm1, dt, (cudaStream_t)stream);
    SWB_API_END
}

int32_t swb_acou_vd_correlate_gradient_m1(int32_t dtype, int32_t flags, const int64_t *n, const double *spacing, void *const grad_m1_stag[2],
                                          const void *const adjv[2], const void *p_it, void *stream)
{
    SWB_API_BEGIN
    SWB_REQUIRE(dtype == SWB_F32 || dtype == SWB_F64, "dtype must be SWB_F32 or SWB_F64");
    vd_correlate_m1(dtype, flags, n, spacing, grad_m1_stag, adjv, p_it, (cudaStream_t)stream);
    SWB_API_END
}

int32_t swb_ela_forward_onestep(const swb_ela_step_args *a)
{
    SWB_API_BEGIN
    SWB_REQUIRE(a != nullptr, "null argument struct");
    ela_step(*a, false);
    SWB_API_END
}

int32_t swb_ela_adjoint_onestep(const swb_ela_step_args *a)
{
    SWB_API_BEGIN
    SWB_REQUIRE(a != nullptr, "null argument struct");
    ela_step(*a, true);
    SWB_API_END
}

int32_t swb_ela_correlate_gradients(const swb_ela_correlate_args *a)
{
    SWB_API_BEGIN
    SWB_REQUIRE(a != nullptr, "null argument struct");
    ela_correlate(*a);
    SWB_API_END
}

// ---- 3. per-shot engine -----------------------------------------------------------------------------
// ---- ABI layout report ---------------------------------------------------------------------------------------------------
namespace {
struct LayoutWriter {
    std::string s = "{";
    bool first_struct = true, first_field = true;
    void begin(const char *name, size_t size)
    {
        s += std::string(first_struct ? "" : ", ") + "\"" + name + "\": {\"size\": " + std::to_string(size) + ", \"fields\": [";
        first_struct = false;
        first_field = true;
    }
    void field(const char *name, size_t off, size_t size)
    {
        s += std::string(first_field ? "" : ", ") + "[\"" + name + "\", " + std::to_string(off) + ", " + std::to_string(size) + "]";
        first_field = false;
    }
    void end() { s += "]}"; }
};
#define SWB_L_BEGIN(S) { typedef S cur_t; w.begin(#S, sizeof(S));
#define SWB_L_F(m) w.field(#m, offsetof(cur_t, m), sizeof(((cur_t *)nullptr)->m));
#define SWB_L_END w.end(); }
std::string build_layout()
{
    LayoutWriter w;
    SWB_L_BEGIN(swb_cpml_axis) SWB_L_F(a) SWB_L_F(a_h) SWB_L_F(b) SWB_L_F(b_h) SWB_L_END
    SWB_L_BEGIN(swb_points) SWB_L_F(n) SWB_L_F(pos) SWB_L_F(tf) SWB_L_F(nt) SWB_L_END
    SWB_L_BEGIN(swb_acou_cd_step_args)
    SWB_L_F(dtype) SWB_L_F(ndim) SWB_L_F(halo) SWB_L_F(flags) SWB_L_F(n) SWB_L_F(spacing) SWB_L_F(pold) SWB_L_F(pcur) SWB_L_F(pnew) SWB_L_F(fact) SWB_L_F(psi) SWB_L_F(xi)
    SWB_L_F(cpml) SWB_L_F(src) SWB_L_F(rec) SWB_L_F(it) SWB_L_F(stream) SWB_L_END
    SWB_L_BEGIN(swb_acou_vd_step_args)
    SWB_L_F(dtype) SWB_L_F(halo) SWB_L_F(flags) SWB_L_F(_pad) SWB_L_F(n) SWB_L_F(spacing) SWB_L_F(pcur) SWB_L_F(vcur) SWB_L_F(fact_m0) SWB_L_F(fact_m1_stag) SWB_L_F(psi)
    SWB_L_F(xi) SWB_L_F(cpml) SWB_L_F(src) SWB_L_F(rec) SWB_L_F(it) SWB_L_F(stream) SWB_L_END
    SWB_L_BEGIN(swb_sinc_points) SWB_L_F(n) SWB_L_F(off) SWB_L_F(ij) SWB_L_F(coef) SWB_L_END
    SWB_L_BEGIN(swb_ela_step_args)
    SWB_L_F(dtype) SWB_L_F(halo) SWB_L_F(flags) SWB_L_F(freetop) SWB_L_F(n) SWB_L_F(spacing) SWB_L_F(dt) SWB_L_F(uold) SWB_L_F(ucur) SWB_L_F(unew) SWB_L_F(sigma)
    SWB_L_F(lambda) SWB_L_F(mu) SWB_L_F(rho_ihalf) SWB_L_F(rho_jhalf) SWB_L_F(mu_ihalf_jhalf) SWB_L_F(psi_dsdx) SWB_L_F(psi_dsdz) SWB_L_F(psi_dudx) SWB_L_F(psi_dudz)
    SWB_L_F(cpml) SWB_L_F(src_kind) SWB_L_F(_pad) SWB_L_F(src_pts) SWB_L_F(srctf) SWB_L_F(nt_tf) SWB_L_F(Mxx) SWB_L_F(Mzz) SWB_L_F(Mxz) SWB_L_F(rec_pts) SWB_L_F(traces)
    SWB_L_F(nt_tr) SWB_L_F(it) SWB_L_F(stream) SWB_L_END
    SWB_L_BEGIN(swb_ela_correlate_args)
    SWB_L_F(dtype) SWB_L_F(flags) SWB_L_F(freetop) SWB_L_F(_pad) SWB_L_F(n) SWB_L_F(spacing) SWB_L_F(dt) SWB_L_F(adjucur) SWB_L_F(u_itm2) SWB_L_F(u_itm1) SWB_L_F(u_it)
    SWB_L_F(lambda) SWB_L_F(mu) SWB_L_F(grad_rho_ihalf) SWB_L_F(grad_rho_jhalf) SWB_L_F(grad_lambda) SWB_L_F(grad_mu) SWB_L_F(grad_mu_ihalf_jhalf) SWB_L_F(stream) SWB_L_END
    SWB_L_BEGIN(swb_sim_desc)
    SWB_L_F(kind) SWB_L_F(dtype) SWB_L_F(ndim) SWB_L_F(device) SWB_L_F(n) SWB_L_F(spacing) SWB_L_F(dt) SWB_L_F(nt) SWB_L_F(halo) SWB_L_F(freetop) SWB_L_F(gradient)
    SWB_L_F(check_freq) SWB_L_F(flags) SWB_L_F(_pad) SWB_L_END
    SWB_L_BEGIN(swb_sinc_points_host) SWB_L_F(n) SWB_L_F(off) SWB_L_F(ij) SWB_L_F(coef) SWB_L_END
    SWB_L_BEGIN(swb_l2_spec) SWB_L_F(observed) SWB_L_F(invcov_diag) SWB_L_F(mask) SWB_L_END
    SWB_L_BEGIN(swb_slab_handle) SWB_L_F(ipc) SWB_L_F(raw) SWB_L_F(nz) SWB_L_F(plane_elems) SWB_L_F(device) SWB_L_F(pid) SWB_L_END
    w.s += "}";
    return w.s;
}
#undef SWB_L_BEGIN
#undef SWB_L_F
#undef SWB_L_END
} // namespace

const char *swb_abi_layout(void)
{
    static const std::string layout = build_layout();
    return layout.c_str();
}

int32_t swb_sim_create(const swb_sim_desc *desc, swb_sim **sim)
{
    SWB_API_BEGIN
    SWB_REQUIRE(desc != nullptr && sim != nullptr, "null argument");
    *sim = nullptr;
    SimBase *impl = nullptr;
    switch (desc->kind) {
    case SWB_ACOU_CD:
        impl = make_acoustic_cd(*desc);
        break;
    case SWB_ACOU_VD:
        impl = make_acoustic_vd(*desc);
        break;
    case SWB_ELA_ISO:
        impl = make_elastic_iso(*desc);
        break;
    default:
        throw Error(SWB_ERR_ARG, "unknown simulation kind");
    }
    swb_sim *s = new swb_sim;
    s->impl.reset(impl);
    *sim = s;
    SWB_API_END
}

int32_t swb_sim_destroy(swb_sim *sim)
{
    SWB_API_BEGIN
    delete sim;
    SWB_API_END
}

int64_t swb_sim_device_bytes(const swb_sim *sim) { return sim ? sim->impl->device_bytes() : 0; }

#define SIM_CALL(expr)                                        \
    SWB_API_BEGIN                                             \
    SWB_REQUIRE(sim != nullptr && sim->impl, "null sim");    \
    expr;                                                     \
    SWB_API_END

int32_t swb_sim_set_material(swb_sim *sim, int32_t nfields, const void *const *host_fields, int32_t interp)
{
    SIM_CALL(sim->impl->set_material(nfields, host_fields, interp, false))
}
int32_t swb_sim_set_material_device(swb_sim *sim, int32_t nfields, const void *const *dev_fields, int32_t interp)
{
    SIM_CALL(sim->impl->set_material(nfields, dev_fields, interp, true))
}
int32_t swb_sim_set_cpml(swb_sim *sim, int32_t axis, const void *a, const void *a_h, const void *b, const void *b_h)
{
    SIM_CALL(sim->impl->set_cpml(axis, a, a_h, b, b_h))
}
int32_t swb_sim_bind_scalar_shot(swb_sim *sim, int64_t nsrc, const int64_t *possrcs, const void *srctf, int64_t nrec, const int64_t *posrecs)
{
    SIM_CALL(sim->impl->bind_scalar_shot(nsrc, possrcs, srctf, nrec, posrecs))
}
int32_t swb_sim_bind_elastic_shot(swb_sim *sim, int32_t src_kind, const swb_sinc_points_host src_pts[2], const void *srctf, const void *Mxx,
                                  const void *Mzz, const void *Mxz, const swb_sinc_points_host rec_pts[2])
{
    SIM_CALL(sim->impl->bind_elastic_shot(src_kind, src_pts, srctf, Mxx, Mzz, Mxz, rec_pts))
}
int32_t swb_sim_forward(swb_sim *sim, void *host_seismograms, int32_t snapevery) { SIM_CALL(sim->impl->forward(host_seismograms, snapevery)) }
int32_t swb_sim_get_snapshot(swb_sim *sim, int64_t it, int32_t field, void *host_out) { SIM_CALL(sim->impl->get_snapshot(it, field, host_out)) }
int32_t swb_sim_gradient_forward(swb_sim *sim, void *host_seismograms) { SIM_CALL(sim->impl->gradient_forward(host_seismograms)) }
int32_t swb_sim_gradient_adjoint(swb_sim *sim, const void *host_adjsrc) { SIM_CALL(sim->impl->gradient_adjoint(host_adjsrc)) }
int32_t swb_sim_gradient_l2(swb_sim *sim, const void *host_observed, void *host_seismograms_or_null, double *misfit_out)
{
    swb_l2_spec spec = {host_observed, nullptr, nullptr};
    SIM_CALL(sim->impl->gradient_l2(spec, host_seismograms_or_null, misfit_out))
}
int32_t swb_sim_gradient_l2_ex(swb_sim *sim, const swb_l2_spec *spec, void *host_seismograms_or_null, double *misfit_out)
{
    SIM_CALL(SWB_REQUIRE(spec != nullptr, "null L2 spec"); sim->impl->gradient_l2(*spec, host_seismograms_or_null, misfit_out))
}
int32_t swb_sim_get_raw_gradient(swb_sim *sim, int32_t which, void *host_out) { SIM_CALL(sim->impl->get_raw_gradient(which, host_out)) }
int32_t swb_sim_accumulate_gradient(swb_sim *sim, int64_t nsrcpos, const void *src_positions, int32_t mute_radius_src, int64_t nrecpos,
                                    const void *rec_positions, int32_t mute_radius_rec)
{
    SIM_CALL(sim->impl->accumulate_gradient(nsrcpos, src_positions, mute_radius_src, nrecpos, rec_positions, mute_radius_rec))
}
int32_t swb_sim_zero_total_gradient(swb_sim *sim) { SIM_CALL(sim->impl->zero_total_gradient()) }
int32_t swb_sim_total_gradient_ptr(swb_sim *sim, int32_t which, void **dev_ptr, size_t *nelem)
{
    SIM_CALL(sim->impl->total_gradient_ptr(which, dev_ptr, nelem))
}
int32_t swb_sim_get_total_gradient(swb_sim *sim, int32_t which, void *host_out) { SIM_CALL(sim->impl->get_total_gradient(which, host_out)) }
int64_t swb_sim_cell_updates(const swb_sim *sim) { return sim ? sim->impl->cell_updates : 0; }
int32_t swb_sim_get_field(swb_sim *sim, const char *name, void *host_out, size_t nbytes)
{
    SIM_CALL(sim->impl->get_field(std::string(name ? name : ""), host_out, nbytes))
}
int32_t swb_sim_stream(swb_sim *sim, void **stream_out) { SIM_CALL(*stream_out = (void *)sim->impl->stream) }
int32_t swb_sim_kernel_timing(swb_sim *sim, int32_t enable, double *ms_total, int64_t *launches)
{
    SIM_CALL(sim->impl->kernel_timing(enable, ms_total, launches))
}

int32_t swb_sim_kernel_timing_class(swb_sim *sim, int32_t cls, double *ms_total, int64_t *launches)
{
    SIM_CALL(sim->impl->kernel_timing_class(cls, ms_total, launches))
}

// ---- 4. multi-GPU (NCCL resolved at run time so that the library loads on hosts without it) ----------
struct swb_comm {
    ncclComm_t comm = nullptr;
    int device = 0;
    int nranks = 1, rank = 0;
};

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi &nccl()
{
    static NcclApi api = [] {
        NcclApi a;
        a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!a.h)
            a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (a.h) {
            a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
            a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
            a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
            a.AllReduce = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
            a.Send = (decltype(a.Send))dlsym(a.h, "ncclSend");
            a.Recv = (decltype(a.Recv))dlsym(a.h, "ncclRecv");
            a.GroupStart = (decltype(a.GroupStart))dlsym(a.h, "ncclGroupStart");
            a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.h, "ncclGroupEnd");
            a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
        }
        return a;
    }();
    if (!api.h || !api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce)
        throw Error(SWB_ERR_NCCL, "libnccl.so.2 could not be loaded");
    return api;
}
void nccl_check(ncclResult_t r, const char *what)
{
    if (r != ncclSuccess)
        throw Error(SWB_ERR_NCCL, std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(r) : "NCCL error"));
}
} // namespace

} // extern "C"

// halo exchange of a z-slab decomposition: one plane to / from each neighbour, all four transfers in one NCCL group
void swb::comm_halo_exchange(swb_comm *comm, const void *send_lo, void *recv_lo, int lower, const void *send_hi, void *recv_hi, int upper, size_t nelem, int dtype,
                             cudaStream_t st)
{
    SWB_REQUIRE(comm != nullptr && comm->comm != nullptr, "null communicator");
    NcclApi &a = nccl();
    SWB_REQUIRE(a.Send && a.Recv && a.GroupStart && a.GroupEnd, "this NCCL build has no point-to-point support");
    const ncclDataType_t t = dtype == SWB_F64 ? ncclDouble : ncclFloat;
    nccl_check(a.GroupStart(), "ncclGroupStart");
    if (lower >= 0) {
        nccl_check(a.Send(send_lo, nelem, t, lower, comm->comm, st), "ncclSend");
        nccl_check(a.Recv(recv_lo, nelem, t, lower, comm->comm, st), "ncclRecv");
    }
    if (upper >= 0) {
        nccl_check(a.Send(send_hi, nelem, t, upper, comm->comm, st), "ncclSend");
        nccl_check(a.Recv(recv_hi, nelem, t, upper, comm->comm, st), "ncclRecv");
    }
    nccl_check(a.GroupEnd(), "ncclGroupEnd");
}

extern "C" {

int32_t swb_comm_unique_id(void *id128)
{
    SWB_API_BEGIN
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id128, &id, 128);
    SWB_API_END
}

int32_t swb_comm_create(const void *id128, int32_t nranks, int32_t rank, int32_t device, swb_comm **comm)
{
    SWB_API_BEGIN
    SWB_REQUIRE(comm != nullptr && id128 != nullptr, "null argument");
    SWB_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
    SWB_CUDA(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, id128, 128);
    swb_comm *c = new swb_comm;
    c->device = device;
    c->nranks = nranks;
    c->rank = rank;
    ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) {
        delete c;
        nccl_check(r, "ncclCommInitRank");
    }
    *comm = c;
    SWB_API_END
}

int32_t swb_comm_destroy(swb_comm *comm)
{
    SWB_API_BEGIN
    if (comm) {
        if (comm->comm)
            nccl().CommDestroy(comm->comm);
        delete comm;
    }
    SWB_API_END
}

int32_t swb_comm_allreduce_sum(swb_comm *comm, void *dev_ptr, size_t nelem, int32_t dtype, void *stream)
{
    SWB_API_BEGIN
    SWB_REQUIRE(comm != nullptr && comm->comm != nullptr, "null communicator");
    SWB_REQUIRE(dtype == SWB_F32 || dtype == SWB_F64, "dtype must be SWB_F32 or SWB_F64");
    SWB_CUDA(cudaSetDevice(comm->device));
    nccl_check(nccl().AllReduce(dev_ptr, dev_ptr, nelem, dtype == SWB_F64 ? ncclDouble : ncclFloat, ncclSum, comm->comm, (cudaStream_t)stream),
               "ncclAllReduce");
    SWB_API_END
}

int32_t swb_sim_allreduce_total_gradient(swb_sim *sim, swb_comm *comm)
{
    SWB_API_BEGIN
    SWB_REQUIRE(sim != nullptr && sim->impl && comm != nullptr && comm->comm != nullptr, "null argument");
    sim->impl->use_device();
    for (int k = 0; k < sim->impl->n_total_gradients(); ++k) {
        void *p;
        size_t n;
        sim->impl->total_gradient_ptr(k, &p, &n);
        nccl_check(nccl().AllReduce(p, p, n, sim->impl->desc.dtype == SWB_F64 ? ncclDouble : ncclFloat, ncclSum, comm->comm, sim->impl->stream),
                   "ncclAllReduce");
    }
    SWB_CUDA(cudaStreamSynchronize(sim->impl->stream));
    SWB_API_END
}

int32_t swb_sim_slab_export(swb_sim *sim, swb_slab_handle *handle_out)
{
    SIM_CALL(SWB_REQUIRE(handle_out != nullptr, "null handle"); sim->impl->slab_export(handle_out))
}

int32_t swb_sim_slab_connect(swb_sim *sim, const swb_slab_handle *lower_or_null, const swb_slab_handle *upper_or_null)
{
    SIM_CALL(sim->impl->slab_connect(lower_or_null, upper_or_null))
}

int32_t swb_sim_set_slab(swb_sim *sim, swb_comm *comm, int32_t lower_rank, int32_t upper_rank)
{
    SWB_API_BEGIN
    SWB_REQUIRE(sim != nullptr && sim->impl, "null simulation handle");
    if (comm) {
        SWB_REQUIRE(lower_rank < comm->nranks && upper_rank < comm->nranks && lower_rank != comm->rank && upper_rank != comm->rank, "bad neighbour rank");
        SWB_REQUIRE(comm->device == sim->impl->desc.device, "communicator and simulation live on different devices");
    }
    sim->impl->use_device();
    sim->impl->set_slab(comm, lower_rank, upper_rank);
    SWB_API_END
}

} // extern "C"
