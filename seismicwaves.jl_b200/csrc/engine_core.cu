// engine_core.cu -- device memory, checkpoint storage and the parts of the per-shot engine that do not
// depend on the physics.
#include "engine.h"
#include <cstdlib>
#include <cuda.h>
#include <algorithm>
#include <cstring>
#include <cmath>

#include "tma.cuh"
namespace swb {

bool pdl_enabled()
{
    static const bool on = std::getenv("SWB_NO_PDL") == nullptr;
    return on;
}

void make_tmap_2d(CUtensorMap *out, int dtype, const void *base, long long ld, long long rows, int box_w, int box_h)
{
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static encode_fn encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SWB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
            throw Error(SWB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        return (encode_fn)fn;
    }();
    const size_t es = dtype == SWB_F64 ? 8 : 4;
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld * es};
    const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)box_h};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, dtype == SWB_F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw Error(SWB_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
}


namespace {
typedef CUresult (*memop32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
memop32_fn driver_memop(const char *name)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SWB_CUDA(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr)
        throw Error(SWB_ERR_CUDA, std::string(name) + " is not available in this driver");
    return (memop32_fn)fn;
}
} // namespace

void stream_write_flag(cudaStream_t st, void *dev_addr, uint32_t value)
{
    static memop32_fn fn = driver_memop("cuStreamWriteValue32");
    const CUresult r = fn((CUstream)st, (CUdeviceptr)dev_addr, value, CU_STREAM_WRITE_VALUE_DEFAULT);
    if (r != CUDA_SUCCESS)
        throw Error(SWB_ERR_CUDA, "cuStreamWriteValue32 failed with CUresult " + std::to_string((int)r));
}

void stream_wait_flag_geq(cudaStream_t st, void *dev_addr, uint32_t value)
{
    static memop32_fn fn = driver_memop("cuStreamWaitValue32");
    const CUresult r = fn((CUstream)st, (CUdeviceptr)dev_addr, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS)
        throw Error(SWB_ERR_CUDA, "cuStreamWaitValue32 failed with CUresult " + std::to_string((int)r));
}

std::atomic<long long> g_launches{0};
std::atomic<long long> g_device_bytes{0};

static thread_local std::string t_last_error;
void set_last_error(const std::string &msg) { t_last_error = msg; }
const char *last_error_cstr() { return t_last_error.c_str(); }

// ---- Fornberg (1998) finite-difference weights ---------------------------------------------------
// Same algorithm the reference uses to build its stencils (src/utils/fdgen.jl:11-47): weights of the
// m-th derivative at 0 on the nodes x[0..n], computed with the published recurrence in double.
static std::vector<double> fornberg_weights(std::vector<double> x, int m)
{
    std::sort(x.begin(), x.end());
    const int np = (int)x.size();
    std::vector<std::vector<double>> c(np, std::vector<double>(m + 1, 0.0));
    double c1 = 1.0, c4 = x[0];
    c[0][0] = 1.0;
    for (int i = 1; i < np; ++i) {
        const int mn = std::min(i, m);
        double c2 = 1.0;
        const double c5 = c4;
        c4 = x[i];
        for (int j = 0; j < i; ++j) {
            const double c3 = x[i] - x[j];
            c2 = c2 * c3;
            if (j == i - 1) {
                for (int k = mn; k >= 1; --k)
                    c[i][k] = c1 * (k * c[i - 1][k - 1] - c5 * c[i - 1][k]) / c2;
                c[i][0] = -c1 * c5 * c[i - 1][0] / c2;
            }
            for (int k = mn; k >= 1; --k)
                c[j][k] = (c4 * c[j][k] - k * c[j][k - 1]) / c3;
            c[j][0] = c4 * c[j][0] / c3;
        }
        c1 = c2;
    }
    std::vector<double> w(np);
    for (int i = 0; i < np; ++i)
        w[i] = c[i][m];
    return w;
}

static std::vector<double> fd_coeffs(int deriv, int order)
{
    const int nnn = order + deriv - 1;
    std::vector<double> nodes(nnn);
    for (int k = 1; k <= nnn; ++k)
        nodes[k - 1] = k - (nnn / 2.0 + 0.5);
    return fornberg_weights(nodes, deriv);
}

const FdWeights &fd_weights()
{
    static const FdWeights w = [] {
        FdWeights f{};
        auto a = fd_coeffs(1, 2), b = fd_coeffs(2, 2), c = fd_coeffs(1, 4);
        std::copy(a.begin(), a.end(), f.d1o2);
        std::copy(b.begin(), b.end(), f.d2o2);
        std::copy(c.begin(), c.end(), f.d1o4);
        return f;
    }();
    return w;
}

// ---- DevBuf -----------------------------------------------------------------------------------------
void DevBuf::alloc(size_t b, cudaStream_t st)
{
    release();
    if (b == 0)
        return;
    cudaError_t e = cudaMalloc(&p, b);
    if (e != cudaSuccess) {
        p = nullptr;
        throw Error(e == cudaErrorMemoryAllocation ? SWB_ERR_NOMEM : SWB_ERR_CUDA,
                    std::string("cudaMalloc of ") + std::to_string(b) + " bytes failed: " + cudaGetErrorString(e));
    }
    bytes = b;
    g_device_bytes.fetch_add((long long)b);
    SWB_CUDA(cudaMemsetAsync(p, 0, b, st));
}

void DevBuf::release()
{
    if (p) {
        cudaFree(p);
        g_device_bytes.fetch_sub((long long)bytes);
    }
    p = nullptr;
    bytes = 0;
}

PinnedBuf::~PinnedBuf()
{
    if (p)
        cudaFreeHost(p);
}
void PinnedBuf::ensure(size_t b)
{
    if (b <= bytes)
        return;
    if (p)
        cudaFreeHost(p);
    p = nullptr;
    bytes = 0;
    SWB_CUDA(cudaMallocHost(&p, b));
    bytes = b;
}

// ---- DeviceCheckpointer -----------------------------------------------------------------------------
DeviceCheckpointer::DeviceCheckpointer(int64_t nt, int64_t cf, std::vector<FieldSpec> fields, cudaStream_t st)
    : nt_(nt), cf_(cf), fields_(std::move(fields)), st_(st)
{
    SWB_REQUIRE(cf >= 1, "check_freq must be positive");
    SWB_REQUIRE(cf < nt, "Checkpointing frequency must be smaller than the number of timesteps!");
    last_ = (nt / cf) * cf;
    curr_ = last_;
    const int nf = (int)fields_.size();
    slots_.resize(nf);
    slabs_.resize(nf);
    bufslabs_.resize(nf);
    for (int f = 0; f < nf; ++f) {
        // slots for it in 0..nt+1 with it % cf == 0, and the width-1 preceding steps (checkpointers.jl:22-34)
        for (int64_t it = 0; it <= nt + 1; ++it)
            if (it % cf == 0)
                for (int64_t itw = it; itw > it - fields_[f].width; --itw)
                    if (!slots_[f].count(itw)) {
                        int64_t idx = (int64_t)slots_[f].size();
                        slots_[f][itw] = idx;
                    }
        const size_t nslots = slots_[f].size();
        for (size_t cb : fields_[f].comp_bytes) {
            DevBuf b;
            b.alloc(cb * nslots, st_);
            bytes_ += b.bytes;
            slabs_[f].push_back(std::move(b));
            DevBuf bb;
            if (fields_[f].buffered) {
                bb.alloc(cb * (size_t)(cf + 1), st_);
                bytes_ += bb.bytes;
            }
            bufslabs_[f].push_back(std::move(bb));
        }
    }
}

std::vector<void *> DeviceCheckpointer::slot_ptrs(int f, int64_t slot) const
{
    std::vector<void *> out;
    for (size_t c = 0; c < fields_[f].comp_bytes.size(); ++c)
        out.push_back((char *)slabs_[f][c].p + fields_[f].comp_bytes[c] * (size_t)slot);
    return out;
}

std::vector<void *> DeviceCheckpointer::buf_ptrs(int f, int64_t k) const
{
    std::vector<void *> out;
    for (size_t c = 0; c < fields_[f].comp_bytes.size(); ++c)
        out.push_back((char *)bufslabs_[f][c].p + fields_[f].comp_bytes[c] * (size_t)k);
    return out;
}

void DeviceCheckpointer::save(int f, const std::vector<const void *> &comps, int64_t it)
{
    const FieldSpec &fs = fields_[f];
    bool ck = false;
    for (int64_t itw = it; itw < it + fs.width; ++itw)
        if (((itw % cf_) + cf_) % cf_ == 0)
            ck = true;
    if (ck) {
        auto s = slots_[f].find(it);
        SWB_REQUIRE(s != slots_[f].end(), "checkpoint slot missing");
        auto dst = slot_ptrs(f, s->second);
        for (size_t c = 0; c < comps.size(); ++c)
            if (fs.comp_bytes[c])
                SWB_CUDA(cudaMemcpyAsync(dst[c], comps[c], fs.comp_bytes[c], cudaMemcpyDeviceToDevice, st_));
    }
    if (fs.buffered && it >= last_) {
        auto dst = buf_ptrs(f, it - last_);
        for (size_t c = 0; c < comps.size(); ++c)
            if (fs.comp_bytes[c])
                SWB_CUDA(cudaMemcpyAsync(dst[c], comps[c], fs.comp_bytes[c], cudaMemcpyDeviceToDevice, st_));
    }
}

std::vector<void *> DeviceCheckpointer::get(int f, int64_t it) const
{
    auto s = slots_[f].find(it);
    if (s != slots_[f].end())
        return slot_ptrs(f, s->second);
    if (!is_buffered(f, it))
        throw Error(SWB_ERR_STATE, "checkpointer: requested step is neither checkpointed nor buffered");
    return buf_ptrs(f, it - curr_);
}

void DeviceCheckpointer::init_recover()
{
    const int64_t old = curr_;
    curr_ -= cf_;
    for (int f = 0; f < (int)fields_.size(); ++f) {
        if (!fields_[f].buffered)
            continue;
        auto b0 = buf_ptrs(f, 0), bl = buf_ptrs(f, cf_);
        auto c0 = slot_ptrs(f, slots_[f].at(curr_)), c1 = slot_ptrs(f, slots_[f].at(old));
        for (size_t c = 0; c < fields_[f].comp_bytes.size(); ++c) {
            SWB_CUDA(cudaMemcpyAsync(b0[c], c0[c], fields_[f].comp_bytes[c], cudaMemcpyDeviceToDevice, st_));
            SWB_CUDA(cudaMemcpyAsync(bl[c], c1[c], fields_[f].comp_bytes[c], cudaMemcpyDeviceToDevice, st_));
        }
    }
}

void DeviceCheckpointer::store_recovered(int f, const std::vector<const void *> &comps, int64_t it)
{
    // buffers[name][it - start_rec_it + 2] with start_rec_it = curr+1 (1-based) == slot it - curr (0-based)
    auto dst = buf_ptrs(f, it - curr_);
    for (size_t c = 0; c < comps.size(); ++c)
        SWB_CUDA(cudaMemcpyAsync(dst[c], comps[c], fields_[f].comp_bytes[c], cudaMemcpyDeviceToDevice, st_));
}

// ---- SimBase ----------------------------------------------------------------------------------------
SimBase::SimBase(const swb_sim_desc &d) : desc(d), esize(d.dtype == SWB_F64 ? 8 : 4)
{
    SWB_REQUIRE(d.dtype == SWB_F32 || d.dtype == SWB_F64, "dtype must be SWB_F32 or SWB_F64");
    SWB_REQUIRE(d.nt > 0, "Number of timesteps must be positive!");
    SWB_REQUIRE(d.dt > 0, "Timestep size must be positive!");
    SWB_REQUIRE(d.halo >= 0, "CPML halo size must be non-negative!");
    for (int k = 0; k < d.ndim; ++k) {
        SWB_REQUIRE(d.n[k] > 0, "All numbers of grid points must be positive!");
        SWB_REQUIRE(d.spacing[k] > 0, "All grid spacings must be positive!");
        SWB_REQUIRE(d.n[k] >= 2 * (int64_t)d.halo + 3, "Number grid points in the dimensions with C-PML boundaries must be at least 2*halo+3!");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        throw Error(SWB_ERR_CUDA, "no CUDA device available: libswb200 has no CPU fallback");
    SWB_REQUIRE(d.device >= 0 && d.device < ndev, "device index out of range");
    SWB_CUDA(cudaSetDevice(d.device));
    SWB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    for (int ax = 0; ax < d.ndim; ++ax) {
        cpml_[ax][0] = dalloc(esize * 2 * (size_t)(d.halo + 1)); // a
        cpml_[ax][1] = dalloc(esize * 2 * (size_t)d.halo);       // a_h
        cpml_[ax][2] = dalloc(esize * 2 * (size_t)(d.halo + 1)); // b
        cpml_[ax][3] = dalloc(esize * 2 * (size_t)d.halo);       // b_h
    }
}

SimBase::~SimBase()
{
    cudaSetDevice(desc.device);
    if (stream) {
        cudaStreamSynchronize(stream);
        cudaStreamDestroy(stream);
    }
    for (auto &e : tev_) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    if (snap_.st) {
        cudaStreamSynchronize(snap_.st);
        cudaStreamDestroy(snap_.st);
        cudaEventDestroy(snap_.ev_fork);
        cudaEventDestroy(snap_.ev_join);
    }
}

DevBuf SimBase::dalloc(size_t bytes)
{
    DevBuf b;
    b.alloc(bytes, stream);
    dev_bytes_ += (int64_t)bytes;
    return b;
}

bool SimBase::ensure(DevBuf &b, size_t bytes)
{
    if (b.p != nullptr && b.bytes == bytes)
        return false;
    dev_bytes_ -= (int64_t)b.bytes;
    b = dalloc(bytes);
    return true;
}

void SimBase::drop_graph(Graph &g)
{
    if (g.exec) {
        cudaStreamSynchronize(stream);
        cudaGraphExecDestroy(g.exec);
    }
    g = Graph();
}

void SimBase::upload(void *dst, const void *src, size_t bytes)
{
    if (bytes == 0)
        return;
    // pageable source: the runtime stages it; synchronous w.r.t. the host buffer, ordered on our stream
    SWB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, stream));
    SWB_CUDA(cudaStreamSynchronize(stream));
}

void SimBase::download(void *dst, const void *src, size_t bytes)
{
    if (bytes == 0)
        return;
    SWB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    SWB_CUDA(cudaStreamSynchronize(stream));
}

size_t SimBase::l2_persist(cudaStream_t st, const void *base, size_t bytes)
{
    // Measured on B200 (tools/ab_l2persist.sh, VD 4096^2): pinning fact_m0 in 79 MB of persisting L2 makes the step SLOWER (forward
    // 107 -> 166 us, adjoint 190 -> 306 us; 48 MB: 108 / 194 us) -- the streaming traffic needs the L2 as its staging buffer more than
    // it gains from one resident array.  Kept as an opt-in experiment (SWB_L2_PERSIST=1), off by default.
    static const bool on = std::getenv("SWB_L2_PERSIST") != nullptr;
    if (!on || bytes == 0)
        return 0;
    use_device();
    int max_persist = 0, max_window = 0;
    SWB_CUDA(cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, desc.device));
    SWB_CUDA(cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, desc.device));
    if (max_persist <= 0 || max_window <= 0)
        return 0;
    size_t want = (size_t)max_persist;
    if (const char *e = std::getenv("SWB_L2_PERSIST_MB"))
        want = std::min<size_t>(want, (size_t)std::atoll(e) << 20);
    SWB_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
    size_t got = 0;
    SWB_CUDA(cudaDeviceGetLimit(&got, cudaLimitPersistingL2CacheSize));
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    const size_t win = std::min(bytes, (size_t)max_window);
    attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
    attr.accessPolicyWindow.num_bytes = win;
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)got / (double)win);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    SWB_CUDA(cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
    return std::min(got, win);
}

void SimBase::d2d(void *dst, const void *src, size_t bytes)
{
    if (bytes)
        SWB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, stream));
}

void SimBase::zero(DevBuf &b)
{
    if (b.bytes)
        SWB_CUDA(cudaMemsetAsync(b.p, 0, b.bytes, stream));
}

void SimBase::set_cpml(int axis, const void *a, const void *a_h, const void *b, const void *b_h)
{
    use_device();
    SWB_REQUIRE(axis >= 0 && axis < desc.ndim, "CPML axis out of range");
    upload(cpml_[axis][0].p, a, cpml_[axis][0].bytes);
    upload(cpml_[axis][1].p, a_h, cpml_[axis][1].bytes);
    upload(cpml_[axis][2].p, b, cpml_[axis][2].bytes);
    upload(cpml_[axis][3].p, b_h, cpml_[axis][3].bytes);
    {
        auto lower_half_zero = [&](const void *v, size_t bytes) {
            const size_t n = bytes / esize / 2;
            for (size_t k = 0; k < n; ++k)
                if ((esize == 8 ? ((const double *)v)[k] : (double)((const float *)v)[k]) != 0.0)
                    return false;
            return true;
        };
        cpml_lo_inactive_[axis] = lower_half_zero(a, cpml_[axis][0].bytes) && lower_half_zero(a_h, cpml_[axis][1].bytes);
    }
    cpml_set_[axis] = true;
    cpml_changed(axis);
}

void SimBase::gradient_l2(const swb_l2_spec &spec, void *host_seis, double *misfit)
{
    gradient_forward(host_seis);
    SWB_REQUIRE(adjsrc_.bytes > 0 && adjsrc_.bytes == traces_.bytes, "no gradient shot bound");
    const size_t nt = (size_t)desc.nt, n = adjsrc_.bytes / esize;
    const void *obs = nullptr, *w = nullptr, *mask = nullptr;
    if (spec.observed) {
        ensure(l2_obs_, adjsrc_.bytes);
        upload(l2_obs_.p, spec.observed, l2_obs_.bytes);
        obs = l2_obs_.p;
    }
    if (spec.invcov_diag) {
        ensure(l2_w_, esize * nt);
        upload(l2_w_.p, spec.invcov_diag, l2_w_.bytes);
        w = l2_w_.p;
    }
    if (spec.mask) {
        ensure(l2_mask_, esize * nt);
        upload(l2_mask_.p, spec.mask, l2_mask_.bytes);
        mask = l2_mask_.p;
    }
    ensure(l2_acc_, sizeof(double));
    SWB_CUDA(cudaMemsetAsync(l2_acc_.p, 0, sizeof(double), stream));
    post_l2_adjsrc(desc.dtype, n, nt, traces_.p, obs, w, mask, adjsrc_.p, l2_acc_.as<double>(), stream);
    adjoint_loop();
    if (misfit)
        download(misfit, l2_acc_.p, sizeof(double));
}

swb_cpml_axis SimBase::cpml_axis(int ax) const
{
    swb_cpml_axis c;
    c.a = cpml_[ax][0].p;
    c.a_h = cpml_[ax][1].p;
    c.b = cpml_[ax][2].p;
    c.b_h = cpml_[ax][3].p;
    return c;
}

void SimBase::bind_scalar_shot(int64_t nsrc, const int64_t *possrcs, const void *srctf, int64_t nrec, const int64_t *posrecs)
{
    use_device();
    SWB_REQUIRE(desc.kind == SWB_ACOU_CD || desc.kind == SWB_ACOU_VD, "scalar shots belong to acoustic simulations");
    SWB_REQUIRE(nsrc > 0, "There must be at least one source!");
    SWB_REQUIRE(nrec > 0, "There must be at least one receiver!");
    for (int64_t s = 0; s < nsrc; ++s)
        for (int d = 0; d < desc.ndim; ++d)
            SWB_REQUIRE(possrcs[s + d * nsrc] >= 1 && possrcs[s + d * nsrc] <= desc.n[d], "source position outside the grid");
    for (int64_t r = 0; r < nrec; ++r)
        for (int d = 0; d < desc.ndim; ++d)
            SWB_REQUIRE(posrecs[r + d * nrec] >= 1 && posrecs[r + d * nrec] <= desc.n[d], "receiver position outside the grid");
    nsrc_ = nsrc;
    nrec_ = nrec;
    bool ch = ensure(possrc_, sizeof(int64_t) * nsrc * desc.ndim);
    ch |= ensure(posrec_, sizeof(int64_t) * nrec * desc.ndim);
    ch |= ensure(srctf_, esize * desc.nt * nsrc);
    ch |= ensure(traces_, esize * desc.nt * nrec);
    if (desc.gradient)
        ch |= ensure(adjsrc_, esize * desc.nt * nrec);
    if (ch)
        pointers_changed();
    upload(possrc_.p, possrcs, possrc_.bytes);
    upload(posrec_.p, posrecs, posrec_.bytes);
    upload(srctf_.p, srctf, srctf_.bytes);
    shot_bound_ = true;
    fwd_done_ = false;
}

void SimBase::bind_elastic_shot(int, const swb_sinc_points_host *, const void *, const void *, const void *, const void *, const swb_sinc_points_host *)
{
    throw Error(SWB_ERR_ARG, "elastic shots belong to elastic simulations");
}

void SimBase::zero_total_gradient()
{
    use_device();
    for (auto &b : total_grad_)
        zero(b);
}

void SimBase::total_gradient_ptr(int which, void **p, size_t *nelem)
{
    SWB_REQUIRE(which >= 0 && which < (int)total_grad_.size(), "gradient component out of range");
    *p = total_grad_[which].p;
    *nelem = total_grad_[which].bytes / esize;
}

void SimBase::get_total_gradient(int which, void *host_out)
{
    use_device();
    SWB_REQUIRE(which >= 0 && which < (int)total_grad_.size(), "gradient component out of range");
    download(host_out, total_grad_[which].p, total_grad_[which].bytes);
}

bool SimBase::snap_setup(int snapevery, const std::vector<size_t> &comp_bytes)
{
    use_device();
    snap_ok_ = false;
    if (snapevery <= 0)
        return false;
    std::vector<size_t> off(comp_bytes.size() + 1, 0);
    for (size_t c = 0; c < comp_bytes.size(); ++c)
        off[c + 1] = off[c] + (comp_bytes[c] + 255) / 256 * 256;
    const int64_t nsnap = desc.nt / snapevery;
    const size_t total = off.back() * (size_t)std::max<int64_t>(nsnap, 1);
    size_t free_b = 0, total_b = 0;
    SWB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    if (total > snap_.dev.bytes && total - snap_.dev.bytes > free_b / 2) // keep half of what is left for the caller: fall back
        return false;
    if (!snap_.st) {
        SWB_CUDA(cudaStreamCreateWithFlags(&snap_.st, cudaStreamNonBlocking));
        SWB_CUDA(cudaEventCreateWithFlags(&snap_.ev_fork, cudaEventDisableTiming));
        SWB_CUDA(cudaEventCreateWithFlags(&snap_.ev_join, cudaEventDisableTiming));
    }
    bool ch = snapevery != snap_.snapevery || off != snap_.comp_off;
    snap_.snapevery = snapevery;
    snap_.nsnap = nsnap;
    snap_.comp_off = off;
    snap_.comp_bytes = comp_bytes;
    ch |= ensure(snap_.dev, total);
    if (total > snap_.host.bytes) {
        snap_.host.ensure(total);
        ch = true;
    }
    snap_ok_ = true;
    return ch;
}

void *SimBase::snap_dev(int64_t it, int comp) const
{
    const int64_t s = it / snap_.snapevery - 1;
    SWB_REQUIRE(snap_ok_ && it % snap_.snapevery == 0 && s >= 0 && s < snap_.nsnap && comp >= 0 && comp + 1 < (int)snap_.comp_off.size(), "snapshot slot out of range");
    return (char *)snap_.dev.p + (size_t)s * snap_.comp_off.back() + snap_.comp_off[comp];
}

void SimBase::snap_drain(int64_t it)
{
    const size_t slot = snap_.comp_off.back(), o = (size_t)(it / snap_.snapevery - 1) * slot;
    SWB_CUDA(cudaEventRecord(snap_.ev_fork, stream));
    SWB_CUDA(cudaStreamWaitEvent(snap_.st, snap_.ev_fork, 0));
    SWB_CUDA(cudaMemcpyAsync((char *)snap_.host.p + o, (const char *)snap_.dev.p + o, slot, cudaMemcpyDeviceToHost, snap_.st));
    snap_.used = true;
}

void SimBase::snap_finish()
{
    if (!snap_.used)
        return;
    SWB_CUDA(cudaEventRecord(snap_.ev_join, snap_.st));
    SWB_CUDA(cudaStreamWaitEvent(stream, snap_.ev_join, 0));
    snap_.used = false;
}

void SimBase::get_snapshot(int64_t it, int field, void *host_out)
{
    if (snap_ok_ && snap_.snapevery > 0 && snapshots_.empty()) { // taken by a fused engine: pinned slots, complete once the sweep's stream is
        SWB_REQUIRE(it >= 1 && it % snap_.snapevery == 0 && it / snap_.snapevery <= snap_.nsnap, "no snapshot stored for this time step");
        SWB_REQUIRE(field >= 0 && field + 1 < (int)snap_.comp_off.size(), "snapshot field out of range");
        sync();
        const size_t o = (size_t)(it / snap_.snapevery - 1) * snap_.comp_off.back() + snap_.comp_off[field];
        std::memcpy(host_out, (const char *)snap_.host.p + o, snap_.comp_bytes[field]);
        return;
    }
    auto s = snapshots_.find(it);
    SWB_REQUIRE(s != snapshots_.end(), "no snapshot stored for this time step");
    SWB_REQUIRE(field >= 0 && field < (int)s->second.size(), "snapshot field out of range");
    std::memcpy(host_out, s->second[field].data(), s->second[field].size());
}

void SimBase::tic(int cls)
{
    tsampled_ = false;
    if (!timing_ || capturing_ || (tcount_++ % 8) != 0) // sample every 8th step: keeps the event overhead out of the timed region
        return;
    tsampled_ = true;
    TimedLaunch t;
    t.cls = cls ? 1 : 0;
    t.n = 1;
    t.n_aux = 0;
    SWB_CUDA(cudaEventCreate(&t.a));
    SWB_CUDA(cudaEventCreate(&t.b));
    SWB_CUDA(cudaEventRecord(t.a, stream));
    tev_.push_back(t);
}

void SimBase::tic_group(int cls, int64_t n, int64_t n_aux)
{
    tsampled_ = false;
    if (!timing_ || capturing_)
        return;
    tsampled_ = true;
    TimedLaunch t;
    t.cls = cls ? 1 : 0;
    t.n = n;
    t.n_aux = n_aux;
    SWB_CUDA(cudaEventCreate(&t.a));
    SWB_CUDA(cudaEventCreate(&t.b));
    SWB_CUDA(cudaEventRecord(t.a, stream));
    tev_.push_back(t);
}

void SimBase::toc()
{
    if (!timing_ || !tsampled_ || tev_.empty())
        return;
    SWB_CUDA(cudaEventRecord(tev_.back().b, stream));
    tsampled_ = false;
}

void SimBase::kernel_timing(int enable, double *ms_total, int64_t *launches)
{
    use_device();
    SWB_CUDA(cudaStreamSynchronize(stream));
    for (auto &e : tev_) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            t_ms_[e.cls] += ms;
            t_n_[e.cls] += e.n;
            t_n_[2] += e.n_aux;
        }
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    tev_.clear();
    if (ms_total)
        *ms_total = t_ms_[0] + t_ms_[1];
    if (launches)
        *launches = t_n_[0] + t_n_[1];
    if (enable == 0 || enable == 1) {
        if (enable == 1 && !timing_) {
            t_ms_[0] = t_ms_[1] = t_ms_[2] = 0;
            t_n_[0] = t_n_[1] = t_n_[2] = 0;
        }
        timing_ = enable == 1;
    }
}

void SimBase::kernel_timing_class(int cls, double *ms_total, int64_t *launches)
{
    kernel_timing(-1, nullptr, nullptr);
    SWB_REQUIRE(cls >= 0 && cls <= 2, "timing class must be 0 (forward), 1 (adjoint) or 2 (re-forward launches inside class-1 time)");
    if (ms_total)
        *ms_total = t_ms_[cls];
    if (launches)
        *launches = t_n_[cls];
}

} // namespace swb
