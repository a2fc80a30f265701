// engine.h -- per-shot engine objects behind swb_sim (one WaveSimulation on one GPU).
#pragma once
#include "common.cuh"
#include "kernels.h"
#include <map>
#include <memory>
#include <vector>

namespace swb {

// RAII device allocation, zero-filled
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) {
            release();
            p = o.p;
            bytes = o.bytes;
            o.p = nullptr;
            o.bytes = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t b, cudaStream_t st = 0);
    void release();
    template <class T>
    T *as() const { return (T *)p; }
};

extern std::atomic<long long> g_device_bytes;

// ------------------------------------------------------------------------------------------------
// Device-resident checkpoint storage with the schedule of the reference's LinearCheckpointer
// (src/utils/checkpointers.jl:1-107): a checkpoint of every registered field at each multiple of
// check_freq (width-w fields also at the w-1 preceding steps), plus a rolling buffer of
// check_freq+1 copies of the buffered fields that holds the segment being correlated.
// A "field" is a list of component arrays (ScalarVariableField = 1, MultiVariableField = n).
// ------------------------------------------------------------------------------------------------
class DeviceCheckpointer {
  public:
    struct FieldSpec {
        std::vector<size_t> comp_bytes;
        int width = 1;
        bool buffered = false;
    };
    DeviceCheckpointer(int64_t nt, int64_t check_freq, std::vector<FieldSpec> fields, cudaStream_t st);
    void reset() { curr_ = last_; }
    // savecheckpoint! (checkpointers.jl:44-56)
    void save(int f, const std::vector<const void *> &comps, int64_t it);
    bool is_saved(int f, int64_t it) const { return is_buffered(f, it) || is_checkpointed(f, it); }
    bool is_buffered(int f, int64_t it) const { return fields_[f].buffered && curr_ <= it && it <= curr_ + cf_; }
    bool is_checkpointed(int f, int64_t it) const { return slots_[f].count(it) != 0; }
    // getsaved (checkpointers.jl:80-85): checkpoint storage preferred over the buffer
    std::vector<void *> get(int f, int64_t it) const;
    // initrecover! (checkpointers.jl:87-94)
    void init_recover();
    // the body of recover! for one re-forwarded step (checkpointers.jl:96-105)
    void store_recovered(int f, const std::vector<const void *> &comps, int64_t it);
    int64_t curr() const { return curr_; }
    int64_t last() const { return last_; }
    int64_t check_freq() const { return cf_; }
    size_t bytes() const { return bytes_; }

  private:
    std::vector<void *> slot_ptrs(int f, int64_t slot) const;
    std::vector<void *> buf_ptrs(int f, int64_t k) const;
    int64_t nt_, cf_, last_, curr_;
    std::vector<FieldSpec> fields_;
    std::vector<std::map<int64_t, int64_t>> slots_; // per field: it -> slot index
    std::vector<std::vector<DevBuf>> slabs_;        // per field, per component: all checkpoint slots
    std::vector<std::vector<DevBuf>> bufslabs_;     // per field, per component: cf+1 buffer slots
    cudaStream_t st_;
    size_t bytes_ = 0;
};

// pinned host staging buffer
struct PinnedBuf {
    void *p = nullptr;
    size_t bytes = 0;
    ~PinnedBuf();
    void ensure(size_t b);
};

// ------------------------------------------------------------------------------------------------
class SimBase {
  public:
    explicit SimBase(const swb_sim_desc &d);
    virtual ~SimBase();
    const swb_sim_desc desc;
    cudaStream_t stream = nullptr;
    size_t esize;
    int64_t cell_updates = 0;
    size_t ncells() const
    {
        size_t c = 1;
        for (int d = 0; d < desc.ndim; ++d)
            c *= (size_t)desc.n[d];
        return c;
    }
    void use_device() const { SWB_CUDA(cudaSetDevice(desc.device)); }

    virtual void set_material(int nfields, const void *const *fields, int interp, bool on_device) = 0;
    void set_cpml(int axis, const void *a, const void *a_h, const void *b, const void *b_h);
    virtual void bind_scalar_shot(int64_t nsrc, const int64_t *possrcs, const void *srctf, int64_t nrec, const int64_t *posrecs);
    virtual void bind_elastic_shot(int src_kind, const swb_sinc_points_host src_pts[2], const void *srctf, const void *Mxx, const void *Mzz,
                                   const void *Mxz, const swb_sinc_points_host rec_pts[2]);
    virtual void forward(void *host_seis, int snapevery) = 0;
    virtual void gradient_forward(void *host_seis) = 0;
    virtual void gradient_adjoint(const void *host_adjsrc) = 0;
    // forward sweep -> residual, window, diagonal covariance and misfit on the device -> adjoint sweep (L2Misfit.jl:24-95)
    void gradient_l2(const swb_l2_spec &spec, void *host_seis, double *misfit);
    virtual void get_raw_gradient(int which, void *host_out) = 0;
    virtual void accumulate_gradient(int64_t nsrcpos, const void *srcpos, int rs, int64_t nrecpos, const void *recpos, int rr) = 0;
    virtual int n_total_gradients() const = 0;
    virtual void get_field(const std::string &name, void *host_out, size_t nbytes) = 0;
    // z-slab domain decomposition (forward only): this sim holds one slab of the last axis, ghost planes included; a negative
    // neighbour rank marks a true domain end
    // the adjoint loop of swgradient_1shot! on the adjoint source in adjsrc_ (after gradient_forward)
    virtual void adjoint_loop() = 0;
    virtual void set_slab(swb_comm *, int, int) { throw Error(SWB_ERR_ARG, "z-slab decomposition is available for the fused 3D acoustic constant-density engine only"); }
    // peer-memory halo exchange of a z-slab decomposition (include/swb200.h, swb_slab_handle)
    virtual void slab_export(swb_slab_handle *) { throw Error(SWB_ERR_ARG, "z-slab decomposition is available for the fused 3D acoustic constant-density engine only"); }
    virtual void slab_connect(const swb_slab_handle *, const swb_slab_handle *) { throw Error(SWB_ERR_ARG, "z-slab decomposition is available for the fused 3D acoustic constant-density engine only"); }
    void zero_total_gradient();
    void total_gradient_ptr(int which, void **p, size_t *nelem);
    void get_total_gradient(int which, void *host_out);
    void get_snapshot(int64_t it, int field, void *host_out);
    void kernel_timing(int enable, double *ms_total, int64_t *launches);
    void kernel_timing_class(int cls, double *ms_total, int64_t *launches);
    int64_t device_bytes() const { return dev_bytes_; }

  protected:
    DevBuf dalloc(size_t bytes);
    // (re)allocate only when the size changes, so that device pointers -- and with them captured CUDA graphs -- survive
    // a re-bind with the same geometry; returns true if the pointer changed
    bool ensure(DevBuf &b, size_t bytes);
    virtual void pointers_changed() {}
    // called at the end of set_cpml (cpml_lo_inactive_[axis] is up to date)
    virtual void cpml_changed(int) {}
    // CUDA-graph replay of a fixed launch sequence: the first call captures `enqueue` on the sim's stream, later calls
    // replay it (SWB_FLAG_NO_GRAPH: always enqueue eagerly).  `updates` = cell-updates the sequence performs.
    struct Graph {
        cudaGraphExec_t exec = nullptr;
        int64_t updates = 0, launches = 0;
    };
    template <class F>
    void run_graph(Graph &g, F &&enqueue);
    void drop_graph(Graph &g);
    // Keep [base, base + bytes) resident in the persisting part of the L2 for every kernel launched on `st` (access-policy window, hit
    // ratio scaled to what the device lets us set aside): a time-invariant array every step re-reads -- a material factor plane -- then
    // stops costing HBM bandwidth.  Opt-in experiment (SWB_L2_PERSIST=1): measured slower on B200, see engine_core.cu; returns 0 when off.
    size_t l2_persist(cudaStream_t st, const void *base, size_t bytes);
    void upload(void *dst, const void *src, size_t bytes);
    void download(void *dst, const void *src, size_t bytes);
    void d2d(void *dst, const void *src, size_t bytes);
    void zero(DevBuf &b);
    void sync() { SWB_CUDA(cudaStreamSynchronize(stream)); }
    swb_cpml_axis cpml_axis(int ax) const;
    // timing of the dominant (stencil) kernel
    void tic(int cls = 0); // cls 0: forward / re-forward step, 1: adjoint step (+ correlation); samples every 8th call
    // time a whole replayed sequence: n launches of class cls (+ n_aux re-forward launches inside an adjoint sequence)
    void tic_group(int cls, int64_t n, int64_t n_aux = 0);
    void toc();

    DevBuf cpml_[3][4]; // a, a_h, b, b_h per axis
    bool cpml_set_[3] = {false, false, false};
    // the lower strip of an axis has a = a_h = 0 (free-surface override of init_bdc!, acou_init_bc.jl:33-39): its memory variables stay 0
    bool cpml_lo_inactive_[3] = {false, false, false};
    // acoustic shot binding
    DevBuf possrc_, posrec_, srctf_, traces_, adjsrc_;
    DevBuf l2_obs_, l2_w_, l2_mask_, l2_acc_; // device-side L2 misfit: observed data, per-sample weights, window mask, accumulator
    int64_t nsrc_ = 0, nrec_ = 0;
    bool shot_bound_ = false;
    bool fwd_done_ = false;
    std::vector<DevBuf> total_grad_;
    std::map<int64_t, std::vector<std::vector<char>>> snapshots_; // it -> field components on the host (step-by-step engines)
    // Snapshots of the fused engines (savesnapshot!, src/utils/snapshotter.jl:33-41) without leaving the graph-captured time loop: a snapshot
    // step copies the dense fields into a device-resident slot (device to device, inside the sweep), and a side stream drains the slots to
    // pinned host memory while the sweep goes on; the two streams join at the end of the sweep.
    struct SnapStore {
        int snapevery = 0;
        int64_t nsnap = 0;
        std::vector<size_t> comp_off; // byte offset of each component inside a slot (+ total at the end)
        std::vector<size_t> comp_bytes;
        DevBuf dev;
        PinnedBuf host;
        cudaStream_t st = nullptr;
        cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
        bool used = false; // a drain was enqueued in the current sweep
    } snap_;
    // (re)sizes the store; returns true when a pointer changed (captured graphs must be dropped).  false + !snap_fits() when the store
    // would not fit the device-memory budget: the caller then takes snapshots the slow way
    bool snap_setup(int snapevery, const std::vector<size_t> &comp_bytes);
    bool snap_ok_ = false;
    void *snap_dev(int64_t it, int comp) const; // where component `comp` of the snapshot of step `it` goes (dense, device)
    void snap_drain(int64_t it);                // enqueue slot -> pinned host on the side stream (after everything enqueued so far)
    void snap_finish();                         // join the side stream back into the sim's stream
    PinnedBuf pin_;
    int64_t dev_bytes_ = 0;
    bool timing_ = false, tsampled_ = false;
    int64_t tcount_ = 0;
    struct TimedLaunch {
        cudaEvent_t a, b;
        int cls;
        int64_t n, n_aux;
    };
    std::vector<TimedLaunch> tev_;
    double t_ms_[3] = {0, 0, 0};
    int64_t t_n_[3] = {0, 0, 0};
    bool capturing_ = false;
};

template <class F>
void SimBase::run_graph(Graph &g, F &&enqueue)
{
    if (desc.flags & SWB_FLAG_NO_GRAPH) {
        enqueue();
        return;
    }
    if (!g.exec) {
        const int64_t cu0 = cell_updates;
        const long long l0 = g_launches.load();
        cudaGraph_t graph = nullptr;
        SWB_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeRelaxed));
        capturing_ = true;
        try {
            enqueue();
            capturing_ = false;
        } catch (...) {
            capturing_ = false;
            cudaStreamEndCapture(stream, &graph);
            if (graph)
                cudaGraphDestroy(graph);
            throw;
        }
        SWB_CUDA(cudaStreamEndCapture(stream, &graph));
        g.updates = cell_updates - cu0;
        g.launches = g_launches.load() - l0;
        cell_updates = cu0; // counted per replay below
        g_launches.fetch_sub(g.launches);
        cudaError_t e = cudaGraphInstantiate(&g.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            g.exec = nullptr;
            throw Error(SWB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        }
    }
    SWB_CUDA(cudaGraphLaunch(g.exec, stream));
    cell_updates += g.updates;
    g_launches.fetch_add(g.launches);
}

// stream-ordered 32-bit flag operations (cuStreamWriteValue32 / cuStreamWaitValue32 resolved through cudaGetDriverEntryPoint; engine_core.cu):
// the write is preceded by a system-wide memory barrier, so everything enqueued on the stream before it -- peer stores over NVLink
// included -- is visible to whoever observes the value; the wait blocks the stream until *addr >= value (wrap-around compare)
void stream_write_flag(cudaStream_t st, void *dev_addr, uint32_t value);
void stream_wait_flag_geq(cudaStream_t st, void *dev_addr, uint32_t value);

// capi.cu (NCCL is resolved there)
void comm_halo_exchange(swb_comm *comm, const void *send_lo, void *recv_lo, int lower, const void *send_hi, void *recv_hi, int upper, size_t nelem, int dtype,
                        cudaStream_t st);

SimBase *make_acoustic_cd(const swb_sim_desc &d);
SimBase *make_acoustic_vd(const swb_sim_desc &d);
SimBase *make_elastic_iso(const swb_sim_desc &d);

} // namespace swb

struct swb_sim {
    std::unique_ptr<swb::SimBase> impl;
};
