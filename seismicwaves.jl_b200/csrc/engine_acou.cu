// engine_acou.cu -- per-shot engines for the acoustic simulations.
//
// Replaces the reference's per-shot drivers swforward_1shot! / swgradient_1shot!
// (src/models/acoustic/acou_forward.jl:22-125, src/models/acoustic/acou_gradient.jl:4-203):
// the whole time loop runs inside the library on the sim's stream, wavefield checkpoints
// stay on the device with the LinearCheckpointer schedule, and the host is only touched for
// the seismograms and (optionally) the adjoint source.
#include "engine.h"
#include "vd_fused.h"
#include "cd_fused.h"
#include "tma.cuh"
#include <cstring>
#include <unistd.h>

namespace swb {

// =====================================================================================================
// Acoustic constant density (2D / 3D)
// =====================================================================================================
class AcousticCD : public SimBase {
  public:
    // own_state = false: a subclass keeps the wavefield state in its own layout (fused engine below)
    explicit AcousticCD(const swb_sim_desc &d, bool own_state = true) : SimBase(d)
    {
        SWB_REQUIRE(d.ndim >= 1 && d.ndim <= 3, "acoustic constant-density engine supports N = 1, 2, 3");
        const size_t nb = ncells() * esize;
        fact_ = dalloc(nb);
        vp_ = dalloc(nb);
        if (own_state) {
            for (int k = 0; k < 3; ++k)
                p_[k] = dalloc(nb);
            alloc_mem(psi_, xi_);
        }
        if (d.gradient) {
            grad_ = dalloc(nb);
            work_ = dalloc(nb);
            total_grad_.push_back(dalloc(nb));
            misfit_acc_ = dalloc(sizeof(double));
        }
        if (d.gradient && own_state) {
            for (int k = 0; k < 3; ++k)
                adj_[k] = dalloc(nb);
            alloc_mem(psi_adj_, xi_adj_);
            std::vector<DeviceCheckpointer::FieldSpec> fs(3);
            fs[0].comp_bytes = {nb};
            fs[0].width = 2;
            fs[0].buffered = true; // "pcur"
            for (int ax = 0; ax < d.ndim; ++ax) {
                fs[1].comp_bytes.push_back(psi_[ax].bytes); // "ψ"
                fs[2].comp_bytes.push_back(xi_[ax].bytes);  // "ξ"
            }
            ckpt_.reset(new DeviceCheckpointer(d.nt, d.check_freq, fs, stream));
            dev_bytes_ += (int64_t)ckpt_->bytes();
        }
        sync();
    }

    void set_material(int nfields, const void *const *fields, int, bool on_device) override
    {
        use_device();
        SWB_REQUIRE(nfields == 1, "acoustic constant density takes one material field (vp)");
        if (on_device)
            d2d(vp_.p, fields[0], vp_.bytes);
        else
            upload(vp_.p, fields[0], vp_.bytes);
        post_cd_fact(desc.dtype, ncells(), vp_.p, desc.dt, fact_.p, stream);
        mat_set_ = true;
    }

    void forward(void *host_seis, int snapevery) override
    {
        begin_shot();
        snapshots_.clear();
        for (int64_t it = 1; it <= desc.nt; ++it) {
            step_forward(it, true);
            if (snapevery > 0 && it % snapevery == 0)
                take_snapshot(it);
        }
        download(host_seis, traces_.p, traces_.bytes);
    }

    void gradient_forward(void *host_seis) override
    {
        SWB_REQUIRE(desc.gradient, "simulation was not built with gradient=true");
        begin_shot();
        ckpt_->reset();
        for (int64_t it = 1; it <= desc.nt; ++it) {
            step_forward(it, true);
            // savecheckpoint! of pcur, ψ, ξ (acou_gradient.jl:39-41)
            ckpt_->save(0, {cur_[1]}, it);
            ckpt_->save(1, mem_ptrs(psi_), it);
            ckpt_->save(2, mem_ptrs(xi_), it);
        }
        if (host_seis)
            download(host_seis, traces_.p, traces_.bytes);
        fwd_done_ = true;
    }

    void gradient_adjoint(const void *host_adjsrc) override
    {
        SWB_REQUIRE(fwd_done_, "gradient_adjoint called before gradient_forward");
        use_device();
        upload(adjsrc_.p, host_adjsrc, adjsrc_.bytes);
        adjoint_loop();
    }

    void get_raw_gradient(int which, void *host_out) override
    {
        use_device();
        SWB_REQUIRE(desc.gradient && which == 0, "acoustic CD has one raw gradient field (grad_vp)");
        download(host_out, grad_.p, grad_.bytes);
    }

    void accumulate_gradient(int64_t nsrcpos, const void *srcpos, int rs, int64_t nrecpos, const void *recpos, int rr) override
    {
        use_device();
        SWB_REQUIRE(desc.gradient, "simulation was not built with gradient=true");
        // gradient = Array(grad_vp); mute sources, then receivers; chain rule; accumulate (acou_gradient.jl:84-93)
        d2d(work_.p, grad_.p, grad_.bytes);
        mute_points(work_.p, nsrcpos, srcpos, rs);
        mute_points(work_.p, nrecpos, recpos, rr);
        post_cd_chain_accumulate(desc.dtype, ncells(), work_.p, vp_.p, total_grad_[0].p, stream);
    }

    int n_total_gradients() const override { return 1; }

    void get_field(const std::string &name, void *host_out, size_t nbytes) override
    {
        use_device();
        const void *src = nullptr;
        size_t b = ncells() * esize;
        if (name == "pcur")
            src = cur_[1];
        else if (name == "pold")
            src = cur_[0];
        else if (name == "fact")
            src = fact_.p;
        else if (name == "adjcur")
            src = acur_[1];
        else if (name == "grad_vp")
            src = grad_.p;
        else
            throw Error(SWB_ERR_ARG, "unknown field name: " + name);
        SWB_REQUIRE(src != nullptr && nbytes == b, "field not available or size mismatch");
        download(host_out, src, b);
    }

  protected:
    void alloc_mem(DevBuf (&psi)[3], DevBuf (&xi)[3])
    {
        for (int ax = 0; ax < desc.ndim; ++ax) {
            size_t other = 1;
            for (int k = 0; k < desc.ndim; ++k)
                if (k != ax)
                    other *= (size_t)desc.n[k];
            psi[ax] = dalloc(esize * other * 2 * (size_t)desc.halo);
            xi[ax] = dalloc(esize * other * 2 * (size_t)(desc.halo + 1));
        }
    }
    std::vector<const void *> mem_ptrs(DevBuf (&m)[3]) const
    {
        std::vector<const void *> v;
        for (int ax = 0; ax < desc.ndim; ++ax)
            v.push_back(m[ax].p);
        return v;
    }
    void mute_points(void *arr, int64_t npos, const void *hostpos, int radius)
    {
        if (radius == 0 || npos == 0)
            return;
        // persistent scratch: cudaMalloc / cudaFree inside the shot loop serialise the device (and were measured at
        // up to a second per call next to instantiated graphs)
        ensure(mute_pos_, esize * npos * desc.ndim);
        upload(mute_pos_.p, hostpos, mute_pos_.bytes);
        post_mute(desc.dtype, desc.ndim, desc.n, desc.spacing, arr, npos, mute_pos_.p, radius, stream);
        sync();
    }

    // reset! (acou_models.jl:221-226): zero every field except fact
    virtual void begin_shot()
    {
        use_device();
        SWB_REQUIRE(mat_set_, "material properties not set");
        SWB_REQUIRE(shot_bound_, "no shot bound");
        for (int ax = 0; ax < desc.ndim; ++ax)
            SWB_REQUIRE(cpml_set_[ax], "C-PML coefficients not set for every axis");
        for (int k = 0; k < 3; ++k)
            zero(p_[k]);
        for (int ax = 0; ax < desc.ndim; ++ax) {
            zero(psi_[ax]);
            zero(xi_[ax]);
        }
        if (desc.gradient) {
            zero(grad_);
            for (int k = 0; k < 3; ++k)
                zero(adj_[k]);
            for (int ax = 0; ax < desc.ndim; ++ax) {
                zero(psi_adj_[ax]);
                zero(xi_adj_[ax]);
            }
        }
        zero(traces_);
        for (int k = 0; k < 3; ++k) {
            cur_[k] = p_[k].p;
            acur_[k] = desc.gradient ? adj_[k].p : nullptr;
        }
        fwd_done_ = false;
    }

    swb_acou_cd_step_args base_args() const
    {
        swb_acou_cd_step_args a;
        std::memset(&a, 0, sizeof(a));
        a.dtype = desc.dtype;
        a.ndim = desc.ndim;
        a.halo = desc.halo;
        a.flags = desc.flags;
        for (int k = 0; k < desc.ndim; ++k) {
            a.n[k] = desc.n[k];
            a.spacing[k] = desc.spacing[k];
            a.cpml[k] = cpml_axis(k);
        }
        a.fact = fact_.p;
        a.stream = stream;
        return a;
    }

    // forward_onestep_CPML! + field rotation (acoustic2D_xPU.jl:78-126)
    void step_forward(int64_t it, bool record)
    {
        swb_acou_cd_step_args a = base_args();
        a.pold = cur_[0];
        a.pcur = cur_[1];
        a.pnew = cur_[2];
        for (int k = 0; k < desc.ndim; ++k) {
            a.psi[k] = psi_[k].p;
            a.xi[k] = xi_[k].p;
        }
        a.src.n = nsrc_;
        a.src.pos = possrc_.as<int64_t>();
        a.src.tf = srctf_.p;
        a.src.nt = desc.nt;
        if (record) {
            a.rec.n = nrec_;
            a.rec.pos = posrec_.as<int64_t>();
            a.rec.tf = traces_.p;
            a.rec.nt = desc.nt;
        }
        a.it = it;
        tic();
        cd_step(a, record);
        toc();
        cur_[0] = cur_[1];
        cur_[1] = cur_[2];
        cur_[2] = cur_[0]; // pnew aliases pold from now on, as in the reference
        cell_updates += (int64_t)ncells();
    }

    // adjoint_onestep_CPML! (acoustic2D_xPU.jl:128-169): residuals injected at the receivers
    void step_adjoint(int64_t it)
    {
        swb_acou_cd_step_args a = base_args();
        a.pold = acur_[0];
        a.pcur = acur_[1];
        a.pnew = acur_[2];
        for (int k = 0; k < desc.ndim; ++k) {
            a.psi[k] = psi_adj_[k].p;
            a.xi[k] = xi_adj_[k].p;
        }
        a.src.n = nrec_;
        a.src.pos = posrec_.as<int64_t>();
        a.src.tf = adjsrc_.p;
        a.src.nt = desc.nt;
        a.it = it;
        tic(1);
        cd_step(a, false);
        toc();
        acur_[0] = acur_[1];
        acur_[1] = acur_[2];
        acur_[2] = acur_[0];
        cell_updates += (int64_t)ncells();
    }

    // adjoint loop with re-forwarding and correlation (acou_gradient.jl:50-82)
    virtual void adjoint_loop()
    {
        const size_t nb = ncells() * esize;
        prescale_residuals(desc.dtype, desc.ndim, desc.n, adjsrc_.p, desc.nt, nrec_, posrec_.as<int64_t>(), fact_.p, stream);
        for (int64_t it = desc.nt; it >= 1; --it) {
            step_adjoint(it);
            if (!ckpt_->is_saved(0, it - 2)) {
                ckpt_->init_recover();
                const int64_t c = ckpt_->curr();
                d2d(cur_[0], ckpt_->get(0, c - 1)[0], nb);
                d2d(cur_[1], ckpt_->get(0, c)[0], nb);
                auto ps = ckpt_->get(1, c), xs = ckpt_->get(2, c);
                for (int ax = 0; ax < desc.ndim; ++ax) {
                    d2d(psi_[ax].p, ps[ax], psi_[ax].bytes);
                    d2d(xi_[ax].p, xs[ax], xi_[ax].bytes);
                }
                for (int64_t rit = c + 1; rit <= c + ckpt_->check_freq() - 1; ++rit) {
                    step_forward(rit, false);
                    ckpt_->store_recovered(0, {cur_[1]}, rit);
                }
            }
            const void *pm2 = ckpt_->get(0, it - 2)[0];
            const void *pm1 = ckpt_->get(0, it - 1)[0];
            const void *p0 = ckpt_->get(0, it)[0];
            cd_correlate(desc.dtype, desc.flags, ncells(), grad_.p, acur_[1], pm2, pm1, p0, desc.dt, stream);
        }
        sync();
    }

    void take_snapshot(int64_t it)
    {
        std::vector<std::vector<char>> comps(1);
        comps[0].resize(ncells() * esize);
        download(comps[0].data(), cur_[1], comps[0].size());
        snapshots_[it] = std::move(comps);
    }

    DevBuf fact_, vp_, p_[3], psi_[3], xi_[3];
    DevBuf grad_, work_, adj_[3], psi_adj_[3], xi_adj_[3], misfit_acc_, obs_, mute_pos_;
    void *cur_[3] = {nullptr, nullptr, nullptr};
    void *acur_[3] = {nullptr, nullptr, nullptr};
    std::unique_ptr<DeviceCheckpointer> ckpt_;
    bool mat_set_ = false;
};

#include "engine_cd_fused.inc"

SimBase *make_acoustic_cd(const swb_sim_desc &d)
{
    // the fused single-launch engine is the default; SWB_FLAG_NO_FUSION selects the one-launch-per-reference-kernel path
    // (1D grids, acoustic1D_xPU.jl, are a few thousand cells: the per-kernel path is all they need)
    if (!(d.flags & SWB_FLAG_NO_FUSION) && d.ndim >= 2)
        return new AcousticCDFused(d);
    return new AcousticCD(d);
}

// =====================================================================================================
// Acoustic variable density, staggered (2D)
// =====================================================================================================
class AcousticVD : public SimBase {
  public:
    // own_state = false: a subclass keeps the wavefield state in its own layout (fused engine below)
    explicit AcousticVD(const swb_sim_desc &d, bool own_state = true) : SimBase(d)
    {
        SWB_REQUIRE(d.ndim == 1 || d.ndim == 2, "acoustic variable-density engine supports N = 1, 2");
        nx_ = d.n[0];
        ny_ = d.ndim == 1 ? 1 : d.n[1]; // a 1D grid is the single-row case of the same kernels (acou_vd.cu)
        nn_[0] = nx_, nn_[1] = ny_;
        nd_ = d.ndim;
        const size_t nb = ncells() * esize, nbx = (size_t)(nx_ - 1) * ny_ * esize, nby = (size_t)nx_ * (ny_ - 1) * esize;
        const size_t h = (size_t)d.halo;
        vp_ = dalloc(nb);
        rho_ = dalloc(nb);
        m0_ = dalloc(nb);
        m1_[0] = dalloc(nbx);
        m1_[1] = dalloc(nby);
        auto mem = [&](DevBuf(&psi)[2], DevBuf(&xi)[2]) {
            psi[0] = dalloc(esize * 2 * h * ny_);
            psi[1] = dalloc(esize * 2 * h * nx_);
            xi[0] = dalloc(esize * 2 * (h + 1) * ny_);
            xi[1] = dalloc(esize * 2 * (h + 1) * nx_);
        };
        if (own_state) {
            p_ = dalloc(nb);
            v_[0] = dalloc(nbx);
            v_[1] = dalloc(nby);
            mem(psi_, xi_);
        }
        if (d.gradient) {
            g0_ = dalloc(nb);
            g1s_[0] = dalloc(nbx);
            g1s_[1] = dalloc(nby);
            g1_ = dalloc(nb);
            work_ = dalloc(nb);
            total_grad_.push_back(dalloc(nb));
            total_grad_.push_back(dalloc(nb));
            misfit_acc_ = dalloc(sizeof(double));
        }
        if (d.gradient && own_state) {
            ap_ = dalloc(nb);
            av_[0] = dalloc(nbx);
            av_[1] = dalloc(nby);
            mem(psi_adj_, xi_adj_);
            std::vector<DeviceCheckpointer::FieldSpec> fs(4);
            fs[0].comp_bytes = {nb};
            fs[0].buffered = true;                                   // "pcur", width 1
            fs[1].comp_bytes = {nbx, nby};                           // "vcur"
            fs[2].comp_bytes = {psi_[0].bytes, psi_[1].bytes};       // "ψ"
            fs[3].comp_bytes = {xi_[0].bytes, xi_[1].bytes};         // "ξ"
            ckpt_.reset(new DeviceCheckpointer(d.nt, d.check_freq, fs, stream));
            dev_bytes_ += (int64_t)ckpt_->bytes();
        }
        sync();
    }

    void set_material(int nfields, const void *const *fields, int interp, bool on_device) override
    {
        use_device();
        SWB_REQUIRE(nfields == 2, "acoustic variable density takes two material fields (vp, rho)");
        SWB_REQUIRE(interp == 0 || interp == 1, "interp must be 0 (arithmetic) or 1 (harmonic)");
        if (on_device) {
            d2d(vp_.p, fields[0], vp_.bytes);
            d2d(rho_.p, fields[1], rho_.bytes);
        } else {
            upload(vp_.p, fields[0], vp_.bytes);
            upload(rho_.p, fields[1], rho_.bytes);
        }
        interp_ = interp;
        post_vd_facts(desc.dtype, nn_, vp_.p, rho_.p, desc.dt, interp, m0_.p, m1_[0].p, m1_[1].p, stream);
        mat_set_ = true;
    }

    void forward(void *host_seis, int snapevery) override
    {
        begin_shot();
        snapshots_.clear();
        for (int64_t it = 1; it <= desc.nt; ++it) {
            step_forward(it, true);
            if (snapevery > 0 && it % snapevery == 0)
                take_snapshot(it);
        }
        download(host_seis, traces_.p, traces_.bytes);
    }

    void gradient_forward(void *host_seis) override
    {
        SWB_REQUIRE(desc.gradient, "simulation was not built with gradient=true");
        begin_shot();
        ckpt_->reset();
        for (int64_t it = 1; it <= desc.nt; ++it) {
            step_forward(it, true);
            save_all(it); // acou_gradient.jl:132-135
        }
        if (host_seis)
            download(host_seis, traces_.p, traces_.bytes);
        fwd_done_ = true;
    }

    void gradient_adjoint(const void *host_adjsrc) override
    {
        SWB_REQUIRE(fwd_done_, "gradient_adjoint called before gradient_forward");
        use_device();
        upload(adjsrc_.p, host_adjsrc, adjsrc_.bytes);
        adjoint_loop();
    }

    void get_raw_gradient(int which, void *host_out) override
    {
        use_device();
        SWB_REQUIRE(desc.gradient && which >= 0 && which < 3, "acoustic VD raw gradients: 0 grad_m0, 1/2 grad_m1_stag");
        DevBuf &b = which == 0 ? g0_ : g1s_[which - 1];
        download(host_out, b.p, b.bytes);
    }

    void accumulate_gradient(int64_t nsrcpos, const void *srcpos, int rs, int64_t nrecpos, const void *recpos, int rr) override
    {
        use_device();
        SWB_REQUIRE(desc.gradient, "simulation was not built with gradient=true");
        // acou_gradient.jl:177-202
        d2d(work_.p, g0_.p, g0_.bytes);
        post_vd_backinterp(desc.dtype, nn_, rho_.p, interp_, g1s_[0].p, g1s_[1].p, g1_.p, stream);
        DevBuf &sp = mute_pos_[0], &rp = mute_pos_[1]; // persistent scratch (no cudaMalloc / cudaFree inside the shot loop)
        if (rs != 0 && nsrcpos > 0) {
            ensure(sp, esize * nsrcpos * nd_);
            upload(sp.p, srcpos, sp.bytes);
            post_mute(desc.dtype, nd_, nn_, desc.spacing, work_.p, nsrcpos, sp.p, rs, stream);
            post_mute(desc.dtype, nd_, nn_, desc.spacing, g1_.p, nsrcpos, sp.p, rs, stream);
        }
        if (rr != 0 && nrecpos > 0) {
            ensure(rp, esize * nrecpos * nd_);
            upload(rp.p, recpos, rp.bytes);
            post_mute(desc.dtype, nd_, nn_, desc.spacing, work_.p, nrecpos, rp.p, rr, stream);
            post_mute(desc.dtype, nd_, nn_, desc.spacing, g1_.p, nrecpos, rp.p, rr, stream);
        }
        post_vd_chain_accumulate(desc.dtype, ncells(), work_.p, g1_.p, vp_.p, rho_.p, total_grad_[0].p, total_grad_[1].p, stream);
        sync();
    }

    int n_total_gradients() const override { return 2; }

    void get_field(const std::string &name, void *host_out, size_t nbytes) override
    {
        use_device();
        const DevBuf *b = nullptr;
        if (name == "pcur")
            b = &p_;
        else if (name == "vx")
            b = &v_[0];
        else if (name == "vy")
            b = &v_[1];
        else if (name == "fact_m0")
            b = &m0_;
        else if (name == "fact_m1_x")
            b = &m1_[0];
        else if (name == "fact_m1_y")
            b = &m1_[1];
        else if (name == "adjpcur")
            b = &ap_;
        else
            throw Error(SWB_ERR_ARG, "unknown field name: " + name);
        SWB_REQUIRE(b->p != nullptr && nbytes == b->bytes, "field not available or size mismatch");
        download(host_out, b->p, b->bytes);
    }

  protected:
    virtual void begin_shot()
    {
        use_device();
        SWB_REQUIRE(mat_set_, "material properties not set");
        SWB_REQUIRE(shot_bound_, "no shot bound");
        SWB_REQUIRE(cpml_set_[0] && (nd_ == 1 || cpml_set_[1]), "C-PML coefficients not set for every axis");
        zero(p_);
        for (int k = 0; k < 2; ++k) {
            zero(v_[k]);
            zero(psi_[k]);
            zero(xi_[k]);
        }
        if (desc.gradient) {
            zero(g0_);
            zero(ap_);
            for (int k = 0; k < 2; ++k) {
                zero(g1s_[k]);
                zero(av_[k]);
                zero(psi_adj_[k]);
                zero(xi_adj_[k]);
            }
        }
        zero(traces_);
        fwd_done_ = false;
    }

    swb_acou_vd_step_args base_args() const
    {
        swb_acou_vd_step_args a;
        std::memset(&a, 0, sizeof(a));
        a.dtype = desc.dtype;
        a.halo = desc.halo;
        a.flags = desc.flags;
        for (int k = 0; k < 2; ++k) {
            a.n[k] = nn_[k];
            a.spacing[k] = desc.spacing[k];
            a.cpml[k] = cpml_axis(k);
            a.fact_m1_stag[k] = m1_[k].p;
        }
        a.fact_m0 = m0_.p;
        a.stream = stream;
        return a;
    }

    void step_forward(int64_t it, bool record)
    {
        swb_acou_vd_step_args a = base_args();
        a.pcur = p_.p;
        for (int k = 0; k < 2; ++k) {
            a.vcur[k] = v_[k].p;
            a.psi[k] = psi_[k].p;
            a.xi[k] = xi_[k].p;
        }
        a.src.n = nsrc_;
        a.src.pos = possrc_.as<int64_t>();
        a.src.tf = srctf_.p;
        a.src.nt = desc.nt;
        if (record) {
            a.rec.n = nrec_;
            a.rec.pos = posrec_.as<int64_t>();
            a.rec.tf = traces_.p;
            a.rec.nt = desc.nt;
        }
        a.it = it;
        tic();
        vd_step(a, false);
        toc();
        cell_updates += (int64_t)ncells();
    }

    void step_adjoint(int64_t it)
    {
        swb_acou_vd_step_args a = base_args();
        a.pcur = ap_.p;
        for (int k = 0; k < 2; ++k) {
            a.vcur[k] = av_[k].p;
            a.psi[k] = psi_adj_[k].p;
            a.xi[k] = xi_adj_[k].p;
        }
        a.src.n = nrec_;
        a.src.pos = posrec_.as<int64_t>();
        a.src.tf = adjsrc_.p;
        a.src.nt = desc.nt;
        a.it = it;
        tic(1);
        vd_step(a, true);
        toc();
        cell_updates += (int64_t)ncells();
    }

    void save_all(int64_t it)
    {
        ckpt_->save(0, {p_.p}, it);
        ckpt_->save(1, {v_[0].p, v_[1].p}, it);
        ckpt_->save(2, {psi_[0].p, psi_[1].p}, it);
        ckpt_->save(3, {xi_[0].p, xi_[1].p}, it);
    }

    // acou_gradient.jl:141-176
    virtual void adjoint_loop()
    {
        prescale_residuals(desc.dtype, nd_, nn_, adjsrc_.p, desc.nt, nrec_, posrec_.as<int64_t>(), m0_.p, stream);
        for (int64_t it = desc.nt; it >= 1; --it) {
            step_adjoint(it);
            if (!ckpt_->is_saved(0, it - 1)) {
                ckpt_->init_recover();
                const int64_t c = ckpt_->curr();
                d2d(p_.p, ckpt_->get(0, c)[0], p_.bytes);
                auto vs = ckpt_->get(1, c), ps = ckpt_->get(2, c), xs = ckpt_->get(3, c);
                for (int k = 0; k < 2; ++k) {
                    d2d(v_[k].p, vs[k], v_[k].bytes);
                    d2d(psi_[k].p, ps[k], psi_[k].bytes);
                    d2d(xi_[k].p, xs[k], xi_[k].bytes);
                }
                for (int64_t rit = c + 1; rit <= c + ckpt_->check_freq() - 1; ++rit) {
                    step_forward(rit, false);
                    ckpt_->store_recovered(0, {p_.p}, rit);
                }
            }
            const void *p_it = ckpt_->get(0, it)[0];
            const void *p_itm1 = ckpt_->get(0, it - 1)[0];
            vd_correlate_m0(desc.dtype, ncells(), g0_.p, ap_.p, p_it, p_itm1, desc.dt, stream);
            void *g[2] = {g1s_[0].p, g1s_[1].p};
            const void *av[2] = {av_[0].p, av_[1].p};
            vd_correlate_m1(desc.dtype, desc.flags, nn_, desc.spacing, g, av, p_it, stream);
        }
        sync();
    }

    virtual void take_snapshot(int64_t it)
    {
        std::vector<std::vector<char>> comps(3);
        const DevBuf *src[3] = {&p_, &v_[0], &v_[1]};
        for (int k = 0; k < 3; ++k) {
            comps[k].resize(src[k]->bytes);
            download(comps[k].data(), src[k]->p, src[k]->bytes);
        }
        snapshots_[it] = std::move(comps);
    }

    int64_t nx_, ny_, nn_[2] = {0, 0};
    int nd_ = 2;
    int interp_ = 0;
    DevBuf vp_, rho_, m0_, m1_[2], p_, v_[2], psi_[2], xi_[2];
    DevBuf g0_, g1s_[2], g1_, work_, ap_, av_[2], psi_adj_[2], xi_adj_[2], misfit_acc_, obs_, mute_pos_[2];
    std::unique_ptr<DeviceCheckpointer> ckpt_;
    bool mat_set_ = false;
};

#include "engine_vd_fused.inc"

SimBase *make_acoustic_vd(const swb_sim_desc &d)
{
    // the fused single-launch engine is the default; SWB_FLAG_NO_FUSION (or a grid smaller than one stencil)
    // selects the one-launch-per-reference-kernel path
    if (!(d.flags & SWB_FLAG_NO_FUSION) && d.ndim == 2 && d.n[0] >= 8 && d.n[1] >= 8)
        return new AcousticVDFused(d);
    return new AcousticVD(d);
}

} // namespace swb
