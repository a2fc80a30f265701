// engine_ela.cu -- per-shot engine of the 2D elastic isotropic P-SV simulation.
//
// Replaces swforward_1shot! / swgradient_1shot! of src/models/elastic/ela_forward.jl:4-159 and
// src/models/elastic/ela_gradient.jl:4-362 (both source kinds): the time loops, the LinearCheckpointer schedule
// ("ucur" with width 2 + the four ψ groups, ela_models.jl:383-396), the re-forwarding and the correlations run inside the
// library on the sim's stream; the host sees seismograms, the adjoint source and the finished gradients only.
#include "engine.h"
#include "ela_fused.h"
#include <algorithm>
#include <cstring>

namespace swb {

class ElasticIso : public SimBase {
  public:
    // own_state = false: a subclass keeps the wavefield state in its own layout (fused engine below); the dense state is
    // then allocated on demand only (snapshots run through this class's step-by-step path)
    explicit ElasticIso(const swb_sim_desc &d, bool own_state = true) : SimBase(d)
    {
        SWB_REQUIRE(d.ndim == 2, "Only elastic 2D is currently implemented.");
        nx_ = d.n[0];
        nz_ = d.n[1];
        SWB_REQUIRE(nx_ >= 6 && nz_ >= 6, "elastic grid must have at least 6 points per axis");
        const size_t h = (size_t)d.halo;
        nb_ = (size_t)nx_ * nz_ * esize;
        nbx_ = (size_t)(nx_ - 1) * nz_ * esize;
        nbz_ = (size_t)nx_ * (nz_ - 1) * esize;
        nbxz_ = (size_t)(nx_ - 1) * (nz_ - 1) * esize;
        // ψ_∂σ∂x {σxx, σxz}, ψ_∂σ∂z {σzz, σxz}, ψ_∂u∂x {ux, uz}, ψ_∂u∂z {ux, uz}  (ela_models.jl:275-299)
        psib_[0][0] = esize * 2 * h * nz_, psib_[0][1] = esize * 2 * (h + 1) * (nz_ - 1);
        psib_[1][0] = esize * nx_ * 2 * h, psib_[1][1] = esize * (nx_ - 1) * 2 * (h + 1);
        psib_[2][0] = esize * 2 * (h + 1) * nz_, psib_[2][1] = esize * 2 * h * (nz_ - 1);
        psib_[3][0] = esize * (nx_ - 1) * 2 * h, psib_[3][1] = esize * nx_ * 2 * (h + 1);
        rho_ = dalloc(nb_), lam_ = dalloc(nb_), mu_ = dalloc(nb_);
        rho_ih_ = dalloc(nbx_), rho_jh_ = dalloc(nbz_), mu_hh_ = dalloc(nbxz_);
        if (own_state) {
            alloc_state(fw_);
            fw_ready_ = true;
        }
        if (d.gradient) {
            if (own_state)
                alloc_state(ad_);
            g_ri_ = dalloc(nbx_), g_rj_ = dalloc(nbz_), g_l_ = dalloc(nb_), g_m_ = dalloc(nb_), g_mh_ = dalloc(nbxz_);
            for (int k = 0; k < 3; ++k) {
                work_[k] = dalloc(nb_);
                total_grad_.push_back(dalloc(nb_));
            }
            if (own_state) {
                std::vector<DeviceCheckpointer::FieldSpec> fs(5);
                fs[0].comp_bytes = {nbx_, nbz_};
                fs[0].width = 2;
                fs[0].buffered = true; // "ucur"
                for (int gq = 0; gq < 4; ++gq)
                    fs[1 + gq].comp_bytes = {psib_[gq][0], psib_[gq][1]};
                ckpt_.reset(new DeviceCheckpointer(d.nt, d.check_freq, fs, stream));
                dev_bytes_ += (int64_t)ckpt_->bytes();
            }
            misfit_acc_ = dalloc(sizeof(double));
        }
        sync();
    }

    // update_matprop! + precomp_elaprop! (ela_models.jl:143-173): fields = {rho, lambda, mu};
    // interp = interp_method_ρ + 2 * interp_method_μ (0 arithmetic, 1 harmonic each)
    void set_material(int nfields, const void *const *fields, int interp, bool on_device) override
    {
        use_device();
        SWB_REQUIRE(nfields == 3, "elastic isotropic takes three material fields (rho, lambda, mu)");
        SWB_REQUIRE(interp >= 0 && interp <= 3, "interp must encode interp_rho + 2*interp_mu with 0 = arithmetic, 1 = harmonic");
        DevBuf *dst[3] = {&rho_, &lam_, &mu_};
        for (int k = 0; k < 3; ++k) {
            if (on_device)
                d2d(dst[k]->p, fields[k], nb_);
            else
                upload(dst[k]->p, fields[k], nb_);
        }
        interp_rho_ = interp & 1;
        interp_mu_ = (interp >> 1) & 1;
        post_ela_props(desc.dtype, desc.n, rho_.p, mu_.p, interp_rho_, interp_mu_, rho_ih_.p, rho_jh_.p, mu_hh_.p, stream);
        mat_set_ = true;
    }

    void bind_elastic_shot(int src_kind, const swb_sinc_points_host src_pts[2], const void *srctf, const void *Mxx, const void *Mzz, const void *Mxz,
                           const swb_sinc_points_host rec_pts[2]) override
    {
        use_device();
        SWB_REQUIRE(src_kind == 1 || src_kind == 2, "src_kind must be 1 (moment tensor) or 2 (external force)");
        SWB_REQUIRE(src_pts[0].n > 0 && src_pts[0].n == src_pts[1].n, "There must be at least one source!");
        SWB_REQUIRE(rec_pts[0].n > 0 && rec_pts[0].n == rec_pts[1].n, "There must be at least one receiver!");
        src_kind_ = src_kind;
        nsrc_ = src_pts[0].n;
        nrec_ = rec_pts[0].n;
        // index ranges of the arrays the lists point into
        const int64_t lim_src[2][2] = {{src_kind == 1 ? nx_ : nx_ - 1, nz_}, {src_kind == 1 ? nx_ - 1 : nx_, nz_ - 1}};
        const int64_t lim_rec[2][2] = {{nx_ - 1, nz_}, {nx_, nz_ - 1}};
        for (int k = 0; k < 2; ++k) {
            upload_list(src_pts[k], src_l_[k], lim_src[k], "source");
            upload_list(rec_pts[k], rec_l_[k], lim_rec[k], "receiver");
        }
        ensure(srctf_, esize * desc.nt * nsrc_ * (src_kind == 2 ? 2 : 1));
        upload(srctf_.p, srctf, srctf_.bytes);
        if (src_kind == 1) {
            SWB_REQUIRE(Mxx && Mzz && Mxz, "moment tensor components missing");
            ensure(mt_[0], esize * nsrc_), ensure(mt_[1], esize * nsrc_), ensure(mt_[2], esize * nsrc_);
            upload(mt_[0].p, Mxx, mt_[0].bytes);
            upload(mt_[1].p, Mzz, mt_[1].bytes);
            upload(mt_[2].p, Mxz, mt_[2].bytes);
        }
        ensure(traces_, esize * desc.nt * 2 * nrec_);
        if (desc.gradient)
            ensure(adjsrc_, esize * desc.nt * 2 * nrec_);
        shot_bound_ = true;
        fwd_done_ = false;
    }

    void forward(void *host_seis, int snapevery) override
    {
        if (!fw_ready_) {
            alloc_state(fw_);
            fw_ready_ = true;
        }
        ElasticIso::begin_shot();
        snapshots_.clear();
        for (int64_t it = 1; it <= desc.nt; ++it) {
            step(fw_, false, it, true);
            if (snapevery > 0 && it % snapevery == 0)
                take_snapshot(it);
        }
        download(host_seis, traces_.p, traces_.bytes);
    }

    void gradient_forward(void *host_seis) override
    {
        SWB_REQUIRE(desc.gradient, "simulation was not built with gradient=true");
        begin_shot();
        gradient_forward_body(host_seis);
    }

    // dense wavefield state, adjoint state and checkpoint storage of the step-by-step path, for a subclass that normally keeps its own
    void ensure_unfused_state()
    {
        if (!fw_ready_) {
            alloc_state(fw_);
            fw_ready_ = true;
        }
        if (desc.gradient && !ckpt_) {
            alloc_state(ad_);
            std::vector<DeviceCheckpointer::FieldSpec> fs(5);
            fs[0].comp_bytes = {nbx_, nbz_};
            fs[0].width = 2;
            fs[0].buffered = true; // "ucur"
            for (int gq = 0; gq < 4; ++gq)
                fs[1 + gq].comp_bytes = {psib_[gq][0], psib_[gq][1]};
            ckpt_.reset(new DeviceCheckpointer(desc.nt, desc.check_freq, fs, stream));
            dev_bytes_ += (int64_t)ckpt_->bytes();
        }
    }

    void gradient_forward_body(void *host_seis)
    {
        ckpt_->reset();
        for (int64_t it = 1; it <= desc.nt; ++it) {
            step(fw_, false, it, true);
            save_all(it); // ela_gradient.jl:77-81
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
        SWB_REQUIRE(desc.gradient && which >= 0 && which < 5, "elastic raw gradients: 0 ρ_ihalf, 1 ρ_jhalf, 2 λ, 3 μ, 4 μ_ihalf_jhalf");
        DevBuf *b[5] = {&g_ri_, &g_rj_, &g_l_, &g_m_, &g_mh_};
        download(host_out, b[which]->p, b[which]->bytes);
    }

    // ela_gradient.jl:155-186: back_interp, mute (sources then receivers), accumulate_gradient! (ela_models.jl:446-450)
    void accumulate_gradient(int64_t nsrcpos, const void *srcpos, int rs, int64_t nrecpos, const void *recpos, int rr) override
    {
        use_device();
        SWB_REQUIRE(desc.gradient, "simulation was not built with gradient=true");
        post_ela_backinterp(desc.dtype, desc.n, rho_.p, mu_.p, interp_rho_, interp_mu_, g_ri_.p, g_rj_.p, g_mh_.p, g_m_.p, work_[0].p, work_[2].p, stream);
        d2d(work_[1].p, g_l_.p, nb_);
        if (rs != 0 && nsrcpos > 0) {
            ensure(mute_pos_[0], esize * nsrcpos * 2);
            upload(mute_pos_[0].p, srcpos, mute_pos_[0].bytes);
            for (int k = 0; k < 3; ++k)
                post_mute(desc.dtype, 2, desc.n, desc.spacing, work_[k].p, nsrcpos, mute_pos_[0].p, rs, stream);
        }
        if (rr != 0 && nrecpos > 0) {
            ensure(mute_pos_[1], esize * nrecpos * 2);
            upload(mute_pos_[1].p, recpos, mute_pos_[1].bytes);
            for (int k = 0; k < 3; ++k)
                post_mute(desc.dtype, 2, desc.n, desc.spacing, work_[k].p, nrecpos, mute_pos_[1].p, rr, stream);
        }
        for (int k = 0; k < 3; ++k) // totgrad["rho" | "lambda" | "mu"] .+= curgrad
            post_axpy(desc.dtype, (size_t)nx_ * nz_, work_[k].p, total_grad_[k].p, stream);
        sync();
    }

    int n_total_gradients() const override { return 3; }

    void get_field(const std::string &name, void *host_out, size_t nbytes) override
    {
        use_device();
        const void *src = nullptr;
        size_t b = 0;
        if (name == "ux")
            src = fw_.u[1][0], b = nbx_;
        else if (name == "uz")
            src = fw_.u[1][1], b = nbz_;
        else if (name == "sxx")
            src = fw_.sig[0].p, b = nb_;
        else if (name == "szz")
            src = fw_.sig[1].p, b = nb_;
        else if (name == "sxz")
            src = fw_.sig[2].p, b = nbxz_;
        else if (name == "rho_ihalf")
            src = rho_ih_.p, b = nbx_;
        else if (name == "rho_jhalf")
            src = rho_jh_.p, b = nbz_;
        else if (name == "mu_ihalf_jhalf")
            src = mu_hh_.p, b = nbxz_;
        else
            throw Error(SWB_ERR_ARG, "unknown field name: " + name);
        SWB_REQUIRE(src != nullptr && nbytes == b, "field not available or size mismatch");
        download(host_out, src, b);
    }

  protected:
    struct State { // one wavefield state (forward or adjoint): three displacement time levels, σ, eight ψ arrays
        DevBuf ubuf[3][2], sig[3], psi[4][2];
        void *u[3][2]; // rotating handles: [0] old, [1] cur, [2] new
    };
    struct DevList {
        DevBuf off, ij, coef;
        int64_t n = 0, npts = 0;
    };

    void alloc_state(State &s)
    {
        for (int k = 0; k < 3; ++k) {
            s.ubuf[k][0] = dalloc(nbx_);
            s.ubuf[k][1] = dalloc(nbz_);
        }
        s.sig[0] = dalloc(nb_), s.sig[1] = dalloc(nb_), s.sig[2] = dalloc(nbxz_);
        for (int gq = 0; gq < 4; ++gq)
            for (int c = 0; c < 2; ++c)
                s.psi[gq][c] = dalloc(psib_[gq][c]);
        reset_handles(s);
    }
    static void reset_handles(State &s)
    {
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 2; ++c)
                s.u[k][c] = s.ubuf[k][c].p;
    }
    void zero_state(State &s)
    {
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 2; ++c)
                zero(s.ubuf[k][c]);
        for (int k = 0; k < 3; ++k)
            zero(s.sig[k]);
        for (int gq = 0; gq < 4; ++gq)
            for (int c = 0; c < 2; ++c)
                zero(s.psi[gq][c]);
        reset_handles(s);
    }

    void upload_list(const swb_sinc_points_host &h, DevList &d, const int64_t lim[2], const char *what)
    {
        SWB_REQUIRE(h.off != nullptr && h.off[0] == 0, "sinc point list: offsets must start at 0");
        const int64_t npts = h.off[h.n];
        for (int64_t p = 0; p < npts; ++p)
            SWB_REQUIRE(h.ij[p] >= 1 && h.ij[p] <= lim[0] && h.ij[p + npts] >= 1 && h.ij[p + npts] <= lim[1],
                        std::string(what) + " sinc point outside the array it is applied to (position too close to the grid edge)");
        d.n = h.n;
        d.npts = npts;
        ensure(d.off, sizeof(int64_t) * (h.n + 1));
        ensure(d.ij, sizeof(int32_t) * 2 * std::max<int64_t>(npts, 1));
        ensure(d.coef, esize * std::max<int64_t>(npts, 1));
        upload(d.off.p, h.off, d.off.bytes);
        if (npts > 0) {
            upload(d.ij.p, h.ij, sizeof(int32_t) * 2 * npts);
            upload(d.coef.p, h.coef, esize * npts);
        }
    }
    static swb_sinc_points dev_points(const DevList &d)
    {
        swb_sinc_points p;
        p.n = d.n;
        p.off = d.off.as<int64_t>();
        p.ij = d.ij.as<int32_t>();
        p.coef = d.coef.p;
        return p;
    }

    // reset! (ela_models.jl:436-442)
    virtual void begin_shot()
    {
        use_device();
        SWB_REQUIRE(mat_set_, "material properties not set");
        SWB_REQUIRE(shot_bound_, "no shot bound");
        SWB_REQUIRE(cpml_set_[0] && cpml_set_[1], "C-PML coefficients not set for every axis");
        zero_state(fw_);
        if (desc.gradient && ckpt_) {
            zero_state(ad_);
            zero(g_ri_), zero(g_rj_), zero(g_l_), zero(g_m_), zero(g_mh_);
        }
        zero(traces_);
        fwd_done_ = false;
    }

    // forward_onestep_CPML! / adjoint_onestep_CPML! + handle rotation (elastic2D_iso_xPU.jl:120-448)
    void step(State &s, bool adjoint, int64_t it, bool record)
    {
        swb_ela_step_args a;
        std::memset(&a, 0, sizeof(a));
        a.dtype = desc.dtype;
        a.halo = desc.halo;
        a.flags = desc.flags;
        a.freetop = desc.freetop;
        for (int k = 0; k < 2; ++k) {
            a.n[k] = desc.n[k];
            a.spacing[k] = desc.spacing[k];
            a.cpml[k] = cpml_axis(k);
            a.uold[k] = s.u[0][k];
            a.ucur[k] = s.u[1][k];
            a.unew[k] = s.u[2][k];
            a.psi_dsdx[k] = s.psi[0][k].p;
            a.psi_dsdz[k] = s.psi[1][k].p;
            a.psi_dudx[k] = s.psi[2][k].p;
            a.psi_dudz[k] = s.psi[3][k].p;
        }
        a.dt = desc.dt;
        for (int k = 0; k < 3; ++k)
            a.sigma[k] = s.sig[k].p;
        a.lambda = lam_.p, a.mu = mu_.p, a.rho_ihalf = rho_ih_.p, a.rho_jhalf = rho_jh_.p, a.mu_ihalf_jhalf = mu_hh_.p;
        if (adjoint) { // residuals as external forces at the receivers' sinc points
            a.src_kind = 2;
            a.src_pts[0] = dev_points(rec_l_[0]);
            a.src_pts[1] = dev_points(rec_l_[1]);
            a.srctf = adjsrc_.p;
        } else {
            a.src_kind = src_kind_;
            a.src_pts[0] = dev_points(src_l_[0]);
            a.src_pts[1] = dev_points(src_l_[1]);
            a.srctf = srctf_.p;
            a.Mxx = mt_[0].p, a.Mzz = mt_[1].p, a.Mxz = mt_[2].p;
            if (record) {
                a.rec_pts[0] = dev_points(rec_l_[0]);
                a.rec_pts[1] = dev_points(rec_l_[1]);
                a.traces = traces_.p;
                a.nt_tr = desc.nt;
            }
        }
        a.nt_tf = desc.nt;
        a.it = it;
        a.stream = stream;
        tic(adjoint ? 1 : 0);
        ela_step(a, adjoint);
        toc();
        for (int c = 0; c < 2; ++c) { // uold <- ucur, ucur <- unew, unew <- uold (aliases from now on, as in the reference)
            s.u[0][c] = s.u[1][c];
            s.u[1][c] = s.u[2][c];
            s.u[2][c] = s.u[0][c];
        }
        cell_updates += (int64_t)ncells();
    }

    std::vector<const void *> psi_ptrs(const State &s, int gq) const { return {s.psi[gq][0].p, s.psi[gq][1].p}; }

    void save_all(int64_t it)
    {
        ckpt_->save(0, {fw_.u[1][0], fw_.u[1][1]}, it);
        for (int gq = 0; gq < 4; ++gq)
            ckpt_->save(1 + gq, psi_ptrs(fw_, gq), it);
    }

    // ela_gradient.jl:93-154
    virtual void adjoint_loop()
    {
        for (int64_t it = desc.nt; it >= 1; --it) {
            step(ad_, true, it, false);
            if (!ckpt_->is_saved(0, it - 2)) {
                ckpt_->init_recover();
                const int64_t c = ckpt_->curr();
                auto uo = ckpt_->get(0, c - 1), uc = ckpt_->get(0, c);
                d2d(fw_.u[0][0], uo[0], nbx_), d2d(fw_.u[0][1], uo[1], nbz_);
                d2d(fw_.u[1][0], uc[0], nbx_), d2d(fw_.u[1][1], uc[1], nbz_);
                for (int gq = 0; gq < 4; ++gq) {
                    auto ps = ckpt_->get(1 + gq, c);
                    for (int k = 0; k < 2; ++k)
                        d2d(fw_.psi[gq][k].p, ps[k], psib_[gq][k]);
                }
                for (int64_t rit = c + 1; rit <= c + ckpt_->check_freq() - 1; ++rit) {
                    step(fw_, false, rit, false);
                    ckpt_->store_recovered(0, {fw_.u[1][0], fw_.u[1][1]}, rit);
                }
            }
            auto uo = ckpt_->get(0, it - 2), uc = ckpt_->get(0, it - 1), un = ckpt_->get(0, it);
            swb_ela_correlate_args ca;
            std::memset(&ca, 0, sizeof(ca));
            ca.dtype = desc.dtype;
            ca.flags = desc.flags;
            ca.freetop = desc.freetop;
            ca.dt = desc.dt;
            for (int k = 0; k < 2; ++k) {
                ca.n[k] = desc.n[k];
                ca.spacing[k] = desc.spacing[k];
                ca.adjucur[k] = ad_.u[1][k];
                ca.u_itm2[k] = uo[k];
                ca.u_itm1[k] = uc[k];
                ca.u_it[k] = un[k];
            }
            ca.lambda = lam_.p, ca.mu = mu_.p;
            ca.grad_rho_ihalf = g_ri_.p, ca.grad_rho_jhalf = g_rj_.p, ca.grad_lambda = g_l_.p, ca.grad_mu = g_m_.p, ca.grad_mu_ihalf_jhalf = g_mh_.p;
            ca.stream = stream;
            ela_correlate(ca);
        }
        sync();
    }

    void take_snapshot(int64_t it)
    { // savesnapshot! of "ucur" and "σ" (ela_forward.jl:64-67)
        std::vector<std::vector<char>> comps(5);
        const void *src[5] = {fw_.u[1][0], fw_.u[1][1], fw_.sig[0].p, fw_.sig[1].p, fw_.sig[2].p};
        const size_t by[5] = {nbx_, nbz_, nb_, nb_, nbxz_};
        for (int k = 0; k < 5; ++k) {
            comps[k].resize(by[k]);
            download(comps[k].data(), src[k], by[k]);
        }
        snapshots_[it] = std::move(comps);
    }

    int64_t nx_, nz_;
    size_t nb_, nbx_, nbz_, nbxz_, psib_[4][2];
    int interp_rho_ = 0, interp_mu_ = 0, src_kind_ = 0;
    DevBuf rho_, lam_, mu_, rho_ih_, rho_jh_, mu_hh_;
    State fw_, ad_;
    DevBuf g_ri_, g_rj_, g_l_, g_m_, g_mh_, work_[3], misfit_acc_, obs_, mt_[3], mute_pos_[2];
    DevList src_l_[2], rec_l_[2];
    std::unique_ptr<DeviceCheckpointer> ckpt_;
    bool mat_set_ = false, fw_ready_ = false;
};

#include "engine_ela_fused.inc"

SimBase *make_elastic_iso(const swb_sim_desc &d)
{
    // the fused engine (stresses on chip, one stencil launch per step) is the default; SWB_FLAG_NO_FUSION selects the
    // one-launch-per-reference-kernel path
    if (!(d.flags & SWB_FLAG_NO_FUSION) && d.n[0] >= 16 && d.n[1] >= 16)
        return new ElasticIsoFused(d);
    return new ElasticIso(d);
}

} // namespace swb
