/*
 * swb200.h -- C ABI of libswb200.so, the B200 (sm_100a) time-stepping engine that sits behind
 * SeismicWaves.jl's backend-module API.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the
 * SeismicWaves.jl v0.9.0 tree).  The Julia side binds these with `ccall` from
 * ext/SeismicWaves_B200BackendExt.jl; tests bind them with ctypes.
 *
 * Conventions
 *  - Every function returns int32 status, 0 = OK.  swb_last_error() returns a thread-local,
 *    NUL-terminated description of the last failure on the calling thread.
 *  - Plain C types only.  Device memory is owned by the library (swb_malloc / sims) or borrowed
 *    from the caller as raw device pointers; host buffers are borrowed for the duration of a call.
 *  - Arrays are column-major (first index fastest) exactly like Julia Arrays.  Grid positions are
 *    1-based int64 as produced by find_nearest_grid_points (src/utils/utils.jl:47-58).
 *  - `dtype` selects Float32 / Float64 storage (the reference's type parameter T).  With Float32
 *    the kernels default to the reference's promotion rule (Float64 literals => Float64
 *    intermediates, one rounding per store; src/utils/fdgen.jl:61-63,131); SWB_FLAG_FAST_F32
 *    switches to pure-Float32 arithmetic.
 *  - Scalars of type T cross the ABI as double (exact for Float32).
 *  - There is no CPU fallback: every compute entry point fails with SWB_ERR_CUDA if no sm_100
 *    device is usable.
 */
#ifndef SWB200_H
#define SWB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWB_ABI_VERSION 2

enum { SWB_OK = 0, SWB_ERR_ARG = 1, SWB_ERR_CUDA = 2, SWB_ERR_STATE = 3, SWB_ERR_NOMEM = 4, SWB_ERR_NCCL = 5 };
enum { SWB_F32 = 0, SWB_F64 = 1 };
/* simulation kinds = the reference's concrete WaveSimulation types */
enum {
    SWB_ACOU_CD = 1, /* AcousticCDCPMLWaveSimulation{T,N}, N = 1,2,3 (src/models/acoustic/acou_models.jl:68) */
    SWB_ACOU_VD = 2, /* AcousticVDStaggeredCPMLWaveSimulation{T,N}, N = 1,2 (src/models/acoustic/acou_models.jl:311) */
    SWB_ELA_ISO = 3  /* ElasticIsoCPMLWaveSimulation{T,2}           (src/models/elastic/ela_models.jl:177) */
};
enum {
    SWB_FLAG_FAST_F32 = 1,   /* Float32 storage AND Float32 arithmetic (default: Float64 intermediates like the reference) */
    SWB_FLAG_NO_FUSION = 2,  /* engine uses the one-kernel-per-reference-kernel path (parity debugging) */
    SWB_FLAG_NO_GRAPH = 4    /* engine does not capture CUDA graphs */
};

const char *swb_last_error(void);
int32_t swb_abi_version(void);
/* Layout of every struct that crosses this ABI, as the compiler that built the library sees it -- a JSON object
 * {"<struct>": {"size": bytes, "fields": [["<member>", offset, size], ...]}, ...} in declaration order.  Bindings written in
 * other languages (ext/SeismicWaves_B200BackendExt.jl, seismicwaves.jl_b200/_lib.py) are checked against it without needing
 * their runtime (tests/test_julia_boundary.py).  Never fails; the string is static. */
const char *swb_abi_layout(void);
/* number of usable sm_100 devices (0 if none); never fails */
int32_t swb_device_count(void);
/* kernel launches issued by this process through the library so far (bench.py's gpu_launches) */
int64_t swb_launch_count(void);
/* Host-only self check (no device needed) of how the fused acoustic CD step divides an (nx, ny, nz) grid among its kernels -- the
 * register-queue march of the interior, the z-marched x / y strip boxes and the per-vector strips / faces (a 2D
 * grid (nx, nz2d) is passed as (nx, 1, nz2d)): every cell must be owned by exactly one (kernel, CTA, in-CTA slot) and every slot must
 * lie inside its CTA's range, because the per-CTA source / receiver lists address cells that way.  esize = 4 / 8; zc = planes per bulk
 * chunk; zpml_lo / zpml_hi = 0 for a z end without C-PML strip (free surface, interior face of a z slab); rim_zc = planes per chunk
 * of the rim march (< 0: the engine's default, 0: no march).  counts[3] receives the cells owned by the three kernel kinds. */
int32_t swb_diag_cd_partition(int32_t esize, int32_t nx, int32_t ny, int32_t nz, int32_t halo, int32_t zc, int32_t zpml_lo, int32_t zpml_hi,
                              int32_t rim_zc, int64_t *counts);

/* ---------------------------------------------------------------------------------------------
 * 1. Device buffers -- backs the Julia `B200Array{T,N}` that plays the role of backend.Data.Array /
 *    backend.zeros (acou_models.jl:112-157, cpmlcoeffs.jl:8-15) and the generic array operations
 *    L3/L4 code applies to backend arrays (copyto!, zero, fill, Array(dev)).
 * ------------------------------------------------------------------------------------------- */
int32_t swb_set_device(int32_t device);
int32_t swb_malloc(void **dev_ptr, size_t nbytes);  /* zero-filled, like backend.zeros */
int32_t swb_free(void *dev_ptr);
int32_t swb_memcpy_h2d(void *dev_dst, const void *host_src, size_t nbytes); /* copyto!(dev, host) / Data.Array(host) */
int32_t swb_memcpy_d2h(void *host_dst, const void *dev_src, size_t nbytes); /* copyto!(host, dev) / Array(dev) */
int32_t swb_memcpy_d2d(void *dev_dst, const void *dev_src, size_t nbytes);  /* copyto!(dev, dev) (fields.jl:20,32-36) */
int32_t swb_fill(void *dev_ptr, int32_t dtype, double value, size_t first, size_t count); /* dev[first+1:first+count] .= value */
int32_t swb_synchronize(void);

/* ---------------------------------------------------------------------------------------------
 * 2. Backend-module functions, one call per reference function, operating on caller-owned device
 *    arrays (the fine-grained drop-in used by the Julia backend modules and by parity tests).
 * ------------------------------------------------------------------------------------------- */

/* C-PML coefficient vectors of one axis, device pointers of type T:
 * a,b: length 2(halo+1); a_h,b_h: length 2halo  (CPMLCoefficientsAxis, src/models/cpmlcoeffs.jl:1-16) */
typedef struct {
    const void *a, *a_h, *b, *b_h;
} swb_cpml_axis;

/* point sources / receivers of an acoustic shot on the device */
typedef struct {
    int64_t n;            /* number of positions */
    const int64_t *pos;   /* device, (n, ndim) column-major, 1-based grid indices */
    void *tf;             /* device, (nt, n) of T: scaled source time functions, or traces to write */
    int64_t nt;           /* leading dimension of tf */
} swb_points;

/* Acoustic constant density, N = 1, 2 or 3.
 * Replaces forward_onestep_CPML! / adjoint_onestep_CPML! of
 * src/models/acoustic/backends/shared/acoustic1D_xPU.jl:63-125, acoustic2D_xPU.jl:78-169 and acoustic3D_xPU.jl:96-219
 * (kernels update_ψ_*!, update_p_CPML!, inject_sources!, record_receivers!).
 * The caller rotates the field handles afterwards exactly as the reference does (acoustic2D_xPU.jl:121-124);
 * pnew may alias pold. */
typedef struct {
    int32_t dtype, ndim, halo, flags;
    int64_t n[3];
    double spacing[3];
    void *pold, *pcur, *pnew;
    const void *fact;
    void *psi[3]; /* extent 2halo along its axis */
    void *xi[3];  /* extent 2(halo+1) along its axis */
    swb_cpml_axis cpml[3];
    swb_points src;  /* injected into pnew: pnew[pos] += tf[it, s] */
    swb_points rec;  /* rec.n = 0 or rec.tf = NULL <=> save_trace = false */
    int64_t it;      /* 1-based time index */
    void *stream;    /* cudaStream_t, NULL = default stream */
} swb_acou_cd_step_args;
int32_t swb_acou_cd_forward_onestep(const swb_acou_cd_step_args *args);
int32_t swb_acou_cd_adjoint_onestep(const swb_acou_cd_step_args *args); /* same operator, never records */

/* correlate_gradient! -- src/models/acoustic/backends/shared/correlate_gradient_xPU.jl:1-10
 * grad += adjcur * (p_itm2 - 2.0*p_itm1 + p_it) / dt^2  (argument order as called at acou_gradient.jl:77-81) */
int32_t swb_acou_cd_correlate_gradient(int32_t dtype, int32_t flags, size_t ncells, void *grad, const void *adjcur,
                                       const void *p_itm2, const void *p_itm1, const void *p_it, double dt, void *stream);

/* prescale_residuals! -- acoustic2D_xPU.jl:64-76 (same in 3D / VD): res[it, r] *= fact[pos_r] */
int32_t swb_prescale_residuals(int32_t dtype, int32_t ndim, const int64_t *n, void *residuals, int64_t nt, int64_t nrec,
                               const int64_t *posrecs, const void *fact, void *stream);

/* Acoustic variable density (staggered), N = 2, and N = 1 as the single-row case n = {nx, 1} (the y arrays are then unused and may be NULL).
 * Replaces forward_onestep_CPML! / adjoint_onestep_CPML! of
 * src/models/acoustic/backends/shared/acoustic2D_VD_xPU.jl:91-178 (update_p_CPML!, inject_sources!,
 * update_vx_CPML!, update_vy_CPML!, record_receivers!) and acoustic1D_VD_xPU.jl:77-126.  All fields are updated in place. */
typedef struct {
    int32_t dtype, halo, flags, _pad;
    int64_t n[2];
    double spacing[2];
    void *pcur;        /* (nx, ny) */
    void *vcur[2];     /* (nx-1, ny), (nx, ny-1) */
    const void *fact_m0;
    const void *fact_m1_stag[2];
    void *psi[2];      /* (2halo, ny), (nx, 2halo) -- memory variables of the v updates */
    void *xi[2];       /* (2(halo+1), ny), (nx, 2(halo+1)) -- memory variables of the p update */
    swb_cpml_axis cpml[2];
    swb_points src;
    swb_points rec;
    int64_t it;
    void *stream;
} swb_acou_vd_step_args;
int32_t swb_acou_vd_forward_onestep(const swb_acou_vd_step_args *args);  /* p, inject, v, record */
int32_t swb_acou_vd_adjoint_onestep(const swb_acou_vd_step_args *args);  /* v, p, inject */

/* correlate_gradient_m0! -- acoustic/backends/shared/correlate_gradient_xPU.jl:12-21: grad_m0 -= adjp*(p_it - p_itm1)/dt */
int32_t swb_acou_vd_correlate_gradient_m0(int32_t dtype, int32_t flags, size_t ncells, void *grad_m0, const void *adjp,
                                          const void *p_it, const void *p_itm1, double dt, void *stream);
/* correlate_gradient_m1! -- acoustic2D_VD_xPU.jl:180-199: grad_m1_d += adjv_d * ∂_d p_it (plain 4-point stencil);
 * n = {nx, 1}: the 1D method, which runs over (2:nx-2) only (acoustic1D_VD_xPU.jl:127-140) */
int32_t swb_acou_vd_correlate_gradient_m1(int32_t dtype, int32_t flags, const int64_t *n, const double *spacing, void *const grad_m1_stag[2],
                                          const void *const adjv[2], const void *p_it, void *stream);

/* Elastic isotropic P-SV, N = 2 (displacement-stress).
 * Replaces forward_onestep_CPML! (moment-tensor and external-force methods), adjoint_onestep_CPML! and
 * correlate_gradients! of src/models/elastic/backends/shared/elastic2D_iso_xPU.jl:120-448 and
 * src/models/elastic/backends/shared/correlate_gradient_xPU.jl:47-83.
 * Sinc-spread point lists (spread_positions, src/utils/utils.jl:168-214) are flattened to CSR:
 * point p of source/receiver s lives at [off[s], off[s+1]). */
typedef struct {
    int64_t n;            /* number of sources / receivers */
    const int64_t *off;   /* device, n+1 offsets */
    const int32_t *ij;    /* device, (npts, 2) column-major, 1-based (i, j) */
    const void *coef;     /* device, npts of T */
} swb_sinc_points;

typedef struct {
    int32_t dtype, halo, flags, freetop;
    int64_t n[2];          /* nx, nz */
    double spacing[2];
    double dt;
    void *uold[2], *ucur[2], *unew[2];   /* ux (nx-1, nz), uz (nx, nz-1); unew may alias uold */
    void *sigma[3];                      /* sxx (nx,nz), szz (nx,nz), sxz (nx-1,nz-1) */
    const void *lambda, *mu, *rho_ihalf, *rho_jhalf, *mu_ihalf_jhalf;
    void *psi_dsdx[2];  /* ψ_∂σxx∂x (2halo, nz),      ψ_∂σxz∂x (2(halo+1), nz-1) */
    void *psi_dsdz[2];  /* ψ_∂σzz∂z (nx, 2halo),      ψ_∂σxz∂z (nx-1, 2(halo+1)) */
    void *psi_dudx[2];  /* ψ_∂ux∂x (2(halo+1), nz),   ψ_∂uz∂x (2halo, nz-1) */
    void *psi_dudz[2];  /* ψ_∂ux∂z (nx-1, 2halo),     ψ_∂uz∂z (nx, 2(halo+1)) */
    swb_cpml_axis cpml[2];
    /* sources: kind 0 = none, 1 = moment tensor (pts[0] on σxx/σzz, pts[1] on σxz; tf (nt, nsrc)),
     *          2 = external force / adjoint source (pts[0] on ux, pts[1] on uz; tf (nt, 2, nsrc)) */
    int32_t src_kind, _pad;
    swb_sinc_points src_pts[2];
    const void *srctf;
    int64_t nt_tf;
    const void *Mxx, *Mzz, *Mxz;  /* device vectors of T, moment tensor components per source */
    /* receivers: rec_pts[0] on ux, rec_pts[1] on uz; traces (nt, 2, nrec); NULL <=> save_trace=false */
    swb_sinc_points rec_pts[2];
    void *traces;
    int64_t nt_tr;
    int64_t it;
    void *stream;
} swb_ela_step_args;
int32_t swb_ela_forward_onestep(const swb_ela_step_args *args);
int32_t swb_ela_adjoint_onestep(const swb_ela_step_args *args);

typedef struct {
    int32_t dtype, flags, freetop, _pad;
    int64_t n[2];
    double spacing[2];
    double dt;
    const void *adjucur[2];
    const void *u_itm2[2], *u_itm1[2], *u_it[2];  /* uold_corr, ucur_corr, unew_corr of ela_gradient.jl:148-153 */
    const void *lambda, *mu;
    void *grad_rho_ihalf, *grad_rho_jhalf, *grad_lambda, *grad_mu, *grad_mu_ihalf_jhalf;
    void *stream;
} swb_ela_correlate_args;
int32_t swb_ela_correlate_gradients(const swb_ela_correlate_args *args);

/* ---------------------------------------------------------------------------------------------
 * 3. Per-shot engine (the performance path): owns all fields of one WaveSimulation on one GPU,
 *    runs whole time loops with fused kernels, device-side checkpoint storage with the
 *    LinearCheckpointer schedule (src/utils/checkpointers.jl) and CUDA-graph replay.
 *    Replaces swforward_1shot! / swgradient_1shot! (src/models/acoustic/acou_forward.jl:22-125,
 *    acou_gradient.jl:4-203, src/models/elastic/ela_forward.jl:4-159, ela_gradient.jl:4-362).
 * ------------------------------------------------------------------------------------------- */
typedef struct swb_sim swb_sim;

typedef struct {
    int32_t kind, dtype, ndim, device;
    int64_t n[3];
    double spacing[3];
    double dt;
    int64_t nt;
    int32_t halo, freetop;
    int32_t gradient;    /* allocate adjoint fields, gradient accumulators and checkpoint storage */
    int32_t check_freq;  /* GradParameters.check_freq (src/models/genparameters.jl:63-78); 1 = keep every step */
    int32_t flags;
    int32_t _pad;
} swb_sim_desc;

int32_t swb_sim_create(const swb_sim_desc *desc, swb_sim **sim);
int32_t swb_sim_destroy(swb_sim *sim);
int64_t swb_sim_device_bytes(const swb_sim *sim);

/* update_matprop! + precompute_fact! / precomp_elaprop! on the device
 * (acou_models.jl:53-60,283-300; ela_models.jl:143-173).  Host arrays of T, grid-sized:
 *   SWB_ACOU_CD: {vp};  SWB_ACOU_VD: {vp, rho};  SWB_ELA_ISO: {rho, lambda, mu}.
 * interp: 0 = ArithmeticAverageInterpolation, 1 = HarmonicAverageInterpolation (for the density / mu staggering). */
int32_t swb_sim_set_material(swb_sim *sim, int32_t nfields, const void *const *host_fields, int32_t interp);
/* same, but the arrays already live on this sim's device (no PCIe traffic) */
int32_t swb_sim_set_material_device(swb_sim *sim, int32_t nfields, const void *const *dev_fields, int32_t interp);

/* copyto!(cpmlcoeffs.a, ...) etc. -- uploads host-computed coefficient vectors for one axis
 * (compute_CPML_coefficientsAxis!, cpmlcoeffs.jl:18-41; free-surface override acou_init_bc.jl:32-39) */
int32_t swb_sim_set_cpml(swb_sim *sim, int32_t axis, const void *a, const void *a_h, const void *b, const void *b_h);

/* Bind an acoustic shot (ScalarShot after possrcrec_scaletf, acou_forward.jl:6-20,67-81):
 * host int64 positions (n, ndim) column-major 1-based, scaled STF (nt, nsrc) of T. */
int32_t swb_sim_bind_scalar_shot(swb_sim *sim, int64_t nsrc, const int64_t *possrcs, const void *srctf, int64_t nrec, const int64_t *posrecs);

/* Bind an elastic shot after possrcrec_scaletf (ela_models.jl:6-90): CSR sinc lists on the host.
 * src_kind 1: moment tensor (lists on σxx/σzz and σxz, srctf (nt, nsrc), M* (nsrc));
 * src_kind 2: external force (lists on ux and uz, srctf (nt, 2, nsrc)). */
typedef struct {
    int64_t n;
    const int64_t *off;  /* host, n+1 */
    const int32_t *ij;   /* host, (npts, 2) column-major 1-based */
    const void *coef;    /* host, npts of T */
} swb_sinc_points_host;
int32_t swb_sim_bind_elastic_shot(swb_sim *sim, int32_t src_kind, const swb_sinc_points_host src_pts[2], const void *srctf,
                                  const void *Mxx, const void *Mzz, const void *Mxz, const swb_sinc_points_host rec_pts[2]);

/* swforward_1shot!: reset!, nt forward steps, seismograms copied to the host buffer
 * ((nt, nrec) acoustic, (nt, 2, nrec) elastic).  snapevery > 0 additionally stores the wavefield every
 * snapevery steps (savesnapshot!, src/utils/snapshotter.jl:33-41) for swb_sim_get_snapshot. */
int32_t swb_sim_forward(swb_sim *sim, void *host_seismograms, int32_t snapevery);
int32_t swb_sim_get_snapshot(swb_sim *sim, int64_t it, int32_t field /*0 = p / ux, 1 = uz*/, void *host_out);

/* swgradient_1shot!, split at the point where the reference hands control to the (pluggable) misfit:
 *   swb_sim_gradient_forward : reset!, forward loop with savecheckpoint! (acou_gradient.jl:33-45) -> seismograms
 *   swb_sim_gradient_adjoint : host adjoint source (= -∂χ/∂u, (nt, nrec) or (nt, 2, nrec)), prescale_residuals!
 *                              (acoustic only), adjoint loop with re-forwarding and zero-lag correlation
 *                              (acou_gradient.jl:55-82).  The correlated fields stay on the device. */
int32_t swb_sim_gradient_forward(swb_sim *sim, void *host_seismograms);
int32_t swb_sim_gradient_adjoint(swb_sim *sim, const void *host_adjsrc);
/* same two phases with an identity-covariance L2 misfit evaluated on the device:
 * observed: host (nt, nrec[,2]) or NULL (= zeros).  misfit_out (may be NULL) receives dot(r, r)/2. */
int32_t swb_sim_gradient_l2(swb_sim *sim, const void *host_observed, void *host_seismograms_or_null, double *misfit_out);

/* The same with the windows and a diagonal inverse covariance of the reference's L2Misfit (src/inversion/misfits/L2Misfit.jl:24-95)
 * applied on the device: r = syn - observed; r = mask .* r (windows, L2Misfit.jl:28-34); adjoint source = -(invcov_diag .* r)
 * (= -∂χ_∂u for a Diagonal invcov, L2Misfit.jl:64-77); misfit = dot(r, invcov_diag .* r) / 2.  Host vectors of T with one entry per
 * time sample; NULL = all ones (no window / identity covariance); observed NULL = zeros.  Dense covariances and other
 * AbstractMisfits use swb_sim_gradient_forward / swb_sim_gradient_adjoint with the adjoint source computed by the caller. */
typedef struct {
    const void *observed;     /* host (nt, nrec) or (nt, 2, nrec) of T, or NULL */
    const void *invcov_diag;  /* host (nt) of T, or NULL */
    const void *mask;         /* host (nt) of T (0 / 1), or NULL */
} swb_l2_spec;
int32_t swb_sim_gradient_l2_ex(swb_sim *sim, const swb_l2_spec *spec, void *host_seismograms_or_null, double *misfit_out);

/* Raw correlated fields of the last shot (what the reference downloads with Array(...)):
 *  SWB_ACOU_CD: 0 grad_vp;  SWB_ACOU_VD: 0 grad_m0, 1 grad_m1_stag[1], 2 grad_m1_stag[2];
 *  SWB_ELA_ISO: 0 grad_ρ_ihalf, 1 grad_ρ_jhalf, 2 grad_λ, 3 grad_μ, 4 grad_μ_ihalf_jhalf. */
int32_t swb_sim_get_raw_gradient(swb_sim *sim, int32_t which, void *host_out);

/* Device-side gradient post-processing of the last shot: back_interp of staggered gradients
 * (src/utils/interpolations.jl:14-28), mutearoundmultiplepoints! around sources then receivers
 * (src/utils/mute_grad.jl:3-80; positions in metres, host (npos, ndim) of T), chain rule
 * (acou_gradient.jl:93,199-202; ela_gradient.jl:155-186), and accumulation into the sim's total gradient
 * (accumulate_gradient!, acou_models.jl:64,302-305).  */
int32_t swb_sim_accumulate_gradient(swb_sim *sim, int64_t nsrcpos, const void *src_positions, int32_t mute_radius_src,
                                    int64_t nrecpos, const void *rec_positions, int32_t mute_radius_rec);
int32_t swb_sim_zero_total_gradient(swb_sim *sim);
/* total gradient component: CD {vp}; VD {vp, rho}; elastic {rho, lambda, mu}.  Device pointer (for an
 * NCCL all-reduce by the caller) or a copy to the host. */
int32_t swb_sim_total_gradient_ptr(swb_sim *sim, int32_t which, void **dev_ptr, size_t *nelem);
int32_t swb_sim_get_total_gradient(swb_sim *sim, int32_t which, void *host_out);

/* bookkeeping for the benchmark: cell-updates (forward + re-forward + adjoint sweeps) executed since creation */
int64_t swb_sim_cell_updates(const swb_sim *sim);
/* debugging / tests: download a named field ("pcur", "vx", "vy", "ux", "uz", "fact", ...) */
int32_t swb_sim_get_field(swb_sim *sim, const char *name, void *host_out, size_t nbytes);
int32_t swb_sim_stream(swb_sim *sim, void **stream_out);
/* timing of the dominant kernel on the sim's own stream (CUDA events): accumulates while enabled */
int32_t swb_sim_kernel_timing(swb_sim *sim, int32_t enable, double *ms_total, int64_t *launches);
/* same totals split by launch class: 0 = forward / re-forward step, 1 = adjoint step (with the fused correlation) */
int32_t swb_sim_kernel_timing_class(swb_sim *sim, int32_t cls, double *ms_total, int64_t *launches);

/* ---------------------------------------------------------------------------------------------
 * 4. Multi-GPU: shots are sharded across ranks by the caller (distribsrcs, src/utils/utils.jl:28-45);
 *    the per-rank total gradients and misfits are summed with one NCCL all-reduce per array.
 *    (The reference has no counterpart: its shot loop is sequential, src/apis/gradient.jl:115-134.)
 * ------------------------------------------------------------------------------------------- */
typedef struct swb_comm swb_comm;
int32_t swb_comm_unique_id(void *id128 /* 128 bytes out */);
int32_t swb_comm_create(const void *id128, int32_t nranks, int32_t rank, int32_t device, swb_comm **comm);
int32_t swb_comm_destroy(swb_comm *comm);
int32_t swb_comm_allreduce_sum(swb_comm *comm, void *dev_ptr, size_t nelem, int32_t dtype, void *stream);
int32_t swb_sim_allreduce_total_gradient(swb_sim *sim, swb_comm *comm);

/* z-slab domain decomposition of a 3D acoustic constant-density FORWARD simulation too large for one GPU (BASELINE
 * config 5; the reference has no counterpart).  The last axis is cut into contiguous slabs, one per rank.  Each rank creates
 * an ordinary SWB_ACOU_CD sim whose n[2] counts its owned planes plus one ghost plane per interior face, and then declares
 * the faces: lower_rank / upper_rank = the neighbour owning the planes below k = 1 / above k = n[2], or -1 at a true domain end
 * (C-PML strip, Dirichlet face).  After every time step the first / last owned plane is sent to the neighbour's ghost plane
 * with NCCL point-to-point transfers on the sim's stream.  Call before swb_sim_bind_scalar_shot; positions are slab-local. */
int32_t swb_sim_set_slab(swb_sim *sim, swb_comm *comm, int32_t lower_rank, int32_t upper_rank);

/* Halo exchange through peer memory instead of NCCL (NVLink / NVSwitch peer stores fused into the step kernels): after swb_sim_set_slab every
 * slab exports a handle of its two rotating pressure planes and of its flag word, the caller passes the handles around (Julia: same
 * process; Python: one all-gather of 256 bytes per rank), and every slab connects to its neighbours' handles.  From then on the kernels
 * that compute the first / last owned plane also store it into the neighbour's ghost plane, and neighbouring slabs are kept within one
 * time step of each other by stream-ordered 32-bit flag writes / waits (cuStreamWriteValue32 / cuStreamWaitValue32) in each other's
 * memory: no collective, no host synchronisation, no extra pass over the planes.  Handles from the same process are used as plain
 * pointers (peer access is enabled between devices), handles from another process are opened with cudaIpcOpenMemHandle.  Every slab of a
 * decomposition must run the same sequence of shots; a slab driven from one host thread per sim (its forward call blocks until its
 * neighbours have caught up).  comm may be NULL in swb_sim_set_slab when this exchange is used. */
typedef struct {
    uint8_t ipc[192];         /* three cudaIpcMemHandle_t (64 bytes each): rotating plane 0, rotating plane 1, flag buffer */
    uint64_t raw[3];          /* the same three device pointers, valid inside the exporting process */
    int64_t nz;               /* planes of the exporting slab (ghost planes included) */
    int64_t plane_elems;      /* elements per plane (row pitch x ny) */
    int32_t device, pid;
} swb_slab_handle;
int32_t swb_sim_slab_export(swb_sim *sim, swb_slab_handle *handle_out);
int32_t swb_sim_slab_connect(swb_sim *sim, const swb_slab_handle *lower_or_null, const swb_slab_handle *upper_or_null);

#ifdef __cplusplus
}
#endif
#endif /* SWB200_H */
