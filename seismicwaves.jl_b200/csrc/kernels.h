// kernels.h -- launchers shared between the physics files, the engine and the C ABI.
#pragma once
#include "common.cuh"

namespace swb {

// acou_cd.cu
void cd_step(const swb_acou_cd_step_args &a, bool record);
void cd_correlate(int dtype, int flags, size_t ncells, void *grad, const void *adj, const void *pm2, const void *pm1, const void *p0, double dt, cudaStream_t st);
void prescale_residuals(int dtype, int ndim, const int64_t *n, void *res, int64_t nt, int64_t nrec, const int64_t *pos, const void *fact, cudaStream_t st);
template <class T>
void launch_inject(T *p, int ndim, const int64_t *n, const swb_points &pts, int64_t it, cudaStream_t st);
template <class T>
void launch_record(const T *p, int ndim, const int64_t *n, const swb_points &pts, int64_t it, cudaStream_t st);

// acou_vd.cu
void vd_step(const swb_acou_vd_step_args &a, bool adjoint);
void vd_correlate_m0(int dtype, size_t ncells, void *g, const void *adjp, const void *p_it, const void *p_itm1, double dt, cudaStream_t st);
void vd_correlate_m1(int dtype, int flags, const int64_t *n, const double *spacing, void *const g[2], const void *const av[2], const void *p, cudaStream_t st);

// ela_iso.cu
void ela_step(const swb_ela_step_args &a, bool adjoint);
void ela_correlate(const swb_ela_correlate_args &a);

// post.cu -- device-side host-prep / post-processing kernels (material factors, mute, back_interp, chain rules, misfit)
void post_cd_fact(int dtype, size_t n, const void *vp, double dt, void *fact, cudaStream_t st);
void post_vd_facts(int dtype, const int64_t *n, const void *vp, const void *rho, double dt, int interp, void *m0, void *m1x, void *m1y, cudaStream_t st);
void post_mute(int dtype, int ndim, const int64_t *n, const double *spacing, void *arr, int64_t npos, const void *dev_positions, int radius, cudaStream_t st);
void post_cd_chain_accumulate(int dtype, size_t n, const void *grad_raw_muted, const void *vp, void *total, cudaStream_t st);
void post_vd_backinterp(int dtype, const int64_t *n, const void *rho, int interp, const void *g1x, const void *g1y, void *g1, cudaStream_t st);
void post_vd_chain_accumulate(int dtype, size_t n, const void *g0, const void *g1, const void *vp, const void *rho, void *tot_vp, void *tot_rho, cudaStream_t st);
void post_ela_props(int dtype, const int64_t *n, const void *rho, const void *mu, int interp_rho, int interp_mu, void *rho_ih, void *rho_jh, void *mu_hh, cudaStream_t st);
void post_ela_backinterp(int dtype, const int64_t *n, const void *rho, const void *mu, int interp_rho, int interp_mu, const void *g_ri, const void *g_rj,
                         const void *g_mh, const void *g_mu, void *out_rho, void *out_mu, cudaStream_t st);
void post_l2_adjsrc(int dtype, size_t n, size_t nt, const void *syn, const void *obs_or_null, const void *w_or_null, const void *mask_or_null, void *adjsrc,
                    double *misfit_accum, cudaStream_t st);
void post_axpy(int dtype, size_t n, const void *x, void *y, cudaStream_t st); // y += x

} // namespace swb
