"""
ORACLE -- test infrastructure, NOT product code.

Host side of the 2D elastic isotropic P-SV restatement (kernels: oracle/swref_elastic.h).  Follows
/root/reference/src/models/elastic/ela_models.jl:6-90,143-173,177-450 (possrcrec_scaletf, precomp_elaprop!, field
allocation, checkpointer set-up), ela_forward.jl:4-159, ela_gradient.jl:4-362, ela_init_bc.jl:7-40 and the Kaiser-sinc
spreading of /root/reference/src/utils/utils.jl:60-214.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.

Parity status: "parity unpinned" beyond the reference's own known-answer criteria (checkpointed == non-checkpointed
gradient, non-zero gradients: test/test_gradient_elastic_homogeneous.jl:19-119).  Third-party / unspecified behaviour
restated here: SpecialFunctions.besseli(0, x) -> scipy.special.i0; the order in which spread_positions lists the
points of one source / receiver (a Dict iteration order in the reference, utils.jl:136-148) -> ascending index order,
first axis fastest; the per-receiver sum (a re-associable @simd reduction, elastic2D_iso_xPU.jl:224-226) -> left to right.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
from scipy.special import i0 as _besseli0

from . import oracle as O
from .oracle import Field, LinearCheckpointer, Params, _fcopy, _p, _sfx, zeros

# --------------------------------------------------------------------------------------
# Kaiser-windowed sinc spreading -- src/utils/utils.jl:60-214
# --------------------------------------------------------------------------------------


def _kaiser(x, r, beta, T):
    """kaiser(x, r, β) (utils.jl:71): arithmetic in T, besseli through scipy (double) rounded to T."""
    if -r <= x <= r:
        arg = T(beta * T(np.sqrt(T(T(1) - T(T(x / r) * T(x / r))))))
        return T(T(_besseli0(float(arg))) / T(_besseli0(float(beta))))
    return 0.0


def _sinc(x, T):
    x = T(x)
    if x == 0:
        return T(1)
    px = T(T(np.pi) * x)
    return T(np.sin(px) / px)


def coeffsinc1d(x0, dx, nx: int, r: int, beta, xstart, mirror: bool, xbl, xbr, T) -> Tuple[List[int], List[float]]:
    """coeffsinc1D (utils.jl:105-154).  Returns 1-based indices (ascending) and coefficients of type T."""
    T = np.dtype(T).type
    x0, dx, beta, xstart, xbl, xbr = T(x0), T(dx), T(beta), T(xstart), T(xbl), T(xbr)
    xs = (xstart + np.arange(nx, dtype=np.float64) * np.float64(dx)).astype(T)  # range(xstart; length, step): exact products rounded to T

    def findnearest(x):
        return int(np.argmin(np.abs(T(x) - xs))) + 1

    i0 = findnearest(x0)
    idxs, coeffs = [], []
    for idx in range(i0 - r - 1, i0 + r + 2):
        xcurr = T(T(T(idx - 1) * dx) + xstart)
        coe = _kaiser(T(xcurr - x0), T(T(r) * dx), beta, T) * _sinc(T(T(xcurr - x0) / dx), T)
        if not abs(float(coe)) <= 1e-15:
            idxs.append(idx)
            coeffs.append(T(coe))
    acc: Dict[int, float] = {}
    for idx, c in zip(idxs, coeffs):
        xcurr = T(T(T(idx - 1) * dx) + xstart)
        if xcurr < xbl:
            k, v = findnearest(T(xbl + T(xbl - xcurr))), (T(-c) if mirror else c)
        elif xcurr > xbr:
            k, v = findnearest(T(xbr - T(xcurr - xbr))), (T(-c) if mirror else c)
        else:
            k, v = idx, c
        acc[k] = T(acc[k] + v) if k in acc else T(v)
    keys = sorted(acc)
    return keys, [acc[k] for k in keys]


def spread_positions(gridsize, spacing, positions: np.ndarray, shift, mirror: bool, T, r: int = 4, beta=6.31):
    """spread_positions (utils.jl:168-214) with freesurfposition = :ongridbound.
    Returns per position an int64 (npts, N) index matrix (1-based) and a T vector of coefficients (tensor product,
    first axis fastest)."""
    T = np.dtype(T).type
    N = len(gridsize)
    extent = [T(T(spacing[n]) * T(gridsize[n] - 1)) for n in range(N)]
    idxs_out, coef_out = [], []
    for p in range(positions.shape[0]):
        per_dim = [coeffsinc1d(positions[p, n], spacing[n], gridsize[n], r, T(beta), shift[n], mirror, T(0), extent[n], T) for n in range(N)]
        assert N == 2
        (ix, cx), (iz, cz) = per_dim
        ij = np.zeros((len(ix) * len(iz), 2), dtype=np.int64, order="F")
        co = np.zeros(len(ix) * len(iz), dtype=T)
        k = 0
        for b in range(len(iz)):
            for a in range(len(ix)):
                ij[k, 0], ij[k, 1] = ix[a], iz[b]
                co[k] = T(cx[a] * cz[b])
                k += 1
        idxs_out.append(ij)
        coef_out.append(co)
    return idxs_out, coef_out


def to_csr(idxs: List[np.ndarray], coefs: List[np.ndarray], T):
    """flatten per-position point lists: offsets (n+1) int64, ij (npts, 2) int32 column-major, coef (npts) of T"""
    off = np.zeros(len(idxs) + 1, dtype=np.int64)
    for k, a in enumerate(idxs):
        off[k + 1] = off[k] + a.shape[0]
    npts = int(off[-1])
    ij = np.zeros((max(npts, 1), 2), dtype=np.int32, order="F")
    co = np.zeros(max(npts, 1), dtype=T)
    for k, (a, c) in enumerate(zip(idxs, coefs)):
        ij[off[k]:off[k + 1], :] = a
        co[off[k]:off[k + 1]] = c
    return off, ij[:npts, :].copy(order="F") if npts else ij, co[:npts].copy() if npts else co


# --------------------------------------------------------------------------------------
# Shots -- src/shots/sources.jl:33-37,54-90, src/shots/receivers.jl:64-110, src/shots/shot.jl:18-44
# --------------------------------------------------------------------------------------


@dataclass
class MomentTensorShot:
    src_positions: np.ndarray  # (nsrc, 2)
    src_tf: np.ndarray  # (nt, nsrc)
    momtens: np.ndarray  # (nsrc, 3): Mxx, Mzz, Mxz
    domfreq: float
    rec_positions: np.ndarray  # (nrec, 2)
    seismograms: Optional[np.ndarray] = None  # (nt, 2, nrec)


@dataclass
class ExternalForceShot:
    src_positions: np.ndarray
    src_tf: np.ndarray  # (nt, 2, nsrc)
    domfreq: float
    rec_positions: np.ndarray
    seismograms: Optional[np.ndarray] = None


class _GeomEla(C.Structure):
    pass


def _make_geom(T):
    R = C.c_float if np.dtype(T) == np.float32 else C.c_double

    class G(C.Structure):
        _fields_ = [("nx", C.c_long), ("nz", C.c_long), ("halo", C.c_long), ("freetop", C.c_int), ("inv_dx", R), ("inv_dz", R), ("dt", R),
                    ("a_x", C.c_void_p), ("a_xh", C.c_void_p), ("b_x", C.c_void_p), ("b_xh", C.c_void_p),
                    ("a_z", C.c_void_p), ("a_zh", C.c_void_p), ("b_z", C.c_void_p), ("b_zh", C.c_void_p)]

    return G


class ElasticIsoSim:
    """ElasticIsoCPMLWaveSimulation{T,2} (ela_models.jl:177-432) on the CPU."""

    U_NAMES = ("uold", "ucur", "unew")
    PSI_NAMES = ("psi_dsdx", "psi_dsdz", "psi_dudx", "psi_dudz")

    def __init__(self, params: Params, gradient: bool = False, check_freq: int = 1, sincinterp: bool = True, interp_rho: str = "arithmetic",
                 interp_mu: str = "arithmetic"):
        self.p = params
        T = self.T = np.dtype(params.dtype).type
        assert len(params.gridsize) == 2, "Only elastic 2D is currently implemented."
        nx, nz = self.n = tuple(int(v) for v in params.gridsize)
        h = params.halo
        ns_cpml = self.n[:-1] if params.freetop else self.n
        assert all(v >= 2 * h + 3 for v in ns_cpml)
        self.dt = T(params.dt)
        self.spacing = tuple(T(s) for s in params.spacing)
        self.sincinterp, self.interp_rho, self.interp_mu = sincinterp, interp_rho, interp_mu
        self.gradient = gradient
        f: Dict[str, Field] = {}

        def ufield():
            return [zeros((nx - 1, nz), T), zeros((nx, nz - 1), T)]

        def psis(prefix=""):
            f[prefix + "psi_dsdx"] = [zeros((2 * h, nz), T), zeros((2 * (h + 1), nz - 1), T)]  # ψ_∂σxx∂x, ψ_∂σxz∂x
            f[prefix + "psi_dsdz"] = [zeros((nx, 2 * h), T), zeros((nx - 1, 2 * (h + 1)), T)]  # ψ_∂σzz∂z, ψ_∂σxz∂z
            f[prefix + "psi_dudx"] = [zeros((2 * (h + 1), nz), T), zeros((2 * h, nz - 1), T)]  # ψ_∂ux∂x, ψ_∂uz∂x
            f[prefix + "psi_dudz"] = [zeros((nx - 1, 2 * h), T), zeros((nx, 2 * (h + 1)), T)]  # ψ_∂ux∂z, ψ_∂uz∂z

        f["sigma"] = [zeros((nx, nz), T), zeros((nx, nz), T), zeros((nx - 1, nz - 1), T)]
        f["uold"], f["ucur"] = ufield(), ufield()
        f["unew"] = ufield()
        f["lam"], f["mu"] = [zeros((nx, nz), T)], [zeros((nx, nz), T)]
        f["rho_ihalf"], f["rho_jhalf"], f["mu_hh"] = [zeros((nx - 1, nz), T)], [zeros((nx, nz - 1), T)], [zeros((nx - 1, nz - 1), T)]
        psis()
        self.ckpt = None
        if gradient:
            f["adjsigma"] = [zeros((nx, nz), T), zeros((nx, nz), T), zeros((nx - 1, nz - 1), T)]
            f["adjuold"], f["adjucur"], f["adjunew"] = ufield(), ufield(), ufield()
            psis("adj")
            f["grad_lam"], f["grad_mu"] = [zeros((nx, nz), T)], [zeros((nx, nz), T)]
            f["grad_rho_ihalf"], f["grad_rho_jhalf"], f["grad_mu_hh"] = [zeros((nx - 1, nz), T)], [zeros((nx, nz - 1), T)], [zeros((nx - 1, nz - 1), T)]
            ck_fields = {k: f[k] for k in ("ucur",) + self.PSI_NAMES}
            self.ckpt = LinearCheckpointer(params.nt, check_freq, ck_fields, ["ucur"], {"ucur": 2})
            self.ckpt.savecheckpoint("ucur", f["uold"], -1)
            self.ckpt.savecheckpoint("ucur", f["ucur"], 0)
            for k in self.PSI_NAMES:
                self.ckpt.savecheckpoint(k, f[k], 0)
        self.f = f
        self.const = ("lam", "mu", "rho_ihalf", "rho_jhalf", "mu_hh")
        self.rho = zeros(self.n, T)
        self.lam = zeros(self.n, T)
        self.mu = zeros(self.n, T)
        self.cpml: List[O.CPMLAxis] = []
        self.cell_updates = 0

    # update_matprop! / precomp_elaprop! -- ela_models.jl:143-173
    def set_matprop(self, rho: np.ndarray, lam: np.ndarray, mu: np.ndarray) -> None:
        T = self.T
        assert rho.shape == lam.shape == mu.shape == self.n
        assert np.all(lam >= 0) and np.all(mu >= 0) and np.all(rho > 0)
        np.copyto(self.rho, np.asarray(rho, dtype=T))
        np.copyto(self.lam, np.asarray(lam, dtype=T))
        np.copyto(self.mu, np.asarray(mu, dtype=T))
        np.copyto(self.f["lam"][0], self.lam)
        np.copyto(self.f["mu"][0], self.mu)
        np.copyto(self.f["rho_ihalf"][0], O.interp(self.interp_rho, self.rho, [0]).astype(T))
        np.copyto(self.f["rho_jhalf"][0], O.interp(self.interp_rho, self.rho, [1]).astype(T))
        np.copyto(self.f["mu_hh"][0], O.interp(self.interp_mu, self.mu, [0, 1]).astype(T))

    def vel_max(self):
        """ela_init_bc.jl:13: maximum(sqrt.((λ .+ 2 .* μ) ./ ρ)), arithmetic in T"""
        T = self.T
        return T(np.max(np.sqrt(((self.lam + T(2) * self.mu).astype(T) / self.rho).astype(T)).astype(T)))

    def init_shot(self, shot) -> None:
        self.cpml = O.init_bdc(self.vel_max(), self.dt, self.p.halo, self.p.rcoef, self.spacing, self.p.freetop, shot.domfreq, self.T)

    def _geom(self):
        G = _make_geom(self.T)
        g = G()
        g.nx, g.nz, g.halo, g.freetop = self.n[0], self.n[1], self.p.halo, int(self.p.freetop)
        g.inv_dx = float(self.T(1) / self.spacing[0])
        g.inv_dz = float(self.T(1) / self.spacing[1])
        g.dt = float(self.dt)
        cx, cz = self.cpml
        g.a_x, g.a_xh, g.b_x, g.b_xh = cx.a.ctypes.data, cx.a_h.ctypes.data, cx.b.ctypes.data, cx.b_h.ctypes.data
        g.a_z, g.a_zh, g.b_z, g.b_zh = cz.a.ctypes.data, cz.a_h.ctypes.data, cz.b.ctypes.data, cz.b_h.ctypes.data
        return g

    def reset(self) -> None:
        for name, fld in self.f.items():
            if name not in self.const:
                for a in fld:
                    a[...] = 0
        if self.ckpt is not None:
            self.ckpt.reset()

    # possrcrec_scaletf -- ela_models.jl:6-90
    def possrcrec_scaletf(self, shot):
        T = self.T
        dx, dz = self.spacing
        half = (T(dx / T(2)), T(dz / T(2)))
        z = T(0)
        sp = np.asarray(shot.src_positions, dtype=T)
        rp = np.asarray(shot.rec_positions, dtype=T)
        if self.sincinterp:
            def spread(pos, shift, mirror):
                return spread_positions(self.n, self.spacing, pos, shift, mirror, T)
            if isinstance(shot, MomentTensorShot):
                src_a = spread(sp, (z, z), True)  # σxx, σzz
                src_b = spread(sp, half, True)  # σxz
            else:
                src_a = spread(sp, (half[0], z), False)  # ux
                src_b = spread(sp, (z, half[1]), False)  # uz
            rec_a = spread(rp, (half[0], z), False)
            rec_b = spread(rp, (z, half[1]), False)
        else:
            si = O.find_nearest_grid_points(sp, self.spacing, T)
            ri = O.find_nearest_grid_points(rp, self.spacing, T)
            one = lambda idx: ([idx[k:k + 1, :].copy(order="F") for k in range(idx.shape[0])], [np.ones(1, dtype=T) for _ in range(idx.shape[0])])
            src_a, src_b, rec_a, rec_b = one(si), one(si), one(ri), one(ri)
        prod = T(np.prod(np.array(self.spacing, dtype=T), dtype=T))
        tf = np.asfortranarray((np.asarray(shot.src_tf, dtype=T) / prod).astype(T))
        return [to_csr(*x, T) for x in (src_a, src_b, rec_a, rec_b)], tf

    # forward_onestep_CPML! -- elastic2D_iso_xPU.jl:120-357 (both source kinds) and adjoint_onestep_CPML! (:359-448)
    def _step(self, pre: str, g, lists, tf, momtens, traces, it: int, src_kind: str) -> None:
        L, sx = O._L(), _sfx(self.T)
        f = self.f
        sig = f[pre + "sigma"]
        uo, uc, un = f[pre + "uold"], f[pre + "ucur"], f[pre + "unew"]
        ps_dsdx, ps_dsdz, ps_dudx, ps_dudz = (f[pre + k] for k in self.PSI_NAMES)
        lam, mu = f["lam"][0], f["mu"][0]
        (s_a, s_b, r_a, r_b) = lists
        nt = self.p.nt
        getattr(L, "ela_update_sxx_szz" + sx)(C.byref(g), _p(sig[0]), _p(sig[1]), _p(uc[0]), _p(uc[1]), _p(lam), _p(mu), _p(ps_dudx[0]), _p(ps_dudz[1]))
        getattr(L, "ela_update_sxz" + sx)(C.byref(g), _p(sig[2]), _p(uc[0]), _p(uc[1]), _p(f["mu_hh"][0]), _p(ps_dudz[0]), _p(ps_dudx[1]))
        if src_kind == "momten":
            nsrc = len(s_a[0]) - 1
            getattr(L, "ela_inject_momten" + sx)(C.byref(g), _p(sig[0]), _p(sig[1]), _p(sig[2]), C.c_long(nsrc), _p(s_a[0]), _p(s_a[1]), _p(s_a[2]), _p(s_b[0]),
                                                 _p(s_b[1]), _p(s_b[2]), _p(momtens[0]), _p(momtens[1]), _p(momtens[2]), _p(tf), C.c_long(nt), C.c_long(it))
        getattr(L, "ela_update_ux" + sx)(C.byref(g), _p(un[0]), _p(uc[0]), _p(uo[0]), _p(sig[0]), _p(sig[2]), _p(f["rho_ihalf"][0]), _p(ps_dsdx[0]), _p(ps_dsdz[1]))
        getattr(L, "ela_update_uz" + sx)(C.byref(g), _p(un[1]), _p(uc[1]), _p(uo[1]), _p(sig[2]), _p(sig[1]), _p(f["rho_jhalf"][0]), _p(ps_dsdx[1]), _p(ps_dsdz[0]))
        if src_kind == "extforce":
            nsrc = len(s_a[0]) - 1
            getattr(L, "ela_inject_extforce" + sx)(C.byref(g), _p(un[0]), _p(un[1]), _p(f["rho_ihalf"][0]), _p(f["rho_jhalf"][0]), C.c_long(nsrc), _p(s_a[0]),
                                                   _p(s_a[1]), _p(s_a[2]), _p(s_b[0]), _p(s_b[1]), _p(s_b[2]), _p(tf), C.c_long(nt), C.c_long(it))
        if traces is not None:
            nrec = len(r_a[0]) - 1
            getattr(L, "ela_record" + sx)(C.byref(g), _p(un[0]), _p(un[1]), C.c_long(nrec), _p(r_a[0]), _p(r_a[1]), _p(r_a[2]), _p(r_b[0]), _p(r_b[1]), _p(r_b[2]),
                                          _p(traces), C.c_long(nt), C.c_long(it))
        # rotate: uold <- ucur, ucur <- unew, unew <- (old) uold
        f[pre + "uold"], f[pre + "ucur"], f[pre + "unew"] = uc, un, uo
        self.cell_updates += int(np.prod(self.n))

    def _momtens(self, shot):
        T = self.T
        if isinstance(shot, MomentTensorShot):
            m = np.asarray(shot.momtens, dtype=T)
            return [np.ascontiguousarray(m[:, k]) for k in range(3)], "momten"
        return None, "extforce"

    # swforward_1shot! -- ela_forward.jl:4-159
    def forward_1shot(self, shot, snapevery: Optional[int] = None):
        lists, tf = self.possrcrec_scaletf(shot)
        momtens, kind = self._momtens(shot)
        nrec = shot.rec_positions.shape[0]
        traces = zeros((self.p.nt, 2, nrec), self.T)
        self.reset()
        g = self._geom()
        snaps = {}
        for it in range(1, self.p.nt + 1):
            self._step("", g, lists, tf, momtens, traces, it, kind)
            if snapevery is not None and it % snapevery == 0:
                snaps[it] = {"ucur": [a.copy(order="F") for a in self.f["ucur"]], "sigma": [a.copy(order="F") for a in self.f["sigma"]]}
        shot.seismograms = traces
        return snaps

    # swgradient_1shot! -- ela_gradient.jl:4-362
    def gradient_1shot(self, shot, misfit, mute_radius_src: int = 0, mute_radius_rec: int = 0) -> Dict[str, np.ndarray]:
        T, L, sx = self.T, O._L(), _sfx(self.T)
        ck = self.ckpt
        assert ck is not None
        lists, tf = self.possrcrec_scaletf(shot)
        momtens, kind = self._momtens(shot)
        nt = self.p.nt
        nrec = shot.rec_positions.shape[0]
        traces = zeros((nt, 2, nrec), T)
        self.reset()
        g = self._geom()
        for it in range(1, nt + 1):
            self._step("", g, lists, tf, momtens, traces, it, kind)
            ck.savecheckpoint("ucur", self.f["ucur"], it)
            for k in self.PSI_NAMES:
                ck.savecheckpoint(k, self.f[k], it)
        shot.seismograms = traces
        adjsrc = np.asfortranarray((-misfit.dchi_du(traces)).astype(T))
        rec_lists = (lists[2], lists[3], lists[2], lists[3])
        for it in range(nt, 0, -1):
            # adjoint step: residuals injected as external forces at the receivers' sinc points (elastic2D_iso_xPU.jl:433-440)
            self._step("adj", g, rec_lists, adjsrc, None, None, it, "extforce")
            self.cell_updates  # (counted in _step)
            if not ck.issaved("ucur", it - 2):
                ck.initrecover()
                _fcopy(self.f["uold"], ck.getsaved("ucur", ck.curr_checkpoint - 1))
                _fcopy(self.f["ucur"], ck.getsaved("ucur", ck.curr_checkpoint))
                for k in self.PSI_NAMES:
                    _fcopy(self.f[k], ck.getsaved(k, ck.curr_checkpoint))

                def rec(recit):
                    self._step("", g, lists, tf, momtens, None, recit, kind)
                    return [("ucur", self.f["ucur"])]

                ck.recover(rec)
            uo, uc, un = ck.getsaved("ucur", it - 2), ck.getsaved("ucur", it - 1), ck.getsaved("ucur", it)
            a = self.f["adjucur"]
            getattr(L, "ela_correlate" + sx)(C.byref(g), _p(a[0]), _p(a[1]), _p(uo[0]), _p(uo[1]), _p(uc[0]), _p(uc[1]), _p(un[0]), _p(un[1]), _p(self.f["lam"][0]),
                                             _p(self.f["mu"][0]), _p(self.f["grad_rho_ihalf"][0]), _p(self.f["grad_rho_jhalf"][0]), _p(self.f["grad_lam"][0]),
                                             _p(self.f["grad_mu"][0]), _p(self.f["grad_mu_hh"][0]))
        grad_lam = self.f["grad_lam"][0].copy(order="F")
        grad_mu = self.f["grad_mu"][0].copy(order="F")
        grad_rho = zeros(self.n, T)
        # gradient_ρ .+= back_interp(ρ_ihalf, 1) .+ back_interp(ρ_jhalf, 2); gradient_μ .+= back_interp(μ_ihalf_jhalf, [1, 2])
        bi = (O.back_interp(self.interp_rho, self.rho, self.f["grad_rho_ihalf"][0], [0]) + O.back_interp(self.interp_rho, self.rho, self.f["grad_rho_jhalf"][0], [1])).astype(T)
        grad_rho = (grad_rho + bi).astype(T)
        grad_mu = (grad_mu + O.back_interp(self.interp_mu, self.mu, self.f["grad_mu_hh"][0], [0, 1])).astype(T)
        srcp = np.asarray(shot.src_positions, dtype=T)
        recp = np.asarray(shot.rec_positions, dtype=T)
        for arr in (grad_rho, grad_lam, grad_mu):
            O.mutearoundmultiplepoints(arr, srcp, self.spacing, mute_radius_src)
        for arr in (grad_rho, grad_lam, grad_mu):
            O.mutearoundmultiplepoints(arr, recp, self.spacing, mute_radius_rec)
        return {"rho": np.asfortranarray(grad_rho), "lambda": np.asfortranarray(grad_lam), "mu": np.asfortranarray(grad_mu)}
