"""ElasticIsoCPMLWaveSimulation{T,2} on the libswb200 engine (src/models/elastic/ela_models.jl:92-450).

Host side = what the reference does in Julia before / after the backend calls: check_matprop, check_numerics, init_bdc!
(ela_init_bc.jl:7-40), possrcrec_scaletf with Kaiser-sinc spreading (ela_models.jl:6-90).  The time loops, checkpointing,
correlation, back-interpolation, muting and accumulation run inside the library.
"""
from __future__ import annotations

import ctypes as C
import logging
import math

import numpy as np

from . import _lib, hostprep
from .api import WaveSimulation, _as_T, _vp
from .types import (ElasticIsoMaterialProperties, ExternalForceShot, ExternalForceSources, MomentTensorShot, MomentTensorSources, VectorReceivers)

log = logging.getLogger("seismicwaves_b200")


def _csr(idxs, coefs, T):
    """per-position point lists -> (offsets int64, ij int32 (npts, 2) column-major, coef T)"""
    off = np.zeros(len(idxs) + 1, dtype=np.int64)
    for k, a in enumerate(idxs):
        off[k + 1] = off[k] + a.shape[0]
    npts = int(off[-1])
    ij = np.zeros((npts, 2), dtype=np.int32, order="F")
    co = np.zeros(npts, dtype=T)
    for k, (a, c) in enumerate(zip(idxs, coefs)):
        ij[off[k]:off[k + 1], :] = a
        co[off[k]:off[k + 1]] = c
    return off, ij, co


class ElasticIsoCPMLWaveSimulation(WaveSimulation):
    kind = _lib.SWB_ELA_ISO
    grad_names = ("rho", "lambda", "mu")
    _dominant_kernel = "ela_fused_kernel (stresses on chip: one stencil launch per time step; adjoint launches carry the five correlations)"

    def __init__(self, params, matprop, cpmlparams, runparams, gradparams, gradient: bool = False, sincinterp: bool = True):
        assert len(params.gridsize) == 2, "Only elastic 2D is currently implemented."
        super().__init__(params, matprop, cpmlparams, runparams, gradparams, gradient=gradient)
        self.sincinterp = sincinterp

    # ---- checks (ela_models.jl:95-141, utils/checks.jl) -------------------------------------------------
    def check_sim_consistency(self, matprop, shots) -> None:
        ts, tr = type(shots[0].srcs), type(shots[0].recs)
        for s in shots:
            if type(s.srcs) is not ts or type(s.recs) is not tr:
                raise TypeError("Types of shots are inconsistent.")
        ok = isinstance(matprop, ElasticIsoMaterialProperties) and issubclass(ts, (MomentTensorSources, ExternalForceSources)) \
            and issubclass(tr, VectorReceivers) and matprop.rho.ndim == self.N and matprop.rho.dtype == self.dtype
        if not ok:
            raise TypeError(f"Types of WaveSimulation, MaterialProperties and Sources/Receivers are inconsistent \n {type(self)}, \n {type(matprop)}, \n {ts}, {tr}")

    def set_wavesim_matprop(self, matprop) -> None:
        T = self.T
        lam, mu, rho = matprop.lam, matprop.mu, matprop.rho
        assert np.all(lam >= 0), "Lamè coefficient λ must be positive!"
        assert np.all(mu >= 0), "Lamè coefficient μ must be positive!"
        assert np.all(rho > 0), "Density must be positive!"
        assert rho.shape == lam.shape == mu.shape == self.gridsize, f"Material property number of grid points must be the same as the wavesim! \n {rho.shape}, {self.gridsize}"
        vp = np.sqrt(((lam + T(2) * mu).astype(T) / rho).astype(T)).astype(T)
        self._vp_max = T(np.max(vp))
        vs = np.sqrt((mu / rho).astype(T))
        vmin = float(np.min(vs))
        self._vel_min = vmin if vmin != 0 else float(np.min(vp))
        tmp = math.sqrt(sum(1.0 / float(s) ** 2 for s in self.spacing))
        courant = float(self._vp_max) * float(self.dt) * tmp * 7 / 6  # 7/6 comes from the higher order stencil
        log.info("Courant number: %s", courant)
        if self.runparams.erroronCFL:
            assert courant < 1, f"Courant condition not satisfied! [{courant}]"
        elif courant > 1:
            log.warning("CFL condition not satisfied! [%s]", courant)
        self.matprop = ElasticIsoMaterialProperties(rho.copy(order="F"), lam.copy(order="F"), mu.copy(order="F"), interp_method_rho=matprop.interp_method_rho,
                                                    interp_method_lam=matprop.interp_method_lam, interp_method_mu=matprop.interp_method_mu)
        arr = (C.c_void_p * 3)(self.matprop.rho.ctypes.data, self.matprop.lam.ctypes.data, self.matprop.mu.ctypes.data)
        code = (0 if matprop.interp_method_rho == "arithmetic" else 1) + 2 * (0 if matprop.interp_method_mu == "arithmetic" else 1)
        _lib.check(self.lib.swb_sim_set_material(self._h, 3, arr, code))

    def init_shot(self, shot) -> None:
        self.check_numerics(shot, self._vel_min)
        self.check_positions(shot.srcs.positions)
        self.check_positions(shot.recs.positions)
        self._set_cpml(self._vp_max, shot.srcs.domfreq)

    # ---- possrcrec_scaletf (ela_models.jl:6-90) -----------------------------------------------------------
    def _lists(self, shot):
        T = self.T
        dx, dz = (T(s) for s in self.spacing)
        half = (T(dx / T(2)), T(dz / T(2)))
        z = T(0)
        sp, rp = _as_T(shot.srcs.positions, T), _as_T(shot.recs.positions, T)
        momten = isinstance(shot, MomentTensorShot)
        if self.sincinterp:
            def spread(pos, shift, mirror):
                return hostprep.spread_positions(self.gridsize, self.spacing, pos, shift, mirror, "ongridbound", T)
            if momten:
                src = (spread(sp, (z, z), True), spread(sp, half, True))  # σxx/σzz, σxz
            else:
                src = (spread(sp, (half[0], z), False), spread(sp, (z, half[1]), False))  # ux, uz
            rec = (spread(rp, (half[0], z), False), spread(rp, (z, half[1]), False))
        else:
            si = hostprep.find_nearest_grid_points(sp, self.spacing, T)
            ri = hostprep.find_nearest_grid_points(rp, self.spacing, T)

            def one(idx):
                return [idx[k:k + 1, :].copy(order="F") for k in range(idx.shape[0])], [np.ones(1, dtype=T) for _ in range(idx.shape[0])]
            src, rec = (one(si), one(si)), (one(ri), one(ri))
        prod = T(np.prod(np.array(self.spacing, dtype=T), dtype=T))
        tf = np.asfortranarray((np.asarray(shot.srcs.tf, dtype=T) / prod).astype(T))
        return [_csr(*x, T) for x in src], [_csr(*x, T) for x in rec], tf

    def _bind(self, shot) -> None:
        T = self.T
        src, rec, tf = self._lists(shot)
        nrec = shot.recs.positions.shape[0]
        assert shot.recs.seismograms.shape == (self.nt, 2, nrec) and shot.recs.seismograms.dtype == self.dtype
        assert tf.shape[0] == self.nt, "source time function length must equal the number of timesteps"
        momten = isinstance(shot, MomentTensorShot)

        def pts(lists):
            arr = (_lib.swb_sinc_points * 2)()
            for k, (off, ij, co) in enumerate(lists):
                arr[k].n = len(off) - 1
                arr[k].off = off.ctypes.data
                arr[k].ij = ij.ctypes.data
                arr[k].coef = co.ctypes.data
            return arr

        s_arr, r_arr = pts(src), pts(rec)
        keep = [src, rec, tf]
        if momten:
            M = [np.array([getattr(m, c) for m in shot.srcs.momtens], dtype=T) for c in ("Mxx", "Mzz", "Mxz")]
            keep.append(M)
            _lib.check(self.lib.swb_sim_bind_elastic_shot(self._h, 1, s_arr, _vp(tf), _vp(M[0]), _vp(M[1]), _vp(M[2]), r_arr))
        else:
            _lib.check(self.lib.swb_sim_bind_elastic_shot(self._h, 2, s_arr, _vp(tf), None, None, None, r_arr))
        self._bound = keep

    def _snapshot_fields(self):
        nx, nz = self.gridsize
        return {"ucur": [(nx - 1, nz), (nx, nz - 1)], "σ": [(nx, nz), (nx, nz), (nx - 1, nz - 1)]}
