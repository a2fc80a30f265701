"""Public API mirror: build_wavesim, swforward!, swmisfit!, swgradient! (src/apis/{build,forward,misfit,gradient}.jl)
on top of the libswb200 per-shot engine.  Function names drop Julia's `!`.

The shot loop, the consistency checks and their error messages follow the reference
(src/apis/forward.jl:71-115, src/apis/gradient.jl:93-136, src/utils/checks.jl:2-94,
src/models/acoustic/acou_models.jl:5-51,240-281); the time loops, checkpointing, correlation and gradient
post-processing run inside the library.
"""
from __future__ import annotations

import ctypes as C
import logging
import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import _lib, hostprep
from .types import (CPMLBoundaryConditionParameters, ElasticIsoMaterialProperties, ExternalForceShot, ExternalForceSources, GradParameters,
                    InputParametersAcoustic, InputParametersElastic, L2Misfit, MomentTensorShot, MomentTensorSources, RunParameters, ScalarReceivers,
                    ScalarShot, ScalarSources, VectorReceivers, VpAcousticCDMaterialProperties, VpRhoAcousticVDMaterialProperties)

log = logging.getLogger("seismicwaves_b200")


def _vp(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _as_T(a, T) -> np.ndarray:
    return np.asfortranarray(np.asarray(a, dtype=T))


class WaveSimulation:
    """One reference WaveSimulation object (AcousticCDCPML / AcousticVDStaggeredCPML / ElasticIsoCPML) bound to one GPU."""

    kind = 0
    grad_names: Tuple[str, ...] = ()

    def __init__(self, params, matprop, cpmlparams: CPMLBoundaryConditionParameters, runparams: RunParameters,
                 gradparams: Optional[GradParameters], gradient: bool = False):
        if runparams.parall != "B200":
            raise ValueError(f"this package only provides the :B200 backend (got parall={runparams.parall!r}); there is no CPU fallback")
        self.lib = _lib.load()
        _lib.require_device()
        self.params, self.cpmlparams, self.runparams, self.gradparams = params, cpmlparams, runparams, gradparams
        self.T = params.dtype.type
        self.dtype = params.dtype
        self.nt, self.dt = params.ntimesteps, params.dt
        self.gridsize, self.spacing = params.gridsize, params.gridspacing
        self.N = len(self.gridsize)
        self.gradient = bool(gradient)
        halo = cpmlparams.halo
        assert halo >= 0, "CPML halo size must be non-negative!"
        ns_cpml = self.gridsize[:-1] if cpmlparams.freeboundtop else self.gridsize
        assert all(n >= 2 * halo + 3 for n in ns_cpml), f"Number grid points in the dimensions with C-PML boundaries must be at least 2*halo+3 = {2 * halo + 3}!"
        if runparams.infoevery is not None:
            assert 1 <= runparams.infoevery <= self.nt, "Infoevery parameter must be positive and less then nt!"
        if runparams.snapevery is not None:
            assert runparams.snapevery < self.nt, "Checkpointing frequency must be smaller than the number of timesteps!"
        d = _lib.swb_sim_desc()
        d.kind = self.kind
        d.dtype = _lib.SWB_F32 if self.dtype == np.float32 else _lib.SWB_F64
        d.ndim = self.N
        d.device = runparams.device
        for k in range(self.N):
            d.n[k] = self.gridsize[k]
            d.spacing[k] = float(self.spacing[k])
        d.dt = float(self.dt)
        d.nt = self.nt
        d.halo = halo
        d.freetop = int(cpmlparams.freeboundtop)
        d.gradient = int(self.gradient)
        d.check_freq = gradparams.check_freq if (gradparams is not None and self.gradient) else 1
        d.flags = ((_lib.SWB_FLAG_FAST_F32 if runparams.fast_f32 else 0) | (0 if runparams.fused else _lib.SWB_FLAG_NO_FUSION)
                   | (0 if getattr(runparams, "graphs", True) else _lib.SWB_FLAG_NO_GRAPH))
        self._h = C.c_void_p()
        _lib.check(self.lib.swb_sim_create(C.byref(d), C.byref(self._h)))
        self.matprop = None
        self.extent = tuple(self.T(self.spacing[k] * self.T(self.gridsize[k] - 1)) for k in range(self.N))

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                self.lib.swb_sim_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:  # interpreter shutdown
            pass

    close = __del__

    # ---- checks mirrored from the reference ---------------------------------------------------------
    _cfl_factor = 1.0

    def check_courant_condition(self, vp) -> None:
        """acou_models.jl:5-17 / 240-251 (7/6 for the 4th-order staggered stencils); vp: the model or its maximum."""
        vel_max = float(np.max(vp))
        tmp = math.sqrt(sum(1.0 / float(s) ** 2 for s in self.spacing))
        courant = vel_max * float(self.dt) * tmp * self._cfl_factor
        log.info("Courant number: %s", courant)
        if self.runparams.erroronCFL:
            assert courant < 1, f"Courant condition not satisfied! [{courant}]"
        elif courant > 1:
            log.warning("CFL condition not satisfied! [%s]", courant)

    def check_numerics(self, shot, vel_min: float) -> None:
        """acou_models.jl:19-38 (points per wavelength, fmax = 2*domfreq)."""
        h_max = float(max(self.spacing))
        fmax = float(shot.srcs.domfreq) * 2.0
        ppw = vel_min / (fmax * h_max)
        min_ppw = self.runparams.minPPW
        log.info("Points per wavelength: %s", ppw)
        if ppw < min_ppw:
            dh0 = round(vel_min / (min_ppw * fmax), 2)
            msg = (f"Not enough points per wavelength (assuming fmax = 2*domfreq)! \n [{round(ppw, 1)} instead of >= {min_ppw}]\n"
                   f"  Grid spacing should be <= {dh0}")
            if self.runparams.erroronPPW:
                raise AssertionError(msg)
            log.warning(msg)

    def check_positions(self, positions: np.ndarray) -> None:
        """utils/checks.jl:55-94."""
        assert positions.shape[1] == self.N, "Positions matrix do not match the dimension of the model!"
        halo = self.cpmlparams.halo
        for s in range(positions.shape[0]):
            for c in range(self.N):
                assert 0 <= positions[s, c] <= self.extent[c], f"Position {positions[s, :]} is not inside the grid!"
                if not (c == self.N - 1 and self.cpmlparams.freeboundtop):
                    w = self.spacing[c] * halo
                    if not (w <= positions[s, c] <= self.extent[c] - w):
                        log.warning("Position %s is inside the CPML region!", positions[s, :])

    def check_sim_consistency(self, matprop, shots) -> None:
        raise NotImplementedError

    # ---- engine plumbing ------------------------------------------------------------------------------
    def _set_cpml(self, vel_max, domfreq) -> None:
        """init_bdc! -- coefficient profiles on the host, uploaded per axis."""
        cp = self.cpmlparams
        if cp.vel_max is not None:
            vel_max = cp.vel_max
        axes = hostprep.init_bdc(vel_max, self.dt, cp.halo, cp.rcoef, self.spacing, cp.freeboundtop, domfreq, self.T)
        self._cpml_host = axes
        for ax, (a, a_h, b, b_h) in enumerate(axes):
            _lib.check(self.lib.swb_sim_set_cpml(self._h, ax, _vp(a), _vp(a_h), _vp(b), _vp(b_h)))

    def cell_updates(self) -> int:
        return int(self.lib.swb_sim_cell_updates(self._h))

    def device_bytes(self) -> int:
        return int(self.lib.swb_sim_device_bytes(self._h))

    def total_gradient_ptr(self, which: int) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        _lib.check(self.lib.swb_sim_total_gradient_ptr(self._h, which, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def zero_total_gradient(self) -> None:
        _lib.check(self.lib.swb_sim_zero_total_gradient(self._h))

    def get_total_gradient(self) -> Dict[str, np.ndarray]:
        out = {}
        for k, name in enumerate(self.grad_names):
            g = np.zeros(self.gridsize, dtype=self.dtype, order="F")
            _lib.check(self.lib.swb_sim_get_total_gradient(self._h, k, _vp(g)))
            out[name] = g
        return out

    def get_raw_gradient(self, which: int, shape) -> np.ndarray:
        g = np.zeros(shape, dtype=self.dtype, order="F")
        _lib.check(self.lib.swb_sim_get_raw_gradient(self._h, which, _vp(g)))
        return g

    def get_field(self, name: str, shape) -> np.ndarray:
        g = np.zeros(shape, dtype=self.dtype, order="F")
        _lib.check(self.lib.swb_sim_get_field(self._h, name.encode(), _vp(g), g.nbytes))
        return g

    def dominant_kernel_name(self) -> str:
        return self._dominant_kernel

    _dominant_kernel = "step"

    def kernel_timing(self, enable: int) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self.lib.swb_sim_kernel_timing(self._h, enable, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def kernel_timing_class(self, cls: int) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self.lib.swb_sim_kernel_timing_class(self._h, cls, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    # ---- per-shot drivers -------------------------------------------------------------------------------
    def set_wavesim_matprop(self, matprop) -> None:
        raise NotImplementedError

    def init_shot(self, shot) -> None:
        raise NotImplementedError

    def _bind(self, shot) -> None:
        raise NotImplementedError

    def _snapshot_fields(self) -> Dict[str, List[tuple]]:
        return {}

    def swforward_1shot(self, shot):
        """swforward_1shot! (acou_forward.jl:22-62,83-125; ela_forward.jl:4-159)."""
        self._bind(shot)
        seis = shot.recs.seismograms
        snapevery = self.runparams.snapevery or 0
        _lib.check(self.lib.swb_sim_forward(self._h, _vp(seis), snapevery))
        snaps = None
        if snapevery:
            snaps = {}
            for it in range(snapevery, self.nt + 1, snapevery):
                snaps[it] = {}
                idx = 0
                for name, shapes in self._snapshot_fields().items():
                    comps = []
                    for shp in shapes:
                        a = np.zeros(shp, dtype=self.dtype, order="F")
                        _lib.check(self.lib.swb_sim_get_snapshot(self._h, it, idx, _vp(a)))
                        comps.append(a)
                        idx += 1
                    snaps[it][name] = comps[0] if len(comps) == 1 else comps
        return snaps

    def swgradient_1shot(self, shot, misfit) -> None:
        """swgradient_1shot! up to and including the raw correlation; the gradient of this shot is then
        post-processed and accumulated on the device by accumulate_gradient()."""
        self._bind(shot)
        seis = shot.recs.seismograms
        spec = misfit._device_spec(self.T) if isinstance(misfit, L2Misfit) and not getattr(self, "force_host_misfit", False) else None
        if spec is not None:  # residual, windows, diagonal covariance on the device: no D2H / H2D round trip between the sweeps
            obs, (w, mask) = _as_T(misfit.observed, self.T), spec
            l2 = _lib.swb_l2_spec(obs.ctypes.data, None if w is None else w.ctypes.data, None if mask is None else mask.ctypes.data)
            _lib.check(self.lib.swb_sim_gradient_l2_ex(self._h, C.byref(l2), _vp(seis), None))
        else:
            _lib.check(self.lib.swb_sim_gradient_forward(self._h, _vp(seis)))
            adjsrc = _as_T(-misfit.dchi_du(shot.recs), self.T)
            _lib.check(self.lib.swb_sim_gradient_adjoint(self._h, _vp(adjsrc)))
        gp = self.gradparams
        sp, rp = _as_T(shot.srcs.positions, self.T), _as_T(shot.recs.positions, self.T)
        _lib.check(self.lib.swb_sim_accumulate_gradient(self._h, sp.shape[0], _vp(sp), gp.mute_radius_src, rp.shape[0], _vp(rp), gp.mute_radius_rec))


class _AcousticBase(WaveSimulation):
    def check_sim_consistency(self, matprop, shots) -> None:
        ts, tr = type(shots[0].srcs), type(shots[0].recs)
        for s in shots:
            if type(s.srcs) is not ts or type(s.recs) is not tr:
                raise TypeError("Types of shots are inconsistent.")
        ok = isinstance(matprop, self._matprop_type) and issubclass(ts, ScalarSources) and issubclass(tr, ScalarReceivers) \
            and matprop.vp.ndim == self.N and matprop.vp.dtype == self.dtype
        if not ok:
            raise TypeError(f"Types of WaveSimulation, MaterialProperties and Sources/Receivers are inconsistent \n {type(self)}, \n {type(matprop)}, \n {ts}, {tr}")

    def init_shot(self, shot) -> None:
        """init_shot! (shots/shot.jl:46-51): check_shot + init_bdc!."""
        self.check_numerics(shot, self._vp_min)
        self.check_positions(shot.srcs.positions)
        self.check_positions(shot.recs.positions)
        self._set_cpml(self._vp_max, shot.srcs.domfreq)

    def _bind_scalar(self, shot, scal_srctf, possrcs, posrecs) -> None:
        self._possrcs, self._posrecs, self._srctf = possrcs, posrecs, scal_srctf
        assert shot.recs.seismograms.shape == (self.nt, posrecs.shape[0]) and shot.recs.seismograms.dtype == self.dtype
        assert scal_srctf.shape[0] == self.nt, "source time function length must equal the number of timesteps"
        _lib.check(self.lib.swb_sim_bind_scalar_shot(self._h, possrcs.shape[0], _vp(possrcs), _vp(scal_srctf), posrecs.shape[0], _vp(posrecs)))


class AcousticCDCPMLWaveSimulation(_AcousticBase):
    """src/models/acoustic/acou_models.jl:68-226."""

    kind = _lib.SWB_ACOU_CD
    grad_names = ("vp",)
    _matprop_type = VpAcousticCDMaterialProperties

    def set_wavesim_matprop(self, matprop) -> None:
        vp = matprop.vp
        assert vp.ndim == self.N, "Material property dimensionality must be the same as the wavesim!"
        assert vp.shape == self.gridsize, f"Material property number of grid points must be the same as the wavesim! \n {vp.shape}, {self.gridsize}"
        vmin, vmax = np.min(vp), np.max(vp)  # one pass each; every check below derives from them
        assert vmin > 0, "Pressure velocity material property must be positive!"
        self.check_courant_condition(vmax)
        self.matprop = VpAcousticCDMaterialProperties(vp.copy(order="F"))
        self._vp_min, self._vp_max = float(vmin), self.T(vmax)  # reductions over the model once per update, not per shot
        arr = (C.c_void_p * 1)(self.matprop.vp.ctypes.data)
        _lib.check(self.lib.swb_sim_set_material(self._h, 1, arr, 0))

    def _bind(self, shot) -> None:
        T = self.T
        possrcs = hostprep.find_nearest_grid_points(shot.srcs.positions, self.spacing, T)
        posrecs = hostprep.find_nearest_grid_points(shot.recs.positions, self.spacing, T)
        tf = hostprep.scale_stf_acoustic_cd(shot.srcs.tf, self.spacing, self.dt, self.matprop.vp, possrcs, T)
        self._bind_scalar(shot, tf, possrcs, posrecs)

    def _snapshot_fields(self):
        return {"pcur": [self.gridsize]}


class AcousticVDStaggeredCPMLWaveSimulation(_AcousticBase):
    """src/models/acoustic/acou_models.jl:311-468."""

    kind = _lib.SWB_ACOU_VD
    grad_names = ("vp", "rho")
    _matprop_type = VpRhoAcousticVDMaterialProperties
    _cfl_factor = 7.0 / 6.0
    _dominant_kernel = "vd_fused_kernel (v update + p update + inject + record [+ correlations], one launch per time step)"

    def set_wavesim_matprop(self, matprop) -> None:
        vp, rho = matprop.vp, matprop.rho
        assert vp.ndim == rho.ndim == self.N, "Material property dimensionality must be the same as the wavesim!"
        assert vp.shape == rho.shape == self.gridsize, f"Material property number of grid points must be the same as the wavesim! \n {vp.shape}, {rho.shape}, {self.gridsize}"
        vmin, vmax = np.min(vp), np.max(vp)  # one pass each; every check below derives from them
        assert vmin > 0, "Pressure velocity material property must be positive!"
        assert np.min(rho) > 0, "Density material property must be positive!"
        self.check_courant_condition(vmax)
        self.matprop = VpRhoAcousticVDMaterialProperties(vp.copy(order="F"), rho.copy(order="F"), interp_method=matprop.interp_method)
        self._vp_min, self._vp_max = float(vmin), self.T(vmax)
        arr = (C.c_void_p * 2)(self.matprop.vp.ctypes.data, self.matprop.rho.ctypes.data)
        _lib.check(self.lib.swb_sim_set_material(self._h, 2, arr, 0 if matprop.interp_method == "arithmetic" else 1))

    def _bind(self, shot) -> None:
        T = self.T
        possrcs = hostprep.find_nearest_grid_points(shot.srcs.positions, self.spacing, T)
        posrecs = hostprep.find_nearest_grid_points(shot.recs.positions, self.spacing, T)
        tf = hostprep.scale_stf_acoustic_vd(shot.srcs.tf, self.spacing, self.dt, self.matprop.vp, self.matprop.rho, possrcs, T)
        self._bind_scalar(shot, tf, possrcs, posrecs)

    def _snapshot_fields(self):
        n = self.gridsize
        return {"pcur": [n], "vcur": [tuple(n[j] - (1 if i == j else 0) for j in range(self.N)) for i in range(self.N)]}


# ---------------------------------------------------------------------------------------------------------
# build_wavesim / swforward! / swmisfit! / swgradient!
# ---------------------------------------------------------------------------------------------------------


def build_wavesim(params, matprop, runparams: Optional[RunParameters] = None, gradparams: Optional[GradParameters] = None, gradient: bool = False, **kw):
    """build_wavesim (src/apis/build.jl:18-88)."""
    runparams = runparams or RunParameters()
    if gradparams is None and gradient:
        gradparams = GradParameters()
    elif gradparams is not None:
        assert gradient, "build_wavesim(...) specifies GradParameters, however, gradient keyword argument is set to false."
    bc = params.boundcond
    if isinstance(params, InputParametersAcoustic) and isinstance(matprop, VpAcousticCDMaterialProperties):
        cls = AcousticCDCPMLWaveSimulation
    elif isinstance(params, InputParametersAcoustic) and isinstance(matprop, VpRhoAcousticVDMaterialProperties):
        cls = AcousticVDStaggeredCPMLWaveSimulation
    elif isinstance(params, InputParametersElastic) and isinstance(matprop, ElasticIsoMaterialProperties):
        from .elastic import ElasticIsoCPMLWaveSimulation

        cls = ElasticIsoCPMLWaveSimulation
    else:
        raise TypeError(f"no WaveSimulation for ({type(params).__name__}, {type(matprop).__name__})")
    return cls(params, matprop, bc, runparams, gradparams, gradient=gradient, **kw)  # kw: sincinterp (elastic, ela_models.jl:208)


def _run_swforward(wavesim: WaveSimulation, matprop, shots):
    """run_swforward! (src/apis/forward.jl:71-115)."""
    log.info(">=====  Forward simulation  ======<")
    wavesim.check_sim_consistency(matprop, shots)
    wavesim.set_wavesim_matprop(matprop)
    takesnapshots = wavesim.runparams.snapevery is not None
    snapshots_per_shot = []
    for s, shot in enumerate(shots):
        log.info("-- Shot #%d --", s + 1)
        wavesim.init_shot(shot)
        snaps = wavesim.swforward_1shot(shot)
        if takesnapshots:
            snapshots_per_shot.append(snaps)
    return snapshots_per_shot if takesnapshots else None


def swforward(params_or_wavesim, matprop, shots, runparams: Optional[RunParameters] = None):
    """swforward! (src/apis/forward.jl:24-66): seismograms are stored in each shot's receivers; returns the
    per-shot snapshots if runparams.snapevery is set, else None."""
    if isinstance(params_or_wavesim, WaveSimulation):
        return _run_swforward(params_or_wavesim, matprop, shots)
    assert runparams is not None, "runparams is a required keyword argument of swforward!"
    wavesim = build_wavesim(params_or_wavesim, matprop, runparams=runparams, gradient=False)
    try:
        return _run_swforward(wavesim, matprop, shots)
    finally:
        wavesim.close()


def swmisfit(params_or_wavesim, matprop, shots, misfit, runparams: Optional[RunParameters] = None, reference_compat: bool = True):
    """swmisfit! (src/apis/misfit.jl:26-104).  The reference's loop `for s in length(shots)` evaluates the
    misfit of shots[1] only (misfit.jl:97-100); reference_compat=True reproduces that, False sums over all shots."""
    if isinstance(params_or_wavesim, WaveSimulation):
        _run_swforward(params_or_wavesim, matprop, shots)
    else:
        assert runparams is not None
        wavesim = build_wavesim(params_or_wavesim, matprop, runparams=runparams, gradient=False)
        try:
            _run_swforward(wavesim, matprop, shots)
        finally:
            wavesim.close()
    log.info("Computing misfit")
    if reference_compat:
        return misfit[0].calcmisfit(shots[0].recs)
    return sum(m.calcmisfit(s.recs) for s, m in zip(shots, misfit))


def _run_swgradient(wavesim: WaveSimulation, matprop, shots, misfit, shot_indices: Optional[Sequence[int]] = None):
    """run_swgradient! (src/apis/gradient.jl:93-136); shot_indices restricts the loop to this rank's shots."""
    log.info(">=====  Gradient computation  ======<")
    assert wavesim.gradient, "the WaveSimulation was not built with gradient=true"
    wavesim.check_sim_consistency(matprop, shots)
    wavesim.set_wavesim_matprop(matprop)
    wavesim.zero_total_gradient()
    totmisfit = 0.0
    idx = range(len(shots)) if shot_indices is None else shot_indices
    for s in idx:
        log.info("Shot #%d", s + 1)
        wavesim.init_shot(shots[s])
        wavesim.swgradient_1shot(shots[s], misfit[s])
        if wavesim.gradparams.compute_misfit:
            totmisfit += misfit[s].calcmisfit(shots[s].recs)
    return totmisfit


def swgradient(params_or_wavesim, matprop, shots, misfit, runparams: Optional[RunParameters] = None, gradparams: Optional[GradParameters] = None):
    """swgradient! (src/apis/gradient.jl:31-88): Dict of gradients (one key per material property), plus the
    misfit value when gradparams.compute_misfit is set."""
    own = not isinstance(params_or_wavesim, WaveSimulation)
    wavesim = build_wavesim(params_or_wavesim, matprop, runparams=runparams or RunParameters(), gradparams=gradparams or GradParameters(),
                            gradient=True) if own else params_or_wavesim
    try:
        totmisfit = _run_swgradient(wavesim, matprop, shots, misfit)
        grad = wavesim.get_total_gradient()
        return (grad, wavesim.T(totmisfit)) if wavesim.gradparams.compute_misfit else grad
    finally:
        if own:
            wavesim.close()
