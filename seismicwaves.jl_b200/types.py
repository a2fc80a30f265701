"""Host-side mirror of the reference's input types (same names, same fields, same checks).

Reference: src/models/acoustic/acou_params.jl, src/models/elastic/ela_params.jl, src/models/bdc_params.jl,
src/models/genparameters.jl, src/models/acoustic/acou_material_properties.jl,
src/models/elastic/ela_material_properties.jl, src/shots/{sources,receivers,shot}.jl,
src/inversion/misfits/L2Misfit.jl.  Arrays are numpy, column-major semantics: index [i, j] means the
same as Julia's [i+1, j+1]; internally everything is converted to Fortran order.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np


def _f(a, dtype=None) -> np.ndarray:
    a = np.asarray(a) if dtype is None else np.asarray(a, dtype=dtype)
    return np.asfortranarray(a)


# ---- boundary conditions / parameters ---------------------------------------------------------


@dataclass
class CPMLBoundaryConditionParameters:
    """src/models/bdc_params.jl:16-27."""

    halo: int = 20
    rcoef: float = 0.0001
    freeboundtop: bool = True
    vel_max: Optional[float] = None


class _InputParameters:
    def __init__(self, ntimesteps: int, dt, gridsize: Sequence[int], gridspacing: Sequence, boundcond: CPMLBoundaryConditionParameters,
                 dtype=None):
        gridsize = tuple(int(g) for g in gridsize)
        N = len(gridsize)
        assert N <= 3, "Dimensionality must be less than or equal to 3!"
        assert len(gridspacing) == N
        assert all(g > 0 for g in gridsize), "All numbers of grid points must be positive!"
        assert all(s > 0 for s in gridspacing), "All grid spacings must be positive!"
        assert ntimesteps > 0, "Number of timesteps must be positive!"
        assert dt > 0, "Timestep size must be positive!"
        if dtype is None:  # the reference's T is the type of dt
            dtype = np.float32 if isinstance(dt, np.float32) else np.float64
        self.dtype = np.dtype(dtype)
        T = self.dtype.type
        self.ntimesteps = int(ntimesteps)
        self.dt = T(dt)
        self.gridsize = gridsize
        self.gridspacing = tuple(T(s) for s in gridspacing)
        self.boundcond = boundcond

    @property
    def ndim(self) -> int:
        return len(self.gridsize)


class InputParametersAcoustic(_InputParameters):
    """src/models/acoustic/acou_params.jl:9-35."""


class InputParametersElastic(_InputParameters):
    """src/models/elastic/ela_params.jl."""


@dataclass
class RunParameters:
    """src/models/genparameters.jl:25-46.  `parall` selects the backend; this package provides :B200 only."""

    parall: str = "B200"
    snapevery: Optional[int] = None
    infoevery: Optional[int] = None
    logger: Optional[object] = None
    erroronCFL: bool = True
    minPPW: int = 10
    erroronPPW: bool = True
    # B200-backend knobs (no reference counterpart)
    device: int = 0
    fast_f32: bool = False   # pure-Float32 arithmetic instead of the reference's Float64 intermediates
    fused: bool = True       # fused engine kernels (False: one launch per reference kernel)
    graphs: bool = True      # replay the per-shot time loops as CUDA graphs (False: eager launches, SWB_FLAG_NO_GRAPH)

    def __post_init__(self):
        assert self.minPPW >= 0


@dataclass
class GradParameters:
    """src/models/genparameters.jl:63-78."""

    mute_radius_src: int = 0
    mute_radius_rec: int = 0
    compute_misfit: bool = False
    check_freq: int = 1

    def __post_init__(self):
        assert self.check_freq > 0


# ---- material properties --------------------------------------------------------------------------


class VpAcousticCDMaterialProperties:
    """src/models/acoustic/acou_material_properties.jl (constant density)."""

    def __init__(self, vp):
        self.vp = _f(vp)


class VpRhoAcousticVDMaterialProperties:
    """src/models/acoustic/acou_material_properties.jl (variable density)."""

    def __init__(self, vp, rho, interp_method: str = "arithmetic"):
        assert interp_method in ("arithmetic", "harmonic")
        self.vp = _f(vp)
        self.rho = _f(rho)
        self.interp_method = interp_method


class ElasticIsoMaterialProperties:
    """src/models/elastic/ela_material_properties.jl:1-24 (keyword names ρ, λ, μ in the reference)."""

    def __init__(self, rho, lam, mu, interp_method_rho: str = "arithmetic", interp_method_lam: str = "arithmetic", interp_method_mu: str = "arithmetic"):
        self.rho = _f(rho)
        self.lam = _f(lam)
        self.mu = _f(mu)
        self.interp_method_rho = interp_method_rho
        self.interp_method_lam = interp_method_lam
        self.interp_method_mu = interp_method_mu


# ---- sources / receivers / shots -------------------------------------------------------------------


class ScalarSources:
    """src/shots/sources.jl:8-26."""

    def __init__(self, positions, tf, domfreq):
        self.positions = _f(positions)
        self.tf = _f(tf)
        assert self.positions.shape[0] > 0, "There must be at least one source!"
        assert self.positions.shape[0] == self.tf.shape[1], "Number of sources do not match between positions and time-functions!"
        self.domfreq = domfreq


@dataclass
class MomentTensor2D:
    """src/shots/sources.jl:33-37."""

    Mxx: float
    Mzz: float
    Mxz: float


class MomentTensorSources:
    """src/shots/sources.jl:54-76."""

    def __init__(self, positions, tf, momtens: List[MomentTensor2D], domfreq):
        self.positions = _f(positions)
        self.tf = _f(tf)
        assert self.positions.shape[0] > 0, "There must be at least one source!"
        assert self.positions.shape[0] == self.tf.shape[1], "Number of sources do not match between positions and time-functions!"
        assert len(momtens) == self.positions.shape[0], "Number of moment tensors must match number of sources!"
        self.momtens = list(momtens)
        self.domfreq = domfreq


class ExternalForceSources:
    """src/shots/sources.jl:78-90; tf is (nt, N, nsrc)."""

    def __init__(self, positions, tf, domfreq):
        self.positions = _f(positions)
        self.tf = _f(tf)
        assert self.positions.shape[0] > 0, "There must be at least one source!"
        assert self.positions.shape[0] == self.tf.shape[2], "Number of sources do not match between positions and time-functions!"
        assert self.tf.shape[1] == self.positions.shape[1], "Number of components do not match between time-functions and positions!"
        self.domfreq = domfreq


class ScalarReceivers:
    """src/shots/receivers.jl:8-53; seismograms (nt, nrec)."""

    def __init__(self, positions, nt: int, dtype=None):
        self.positions = _f(positions)
        assert self.positions.shape[0] > 0, "There must be at least one receiver!"
        self.seismograms = np.zeros((nt, self.positions.shape[0]), dtype=dtype or self.positions.dtype, order="F")


class VectorReceivers:
    """src/shots/receivers.jl:64-110; seismograms (nt, ndim, nrec)."""

    def __init__(self, positions, nt: int, ndim: int = 2, dtype=None):
        self.positions = _f(positions)
        assert self.positions.shape[0] > 0, "There must be at least one receiver!"
        self.seismograms = np.zeros((nt, ndim, self.positions.shape[0]), dtype=dtype or self.positions.dtype, order="F")


@dataclass
class ScalarShot:
    """src/shots/shot.jl:9-16."""

    srcs: ScalarSources
    recs: ScalarReceivers


@dataclass
class MomentTensorShot:
    """src/shots/shot.jl:23-30."""

    srcs: MomentTensorSources
    recs: VectorReceivers


@dataclass
class ExternalForceShot:
    """src/shots/shot.jl:37-44."""

    srcs: ExternalForceSources
    recs: VectorReceivers


# ---- misfit ------------------------------------------------------------------------------------------


class L2Misfit:
    """src/inversion/misfits/L2Misfit.jl:2-95.  invcov: None (identity), 1-D (Diagonal) or dense (nt, nt)."""

    def __init__(self, observed, invcov=None, windows: Optional[List[Tuple[int, int]]] = None):
        observed = _f(observed)
        if observed.ndim not in (2, 3):
            raise ValueError("Observed data must be a 2D or 3D array!")
        nt = observed.shape[0]
        if invcov is not None:
            invcov = np.asarray(invcov)
            if invcov.ndim == 2:
                assert invcov.shape[0] == invcov.shape[1], "Inverse covariance matrix must be square!"
            assert invcov.shape[0] == nt, "Size of inverse covariance matrix must match the number of timesteps!"
        windows = list(windows or [])
        assert all(1 <= a <= b <= nt for (a, b) in windows), "Windows indices must be between 1 and maximum number of timesteps!"
        self.observed = observed.copy(order="F")
        self.invcov = invcov
        self.windows = windows

    def _is_plain(self) -> bool:
        """identity covariance, no windows: eligible for the device-side misfit path"""
        if self.windows:
            return False
        if self.invcov is None:
            return True
        ic = self.invcov
        if ic.ndim == 1:
            return bool(np.all(ic == 1))
        return bool(np.array_equal(ic, np.eye(ic.shape[0], dtype=ic.dtype)))

    def _device_spec(self, T):
        """(invcov_diag or None, mask or None) if the misfit can be evaluated by the engine (identity / diagonal inverse covariance,
        any windows: swb_sim_gradient_l2_ex), else None (dense covariance -> host path)."""
        ic = self.invcov
        w = None
        if ic is not None:
            if ic.ndim == 1:
                w = ic
            elif np.array_equal(ic, np.diag(np.diag(ic))):
                w = np.diag(ic)
            else:
                return None
            w = None if bool(np.all(w == 1)) else np.ascontiguousarray(w, dtype=T)
        mask = None
        if self.windows:
            mask = np.zeros(self.observed.shape[0], dtype=T)
            for (a, b) in self.windows:
                mask[a - 1:b] = 1
        return w, mask

    def _residuals(self, seis: np.ndarray) -> np.ndarray:
        res = (seis - self.observed).astype(seis.dtype)
        if self.windows:
            mask = np.zeros(res.shape[0], dtype=seis.dtype)
            for (a, b) in self.windows:
                mask[a - 1:b] = 1
            res = (mask.reshape((-1,) + (1,) * (res.ndim - 1)) * res).astype(seis.dtype)
        return res

    def _invcov_mul(self, r2d: np.ndarray) -> np.ndarray:
        if self.invcov is None:
            return r2d
        if self.invcov.ndim == 1:
            return (self.invcov.reshape(-1, 1) * r2d).astype(r2d.dtype)
        return (self.invcov @ r2d).astype(r2d.dtype)

    def calcmisfit(self, recs) -> float:
        """calcmisfit (L2Misfit.jl:24-60): dot(res, invcov, res)/2."""
        res = self._residuals(recs.seismograms)
        if res.ndim == 2:
            return float(np.sum(res * self._invcov_mul(res), dtype=np.float64) / 2)
        tot = 0.0
        for i in range(res.shape[1]):
            r = res[:, i, :]
            tot += float(np.sum(r * self._invcov_mul(r), dtype=np.float64))
        return tot / 2

    def dchi_du(self, recs) -> np.ndarray:
        """∂χ_∂u (L2Misfit.jl:64-95)."""
        res = self._residuals(recs.seismograms)
        if res.ndim == 2:
            return np.asfortranarray(self._invcov_mul(res))
        out = np.empty_like(res, order="F")
        for d in range(res.shape[1]):
            out[:, d, :] = self._invcov_mul(res[:, d, :])
        return out
