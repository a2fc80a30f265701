"""seismicwaves.jl_b200 -- B200-native (sm_100a) finite-difference time-stepping engine behind the
SeismicWaves.jl API, plus the host-side mirror of the reference interface used to drive it from Python.

The directory name contains a dot, so import it through the repo-root shim:  `import swb200`.
"""
from . import _lib, hostprep  # noqa: F401
from ._lib import SwbError, device_count  # noqa: F401
from .api import (AcousticCDCPMLWaveSimulation, AcousticVDStaggeredCPMLWaveSimulation, WaveSimulation, build_wavesim, swforward, swgradient,  # noqa: F401
                  swmisfit)
from .elastic import ElasticIsoCPMLWaveSimulation  # noqa: F401
from .hostprep import distribsrcs, gaussderivstf, gaussstf, rickerstf  # noqa: F401
from .types import (CPMLBoundaryConditionParameters, ElasticIsoMaterialProperties, ExternalForceShot, ExternalForceSources, GradParameters,  # noqa: F401
                    InputParametersAcoustic, InputParametersElastic, L2Misfit, MomentTensor2D, MomentTensorShot, MomentTensorSources, RunParameters,
                    ScalarReceivers, ScalarShot, ScalarSources, VectorReceivers, VpAcousticCDMaterialProperties, VpRhoAcousticVDMaterialProperties)

# Julia-style aliases
swforward_ = swforward
swmisfit_ = swmisfit
swgradient_ = swgradient
