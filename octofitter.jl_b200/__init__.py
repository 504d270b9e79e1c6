"""octofitter.jl_b200 — B200-native (sm_100a) hot path of Octofitter.jl behind a C ABI.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C-ABI shim, built into
`lib/libocto_b200.so`) and the host-side mirror of the reference interface (`model.py`).
The directory name is not a Python identifier; import it through the repo-root loader
`octofitter_jl_b200` (see octofitter_jl_b200.py).
"""
from ._abi import (OctoConstants, OctoLayout, OctoObsBlock, OctoPrior, OctoInputDef, default_constants, load_library, pack,
                   EXPORTED_SYMBOLS, LIB_PATH,
                   KIND_ASTROM_RADEC, KIND_ASTROM_PASEP, KIND_RV_STAR_ABS, KIND_RV_STAR_MARGIN, KIND_RV_PLANET_REL,
                   KIND_HGCA_INSTANT)
from .model import (Normal, Uniform, LogUniform, Sine, truncated, UniformCircular, θ_at_epoch_to_tperi,
                    theta_at_epoch_to_tperi, Table, PlanetRelAstromObs, PlanetRelAstromLikelihood, ObsPriorAstromONeil2019, HGCAInstantaneousObs, StarAbsoluteRVObs,
                    StarAbsoluteRVLikelihood, MarginalizedStarAbsoluteRVObs, MarginalizedStarAbsoluteRVLikelihood,
                    PlanetRelativeRVObs, PlanetRelativeRVLikelihood, Planet, System, ModelSpec, LogDensityModel, OctoError)
from .pt import ParallelTempering
from .samplers import batched_hmc, batched_parallel_tempering, batched_slice_sampler, batched_slice_parallel_tempering, device_hmc, device_parallel_tempering, device_parallel_tempering_dist, diagonal_metric, hmc_random, octofit, octofit_rejection

__all__ = [n for n in dir() if not n.startswith("_")]
