"""pgopt_b200 — B200-native (sm_100a) particle-Gibbs engine for PGopt's data-parallel hot path.

Host-side mirror of the reference's Julia API (Julia/src/PGopt.jl:20 exports) over the C ABI of
libpgopt_b200.so.  The Julia `ccall` shim with the same signatures is julia/PGoptB200.jl.
"""
from .api import (PG_sample, PGSampleBatch, DeviceSamples, MultiGPUEngine, pinned_empty, KnownBasisVB, HilbertGPBasis, LinearMeasurement, MvNormal, ParticleGibbsEngine,
                  particle_Gibbs, systematic_resampling, MNIW_sample, scenario_rollout, test_prediction, scenario_constraint_max,
                  hilbert_gp_prior_V)
from ._lib import PgoptError

__all__ = ["PG_sample", "PGSampleBatch", "DeviceSamples", "MultiGPUEngine", "pinned_empty", "KnownBasisVB", "HilbertGPBasis", "LinearMeasurement", "MvNormal", "ParticleGibbsEngine",
           "particle_Gibbs", "systematic_resampling", "MNIW_sample", "scenario_rollout", "test_prediction", "scenario_constraint_max", "hilbert_gp_prior_V",
           "PgoptError"]
