"""gen_b200: B200-native sequential Monte Carlo and importance sampling behind Gen.jl's
src/inference API. The compute path is libgensmc.so (hand-written sm_100a CUDA, C ABI in
include/gen_b200.h); this package is the Python mirror of the Julia host interface."""
from ._lib import GsmcError, load
from .choicemap import ChoiceMap, NoChange, UnknownChange, choicemap, merge
from .inference import (DeviceTrace, DeviceTraces, ParticleFilterState, get_log_weights, get_traces,
                        importance_resampling, importance_sampling, initialize_particle_filter, log_ml_estimate, maybe_resample_,
                        maybe_resample_b, particle_filter_step_, particle_filter_step_b, sample_unweighted_traces)
from .pmmh import ParticleFilterCombinator, PFCombinatorTrace, pmmh
from .models import (HMM, BearingsOnly, DeviceModel, DeviceProposal, LinearGaussianSSM, LinearRegression, NormalNormal,
                     OutlierRegression, StochasticVolatility, UniformNormal)

__all__ = [n for n in dir() if not n.startswith("_")]
