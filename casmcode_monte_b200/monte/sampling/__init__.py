"""Mirror of the ``libcasm.monte.sampling`` names used by the Ising SGC path
(python/libcasm/monte/sampling/__init__.py:4-60)."""
from .._ext import ext as _ext
from ._requested_precision_constructor import RequestedPrecisionConstructor, converge

for _n in (
    "BasicStatistics BasicStatisticsCalculator CompletionCheck CompletionCheckParams CompletionCheckResults "
    "ConvergenceCheckResults CutoffCheckParams EquilibrationCheckResults IndividualConvergenceResult "
    "IndividualEquilibrationResult RequestedPrecision RequestedPrecisionMap Sampler SamplerComponent SamplerMap "
    "StateSamplingFunction StateSamplingFunctionMap jsonSampler jsonSamplerMap jsonStateSamplingFunction "
    "jsonStateSamplingFunctionMap all_minimums_met any_maximum_met colmajor_component_names "
    "default_component_names default_equilibration_check get_n_samples matrix_as_vector scalar_as_vector "
    "vector_as_vector SAMPLE_MODE SAMPLE_METHOD SamplingParams"
).split():
    globals()[_n] = getattr(_ext, _n)
del _n
