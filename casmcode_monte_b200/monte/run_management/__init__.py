"""Run management for the device-backed Ising SGC path: sampling fixtures, the
run manager and the occupation-Metropolis driver built on them.

The reference keeps these in C++ only (include/casm/monte/run_management/*.hh,
include/casm/monte/methods/occupation_metropolis.hh; libcasm-clexmonte binds
them for its own calculators).  Names, members and meaning follow those headers;
``SamplingParams`` / ``SAMPLE_MODE`` / ``SAMPLE_METHOD`` are also exported from
``monte.sampling`` as in python/src/monte_sampling.cpp:414-742.
"""
from .._ext import ext as _ext
from ._json_results_io import jsonResultsIO

for _n in (
    "SAMPLE_MODE SAMPLE_METHOD SamplingParams MonteCounter Results ResultsAnalysisFunction "
    "ResultsAnalysisFunctionMap SamplingFixtureParams SamplingFixture RunManager occupation_metropolis "
    "make_heat_capacity_f make_susceptibility_f sample_at stochastic_count_step stochastic_time_step"
).split():
    globals()[_n] = getattr(_ext, _n)
del _n

__all__ = [
    "SAMPLE_MODE", "SAMPLE_METHOD", "SamplingParams", "MonteCounter", "Results", "ResultsAnalysisFunction",
    "ResultsAnalysisFunctionMap", "SamplingFixtureParams", "SamplingFixture", "RunManager",
    "occupation_metropolis", "make_heat_capacity_f", "make_susceptibility_f", "sample_at",
    "stochastic_count_step", "stochastic_time_step", "jsonResultsIO",
]
