"""Mirror of ``libcasm.monte.methods`` for this path
(python/libcasm/monte/methods/__init__.py:3-9): the acceptance rule
(include/casm/monte/methods/metropolis.hh:26-35) and the generic
callback-driven loop with its data structure
(include/casm/monte/methods/basic_occupation_metropolis.hh)."""
from .._ext import ext as _ext

BasicOccupationMetropolisData = _ext.SemiGrandCanonicalData
basic_occupation_metropolis = _ext.basic_occupation_metropolis
metropolis_acceptance = _ext.metropolis_acceptance

__all__ = ["BasicOccupationMetropolisData", "basic_occupation_metropolis", "metropolis_acceptance"]
