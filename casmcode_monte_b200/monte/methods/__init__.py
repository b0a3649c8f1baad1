"""Mirror of ``libcasm.monte.methods`` for this path: the acceptance rule
(include/casm/monte/methods/metropolis.hh:26-35)."""
from .._ext import ext as _ext

metropolis_acceptance = _ext.metropolis_acceptance

__all__ = ["metropolis_acceptance"]
