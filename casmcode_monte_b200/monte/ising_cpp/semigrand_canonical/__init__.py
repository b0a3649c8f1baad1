"""Mirror of ``libcasm.monte.ising_cpp.semigrand_canonical``
(python/libcasm/monte/ising_cpp/semigrand_canonical/__init__.py:3-10)."""
from ..._ext import ext as _ext

SemiGrandCanonicalCalculator = _ext.SemiGrandCanonicalCalculator
SemiGrandCanonicalConditions = _ext.SemiGrandCanonicalConditions
SemiGrandCanonicalData = _ext.SemiGrandCanonicalData
SemiGrandCanonicalEventGenerator = _ext.SemiGrandCanonicalEventGenerator
SemiGrandCanonicalPotential = _ext.SemiGrandCanonicalPotential
default_write_status = _ext.default_write_status

__all__ = [
    "SemiGrandCanonicalCalculator",
    "SemiGrandCanonicalConditions",
    "SemiGrandCanonicalData",
    "SemiGrandCanonicalEventGenerator",
    "SemiGrandCanonicalPotential",
    "default_write_status",
]
