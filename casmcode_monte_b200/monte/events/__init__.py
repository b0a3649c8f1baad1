"""Mirror of the ``libcasm.monte.events`` names on this path."""
from .._ext import ext as _ext

Conversions = _ext.Conversions
IntVector = _ext.IntVector
LongVector = _ext.LongVector
OccEvent = _ext.OccEvent

__all__ = ["Conversions", "IntVector", "LongVector", "OccEvent"]
