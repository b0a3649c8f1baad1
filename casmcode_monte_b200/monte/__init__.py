"""Drop-in for the part of ``libcasm.monte`` the Ising SGC path uses
(python/libcasm/monte/__init__.py:3-8 of the reference)."""
from ._ext import ext as _ext

KB = _ext.KB
MethodLog = _ext.MethodLog
RandomNumberEngine = _ext.RandomNumberEngine
RandomNumberGenerator = _ext.RandomNumberGenerator
ValueMap = _ext.ValueMap

__all__ = ["KB", "MethodLog", "RandomNumberEngine", "RandomNumberGenerator", "ValueMap"]
