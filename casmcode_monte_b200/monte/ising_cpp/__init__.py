"""Mirror of ``libcasm.monte.ising_cpp`` (python/libcasm/monte/ising_cpp/__init__.py:3-9)."""
from .._ext import ext as _ext

IsingConfiguration = _ext.IsingConfiguration
IsingFormationEnergy = _ext.IsingFormationEnergy
IsingParamComposition = _ext.IsingParamComposition
IsingState = _ext.IsingState
IsingSystem = _ext.IsingSystem

__all__ = ["IsingConfiguration", "IsingFormationEnergy", "IsingParamComposition", "IsingState", "IsingSystem"]
