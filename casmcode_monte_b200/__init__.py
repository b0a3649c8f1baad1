"""B200-native (sm_100a) implementation of libcasm-monte's Ising semi-grand-
canonical Metropolis hot path, behind a C ABI (include/casm_monte_gpu.h).

Layers (all thin, the product is the CUDA library):
  csrc/cmg_device.cuh, csrc/cmg_capi.cu   hand-written kernels + the C ABI
  _capi.py, lattice.py                    ctypes binding / numpy wrapper
  monte/                                  host-side mirror of the reference's
                                          libcasm.monte API for this path
"""
from ._capi import KB, LIB_PATH, CmgError  # noqa: F401
from .lattice import (  # noqa: F401
    MODE_CHECKERBOARD,
    MODE_SERIAL_REFERENCE,
    Q_FORMATION_ENERGY,
    Q_PARAM_COMPOSITION,
    Q_POTENTIAL_ENERGY,
    IsingLatticeGPU,
)

__all__ = [
    "IsingLatticeGPU",
    "CmgError",
    "KB",
    "MODE_CHECKERBOARD",
    "MODE_SERIAL_REFERENCE",
    "Q_PARAM_COMPOSITION",
    "Q_FORMATION_ENERGY",
    "Q_POTENTIAL_ENERGY",
]
