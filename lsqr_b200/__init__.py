"""lsqr_b200 -- B200-native LSQR engine behind the API of jacobwilliams/LSQR.

The package is a thin host layer over ``lib/liblsqr_b200.so`` (hand-written sm_100a CUDA, C ABI in
``include/lsqr_b200.h``).  Importing it loads the shared library and raises ImportError if the
library has not been built: there is no CPU path.
"""
from . import _lib
from ._lib import LsqrError, LIB_PATH

_lib.load()   # fail loudly when the CUDA extension is missing

from .solver import (LsqrSolverEz, LsqrSolver, LsqrSolverHost, EzAsOperator, SolveResult,   # noqa: E402
                     dnrm2, ddot, dscal, dcopy)


def device_count() -> int:
    return _lib.load().lsqr_b200_device_count()


def version() -> int:
    return _lib.load().lsqr_b200_version()


__all__ = ["LsqrSolverEz", "LsqrSolver", "LsqrSolverHost", "EzAsOperator", "SolveResult", "LsqrError", "LIB_PATH",
           "dnrm2", "ddot", "dscal", "dcopy", "device_count", "version"]
