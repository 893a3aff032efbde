"""eigenkernel_b200 -- B200-native (sm_100a) dense FP64 symmetric eigensolve behind EigenKernel's
`-s <solver>` boundary (reference src/solver_main.f90:52-99).  Hand-written CUDA in csrc/, flat C-ABI in
include/ekb200.h, host-side mirror of the reference's solver interface in solver.py."""
from ._lib import Ekb200Error, LIB_PATH, exported_symbols, load  # noqa: F401

__all__ = ["Ekb200Error", "LIB_PATH", "exported_symbols", "load"]
