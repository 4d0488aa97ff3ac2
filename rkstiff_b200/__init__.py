"""rkstiff_b200: B200-native engine for rkstiff's diagonal ETD/IF Runge-Kutta stepping path.

Drop-in solver classes with the reference API (IF34, ETD34, ETD35, IF45DP, ETD4, ETD5, IF4 taking
``(lin_op, nl_func, SolverConfig)``, ``evolve``/``step``, ``solver.t``/``solver.u``) whose numeric work
runs in hand-written sm_100a CUDA kernels behind a C ABI (include/rkstiff_b200.h).  There is no CPU
fallback: importing the package loads ``librkstiff_b200.so`` and fails loudly if it is missing.
"""
from . import _abi  # noqa: F401  (loads the CUDA library or raises)
from . import (derivatives, etd, etd4, etd5, etd34, etd35, grids, if4, if34, if45dp, models, solver,  # noqa: F401
               solveras, solvercs)
from .etd import ETDConfig
from .etd4 import ETD4
from .etd5 import ETD5
from .etd34 import ETD34
from .etd35 import ETD35
from .if4 import IF4
from .if34 import IF34
from .if45dp import IF45DP
from .solveras import SolverConfig

__version__ = "0.1.0"
__all__ = ["IF4", "ETD4", "ETD5", "IF34", "ETD34", "ETD35", "IF45DP", "SolverConfig", "ETDConfig"]
