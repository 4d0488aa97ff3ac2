"""SolverType enum (rkstiff/util/solver_type.py)."""
from __future__ import annotations

from enum import Enum, auto


class SolverType(Enum):
    CONSTANT_STEP = auto()
    ADAPTIVE_STEP = auto()
    CS = CONSTANT_STEP
    AS = ADAPTIVE_STEP

    def __str__(self) -> str:
        return self.name.replace("_", " ").title()

    @classmethod
    def from_solver(cls, solver) -> "SolverType":
        from ..solveras import BaseSolverAS
        from ..solvercs import BaseSolverCS
        if isinstance(solver, BaseSolverAS):
            return cls.ADAPTIVE_STEP
        if isinstance(solver, BaseSolverCS):
            return cls.CONSTANT_STEP
        raise TypeError(f"Cannot determine solver type for {type(solver).__name__}. "
                        "Solver must inherit from BaseSolverCS or BaseSolverAS.")
