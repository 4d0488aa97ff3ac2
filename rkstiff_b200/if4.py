"""IF4: fixed-step RK4 in integrating-factor form (rkstiff/if4.py:240-245), diagonal operators only."""
from __future__ import annotations

from typing import Union

from .solvercs import BaseSolverCS


class IF4(BaseSolverCS):
    METHOD = "IF4"

    def __init__(self, lin_op, nl_func, loglevel: Union[str, int] = "WARNING", group=None) -> None:
        super().__init__(lin_op, nl_func, loglevel, group=group)
        self._h_coeff = None
