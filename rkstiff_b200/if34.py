"""IF34: adaptive RK4(3) in integrating-factor form (rkstiff/if34.py:289-296), diagonal operators only."""
from __future__ import annotations

from typing import Optional, Union

from .solveras import BaseSolverAS, SolverConfig


class IF34(BaseSolverAS):
    METHOD = "IF34"

    def __init__(self, lin_op, nl_func, config: Optional[SolverConfig] = None, diagonalize: bool = False,
                 loglevel: Union[str, int] = "WARNING", group=None) -> None:
        if diagonalize:
            raise NotImplementedError("diagonalize=True (dense lin_op) is outside the diagonal hot path")
        super().__init__(lin_op, nl_func, config=config, loglevel=loglevel, group=group)
        self._h_coeff = None
