"""IF34: adaptive RK4(3) in integrating-factor form (rkstiff/if34.py:289-296), diagonal operators only."""
from __future__ import annotations

from typing import Optional, Union

from .solveras import BaseSolverAS, SolverConfig


class IF34(BaseSolverAS):
    METHOD = "IF34"

    def __init__(self, lin_op, nl_func, config: Optional[SolverConfig] = None, diagonalize: bool = False,
                 loglevel: Union[str, int] = "WARNING", group=None) -> None:
        super().__init__(lin_op, nl_func, config=config, loglevel=loglevel, group=group)
        # diagonalize=True: a 2-D lin_op is a DENSE matrix, diagonalised once on the host; a 1-D lin_op is
        # already diagonal and the flag is ignored, as in the reference (tests/test_etd35.py:202-205)
        if diagonalize and lin_op.dim() >= 2:
            if group is not None:
                raise ValueError("diagonalize=True is a single-process mode")
            self._init_diagonalized(lin_op)
        self._h_coeff = None
