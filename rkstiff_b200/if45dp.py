"""IF45DP: adaptive Dormand-Prince 5(4) in integrating-factor form (rkstiff/if45dp.py:47-70).

``r4_fix=False`` (default) reproduces the reference's error weight r4 = 17 h e^{z/5}/1920
(if45dp.py:234); ``r4_fix=True`` uses the Dormand-Prince value 71/1920.
"""
from __future__ import annotations

from typing import Optional, Union

from .solveras import BaseSolverAS, SolverConfig


class IF45DP(BaseSolverAS):
    METHOD = "IF45DP"

    def __init__(self, lin_op, nl_func, config: Optional[SolverConfig] = None,
                 loglevel: Union[str, int] = "WARNING", group=None, r4_fix: bool = False) -> None:
        self.r4_fix = bool(r4_fix)
        super().__init__(lin_op, nl_func, config=config, loglevel=loglevel, group=group)
        self._h_coeff = None
