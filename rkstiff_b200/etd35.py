"""ETD35: adaptive exponential time differencing (rkstiff/etd35.py:784-792), diagonal operators only."""
from __future__ import annotations

from typing import Optional, Union

from .etd import ETDAS, ETDConfig, SolverConfig  # noqa: F401  (re-exported like the reference module)


class ETD35(ETDAS):
    METHOD = "ETD35"

    def __init__(self, lin_op, nl_func, config: Optional[SolverConfig] = None,
                 etd_config: Optional[ETDConfig] = None, diagonalize: bool = False,
                 loglevel: Union[str, int] = "WARNING", group=None) -> None:
        if diagonalize:
            raise NotImplementedError("diagonalize=True (dense lin_op) is outside the diagonal hot path")
        super().__init__(lin_op, nl_func, config=config, etd_config=etd_config, loglevel=loglevel, group=group)
