"""ETD5: fixed-step exponential time differencing (rkstiff/etd5.py:535-541), diagonal operators only."""
from __future__ import annotations

from typing import Optional, Union

from .etd import ETDCS, ETDConfig


class ETD5(ETDCS):
    METHOD = "ETD5"

    def __init__(self, lin_op, nl_func, etd_config: Optional[ETDConfig] = None,
                 loglevel: Union[str, int] = "WARNING", group=None) -> None:
        super().__init__(lin_op, nl_func, etd_config=etd_config, loglevel=loglevel, group=group)
