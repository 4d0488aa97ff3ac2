"""Logger plumbing with the reference's names and semantics (rkstiff/util/loghelper.py:107-219):
one logger per solver class, named ``rkstiff.<ClassName>``, with a single stream handler."""
from __future__ import annotations

import logging
from typing import Union

_FORMAT = "%(asctime)s - %(name)s - %(levelname)s - %(message)s"
_NAMES = {logging.DEBUG: "DEBUG", logging.INFO: "INFO", logging.WARNING: "WARNING", logging.ERROR: "ERROR",
          logging.CRITICAL: "CRITICAL"}


def _parse_loglevel(loglevel: Union[str, int]) -> int:
    if isinstance(loglevel, str):
        level = getattr(logging, loglevel.upper(), None)
        if not isinstance(level, int):
            raise ValueError(f"Invalid log level: {loglevel}")
        return level
    return loglevel


def get_level_name(level: int) -> str:
    return _NAMES.get(level, f"Level {level}")


def setup_logger(name: str, loglevel: Union[str, int] = "WARNING") -> logging.Logger:
    logger = logging.getLogger(name)
    logger.setLevel(_parse_loglevel(loglevel))
    if not logger.handlers:
        handler = logging.StreamHandler()
        handler.setFormatter(logging.Formatter(_FORMAT, datefmt="%Y-%m-%d %H:%M:%S"))
        logger.addHandler(handler)
    return logger


def set_log_level(logger: logging.Logger, loglevel: Union[str, int]) -> None:
    logger.setLevel(_parse_loglevel(loglevel))


def get_solver_logger(solver_class: type, loglevel: Union[str, int] = "WARNING") -> logging.Logger:
    return setup_logger(f"rkstiff.{solver_class.__name__}", loglevel)
