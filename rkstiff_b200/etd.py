"""ETDConfig, the psi functions and the ETD base classes (rkstiff/etd.py:46-278)."""
from __future__ import annotations

from typing import Optional, Union

import torch

from .solveras import BaseSolverAS, SolverConfig
from .solvercs import BaseSolverCS


class ETDConfig:
    """Contour / cutoff parameters of the psi-function evaluation (etd.py:81-131)."""

    def __init__(self, modecutoff: float = 0.01, contour_points: int = 32, contour_radius: float = 1.0) -> None:
        self.modecutoff = modecutoff
        self.contour_points = contour_points
        self.contour_radius = contour_radius

    @property
    def modecutoff(self) -> float:
        return self._modecutoff

    @modecutoff.setter
    def modecutoff(self, value: float) -> None:
        if value > 1.0 or value <= 0:
            raise ValueError(f"modecutoff must be between 0.0 and 1.0 but is {value}")
        self._modecutoff = value

    @property
    def contour_points(self) -> int:
        return self._contour_points

    @contour_points.setter
    def contour_points(self, value: int) -> None:
        if not isinstance(value, int):
            raise TypeError(f"contour_points must be an integer but is {value}")
        if value <= 1:
            raise ValueError(f"contour_points must be an integer greater than 1 but is {value}")
        self._contour_points = value

    @property
    def contour_radius(self) -> float:
        return self._contour_radius

    @contour_radius.setter
    def contour_radius(self, value: float) -> None:
        if value <= 0:
            raise ValueError(f"contour_radius must greater than 0 but is {value}")
        self._contour_radius = value


def psi1(z: torch.Tensor) -> torch.Tensor:
    """(e^z - 1)/z  (etd.py:148).  Convenience only; the engine evaluates psi in CUDA (csrc/coeffs.cuh)."""
    return (torch.exp(z) - 1) / z


def psi2(z: torch.Tensor) -> torch.Tensor:
    """2 (e^z - 1 - z)/z^2  (etd.py:165)."""
    return 2 * (torch.exp(z) - 1 - z) / z ** 2


def psi3(z: torch.Tensor) -> torch.Tensor:
    """6 (e^z - 1 - z - z^2/2)/z^3  (etd.py:182)."""
    return 6 * (torch.exp(z) - 1 - z - z ** 2 / 2) / z ** 3


class ETDAS(BaseSolverAS):
    def __init__(self, lin_op, nl_func, config: Optional[SolverConfig] = None,
                 etd_config: Optional[ETDConfig] = None, loglevel: Union[str, int] = "WARNING", group=None) -> None:
        super().__init__(lin_op, nl_func, config, loglevel=loglevel, group=group)
        self.etd_config = etd_config if etd_config is not None else ETDConfig()
        self._h_coeff = None


class ETDCS(BaseSolverCS):
    def __init__(self, lin_op, nl_func, etd_config: Optional[ETDConfig] = None,
                 loglevel: Union[str, int] = "WARNING", group=None) -> None:
        super().__init__(lin_op, nl_func, loglevel=loglevel, group=group)
        self.etd_config = etd_config if etd_config is not None else ETDConfig()
        self._h_coeff = None
