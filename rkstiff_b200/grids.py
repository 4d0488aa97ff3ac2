"""Grid constructors on the device: the Fourier ones (rkstiff/grids.py:40-131) and the Chebyshev pair
(rkstiff/grids.py:134-220) that the dense-operator models of ``diagonalize=True`` use.  Setup-time helpers."""
from __future__ import annotations

import math
from typing import Tuple

import torch


def _check(n: int) -> None:
    if not isinstance(n, int):
        raise TypeError("n must be an integer.")
    if n <= 2 or n % 2 != 0:
        raise ValueError("n must be an even integer greater than 2.")


def construct_x_kx_rfft(n: int, a: float = 0.0, b: float = 2 * math.pi, device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """x = a + dx*arange(n), kx = 2 pi rfftfreq(n, dx) as float64 tensors."""
    _check(n)
    dx = (b - a) / n
    x = torch.arange(n, dtype=torch.float64, device=device) * dx + a
    kx = 2 * math.pi * torch.fft.rfftfreq(n, d=dx, dtype=torch.float64, device=device)
    return x, kx


def construct_x_kx_fft(n: int, a: float = 0.0, b: float = 2 * math.pi, device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """x = a + dx*arange(n), kx = 2 pi fftfreq(n, dx) as float64 tensors."""
    _check(n)
    dx = (b - a) / n
    x = torch.arange(n, dtype=torch.float64, device=device) * dx + a
    kx = 2 * math.pi * torch.fft.fftfreq(n, d=dx, dtype=torch.float64, device=device)
    return x, kx


def construct_x_cheb(n: int, a: float = -1.0, b: float = 1.0, device="cuda") -> torch.Tensor:
    """The n + 1 Chebyshev-Gauss-Lobatto points of [a, b], ascending from a to b, as NumPy's ``chebpts2``
    orders them (rkstiff/grids.py:134-176)."""
    if not isinstance(n, int):
        raise TypeError("n must be an integer.")
    if n < 2:
        raise ValueError("n must be >= 2.")
    j = torch.arange(n + 1, dtype=torch.float64, device=device)
    x = torch.sin(math.pi * (2 * j - n) / (2 * n))            # = -cos(pi j / n), symmetric form (chebpts2)
    return a + (b - a) * (x + 1.0) / 2.0


def construct_x_dx_cheb(n: int, a: float = -1.0, b: float = 1.0, device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """Chebyshev points and the (n+1) x (n+1) differentiation matrix D, D @ f ~ df/dx (rkstiff/grids.py:178-220):
    D_ij = c_i / (c_j (x_i - x_j)) off the diagonal, c = (2, 1, ..., 1, 2) * (-1)^i, rows summing to zero."""
    x = construct_x_cheb(n, a, b, device)
    c = torch.ones(n + 1, dtype=torch.float64, device=device)
    c[0] = c[-1] = 2.0
    c = c * (-1.0) ** torch.arange(n + 1, dtype=torch.float64, device=device)
    dx = x[:, None] - x[None, :]
    d = torch.outer(c, 1.0 / c) / (dx + torch.eye(n + 1, dtype=torch.float64, device=device))
    d = d - torch.diag(d.sum(dim=1))
    return x, d

