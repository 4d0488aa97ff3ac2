"""Fourier grid constructors on the device (rkstiff/grids.py:40-131): setup-time helpers."""
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
