"""Spectral derivatives on the device: the same calls as the reference's
``rkstiff/derivatives.py:47-179`` (``dx_rfft``, ``dx_fft``) for torch tensors.  Convenience for
post-processing snapshots (what the demos do after ``evolve``); not part of the stepping path, so the
transforms are torch's.  A leading batch dimension is allowed: the transform runs over the last axis.
"""
from __future__ import annotations

import torch


def _check_order(n) -> None:
    if not isinstance(n, int) or isinstance(n, bool):
        raise TypeError(f"derivative order n must be an integer, it is {n}")
    if n < 0:
        raise ValueError(f"derivative order n must be non-negative, it is {n}")


def dx_rfft(kx: torch.Tensor, u: torch.Tensor, n: int = 1) -> torch.Tensor:
    """n-th derivative of a real array: irfft((i kx)^n rfft(u)) (reference derivatives.py:47-123)."""
    _check_order(n)
    if u.is_complex():
        raise TypeError("dx_rfft requires real-valued input. Use dx_fft for complex arrays.")
    if u.numel() == 0:
        return torch.empty(0, dtype=torch.float64, device=u.device)
    if n == 0:
        return u
    u_fft = torch.fft.rfft(u, dim=-1)
    if tuple(kx.shape) != (u_fft.shape[-1],):
        raise ValueError(f"kx shape {tuple(kx.shape)} does not match rFFT output shape {(u_fft.shape[-1],)}. "
                         "For input size N, kx should have size N//2 + 1.")
    return torch.fft.irfft((1j * kx.to(u_fft.device)) ** n * u_fft, n=u.shape[-1], dim=-1)


def dx_fft(kx: torch.Tensor, u: torch.Tensor, n: int = 1) -> torch.Tensor:
    """n-th derivative of a complex periodic array: ifft((i kx)^n fft(u)) (reference derivatives.py:126-179)."""
    _check_order(n)
    if n == 0:
        return u
    u_fft = torch.fft.fft(u, dim=-1)
    if tuple(kx.shape) != (u_fft.shape[-1],):
        raise ValueError(f"kx shape {tuple(kx.shape)} must match FFT output {(u_fft.shape[-1],)}")
    return torch.fft.ifft((1j * kx.to(u_fft.device)) ** n * u_fft, dim=-1)
