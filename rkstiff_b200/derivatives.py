"""Spectral derivatives on the device: the same calls as the reference's
``rkstiff/derivatives.py:47-179`` (``dx_rfft``, ``dx_fft``) for torch tensors, with a leading batch allowed (the
transform runs over the last axis).

CUDA tensors whose last axis is a power of two in 16 ... 8192 points run on the engine's own K4 transform pair
(``csrc/fft_fast.cuh``: ``DerivModel`` / ``DerivPairModel`` behind ``rks_rows_*``): one kernel reads the rows,
transforms them forward, multiplies by ``(i kx)^n`` and transforms back -- one read and one write of the array, no
intermediate spectrum in HBM.  Real rows go two at a time through one complex transform (``z = a + i b``; the
multiplier is Hermitian, so the result is ``a' + i b'``).  Other lengths on the device use ``torch.fft``; host
tensors (argument checks, the reference's doctests) use ``torch.fft`` on the host.
"""
from __future__ import annotations

from ctypes import byref, c_void_p

import torch


def _check_order(n) -> None:
    if not isinstance(n, int) or isinstance(n, bool):
        raise TypeError(f"derivative order n must be an integer, it is {n}")
    if n < 0:
        raise ValueError(f"derivative order n must be non-negative, it is {n}")


def _engine_length(npts: int) -> bool:
    return 16 <= npts <= 8192 and npts & (npts - 1) == 0


def _ik_power(kx: torch.Tensor, order: int) -> torch.Tensor:
    """(i kx)^order with exact zeros in the part that vanishes (NumPy's repeated complex products give the same)."""
    mag = kx.to(torch.float64) ** order
    quarter = order % 4
    zero = torch.zeros_like(mag)
    re, im = ((mag, zero), (zero, mag), (-mag, zero), (zero, -mag))[quarter]
    return torch.complex(re, im)


class SpectralDerivative:
    """Reusable handle: ``d = SpectralDerivative(kx, npts, order, real=True); ux = d(u)`` for CUDA tensors whose last
    axis has ``npts`` points (a power of two in 16 ... 8192).  ``kx`` as for ``dx_rfft`` (real=True: npts//2 + 1
    wavenumbers) or ``dx_fft`` (real=False: npts wavenumbers in FFT order)."""

    def __init__(self, kx: torch.Tensor, npts: int, order: int = 1, real: bool = True) -> None:
        from . import _abi
        _check_order(order)
        if not kx.is_cuda:
            raise ValueError("SpectralDerivative needs kx on a CUDA device")
        if not _engine_length(npts):
            raise ValueError("SpectralDerivative: the row length must be a power of two in 16 ... 8192")
        self.npts, self.order, self.real = int(npts), int(order), bool(real)
        self.device = kx.device
        m = _ik_power(kx, order)
        if real:
            if tuple(kx.shape) != (npts // 2 + 1,):
                raise ValueError(f"kx shape {tuple(kx.shape)} does not match rFFT output shape {(npts // 2 + 1,)}. "
                                 "For input size N, kx should have size N//2 + 1.")
            half = m.clone()
            half[-1] = half[-1].real + 0j                 # irfft keeps the real part of the Nyquist coefficient
            m = torch.cat([half, half[1:-1].flip(0).conj()])
        elif tuple(kx.shape) != (npts,):
            raise ValueError(f"kx shape {tuple(kx.shape)} must match FFT output {(npts,)}")
        # the kernel works on conjugated data (fft_fast.cuh): table = conj(M) / n
        table = torch.view_as_real((m.conj() / npts).contiguous()).contiguous()
        self._abi = _abi
        self._h = c_void_p()
        model = _abi.MODEL_DERIV_RFFT_PAIR if real else _abi.MODEL_DERIV_FFT
        with torch.cuda.device(self.device):
            st = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _abi.check(_abi.lib.rks_rows_create(byref(self._h), model, self.npts, c_void_p(table.data_ptr()), 0.0, st))
        self._table = table                                  # alive until the stream-ordered copy has run

    def __call__(self, u: torch.Tensor) -> torch.Tensor:
        if u.shape[-1] != self.npts or u.device != self.device:
            raise ValueError(f"SpectralDerivative: the last axis must have {self.npts} points on {self.device}")
        rows = u.numel() // self.npts
        st = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if self.real:
            x = u.to(torch.float64).reshape(rows, self.npts).contiguous()
            if rows & 1:                                      # rows go in pairs: pad with a zero row
                x = torch.cat([x, torch.zeros(1, self.npts, dtype=torch.float64, device=self.device)])
            out = torch.empty_like(x)
            self._abi.check(self._abi.lib.rks_rows_apply(self._h, c_void_p(x.data_ptr()), c_void_p(out.data_ptr()),
                                                         x.shape[0] // 2, st))
            return out[:rows].reshape(u.shape)
        x = u.to(torch.complex128).reshape(rows, self.npts).contiguous()
        out = torch.empty_like(x)
        self._abi.check(self._abi.lib.rks_rows_apply(self._h, c_void_p(x.data_ptr()), c_void_p(out.data_ptr()), rows, st))
        return out.reshape(u.shape)

    def __del__(self):
        try:
            if self._h:
                self._abi.lib.rks_rows_destroy(self._h)
                self._h = None
        except Exception:
            pass


def dx_rfft(kx: torch.Tensor, u: torch.Tensor, n: int = 1) -> torch.Tensor:
    """n-th derivative of a real array: irfft((i kx)^n rfft(u)) (reference derivatives.py:47-123)."""
    _check_order(n)
    if u.is_complex():
        raise TypeError("dx_rfft requires real-valued input. Use dx_fft for complex arrays.")
    if u.numel() == 0:
        return torch.empty(0, dtype=torch.float64, device=u.device)
    if n == 0:
        return u
    npts = u.shape[-1]
    if tuple(kx.shape) != (npts // 2 + 1,):
        raise ValueError(f"kx shape {tuple(kx.shape)} does not match rFFT output shape {(npts // 2 + 1,)}. "
                         "For input size N, kx should have size N//2 + 1.")
    if u.is_cuda and _engine_length(npts):
        return SpectralDerivative(kx.to(u.device), npts, n, real=True)(u)
    u_fft = torch.fft.rfft(u, dim=-1)
    return torch.fft.irfft((1j * kx.to(u_fft.device)) ** n * u_fft, n=npts, dim=-1)


def dx_fft(kx: torch.Tensor, u: torch.Tensor, n: int = 1) -> torch.Tensor:
    """n-th derivative of a complex periodic array: ifft((i kx)^n fft(u)) (reference derivatives.py:126-179)."""
    _check_order(n)
    if n == 0:
        return u
    npts = u.shape[-1]
    if tuple(kx.shape) != (npts,):
        raise ValueError(f"kx shape {tuple(kx.shape)} must match FFT output {(npts,)}")
    if u.is_cuda and _engine_length(npts):
        return SpectralDerivative(kx.to(u.device), npts, n, real=False)(u)
    u_fft = torch.fft.fft(u, dim=-1)
    return torch.fft.ifft((1j * kx.to(u_fft.device)) ** n * u_fft, dim=-1)
