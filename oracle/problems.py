"""NumPy problem zoo for the oracle, the golden generator and the CPU baseline.

TEST INFRASTRUCTURE ONLY (see oracle/rk_oracle.py header).  Every builder returns a
``Problem`` whose ``nl_func`` is the NumPy closure the reference would be given and whose
``model``/``params`` name the fused CUDA nonlinearity that computes the same thing.

Formulations (reference file:line):
  KS ........ README.md:86-104, demos/ks.ipynb:126-131
  KdV ....... rkstiff/models.py:113-145, tests/testing_util.py:42-56
  Burgers ... rkstiff/models.py:153-194, tests/testing_util.py:28-39
  NLS 1-D ... demos/nls.ipynb (L = -i k^2, N = i*gamma*F{|u|^2 u}, gamma = 2)
  grids ..... rkstiff/grids.py:40-131
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import numpy as np


@dataclass
class Problem:
    name: str
    lin_op: np.ndarray                 # (n_c,) float64 or complex128
    nl_func: Callable[[np.ndarray], np.ndarray]
    u0: np.ndarray                     # (n_c,) or (B, n_c) complex128 spectrum
    model: str                         # fused-NL id: "ks" | "burgers" | "kdv" | "nls"
    n: int                             # real-space grid points per trajectory
    kx: np.ndarray                     # wavenumbers (rfft or fft layout)
    params: Dict[str, float] = field(default_factory=dict)
    x: Optional[np.ndarray] = None


def x_kx_rfft(n: int, a: float, b: float):
    """grids.py:40-85: x = arange(a, b, dx), kx = 2 pi rfftfreq(n, dx)."""
    dx = (b - a) / n
    return np.arange(a, b, dx), 2 * np.pi * np.fft.rfftfreq(n, d=dx)


def x_kx_fft(n: int, a: float, b: float):
    """grids.py:88-131."""
    dx = (b - a) / n
    return np.arange(a, b, dx), 2 * np.pi * np.fft.fftfreq(n, d=dx)


def _uux_nl(kx: np.ndarray, c: float):
    """N(u^) = -c * rfft(irfft(u^) * irfft(i kx u^)) along the last axis (models.py:140-143)."""
    def nl(uf):
        u = np.fft.irfft(uf, axis=-1)
        ux = np.fft.irfft(1j * kx * uf, axis=-1)
        if c == 1.0:
            return -np.fft.rfft(u * ux, axis=-1)
        return -c * np.fft.rfft(u * ux, axis=-1)
    return nl


def ks(n: int = 1024, batch: int = 0, seed: int = 0) -> Problem:
    """Kuramoto-Sivashinsky on [0, 32 pi): L = k^2(1-k^2), N = -F{u u_x}."""
    x, kx = x_kx_rfft(n, 0.0, 32.0 * np.pi)
    lin = kx ** 2 * (1 - kx ** 2)
    if batch:
        phi = np.random.default_rng(seed).uniform(0.0, 2 * np.pi, size=(batch, 1))
        u0 = np.cos(x[None, :] / 16 + phi) * (1.0 + np.sin(x[None, :] / 16))
    else:
        u0 = np.cos(x / 16) * (1.0 + np.sin(x / 16))
    return Problem("ks", lin, _uux_nl(kx, 1.0), np.fft.rfft(u0, axis=-1), "ks", n, kx, {"c": 1.0}, x)


def kdv_soliton_profile(x, ampl=0.5, x0=0.0, t=0.0):
    """models.py:32: 0.5 a^2 sech^2(a (x - x0 - a^2 t)/2)."""
    return 0.5 * ampl ** 2 / np.cosh(ampl * (x - x0 - ampl ** 2 * t) / 2) ** 2


def kdv(n: int = 256, batch: int = 0, seed: int = 0) -> Problem:
    """KdV soliton test problem of tests/testing_util.py:42-56: L = i k^3, N = -6 F{u u_x}."""
    x, kx = x_kx_rfft(n, -30.0, 30.0)
    lin = 1j * kx ** 3
    if batch:
        rng = np.random.default_rng(seed)
        a = rng.uniform(0.8, 1.2, size=(batch, 1))
        x0 = rng.uniform(-8.0, -2.0, size=(batch, 1))
        u0 = kdv_soliton_profile(x[None, :], a, x0)
    else:
        u0 = kdv_soliton_profile(x, 1.0, -5.0)
    return Problem("kdv", lin, _uux_nl(kx, 6.0), np.fft.rfft(u0, axis=-1), "kdv", n, kx, {"c": 6.0}, x)


def burgers(n: int = 1024, mu: float = 0.0005, batch: int = 0, seed: int = 0) -> Problem:
    """Viscous Burgers of tests/testing_util.py:28-39: L = -mu k^2, N = -F{u u_x}."""
    x, kx = x_kx_rfft(n, -np.pi, np.pi)
    lin = -mu * kx ** 2
    if batch:
        amp = np.random.default_rng(seed).uniform(0.5, 1.5, size=(batch, 1))
        u0 = amp * np.exp(-10 * np.sin(x[None, :] / 2) ** 2)
    else:
        u0 = np.exp(-10 * np.sin(x / 2) ** 2)
    return Problem("burgers", lin, _uux_nl(kx, 1.0), np.fft.rfft(u0, axis=-1), "burgers", n, kx,
                   {"c": 1.0, "mu": mu}, x)


def nls(n: int = 8192, batch: int = 0, seed: int = 2, gamma: float = 2.0,
        half_width: float = 40.0 * np.pi) -> Problem:
    """Focusing NLS u_t = i u_xx + i gamma |u|^2 u on [-W, W): L = -i k^2, N = i gamma F{|u|^2 u}.

    Batch members are solitons eta sech(eta (x - x0)) exp(i c x) (cfg 2).
    """
    x, kx = x_kx_fft(n, -half_width, half_width)
    lin = -1j * kx ** 2

    def nl(uf):
        f = np.fft.ifft(uf, axis=-1)
        f2 = f.real ** 2 + f.imag ** 2
        return 1j * gamma * np.fft.fft(f2 * f, axis=-1)

    if batch:
        rng = np.random.default_rng(seed)
        eta = rng.uniform(0.5, 1.5, size=(batch, 1))
        x0 = rng.uniform(-20.0, 20.0, size=(batch, 1))
        c = rng.uniform(-0.5, 0.5, size=(batch, 1))
        u0 = eta / np.cosh(eta * (x[None, :] - x0)) * np.exp(1j * c * x[None, :])
    else:
        u0 = 1.0 / np.cosh(x) * np.exp(0.25j * x)
    return Problem("nls", lin, nl, np.fft.fft(u0, axis=-1), "nls", n, kx, {"gamma": gamma}, x)


BUILDERS = {"ks": ks, "kdv": kdv, "burgers": burgers, "nls": nls}


# ----------------------------------------------------------------------------------------------
# N-D grids, formulated the way the reference's demos do it: lin_op and u FLATTENED to 1-D and
# reshaped inside nl_func (demos/nls.ipynb:496-511).  These are the oracles for
# "lin_op shaped like u" on the engine side.
# ----------------------------------------------------------------------------------------------
def allen_cahn_2d(n: int = 256, eps: float = 0.01, seed: int = 1234) -> Problem:
    """Periodic Allen-Cahn u_t = eps lap(u) + u - u^3 on [0, 2 pi)^2, rfft2 half spectrum (n, n/2+1):
    L = 1 - eps (kx^2 + ky^2), N = -rfft2(irfft2(u^)^3)   (SURVEY.md 8d cfg 4)."""
    x = np.arange(n) * (2 * np.pi / n)
    ky = 2 * np.pi * np.fft.fftfreq(n, d=2 * np.pi / n)
    kx = 2 * np.pi * np.fft.rfftfreq(n, d=2 * np.pi / n)
    KY, KX = np.meshgrid(ky, kx, indexing="ij")
    shape = KX.shape
    lin = (1.0 - eps * (KX ** 2 + KY ** 2)).ravel()
    rng = np.random.default_rng(seed)
    Y, X = np.meshgrid(x, x, indexing="ij")
    u0 = np.zeros((n, n))
    for _ in range(16):
        g, m, p, th = rng.standard_normal(), rng.integers(-4, 5), rng.integers(-4, 5), rng.uniform(0, 2 * np.pi)
        u0 += 0.1 * g * np.cos(m * X + p * Y + th)

    def nl(uf):
        u = np.fft.irfft2(uf.reshape(shape), s=(n, n))
        return -np.fft.rfft2(u ** 3).ravel()

    prob = Problem("allen_cahn_2d", lin, nl, np.fft.rfft2(u0).ravel(), "none", n * n, kx, {"eps": eps}, x)
    prob.params["shape"] = shape
    return prob


def nls_3d(n: int = 32, gamma: float = 2.0, half_width: float = 6.0) -> Problem:
    """3-D NLS u_t = i lap(u) + i gamma |u|^2 u on [-W, W)^3: L = -i |k|^2, N = i gamma fftn(|f|^2 f)
    with a Gaussian initial condition (SURVEY.md 8d cfg 5)."""
    x, k = x_kx_fft(n, -half_width, half_width)
    KX, KY, KZ = np.meshgrid(k, k, k, indexing="ij")
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    shape = (n, n, n)
    lin = (-1j * (KX ** 2 + KY ** 2 + KZ ** 2)).ravel()
    u0 = np.exp(-(X ** 2 + Y ** 2 + Z ** 2)).astype(np.complex128)

    def nl(uf):
        f = np.fft.ifftn(uf.reshape(shape))
        f2 = f.real ** 2 + f.imag ** 2
        return (1j * gamma * np.fft.fftn(f2 * f)).ravel()

    prob = Problem("nls_3d", lin, nl, np.fft.fftn(u0).ravel(), "none", n ** 3, k, {"gamma": gamma}, x)
    prob.params["shape"] = shape
    return prob


# ----------------------------------------------------------------------------------------------
# Fourier-diagonal models without a reference formulation: the oracle is the
# reference's solver classes (restated in rk_oracle.py) driven by these NumPy closures.
# ----------------------------------------------------------------------------------------------
def allen_cahn_1d(n: int = 256, eps: float = 0.01, batch: int = 0, seed: int = 5) -> Problem:
    """Periodic 1-D Allen-Cahn u_t = eps u_xx + u - u^3 on [0, 2 pi): L = 1 - eps k^2 (the +u goes into L,
    like rkstiff/models.py:240-244), N = -rfft(irfft(u^)^3)."""
    x, kx = x_kx_rfft(n, 0.0, 2 * np.pi)
    lin = 1.0 - eps * kx ** 2
    rng = np.random.default_rng(seed)

    def ic():
        u = np.zeros(n)
        for m in range(1, 6):
            u += 0.3 * rng.standard_normal() * np.cos(m * x + rng.uniform(0, 2 * np.pi))
        return u

    u0 = np.stack([ic() for _ in range(batch)]) if batch else ic()

    def nl(uf):
        u = np.fft.irfft(uf, axis=-1)
        return -np.fft.rfft(u ** 3, axis=-1)

    return Problem("allen_cahn_1d", lin, nl, np.fft.rfft(u0, axis=-1), "cubic", n, kx, {"c": -1.0, "eps": eps}, x)


def sine_gordon(n: int = 256, batch: int = 0, half_width: float = 20.0, seed: int = 7) -> Problem:
    """Sine-Gordon phi_tt = phi_xx - sin(phi) in first-order complex form psi = phi_t + i Omega phi,
    Omega = sqrt(1 + k^2):  L = i Omega,  N(psi^) = fft(phi - sin phi),
    phi^(k) = (psi^(k) - conj(psi^(-k))) / (2 i Omega(k)).  Initial data: breathers at rest."""
    x, k = x_kx_fft(n, -half_width, half_width)
    omega = np.sqrt(1.0 + k ** 2)
    lin = 1j * omega
    rev = (-np.arange(n)) % n

    def nl(pf):
        phi_hat = (pf - np.conj(pf[..., rev])) / (2j * omega)
        phi = np.fft.ifft(phi_hat, axis=-1).real
        return np.fft.fft(phi - np.sin(phi), axis=-1)

    def breather(w, x0):
        s = np.sqrt(1 - w * w)
        return 4 * np.arctan(s / w / np.cosh(s * (x - x0)))

    if batch:
        rng = np.random.default_rng(seed)
        phi0 = np.stack([breather(rng.uniform(0.3, 0.8), rng.uniform(-5, 5)) for _ in range(batch)])
    else:
        phi0 = breather(0.5, 0.0)
    psi0 = 1j * omega * np.fft.fft(phi0, axis=-1)            # phi_t = 0
    return Problem("sine_gordon", lin, nl, psi0, "sine_gordon", n, omega, {}, x)


def dense_advection_diffusion(n: int = 16, nu: float = 0.05, c: float = 0.5, omega: float = 3.0) -> Problem:
    """A small DENSE-operator problem for ``diagonalize=True``: two fields (p, q) on (0, 1), homogeneous
    Dirichlet ends, centred differences, each with A = nu d_xx - c d_x (non-symmetric: the eigenvector
    matrix is not orthogonal, cond(S) ~ 1e2) and rotating into each other at rate omega (complex
    eigenvalue pairs lambda_k +- i omega):  L = [[A, omega I], [-omega I, A]],  N(u) = u - u^3."""
    dx = 1.0 / (n + 1)
    x = np.arange(1, n + 1) * dx
    lo = nu / dx ** 2 + c / (2 * dx)
    hi = nu / dx ** 2 - c / (2 * dx)
    amat = np.diag(np.full(n, -2 * nu / dx ** 2)) + np.diag(np.full(n - 1, lo), -1) + np.diag(np.full(n - 1, hi), 1)
    eye = np.eye(n)
    lin = np.block([[amat, omega * eye], [-omega * eye, amat]])

    def nl(u):
        return u - u ** 3

    u0 = np.concatenate([0.8 * np.sin(np.pi * x) + 0.3 * np.sin(3 * np.pi * x), 0.5 * np.sin(2 * np.pi * x)])
    return Problem("dense_advdiff", lin, nl, u0.astype(np.complex128), "none", 2 * n, x, {"nu": nu, "c": c, "omega": omega}, x)


def allen_cahn_cheb(n: int = 20, eps: float = 0.01) -> Problem:
    """The reference's own dense-operator test problem (tests/testing_util.py:12-25, models.py:202-262):
    Allen-Cahn on n + 1 Chebyshev points, w = u - x on the interior, lin_op = eps D^2 + I (dense)."""
    j = np.arange(n + 1)
    x = np.polynomial.chebyshev.chebpts2(n + 1)
    c = np.r_[2, np.ones(n - 1), 2] * np.power(-1.0, j)
    dmat = np.outer(c, 1.0 / c) / ((x[:, None] - x[None, :]) + np.eye(n + 1))
    dmat = dmat - np.diag(dmat.sum(axis=1))
    lin = (eps * dmat.dot(dmat) + np.eye(n + 1))[1:-1, 1:-1]
    xi = x[1:-1]

    def nl(w):
        return np.asarray(xi - np.power(w + xi, 3), dtype=np.complex128).ravel()

    u0 = 0.53 * x + 0.47 * np.sin(-1.5 * np.pi * x)
    return Problem("allen_cahn_cheb", lin, nl, (u0 - x)[1:-1].astype(np.complex128), "none", n - 1, xi,
                   {"eps": eps, "u0int": u0[1:-1]}, xi)

