"""Model zoo: linear operators and nonlinear terms of the reference's spectral equations, with the
nonlinear term available both as a fused CUDA kernel (K4, csrc/fft.cuh) and as a torch callable.

Reference formulations (file:line):
  kdv_ops ........ rkstiff/models.py:113-145   L = i k^3,        N = -6 F{u u_x}
  burgers_ops .... rkstiff/models.py:153-194   L = -mu k^2,      N = -F{u u_x}
  ks_ops ......... README.md:86-97             L = k^2 (1-k^2),  N = -F{u u_x}
  nls_ops ........ demos/nls.ipynb             L = -i k^2,       N = i gamma F{|u|^2 u}
  kdv_soliton .... rkstiff/models.py:32, kdv_multi_soliton models.py:76-110
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _abi


class FusedNL:
    """Handle of a fused spectral nonlinearity.

    Passed as ``nl_func`` it makes the solver run the hand-written shared-memory FFT kernel;
    called like a function it evaluates the same expression with ``torch.fft`` (used by tests
    and wherever a plain callable is wanted).
    """

    def __init__(self, model_id: int, n: int, kx: Optional[torch.Tensor], param: float, name: str) -> None:
        if n < 16 or n & (n - 1):
            raise ValueError("fused nonlinearities need a power-of-two grid with n >= 16")
        self.model_id = model_id
        self.n = int(n)
        self.kx = kx
        self.param = float(param)
        self.name = name

    def __call__(self, uf: torch.Tensor) -> torch.Tensor:
        if self.model_id == _abi.MODEL_UUX_RFFT:
            u = torch.fft.irfft(uf, n=self.n, dim=-1)
            ux = torch.fft.irfft(1j * self.kx * uf, n=self.n, dim=-1)
            return -self.param * torch.fft.rfft(u * ux, dim=-1)
        if self.model_id == _abi.MODEL_CUBIC_RFFT:
            u = torch.fft.irfft(uf, n=self.n, dim=-1)
            return self.param * torch.fft.rfft(u * u * u, dim=-1)
        if self.model_id == _abi.MODEL_SINE_GORDON:
            rev = (-torch.arange(self.n, device=uf.device)) % self.n
            phi_hat = (uf - torch.conj(uf[..., rev])) / (2j * self.kx)
            phi = torch.fft.ifft(phi_hat, dim=-1).real
            return torch.fft.fft(phi - torch.sin(phi), dim=-1)
        f = torch.fft.ifft(uf, dim=-1)
        f2 = f.real ** 2 + f.imag ** 2
        return 1j * self.param * torch.fft.fft(f2 * f, dim=-1)

    def __repr__(self) -> str:
        return f"FusedNL({self.name}, n={self.n}, param={self.param})"


class FusedGridNL(FusedNL):
    """Handle of a fused nonlinearity on a 2-D / 3-D spectral grid (``rks_set_model_nd``): passed as ``nl_func`` the
    engine launches the strided-axis and row FFT kernels itself, predicated on the device, so whole adaptive trials
    are enqueued and graph-replayed without a host sync (the N-D closures of demos/nls.ipynb:496-511 inside the
    trial loop solveras.py:379-410).  Called like a function it runs the same kernels from Python."""

    def __init__(self, model_id: int, grid: Sequence[int], param: float, name: str, compose) -> None:
        self.model_id = model_id
        self.grid = tuple(int(g) for g in grid)
        self.n = self.grid[-1]
        self.kx = None
        self.param = float(param)
        self.name = name
        self._compose = compose
        self.supports_out = True

    def __call__(self, uf: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        return self._compose(uf, out=out)

    def __repr__(self) -> str:
        return f"FusedGridNL({self.name}, grid={self.grid}, param={self.param})"


def _n_from_rfft_kx(kx: torch.Tensor) -> int:
    return 2 * (kx.shape[-1] - 1)


def kdv_ops(kx: torch.Tensor) -> Tuple[torch.Tensor, FusedNL]:
    """KdV u_t = -u_xxx - 6 u u_x in rfft space (models.py:113-145)."""
    lin_op = 1j * kx.to(torch.complex128) ** 3
    return lin_op, FusedNL(_abi.MODEL_UUX_RFFT, _n_from_rfft_kx(kx), kx, 6.0, "kdv")


def burgers_ops(kx: torch.Tensor, mu: float) -> Tuple[torch.Tensor, FusedNL]:
    """Viscous Burgers u_t = mu u_xx - u u_x in rfft space (models.py:153-194)."""
    lin_op = -mu * kx ** 2
    return lin_op, FusedNL(_abi.MODEL_UUX_RFFT, _n_from_rfft_kx(kx), kx, 1.0, "burgers")


def ks_ops(kx: torch.Tensor) -> Tuple[torch.Tensor, FusedNL]:
    """Kuramoto-Sivashinsky u_t = -u_xx - u_xxxx - u u_x in rfft space (README.md:86-97)."""
    lin_op = kx ** 2 * (1 - kx ** 2)
    return lin_op, FusedNL(_abi.MODEL_UUX_RFFT, _n_from_rfft_kx(kx), kx, 1.0, "ks")


def nls_ops(kx: torch.Tensor, gamma: float = 2.0) -> Tuple[torch.Tensor, FusedNL]:
    """Cubic NLS u_t = i u_xx + i gamma |u|^2 u in fft space (demos/nls.ipynb)."""
    lin_op = -1j * kx.to(torch.complex128) ** 2
    return lin_op, FusedNL(_abi.MODEL_NLS_FFT, kx.shape[-1], kx, gamma, "nls")


def allen_cahn_1d_ops(kx: torch.Tensor, eps: float = 0.01) -> Tuple[torch.Tensor, FusedNL]:
    """Periodic 1-D Allen-Cahn u_t = eps u_xx + u - u^3 in rfft space: L = 1 - eps k^2 (the +u goes into
    L, mirroring the split of rkstiff/models.py:240-244), N = -F{u^3}."""
    lin_op = 1.0 - eps * kx ** 2
    return lin_op, FusedNL(_abi.MODEL_CUBIC_RFFT, _n_from_rfft_kx(kx), None, -1.0, "allen_cahn")


def sine_gordon_ops(kx: torch.Tensor) -> Tuple[torch.Tensor, FusedNL]:
    """Sine-Gordon phi_tt = phi_xx - sin(phi) on the full fft grid in first-order complex form
    psi = phi_t + i Omega phi, Omega = sqrt(1 + k^2) (the reference's own SG demo is a Chebyshev dense
    system; this Fourier-diagonal restatement is SURVEY.md 8f-1):  L = i Omega,  N = F{phi - sin phi}.
    State: psi^ = F{phi_t} + i Omega F{phi};  phi^ = (psi^(k) - conj(psi^(-k))) / (2 i Omega)."""
    omega = torch.sqrt(1.0 + kx.to(torch.float64) ** 2)
    lin_op = 1j * omega.to(torch.complex128)
    return lin_op, FusedNL(_abi.MODEL_SINE_GORDON, kx.shape[-1], omega, 0.0, "sine_gordon")


def kdv_soliton(x: torch.Tensor, ampl: float = 0.5, x0: float = 0.0, t: float = 0.0) -> torch.Tensor:
    """Single KdV soliton 0.5 a^2 sech^2(a (x - x0 - a^2 t)/2) (models.py:32)."""
    return 0.5 * ampl ** 2 / torch.cosh(ampl * (x - x0 - ampl ** 2 * t) / 2) ** 2


def kdv_multi_soliton(x: torch.Tensor, ampl: Sequence[float], x0: Sequence[float], t: float = 0.0) -> torch.Tensor:
    """Superposition of KdV solitons (models.py:76-110)."""
    if len(x0) != len(ampl):
        raise ValueError("Lengths of ampl and x0 must match.")
    out = torch.zeros_like(x)
    for a, c in zip(ampl, x0):
        out = out + kdv_soliton(x, a, c, t)
    return out


def allen_cahn_ops(x: torch.Tensor, d_cheb_matrix: torch.Tensor, epsilon: float = 0.01):
    """Allen-Cahn on a Chebyshev grid with the ends pinned to u(+-1) = +-1, written for w = u - x
    (rkstiff/models.py:202-262): DENSE ``lin_op = eps D^2 + I`` without its boundary rows/columns and
    ``nl_func(w) = x - (w + x)^3`` on the interior points.  Use with ``diagonalize=True``."""
    d2 = d_cheb_matrix @ d_cheb_matrix
    lin_op = (epsilon * d2 + torch.eye(d2.shape[0], dtype=d2.dtype, device=d2.device))[1:-1, 1:-1].contiguous()
    xi = x[1:-1]

    def nl_func(w: torch.Tensor) -> torch.Tensor:
        return (xi - (w + xi) ** 3).to(torch.complex128)

    return lin_op, nl_func


# ----------------------------------------------------------------------------------------------
# N-D Fourier-diagonal models.  lin_op has the shape of the spectral grid and u
# that shape (plus optional leading batch dims): the engine's "lin_op shaped like u" path.  The
# nonlinear term is a Python callable that composes the engine's own kernels: AxisFFT over the strided
# axes (csrc/fft_axis.cuh) around RowNL, the fused last-axis kernel.  Grids whose axis lengths are not
# powers of two (16..4096 strided, 16..16384 last axis) fall back to torch.fft + the pointwise kernel.
# ----------------------------------------------------------------------------------------------
def pointwise_(model_id: int, x: torch.Tensor, p0: float) -> torch.Tensor:
    """In-place pointwise nonlinearity of an N-D model in one CUDA kernel (rks_pointwise):
    MODEL_NLS_FFT: x <- i p0 |x|^2 x (complex128);  MODEL_CUBIC_RFFT: x <- p0 x^3 (float64)."""
    from ctypes import c_void_p
    if not x.is_contiguous():
        raise ValueError("pointwise_ needs a contiguous tensor")
    st = c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    _abi.check(_abi.lib.rks_pointwise(model_id, c_void_p(x.data_ptr()), c_void_p(x.data_ptr()), x.numel(), float(p0), st))
    return x


class RowNL:
    """Fused ``F{ N( F^-1{ . } ) }`` along the LAST axis of any contiguous array (rks_rows_*): the
    innermost-axis part of an N-D nonlinear term, done by the engine's hand-written FFT kernel in one
    read + one write.  The outer axes are transformed by the caller (AxisFFT)."""

    def __init__(self, model_id: int, n: int, kx: Optional[torch.Tensor], p0: float, device) -> None:
        from ctypes import byref, c_void_p
        self.model_id, self.n, self.p0 = model_id, int(n), float(p0)
        self.n_c = n // 2 + 1 if model_id in (_abi.MODEL_UUX_RFFT, _abi.MODEL_CUBIC_RFFT) else n
        self.device = torch.device(device)
        self._h = c_void_p()
        kxd = kx.to(device=self.device, dtype=torch.float64).contiguous() if kx is not None else None
        with torch.cuda.device(self.device):
            st = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _abi.check(_abi.lib.rks_rows_create(byref(self._h), model_id, self.n,
                                                c_void_p(kxd.data_ptr()) if kxd is not None else None, self.p0, st))

    def __call__(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        from ctypes import c_void_p
        if x.dtype != torch.complex128 or not x.is_contiguous() or x.shape[-1] != self.n_c:
            raise ValueError(f"RowNL needs a contiguous complex128 array with last dimension {self.n_c}")
        if out is None:
            out = torch.empty_like(x)
        elif not out.is_contiguous() or out.shape != x.shape or out.dtype != x.dtype:
            raise ValueError("RowNL: out must be a contiguous array shaped like the input")
        st = c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _abi.check(_abi.lib.rks_rows_apply(self._h, c_void_p(x.data_ptr()), c_void_p(out.data_ptr()),
                                           x.numel() // self.n_c, st))
        return out

    def __del__(self):
        try:
            if self._h:
                _abi.lib.rks_rows_destroy(self._h)
                self._h = None
        except Exception:
            pass


class AxisFFT:
    """Hand-written FP64 transform along a NON-last axis of a contiguous complex128 array (rks_axis_*,
    csrc/fft_axis.cuh), in place.  ``inverse_`` leaves the axis in digit-reversed order and ``forward_``
    expects that order, so the pair brackets a pointwise nonlinearity exactly like ifft / fft do."""

    def __init__(self, n: int, device) -> None:
        from ctypes import byref, c_void_p
        self.n = int(n)
        self.device = torch.device(device)
        self._h = c_void_p()
        with torch.cuda.device(self.device):
            st = c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _abi.check(_abi.lib.rks_axis_create(byref(self._h), self.n, st))

    @staticmethod
    def supported(n: int) -> bool:
        return 16 <= n <= 4096 and n & (n - 1) == 0

    def _apply(self, x: torch.Tensor, dim: int, inverse: int, out: Optional[torch.Tensor]) -> torch.Tensor:
        from ctypes import c_void_p
        dim = dim % x.dim()
        if x.dtype != torch.complex128 or not x.is_contiguous() or x.shape[dim] != self.n or dim == x.dim() - 1:
            raise ValueError(f"AxisFFT needs a contiguous complex128 array with {self.n} points along a non-last axis")
        if out is None:
            out = x
        elif out.dtype != x.dtype or out.shape != x.shape or not out.is_contiguous():
            raise ValueError("AxisFFT: out must be a contiguous array shaped like the input")
        outer = 1
        for d in x.shape[:dim]:
            outer *= int(d)
        inner = 1
        for d in x.shape[dim + 1:]:
            inner *= int(d)
        st = c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _abi.check(_abi.lib.rks_axis_apply(self._h, c_void_p(x.data_ptr()), c_void_p(out.data_ptr()), outer, inner,
                                           inverse, st))
        return out

    def chunked_(self, x: torch.Tensor, inverse: bool) -> torch.Tensor:
        """In-place transform of an axis that arrives split into G row blocks stored block-major: ``x`` has shape
        ``(G, outer, n / G, ...)`` (what the all-to-all of a slab decomposition delivers, dist_fft.py)."""
        from ctypes import c_void_p
        if x.dtype != torch.complex128 or not x.is_contiguous() or x.dim() < 4 or x.shape[0] * x.shape[2] != self.n:
            raise ValueError(f"AxisFFT.chunked_ needs a contiguous complex128 array (G, outer, {self.n}/G, inner...)")
        inner = 1
        for d in x.shape[3:]:
            inner *= int(d)
        st = c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _abi.check(_abi.lib.rks_axis_apply_chunked(self._h, c_void_p(x.data_ptr()), c_void_p(x.data_ptr()), int(x.shape[1]),
                                                   inner, int(x.shape[0]), int(bool(inverse)), st))
        return x

    def scatter_(self, x: torch.Tensor, out_bases: torch.Tensor, outer: int, inner: int, in_chunks: int, inverse: bool) -> None:
        """Transform whose output rows go to ``out_bases.numel()`` destination blocks (``rks_axis_apply_scatter``):
        ``out_bases`` is an int64 device tensor of block addresses, peer mappings included -- the stores of the last
        level are the exchange of a slab-decomposed grid (dist_fft.py).  ``x``: ``[outer][n][inner]`` (in_chunks = 1)
        or chunk-major ``[in_chunks][outer][n / in_chunks][inner]``."""
        from ctypes import c_void_p
        if x.dtype != torch.complex128 or not x.is_contiguous() or x.numel() != outer * self.n * inner:
            raise ValueError("AxisFFT.scatter_ needs a contiguous complex128 array of outer * n * inner elements")
        if out_bases.dtype != torch.int64 or not out_bases.is_cuda or not out_bases.is_contiguous():
            raise ValueError("AxisFFT.scatter_: out_bases must be a contiguous int64 CUDA tensor")
        st = c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _abi.check(_abi.lib.rks_axis_apply_scatter(self._h, c_void_p(x.data_ptr()), c_void_p(out_bases.data_ptr()), int(outer),
                                                   int(inner), int(in_chunks), int(out_bases.numel()), int(bool(inverse)), st))

    def inverse_(self, x: torch.Tensor, dim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """ifft along ``dim`` (scaled 1/n), rows left in digit-reversed order; in place unless ``out`` is given."""
        return self._apply(x, dim, 1, out)

    def forward_(self, x: torch.Tensor, dim: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fft along ``dim`` of data in the digit-reversed order ``inverse_`` produces."""
        return self._apply(x, dim, 0, out)

    def __del__(self):
        try:
            if self._h:
                _abi.lib.rks_axis_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _pow2_in_range(n: int) -> bool:
    return 16 <= n <= 16384 and n & (n - 1) == 0


def allen_cahn_fourier_ops(n: int, eps: float = 0.01, length: float = 2 * 3.141592653589793, device="cuda"):
    """Periodic 2-D Allen-Cahn u_t = eps lap(u) + u - u^3 on an n x n grid, rfft2 half spectrum:
    L = 1 - eps |k|^2 (float64, shape (n, n/2+1)), N(u^) = -rfft2(irfft2(u^)^3).  The +u goes into L,
    mirroring the split of rkstiff/models.py:240-244."""
    d = length / n
    ky = 2 * 3.141592653589793 * torch.fft.fftfreq(n, d=d, dtype=torch.float64, device=device)
    kx = 2 * 3.141592653589793 * torch.fft.rfftfreq(n, d=d, dtype=torch.float64, device=device)
    lin_op = 1.0 - eps * (kx[None, :] ** 2 + ky[:, None] ** 2)

    rows = RowNL(_abi.MODEL_CUBIC_RFFT, n, None, -1.0, device) if _pow2_in_range(n) else None
    ycol = AxisFFT(n, device) if rows is not None and AxisFFT.supported(n) else None

    def nl_func(uf: torch.Tensor, out=None) -> torch.Tensor:
        if ycol is not None and uf.is_contiguous():
            # all three transforms are the engine's kernels: [inverse over y, strided columns] . [c2r, cube,
            # r2c along x: one fused kernel] . [forward over y]: 3 kernels, 6 passes over the half spectrum
            work = out if out is not None else torch.empty_like(uf)
            ycol.inverse_(uf, -2, out=work)
            rows(work, out=work)
            return ycol.forward_(work, -2)
        if rows is not None:
            # irfft2 / cube / rfft2 = [ifft over y] . [c2r, cube, r2c along x: ONE fused kernel] . [fft over y]
            a = torch.fft.ifft(uf, dim=-2).contiguous()      # (a transform along a non-last axis may return strided)
            return torch.fft.fft(rows(a, out=a), dim=-2, out=out)
        u = torch.fft.irfft2(uf, s=(n, n)).contiguous()
        return torch.fft.rfft2(pointwise_(_abi.MODEL_CUBIC_RFFT, u, -1.0), out=out)      # -(u^3), one kernel

    nl_func.supports_out = True          # the engine lets the transform write N_j in place (no copy pass)
    if ycol is not None:
        return lin_op, FusedGridNL(_abi.MODEL_CUBIC_RFFT, (n, n), -1.0, "allen_cahn_2d", nl_func)
    return lin_op, nl_func


def nls_nd_ops(k_axes: Sequence[torch.Tensor], gamma: float = 2.0):
    """N-D cubic NLS u_t = i lap(u) + i gamma |u|^2 u (demos/nls.ipynb:496-508 is the 2-D case):
    L = -i sum_d k_d^2 on the full fft grid, N(u^) = i gamma fftn(|f|^2 f), f = ifftn(u^)."""
    nd = len(k_axes)
    k2 = 0
    for d, k in enumerate(k_axes):
        shape = [1] * nd
        shape[d] = k.shape[0]
        k2 = k2 + (k.to(torch.float64) ** 2).reshape(shape)
    lin_op = -1j * k2.to(torch.complex128)
    dims = tuple(range(-nd, 0))

    n_last = int(k_axes[-1].shape[0])
    # power-of-two grids: every transform is the engine's own (10 passes per 3-D evaluation where ifftn + pointwise
    # + fftn cost 14; 512^3 ETD35: 85 -> 53 ms per trial); other sizes: torch.fft around the fused last axis (2-D)
    # or around the pointwise kernel
    device = k_axes[-1].device
    sizes = [int(k.shape[0]) for k in k_axes]
    own = _pow2_in_range(n_last) and all(AxisFFT.supported(s) for s in sizes[:-1])
    rows = RowNL(_abi.MODEL_NLS_FFT, n_last, None, gamma, device) if (own or nd == 2) and _pow2_in_range(n_last) else None
    cols = {s: AxisFFT(s, device) for s in set(sizes[:-1])} if own else None
    outer = dims[:-1]

    def nl_func(uf: torch.Tensor, out=None) -> torch.Tensor:
        if cols is not None and uf.is_contiguous() and uf.dim() >= nd:
            # every transform is the engine's: inverse over the strided axes (data stay digit-reversed along
            # them), then ONE fused kernel for the last axis (inverse, i gamma |f|^2 f, forward), then the
            # forward transforms over the strided axes: 2 nd - 1 kernels, each one read + one write
            work = out if out is not None else torch.empty_like(uf)
            src = uf
            for d in range(nd - 1):
                cols[sizes[d]].inverse_(src, d - nd, out=work)
                src = work
            rows(work, out=work)
            for d in range(nd - 2, -1, -1):
                cols[sizes[d]].forward_(work, d - nd)
            return work
        if rows is not None:
            # the innermost axis (inverse transform, i gamma |f|^2 f, forward transform) is ONE fused kernel
            a = torch.fft.ifftn(uf, dim=outer).contiguous()
            return torch.fft.fftn(rows(a, out=a), dim=outer, out=out)
        f = torch.fft.ifftn(uf, dim=dims).contiguous()
        return torch.fft.fftn(pointwise_(_abi.MODEL_NLS_FFT, f, gamma), dim=dims, out=out)   # F{i gamma |f|^2 f}

    nl_func.supports_out = True
    if cols is not None and nd in (2, 3):
        return lin_op, FusedGridNL(_abi.MODEL_NLS_FFT, sizes, gamma, f"nls_{nd}d", nl_func)
    return lin_op, nl_func
