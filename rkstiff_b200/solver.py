"""BaseSolver: the API shell shared by every solver class (rkstiff/solver.py:30-213).

Differences from the reference, all extensions:
  * ``lin_op`` and ``u`` are CUDA tensors (complex128 state, float64 or complex128 operator);
  * ``lin_op`` is a *diagonal* operator: it may be ``(n,)`` broadcast over leading batch dimensions of
    ``u`` or have any rank matching the trailing dimensions of ``u`` (N-D grids).  A dense square matrix
    is accepted by IF34 / ETD34 / ETD35 with ``diagonalize=True`` (eigen-decomposition on the host, the
    engine steps the eigenbasis state); the reference's matrix-exponential strategies are out of scope;
  * ``nl_func`` is a torch callable or a fused-kernel handle from :mod:`rkstiff_b200.models`.
"""
from __future__ import annotations

import os
from abc import ABC, abstractmethod
from typing import Callable, Dict, Optional, Union

import torch

from . import _abi
from ._engine import Engine
from .util.loghelper import get_level_name, get_solver_logger, set_log_level


class BaseSolver(ABC):
    METHOD: str = ""

    def __init__(self, lin_op: torch.Tensor, nl_func: Callable[[torch.Tensor], torch.Tensor],
                 loglevel: Union[str, int] = "WARNING", group=None) -> None:
        if not torch.is_tensor(lin_op):
            raise TypeError("lin_op must be a torch tensor on a CUDA device")
        if lin_op.dim() < 1:
            raise ValueError("lin_op must have at least one dimension")
        if not lin_op.is_cuda:
            raise ValueError("lin_op must live on a CUDA device: rkstiff_b200 has no CPU path")
        if lin_op.dtype not in (torch.float64, torch.complex128):
            raise TypeError("lin_op must be float64 or complex128")
        self.lin_op = lin_op
        self.nl_func = nl_func
        self.logger = get_solver_logger(self.__class__, loglevel)
        self.logger.info("Initialized %s solver", self.__class__.__name__)
        self.t, self.u = [], []
        #: where evolve() keeps the snapshots of solver.u: "cuda" (device clones) or "cpu" (pinned host
        #: tensors filled by asynchronous device-to-host copies on a side stream, so that storing multi-GB
        #: states neither stalls the stepping stream nor fills HBM; call solver.sync_snapshots() before use)
        self.snapshot_device = "cuda"
        #: coefficient storage of N-D grids: "auto" (grids of >= Engine.DEDUPE_MIN_MODES modes: per-axis exponential
        #: tables for IF methods with a sum-separable lin_op, else one record per DISTINCT lin_op value), "arrays"
        #: (one entry per mode), "indexed", "separable".  Read when the plan is built.
        self.coef_storage = os.environ.get("RKS_COEF_STORAGE", "auto")
        self._snap_stream = None
        self._diag = True
        # diagonalize=True (dense lin_op): eigen-decomposition done once on the host; the engine steps the
        # eigenbasis state with lin_op = eigenvalues (solveras.py:_init_diagonalized)
        self._S = self._Sinv = self._eig = None
        self._nl_eig = None
        self._group = group
        self._engines: Dict[tuple, Engine] = {}
        self._engine: Optional[Engine] = None
        self.logger.debug("Linear operator shape: %s, diagonal: %s", tuple(lin_op.shape), self._diag)

    # -- engine plumbing ------------------------------------------------------------------
    def _rks_config(self) -> _abi.RksConfig:
        cfg = getattr(self, "config", None)
        etd = getattr(self, "etd_config", None)
        return _abi.RksConfig(
            epsilon=cfg.epsilon if cfg else 1e-4, incr_f=cfg.incr_f if cfg else 1.25,
            decr_f=cfg.decr_f if cfg else 0.85, safety_f=cfg.safety_f if cfg else 0.8,
            adapt_cutoff=cfg.adapt_cutoff if cfg else 0.01, minh=cfg.minh if cfg else 1e-16,
            modecutoff=etd.modecutoff if etd else 0.01, contour_radius=etd.contour_radius if etd else 1.0,
            contour_points=etd.contour_points if etd else 32,
            if45dp_r4_fix=int(bool(getattr(self, "r4_fix", False))))

    def _fused(self):
        from .models import FusedNL
        if self._S is not None:
            return None
        return self.nl_func if isinstance(self.nl_func, FusedNL) else None

    def _callable(self):
        """None when the fused CUDA nonlinearity is used, else the torch callable."""
        if self._S is not None:
            if self._nl_eig is None:
                # N'(k) = S^-1 N(S k)  (etd35.py:463): two dense matrix-vector products around the user's callable
                self._nl_eig = lambda k: self._gemv(self._Sinv, self.nl_func(self._gemv(self._S, k)))
            return self._nl_eig
        return None if self._fused() is not None else self.nl_func

    @staticmethod
    def _gemv(mat: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        """mat @ x for a dense complex128 matrix and a vector (or a batch of row vectors) on the engine's own kernel
        (rks_gemv: one warp per matrix row), on the current stream."""
        from ctypes import c_void_p
        n = mat.shape[0]
        xv = x.to(torch.complex128).contiguous()
        if xv.shape[-1] != n:
            raise ValueError(f"state of {xv.shape[-1]} points does not match the {n} x {n} operator")
        y = torch.empty_like(xv)
        st = c_void_p(torch.cuda.current_stream(mat.device).cuda_stream)
        _abi.check(_abi.lib.rks_gemv(c_void_p(mat.data_ptr()), c_void_p(xv.data_ptr()), c_void_p(y.data_ptr()), n,
                                     xv.numel() // n, st))
        return y

    def _to_eig(self, u: torch.Tensor) -> torch.Tensor:
        return u if self._S is None else self._gemv(self._Sinv, u)

    def _to_phys(self, v: torch.Tensor) -> torch.Tensor:
        return v if self._S is None else self._gemv(self._S, v)

    def _get_engine(self, u: torch.Tensor) -> Engine:
        key = tuple(u.shape)
        eng = self._engines.get(key)
        if (eng is None and self._S is None and self.lin_op.dim() == 2 and u.dim() == 1
                and self.lin_op.shape[0] == self.lin_op.shape[1] == u.shape[0]):
            # the reference reads ANY 2-D lin_op as a dense matrix (solver.py:129-135); here a 2-D lin_op is a
            # diagonal operator on a 2-D grid unless diagonalize=True was given
            raise ValueError("lin_op is a square matrix and u a vector: dense operators need diagonalize=True "
                             "(IF34 / ETD34 / ETD35 only); without it a 2-D lin_op is an elementwise operator "
                             "on a 2-D grid and u must have the same trailing shape")
        if eng is None:
            eng = Engine(self.METHOD, self.lin_op if self._eig is None else self._eig, u.shape, self._rks_config(),
                         fused=self._fused(), group=self._group, coef_storage=self.coef_storage)
            if self._S is not None:
                eng.norm_map = self._to_phys          # the controller's |u+| is the physical one (etd35.py:495)
            self._engines = {key: eng}           # one live plan per solver: a new shape replaces the old
            self._cfg_sig = self._config_signature()
        elif self._cfg_sig != self._config_signature():
            # config objects are read live on every trial by the reference (solveras.py:452-454)
            eng.set_config(self._rks_config())
            self._cfg_sig = self._config_signature()
        self._engine = eng
        return eng

    def close(self) -> None:
        """Release device-side resources tied to a process group: the sharded shared-dt trial is a captured CUDA
        graph that holds NCCL kernels and must be gone before ``torch.distributed.destroy_process_group()``."""
        for eng in self._engines.values():
            eng.close()

    def _config_signature(self):
        c = self._rks_config()
        return tuple(getattr(c, f) for f, _ in c._fields_)

    # -- snapshot pipeline -----------------------------------------------------------------
    def _store_snapshot(self, t: float, u_dev: torch.Tensor) -> None:
        """Append (t, u) to solver.t / solver.u; u_dev is a device tensor the caller will not modify."""
        self.t.append(t)
        if self.snapshot_device == "cuda":
            self.u.append(u_dev)
            return
        if self.snapshot_device != "cpu":
            raise ValueError("snapshot_device must be 'cuda' or 'cpu'")
        if self._snap_stream is None:
            self._snap_stream = torch.cuda.Stream(device=u_dev.device)
        host = torch.empty(u_dev.shape, dtype=u_dev.dtype, pin_memory=True)
        self._snap_stream.wait_stream(torch.cuda.current_stream(u_dev.device))
        with torch.cuda.stream(self._snap_stream):
            host.copy_(u_dev, non_blocking=True)
        u_dev.record_stream(self._snap_stream)            # keep the device clone alive until the copy ran
        self.u.append(host)

    def sync_snapshots(self) -> None:
        """Wait for outstanding snapshot copies (snapshot_device == "cpu")."""
        if self._snap_stream is not None:
            self._snap_stream.synchronize()

    # -- reference API ---------------------------------------------------------------------
    @property
    @abstractmethod
    def solver_type(self):
        """SolverType.CONSTANT_STEP or SolverType.ADAPTIVE_STEP."""

    def set_loglevel(self, loglevel: Union[str, int]) -> None:
        set_log_level(self.logger, loglevel)
        self.logger.info("Log level changed to %s", get_level_name(self.logger.level))

    @abstractmethod
    def reset(self) -> None:
        """Clear stored snapshots and internal stepping state."""

    @abstractmethod
    def _reset(self) -> None:
        """Method-specific part of reset()."""
