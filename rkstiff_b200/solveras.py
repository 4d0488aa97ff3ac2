"""SolverConfig and BaseSolverAS: the adaptive-step driver (rkstiff/solveras.py:27-650) on the
CUDA engine.

The accept/reject state machine, the masked error norms and the evolve() bookkeeping
(`tc += h`, end clamp, snapshot cadence) run on the device (csrc/errctl.cuh); this class
enqueues trials and reads the control block back.  With a fused nonlinearity evolve() syncs
once per chunk of trials, not once per trial; with a torch ``nl_func`` it syncs once per trial
(the callable needs the buffer roles of the previous accept).
"""
from __future__ import annotations

import math
import weakref
from typing import Callable, List, Optional, Tuple, Union

import torch

from . import _abi
from ._engine import Engine
from .solver import BaseSolver
from .util.solver_type import SolverType


class SolverConfig:
    """Adaptive-step parameters with the reference's validation (solveras.py:71-169)."""

    def __init__(self, epsilon: float = 1e-4, incr_f: float = 1.25, decr_f: float = 0.85, safety_f: float = 0.8,
                 adapt_cutoff: float = 0.01, minh: float = 1e-16) -> None:
        self.epsilon = epsilon
        self.incr_f = incr_f
        self.decr_f = decr_f
        self.safety_f = safety_f
        self.adapt_cutoff = adapt_cutoff
        self.minh = minh

    @property
    def epsilon(self) -> float:
        return self._epsilon

    @epsilon.setter
    def epsilon(self, value: float) -> None:
        if value <= 0:
            raise ValueError(f"epsilon must be positive but is {value}")
        self._epsilon = float(value)

    @property
    def incr_f(self) -> float:
        return self._incr_f

    @incr_f.setter
    def incr_f(self, value: float) -> None:
        if value <= 1.0:
            raise ValueError(f"incr_f must be > 1.0 but is {value}")
        self._incr_f = float(value)

    @property
    def decr_f(self) -> float:
        return self._decr_f

    @decr_f.setter
    def decr_f(self, value: float) -> None:
        if value >= 1.0:
            raise ValueError(f"decr_f must be < 1.0 but is {value}")
        self._decr_f = float(value)

    @property
    def safety_f(self) -> float:
        return self._safety_f

    @safety_f.setter
    def safety_f(self, value: float) -> None:
        if value > 1.0:
            raise ValueError(f"safety_f must be <= 1.0 but is {value}")
        self._safety_f = float(value)

    @property
    def adapt_cutoff(self) -> float:
        return self._adapt_cutoff

    @adapt_cutoff.setter
    def adapt_cutoff(self, value: float) -> None:
        if value >= 1.0:
            raise ValueError(f"adapt_cutoff must be < 1.0 but is {value}")
        self._adapt_cutoff = float(value)

    @property
    def minh(self) -> float:
        return self._minh

    @minh.setter
    def minh(self, value: float) -> None:
        if value <= 0:
            raise ValueError(f"minh must be positive but is {value}")
        self._minh = float(value)


class BaseSolverAS(BaseSolver):
    """Adaptive-step solver base: step(), evolve(), exceptions and limits of the reference."""

    class SolverError(RuntimeError):
        pass

    class MaxLoopsExceeded(SolverError):
        pass

    class MinimumStepReached(SolverError):
        pass

    MAX_LOOPS = 50      # enforced on the device (csrc/errctl.cuh), mirrored here for API parity
    MAX_S = 4.0
    MIN_S = 0.25

    #: trials enqueued between two control-block reads on the fused path
    CHUNK = 32
    #: upper bound on the bytes a snapshot ring may take
    RING_BYTES = 2 << 30

    def __init__(self, lin_op, nl_func, config: Optional[SolverConfig] = None,
                 loglevel: Union[str, int] = "WARNING", group=None) -> None:
        super().__init__(lin_op, nl_func, loglevel, group=group)
        # a fresh SolverConfig per solver: the reference's shared default instance leaks state
        self.config = config if config is not None else SolverConfig()
        self.logger.debug("Adaptive configuration: epsilon=%s, incr_f=%s, decr_f=%s, safety_f=%s",
                          self.config.epsilon, self.config.incr_f, self.config.decr_f, self.config.safety_f)
        self._accept = False
        self._stepping = False
        self._last_out = None
        self.trial_log: List[Tuple[float, float, bool, float]] = []   # (h, s, accepted, t_after) of the last run

    def _init_diagonalized(self, matrix: torch.Tensor) -> None:
        """diagonalize=True for a dense square ``lin_op`` (etd35.py:404-424, etd34.py:233-247, if34.py:156-170):
        L = S diag(w) S^-1 with NumPy/LAPACK on the host, exactly as the reference does; the engine then steps
        v = S^-1 u with the diagonal operator w and N'(k) = S^-1 N(S k)."""
        import numpy as np
        mat = matrix.detach().cpu().numpy()
        if mat.ndim != 2 or mat.shape[0] != mat.shape[1]:
            raise ValueError("Cannot diagonalize a 1D system")
        cond = np.linalg.cond(mat)
        if cond > 1e16:
            raise ValueError("Linear operator is non-invertible")
        if cond > 1000:
            self.logger.warning("Linear matrix array has a large condition number of %.2f, method may be unstable", cond)
        eig_vals, s_mat = np.linalg.eig(mat)
        dev = matrix.device
        self._S = torch.from_numpy(np.ascontiguousarray(s_mat.astype(np.complex128))).to(dev)
        self._Sinv = torch.from_numpy(np.ascontiguousarray(np.linalg.inv(s_mat).astype(np.complex128))).to(dev)
        self._eig = torch.from_numpy(np.ascontiguousarray(eig_vals.astype(np.complex128))).to(dev)
        self._diag = False

    @property
    def solver_type(self) -> SolverType:
        return SolverType.ADAPTIVE_STEP

    def reset(self) -> None:
        self.logger.debug("Resetting adaptive solver state")
        self.t, self.u = [], []
        self._accept = False
        self._reset()

    def _reset(self) -> None:
        self._h_coeff = None
        self._stepping = False
        self._last_out = None
        self.trial_log = []

    def _q(self) -> int:
        return 5 if self.METHOD == "IF45DP" else 4

    # -- failure reporting ------------------------------------------------------------------
    def _raise_on_failure(self, status: int) -> None:
        if status == _abi.CTRL_MAX_LOOPS:
            msg = ("Solver failed: adaptive step made too many attempts to find a step size with an "
                   "acceptible amount of error.")
            self.logger.error(msg)
            raise self.MaxLoopsExceeded(msg)
        if status == _abi.CTRL_MIN_STEP:
            msg = "Solver failed: adaptive step reached minimum step size"
            self.logger.error(msg)
            raise self.MinimumStepReached(msg)

    def _drain_log(self, eng, drained: int) -> int:
        """Append the trial records written since `drained` to trial_log; emit the reference's log lines."""
        count = eng.ctrl.log_count - drained
        for r in eng.read_log(drained, count):
            self.trial_log.append((r.h, r.s, bool(r.accepted), r.t_after))
            if not r.accepted and (math.isnan(r.s) or math.isinf(r.s)):
                self.logger.warning("inf or nan number encountered: reducing step size to %s", r.h)
            self.logger.debug("Computed s=%s for h=%s (%s)", r.s, r.h, "accepted" if r.accepted else "rejected")
        return eng.ctrl.log_count

    # -- step() ------------------------------------------------------------------------------
    def step(self, u: torch.Tensor, h_suggest: float) -> Tuple[torch.Tensor, float, float]:
        h = h_suggest
        assert h >= 0.0
        self.logger.debug("Starting step with h_suggest=%s", h_suggest)
        eng = self._get_engine(u)
        if not self._stepping:
            eng.begin(0.0, math.inf, h, 0, True, keep_fsal=False)
            self._stepping = True
            self._drained = 0
            eng.set_u(self._to_eig(u))
        else:
            eng.set_h(h)
            tag = self._last_out
            # identity of the tensor object (a weak reference: a recycled address can never match), unmodified since
            if not (tag is not None and tag[0] is eng and tag[1]() is u and tag[2] == u._version):
                eng.set_u(self._to_eig(u))
        nl = self._callable()
        while True:
            eng.enqueue_trial(nl)
            c = eng.read_ctrl()
            self._drained = self._drain_log(eng, self._drained)
            if c.status == _abi.CTRL_DONE:
                break
            self._raise_on_failure(c.status)
        self._accept = True
        self._h_coeff = c.h_coeff
        out = self._to_phys(eng.get_u())
        self._last_out = (eng, weakref.ref(out), out._version)
        self.logger.debug("Step accepted, returning h=%s, h_suggest=%s", c.h_last, c.h)
        return out, c.h_last, c.h

    # -- independent-dt ensembles (BASELINE cfg 2b) ---------------------------------------------
    def evolve_independent(self, u: torch.Tensor, t0: float, tf: float, h_init: Optional[float] = None,
                           keep_log: bool = True):
        """Evolve every row of ``u`` (shape ``(B, n_c)``) as its own adaptive problem: B reference solvers
        (solveras.py:279-325 per trajectory), each with its own dt sequence, accept/reject decisions and
        coefficient arrays, stepped together by one set of kernel launches.  Needs a fused nonlinearity.

        Returns ``(u_final, logs)`` with ``logs[b] = [(h, s, accepted, t_after), ...]`` of trajectory b
        (empty lists with ``keep_log=False``).  A trajectory that fails raises the reference's exception.
        ``solver.t`` / ``solver.u`` are not filled: rows are at different times in between."""
        if self._fused() is None:
            raise ValueError("evolve_independent needs a fused nonlinearity (rkstiff_b200.models.*_ops)")
        if u.dim() != 2:
            raise ValueError("u must have shape (batch, n_c)")
        self.reset()
        if h_init is None:
            h_init = (tf - t0) / 100.0
        h = h_init
        if t0 + h > tf:
            h = tf - t0
        nrows = int(u.shape[0])
        logs = [[] for _ in range(nrows)]
        if not t0 < tf:
            return u, logs
        key = ("independent",) + tuple(u.shape)
        eng = self._engines.get(key)
        if eng is None:
            eng = Engine(self.METHOD, self.lin_op, u.shape, self._rks_config(), fused=self._fused(), independent=True)
            self._engines = {key: eng}
            self._cfg_sig = self._config_signature()
        elif self._cfg_sig != self._config_signature():
            eng.set_config(self._rks_config())
            self._cfg_sig = self._config_signature()
        self._engine = eng
        eng.begin(t0, tf, h, 0, False, keep_fsal=False)
        eng.set_u(u)
        chunk = min(self.CHUNK, _abi.ROW_LOG_CAP)
        drained = [0] * nrows
        while True:
            eng.run_trials(chunk)
            c = eng.read_ctrl()
            if keep_log:
                rows = eng.read_rows()
                ring = eng.row_logs()
                for b in range(nrows):
                    for i in range(drained[b], rows[b].log_count):
                        r = ring[b, i % _abi.ROW_LOG_CAP]
                        logs[b].append((float(r["h"]), float(r["s"]), bool(r["accepted"]), float(r["t_after"])))
                    drained[b] = rows[b].log_count
            if c.status != _abi.CTRL_RUNNING:
                break
        self._raise_on_failure(c.status)
        self.logger.info("Independent evolution of %d trajectories complete", nrows)
        return eng.get_u(), logs

    # -- evolve() ----------------------------------------------------------------------------
    def evolve(self, u: torch.Tensor, t0: float, tf: float, h_init: Optional[float] = None,
               store_data: bool = True, store_freq: int = 1) -> torch.Tensor:
        self.reset()
        self.logger.info("Starting evolution from t=%s to t=%s", t0, tf)
        if store_data:
            self.t.append(t0)
            self.u.append(u)
        if h_init is None:
            h_init = (tf - t0) / 100.0
        h = h_init
        self.logger.debug("Initial step size h=%s, store_freq=%s", h, store_freq)
        if t0 + h > tf:
            h = tf - t0
        if not t0 < tf:
            return u                                  # loop body never runs (tests/test_etd35.py:124-133)
        eng = self._get_engine(u)
        eng.begin(t0, tf, h, store_freq if store_data else 0, False, keep_fsal=False)
        eng.set_u(self._to_eig(u))
        nl = self._callable()
        fused = nl is None
        chunk = self.CHUNK if fused else 1
        ring = ring_t = None
        if store_data:
            sf = max(1, store_freq)
            need = -(-chunk // sf)                     # most snapshots `chunk` trials can produce
            cap_mem = max(1, self.RING_BYTES // max(1, eng.state_bytes))
            if need > cap_mem:                         # large states: shorten the chunk instead
                need = cap_mem
                chunk = max(1, min(chunk, cap_mem * sf))
            ring = torch.empty((need,) + tuple(eng.u_shape), dtype=torch.complex128, device=eng.device)
            ring_t = torch.empty(need, dtype=torch.float64, device=eng.device)
        drained = snaps = reported = 0
        while True:
            if fused:
                eng.run_trials(chunk, ring, ring_t)
            else:
                eng.enqueue_trial(nl, ring, ring_t)
            c = eng.read_ctrl()
            drained = self._drain_log(eng, drained)
            while c.step_count // 100 > reported:
                reported += 1
                self.logger.info("Progress: t=%.6f/%.6f (%.1f%%), steps>=%d", c.t, tf, 100 * c.t / tf, reported * 100)
            if store_data and c.snap_count > snaps:
                times = ring_t.cpu()
                cap = ring.shape[0]
                for i in range(snaps, c.snap_count):
                    self._store_snapshot(float(times[i % cap]), self._to_phys(ring[i % cap].clone()))
                    self.logger.debug("Stored solution at t=%.6f", self.t[-1])
                snaps = c.snap_count
            if c.status != _abi.CTRL_RUNNING:
                break
        self._raise_on_failure(c.status)
        self._accept = bool(c.accept)
        self._h_coeff = c.h_coeff
        self.sync_snapshots()
        self.logger.info("Evolution complete after %d steps", c.step_count)
        self.logger.info("Stored %d solution snapshots", len(self.u))
        return self._to_phys(eng.get_u())
