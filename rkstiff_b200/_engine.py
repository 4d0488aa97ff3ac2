"""Host-side wrapper of one engine plan (include/rkstiff_b200.h) for one (method, lin_op, u-shape).

PyTorch is used for device memory (the plan's workspace is one uint8 tensor), streams and,
for multi-GPU shared-dt ensembles, ``torch.distributed``; every numeric operation of the
stepping path runs in the CUDA library.
"""
from __future__ import annotations

import ctypes
import math
from ctypes import byref, c_double, c_void_p
from typing import Callable, List, Optional, Tuple

import torch

from . import _abi
from ._abi import RksConfig, RksCtrl, RksTrialRec, check, lib


def _stream(device) -> c_void_p:
    return c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _bits(x: torch.Tensor) -> torch.Tensor:
    return x.contiguous().view(torch.int64)


def distinct_values(lin: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The distinct entries of ``lin`` (by bit pattern) and, per element, the int32 index of its entry.
    A Fourier-grid operator has few of them (|k|^2 takes <= 3 (n/2)^2 values on n^3 points)."""
    flat = lin.reshape(-1)
    if flat.dtype == torch.complex128:
        ur, ir = torch.unique(_bits(flat.real), return_inverse=True)
        ui, ii = torch.unique(_bits(flat.imag), return_inverse=True)
        key, inv = torch.unique(ir * ui.numel() + ii, return_inverse=True)
        del ir, ii
        vals = torch.complex(ur[key // ui.numel()].view(torch.float64), ui[key % ui.numel()].view(torch.float64))
    else:
        ub, inv = torch.unique(_bits(flat), return_inverse=True)
        vals = ub.view(torch.float64)
    return vals.contiguous(), inv.to(torch.int32).contiguous()


def separable_terms(lin: torch.Tensor) -> Optional[torch.Tensor]:
    """``lin[i0, i1(, i2)] == a_0[i0] + a_1[i1] (+ a_2[i2])`` up to rounding?  Returns the concatenated per-axis
    terms (the constant folded into a_0) or None.  Fourier-grid operators are of this form: c - eps |k|^2."""
    nd = lin.dim()
    if nd not in (2, 3):
        return None
    origin = lin[(0,) * nd]
    terms, recon, mag = [], None, None
    for d in range(nd):
        idx = [0] * nd
        idx[d] = slice(None)
        a = lin[tuple(idx)].clone()
        if d > 0:
            a = a - origin
        shape = [1] * nd
        shape[d] = lin.shape[d]
        recon = a.reshape(shape) if recon is None else recon + a.reshape(shape)
        mag = a.abs().reshape(shape) if mag is None else mag + a.abs().reshape(shape)
        terms.append(a)
    eps = torch.finfo(torch.float64).eps
    ok = bool(((recon - lin).abs() <= 4 * eps * (mag + lin.abs())).all())
    del recon, mag
    return torch.cat(terms).contiguous() if ok else None


class Engine:
    """One plan: owns the workspace tensor and exposes the strategy-object operations
    (update_coeffs / n1_init / update_stages of the reference) plus the device controller."""

    #: grids with at least this many modes per trajectory get their coefficients by distinct value / per axis
    #: instead of as full-size arrays (coef_storage="auto")
    DEDUPE_MIN_MODES = 1 << 16

    def __init__(self, method: str, lin_op: torch.Tensor, u_shape: torch.Size, cfg: RksConfig,
                 fused=None, group=None, independent: bool = False, coef_storage: str = "auto"):
        if not lin_op.is_cuda:
            raise ValueError("lin_op must be a CUDA tensor: rkstiff_b200 has no CPU path")
        if lin_op.dtype not in (torch.float64, torch.complex128):
            raise TypeError("lin_op must be float64 or complex128")
        nd = lin_op.dim()
        if tuple(u_shape[len(u_shape) - nd:]) != tuple(lin_op.shape):
            raise ValueError(f"lin_op shape {tuple(lin_op.shape)} must match the trailing dims of u {tuple(u_shape)}")
        self.method = method
        self.mid = _abi.METHOD_IDS[method]
        self.device = lin_op.device
        self.u_shape = torch.Size(u_shape)
        self.n_c = lin_op.numel()
        self.batch = 1
        for d in u_shape[: len(u_shape) - nd]:
            self.batch *= int(d)
        self.adaptive = bool(lib.rks_is_adaptive(self.mid))
        self.stages = lib.rks_num_stages(self.mid)
        self.n_nl = lib.rks_num_nl_buffers(self.mid)
        self.fsal = method in ("IF34", "ETD34", "IF45DP")
        self.group = group
        self.fused = fused
        self.independent = bool(independent)
        if self.independent:
            # one controller / dt / coefficient set per trajectory (include/rkstiff_b200.h, cfg 2b)
            if not self.adaptive:
                raise ValueError("independent dt needs an adaptive method")
            if fused is None:
                raise ValueError("independent dt needs a fused nonlinearity (models.*_ops)")
            if group is not None:
                raise ValueError("independent trajectories need no process group: shard u0 instead")
            if nd != 1 or len(u_shape) != 2:
                raise ValueError("independent dt takes u of shape (batch, n_c) and a 1-D lin_op")
        lin = lin_op.contiguous()
        is_cx = lin.dtype == torch.complex128
        self.lin_complex = is_cx      # coefficient arrays are real only for IF methods with a real lin_op
        with torch.cuda.device(self.device):
            if self.independent:
                nbytes = lib.rks_workspace_bytes_independent(self.mid, self.batch, self.n_c, int(is_cx))
            else:
                nbytes = lib.rks_workspace_bytes(self.mid, self.batch, self.n_c, self.n_c, int(is_cx))
            if nbytes == 0:
                raise ValueError("invalid plan geometry")
            self.coef_storage = "arrays"
            extra = self._coef_strategy(lin, coef_storage) if not self.independent else None
            if extra is not None and extra[0] == "separable":
                dims = (ctypes.c_int64 * nd)(*[int(d) for d in lin.shape])
                nbytes = lib.rks_workspace_bytes_separable(self.mid, self.batch, nd, dims, int(is_cx))
            elif extra is not None:
                nbytes = lib.rks_workspace_bytes_indexed(self.mid, self.batch, self.n_c, extra[1].numel(), int(is_cx))
            if nbytes == 0:
                raise ValueError("invalid plan geometry")
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.plan = c_void_p()
            self._cfg = cfg
            if extra is not None and extra[0] == "separable":
                check(lib.rks_plan_create_separable(byref(self.plan), self.mid, self.batch, nd, dims,
                                                    c_void_p(extra[1].data_ptr()), int(is_cx), byref(cfg),
                                                    c_void_p(self.ws.data_ptr()), nbytes, _stream(self.device)))
                self.coef_storage = "separable"
            elif extra is not None:
                check(lib.rks_plan_create_indexed(byref(self.plan), self.mid, self.batch, self.n_c,
                                                  c_void_p(extra[1].data_ptr()), int(is_cx), extra[1].numel(),
                                                  c_void_p(extra[2].data_ptr()), byref(cfg),
                                                  c_void_p(self.ws.data_ptr()), nbytes, _stream(self.device)))
                self.coef_storage = "indexed"
            elif self.independent:
                check(lib.rks_plan_create_independent(byref(self.plan), self.mid, self.batch, self.n_c,
                                                      c_void_p(lin.data_ptr()), int(is_cx), byref(cfg),
                                                      c_void_p(self.ws.data_ptr()), nbytes, _stream(self.device)))
            else:
                check(lib.rks_plan_create(byref(self.plan), self.mid, self.batch, self.n_c, c_void_p(lin.data_ptr()),
                                          int(is_cx), self.n_c, byref(cfg), c_void_p(self.ws.data_ptr()), nbytes,
                                          _stream(self.device)))
            if fused is not None and hasattr(fused, "grid"):
                if tuple(lin.shape) != tuple(fused.grid[:-1]) + (self.n_c // max(1, math.prod(fused.grid[:-1])),):
                    raise ValueError(f"lin_op shape {tuple(lin.shape)} does not match the model's grid {fused.grid}")
                grid = (ctypes.c_int64 * len(fused.grid))(*fused.grid)
                check(lib.rks_set_model_nd(self.plan, fused.model_id, len(fused.grid), grid, float(fused.param),
                                           _stream(self.device)))
            elif fused is not None:
                kx = fused.kx.to(device=self.device, dtype=torch.float64).contiguous() if fused.kx is not None else None
                params = (c_double * 1)(float(fused.param))
                check(lib.rks_set_model(self.plan, fused.model_id, fused.n,
                                        c_void_p(kx.data_ptr()) if kx is not None else None, params, 1,
                                        _stream(self.device)))
        self.state_bytes = self.batch * self.n_c * 16
        self.ctrl = RksCtrl()
        self.ctrl.need_n1 = 1
        self._h_host: Optional[float] = None          # fixed-step: h the coefficient arrays hold
        self._h_set: Optional[float] = None
        self._n1_ready = False
        self._red = None
        #: diagonalize=True: maps the plan's (eigenbasis) u+ to the physical array whose magnitudes drive the
        #: error controller (rks_norm_override); None for diagonal operators
        self.norm_map: Optional[Callable] = None
        self._norm_keep = None
        self._group_graphs = {}                       # sharded shared-dt trial captured with its NCCL all-reduces
        self._group_graph_ok = True
        self._group_graph_launches = 0
        self._replayed_launches = 0

    def _coef_strategy(self, lin: torch.Tensor, want: str):
        """How the coefficient arrays of this plan are stored: None = one entry per mode (1-D rows:
        shared by the batch, a few KB), ("separable", terms) for IF methods whose N-D lin_op is a sum of per-axis
        terms, ("indexed", values, index) when lin_op has few distinct values."""
        if want not in ("auto", "arrays", "indexed", "separable"):
            raise ValueError("coef_storage must be 'auto', 'arrays', 'indexed' or 'separable'")
        if want == "arrays" or (want == "auto" and (lin.dim() < 2 or self.n_c < self.DEDUPE_MIN_MODES)):
            return None
        if want in ("auto", "separable") and self.method in ("IF4", "IF34", "IF45DP") and lin.dim() in (2, 3):
            terms = separable_terms(lin)
            if terms is not None:
                return ("separable", terms)
            if want == "separable":
                raise ValueError("lin_op is not a sum of per-axis terms")
        elif want == "separable":
            raise ValueError("separable coefficient tables need an IF method and a 2-D or 3-D lin_op")
        if self.batch > 65535:
            return None
        values, index = distinct_values(lin)
        if want == "auto" and values.numel() * 4 > self.n_c:
            return None                                   # too few repeats to pay for the gather
        return ("indexed", values, index)

    def close(self) -> None:
        """Release the captured sharded-trial graphs (they hold NCCL kernels: call this, or drop the solver, before
        ``torch.distributed.destroy_process_group()``)."""
        self._group_graphs = {}
        self._group_graph_ok = False

    def __del__(self):
        try:
            self._group_graphs = {}
            if getattr(self, "plan", None):
                lib.rks_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # -- plumbing -------------------------------------------------------------------------
    @property
    def st(self) -> c_void_p:
        return _stream(self.device)

    def _view(self, ptr: int) -> torch.Tensor:
        off = ptr - self.ws.data_ptr()
        return self.ws[off: off + self.state_bytes].view(torch.complex128).view(self.u_shape)

    def array(self, name: str) -> int:
        p = lib.rks_array(self.plan, name.encode())
        if not p:
            raise KeyError(name)
        return p

    def state_view(self, name: str) -> torch.Tensor:
        return self._view(self.array(name))

    def coef_view(self, name: str) -> torch.Tensor:
        real = self.method in ("IF4", "IF34", "IF45DP") and not self.lin_complex
        off = self.array(name) - self.ws.data_ptr()
        nb = self.n_c * (8 if real else 16)
        return self.ws[off: off + nb].view(torch.float64 if real else torch.complex128)

    def set_config(self, cfg: RksConfig) -> None:
        self._cfg = cfg
        check(lib.rks_set_config(self.plan, byref(cfg), self.st))
        self._h_host = None

    def launches_raw(self) -> int:
        return int(lib.rks_kernel_launches(self.plan))

    def launches(self) -> int:
        """Kernel launches of this plan, including those replayed from the captured sharded-trial graph."""
        return self.launches_raw() + self._replayed_launches

    # -- state ----------------------------------------------------------------------------
    def begin(self, t0: float, tf: float, h: float, store_freq: int, step_mode: bool, keep_fsal: bool = False):
        check(lib.rks_begin(self.plan, t0, tf, h, int(store_freq), int(step_mode), int(keep_fsal), self.st))
        self._h_set = h
        if not keep_fsal:
            self._h_host = None
            self._n1_ready = False
            self.ctrl.need_n1 = 1
            self.ctrl.n_sel = 0
            self.ctrl.log_count = 0
            self.ctrl.snap_count = 0

    def set_h(self, h: float) -> None:
        check(lib.rks_set_h(self.plan, h, self.st))
        self._h_set = h

    def _as_state(self, u: torch.Tensor) -> torch.Tensor:
        if tuple(u.shape) != tuple(self.u_shape):
            raise ValueError(f"u has shape {tuple(u.shape)}, plan was built for {tuple(self.u_shape)}")
        if not u.is_cuda:
            raise ValueError("u must be a CUDA tensor")
        if u.dtype != torch.complex128:
            u = u.to(torch.complex128)
        return u.contiguous()

    def set_u(self, u: torch.Tensor) -> None:
        u = self._as_state(u)
        check(lib.rks_set_u(self.plan, c_void_p(u.data_ptr()), self.st))

    def get_u(self) -> torch.Tensor:
        out = torch.empty(self.u_shape, dtype=torch.complex128, device=self.device)
        check(lib.rks_get_u(self.plan, c_void_p(out.data_ptr()), self.st))
        return out

    # -- reference strategy operations ----------------------------------------------------
    def update_coeffs(self) -> None:
        check(lib.rks_update_coeffs(self.plan, self.st))

    def stage(self, s: int) -> None:
        check(lib.rks_stage(self.plan, s, self.st))

    def nl(self, j: int) -> None:
        check(lib.rks_nl(self.plan, j, self.st))

    def nl_apply(self, j: int, nl_func: Callable) -> None:
        """N_j = nl_func(input_j) for a caller-supplied torch callable (roles from the last read_ctrl)."""
        src = self._view(lib.rks_nl_input(self.plan, j))
        dst = self._view(lib.rks_nl_output(self.plan, j))
        if getattr(nl_func, "supports_out", False):
            res = nl_func(src, out=dst)                  # the callable writes N_j itself (no copy pass)
            if res.data_ptr() == dst.data_ptr():
                return
        else:
            res = nl_func(src)
        if not torch.is_tensor(res):
            raise TypeError("nl_func must return a torch tensor")
        dst.copy_(res.reshape(self.u_shape))

    def _red_view(self) -> torch.Tensor:
        if self._red is None:
            off = lib.rks_reduction_scalars(self.plan) - self.ws.data_ptr()
            self._red = self.ws[off: off + 24].view(torch.float64)
        return self._red

    def error_control(self) -> None:
        if self.norm_map is not None:
            uplus = self._view(lib.rks_nl_input(self.plan, self.stages + 1))       # U[1 - u_sel]
            self._norm_keep = self.norm_map(uplus).contiguous()
            check(lib.rks_norm_override(self.plan, c_void_p(self._norm_keep.data_ptr()), self.st))
        if self.group is None:
            check(lib.rks_error_control(self.plan, self.st))
            return
        # shared-dt ensemble sharded by batch: reference-exact global norms (solveras.py:451-454)
        from .dist import allreduce_error_scalars
        allreduce_error_scalars(self._red_view(), self.group,
                                between=lambda: check(lib.rks_error_sums(self.plan, self.st)))
        check(lib.rks_controller(self.plan, self.st))

    def read_ctrl(self) -> RksCtrl:
        check(lib.rks_read_ctrl(self.plan, byref(self.ctrl), self.st))
        return self.ctrl

    def read_log(self, first: int, count: int) -> List[RksTrialRec]:
        if count <= 0:
            return []
        buf = (RksTrialRec * count)()
        check(lib.rks_read_log(self.plan, buf, first, count, self.st))
        return list(buf)

    # -- independent-dt ensembles -----------------------------------------------------------
    def read_rows(self):
        """Control block of every trajectory of an independent-dt plan (syncs)."""
        rows = (RksCtrl * self.batch)()
        check(lib.rks_read_rows(self.plan, rows, self.batch, self.st))
        return rows

    def row_logs(self):
        """Host copy of every row's trial-record ring, as a (batch, ROW_LOG_CAP) structured array (syncs)."""
        import numpy as np
        rec = np.dtype([("h", "<f8"), ("s", "<f8"), ("t_after", "<f8"), ("accepted", "<i4"), ("pad", "<i4")])
        off = self.array("row_logs") - self.ws.data_ptr()
        nb = self.batch * _abi.ROW_LOG_CAP * rec.itemsize
        raw = self.ws[off: off + nb].cpu().numpy()
        return raw.view(rec).reshape(self.batch, _abi.ROW_LOG_CAP)

    # -- whole trials ---------------------------------------------------------------------
    def enqueue_trial(self, nl_func: Optional[Callable], ring=None, ring_t=None) -> None:
        """One adaptive trial.  nl_func None => fused NL kernels (device predicated, no sync needed);
        otherwise the torch callable (roles/need_n1 taken from the last read_ctrl)."""
        S = self.stages
        self.update_coeffs()
        if nl_func is None:
            self.nl(1)
        elif self.ctrl.need_n1:
            self.nl_apply(1, nl_func)
        for s in range(1, S + 1):
            if nl_func is None:
                check(lib.rks_stage_nl(self.plan, s, self.st))      # fused K1+K4 where available
            else:
                self.stage(s)
                if s < S or self.fsal:
                    self.nl_apply(s + 1, nl_func)
        self.error_control()
        if ring is not None:
            check(lib.rks_snapshot(self.plan, c_void_p(ring.data_ptr()), c_void_p(ring_t.data_ptr()), ring.shape[0],
                                   self.st))

    def run_trials(self, k: int, ring=None, ring_t=None) -> None:
        """k fused trials without host sync (kernels after the final accept are predicated off)."""
        if self.group is None:
            check(lib.rks_run_trials(self.plan, k, c_void_p(ring.data_ptr()) if ring is not None else None,
                                     c_void_p(ring_t.data_ptr()) if ring is not None else None,
                                     ring.shape[0] if ring is not None else 0, self.st))
        else:
            self._run_trials_group(k, ring, ring_t)

    def _run_trials_group(self, k: int, ring, ring_t) -> None:
        """Sharded shared-dt ensemble: one trial = ~17 launches plus two NCCL all-reduces of the error scalars.
        The first trial runs eagerly (it creates the communicator and every lazily initialised kernel attribute);
        the trial is then captured ONCE into a CUDA graph -- NCCL collectives included -- and replayed, so the
        per-trial host cost is one graph launch, as on the single-GPU path (rks_run_trials).  RKS_GROUP_GRAPH=0
        or a failed capture keeps the eager loop."""
        import os
        key = (ring.data_ptr(), ring_t.data_ptr(), ring.shape[0]) if ring is not None else None
        graph = self._group_graphs.get(key)
        if graph is None and self._group_graph_ok and os.environ.get("RKS_GROUP_GRAPH", "1")[:1] != "0" and k > 1:
            self.enqueue_trial(None, ring, ring_t)
            k -= 1
            try:
                torch.cuda.synchronize(self.device)
                before = self.launches_raw()
                g = torch.cuda.CUDAGraph()
                # thread_local: the NCCL watchdog thread keeps polling its events while this thread captures
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self.enqueue_trial(None, ring, ring_t)
                self._group_graph_launches = self.launches_raw() - before
                self._replayed_launches -= self._group_graph_launches      # the capture itself launched nothing
                self._group_graphs[key] = graph = (g, ring, ring_t)         # keeps the ring alive with the graph
            except Exception:                                               # noqa: BLE001
                self._group_graph_ok = False
                torch.cuda.synchronize(self.device)
        if graph is not None:
            for _ in range(k):
                graph[0].replay()
            self._replayed_launches += k * self._group_graph_launches
            return
        for _ in range(k):
            self.enqueue_trial(None, ring, ring_t)

    # -- fixed step -----------------------------------------------------------------------
    def ensure_fixed_coeffs(self, h: float) -> None:
        if self._h_set != h:
            self.set_h(h)
        if self._h_host != h:                      # exact float equality, etd4.py:392
            self.update_coeffs()
            self._h_host = h

    def fixed_step(self, nl_func: Optional[Callable]) -> None:
        S = self.stages
        if not self._n1_ready:
            if nl_func is None:
                self.nl(1)
            else:
                self.nl_apply(1, nl_func)
            self._n1_ready = True
        if nl_func is None:
            check(lib.rks_run_fixed(self.plan, 1, self.st))
            return
        for s in range(1, S + 1):
            self.stage(s)
            self.nl_apply(s + 1 if s < S else 1, nl_func)

    def run_fixed(self, nsteps: int) -> None:
        if nsteps <= 0:
            return
        if not self._n1_ready:
            self.nl(1)
            self._n1_ready = True
        check(lib.rks_run_fixed(self.plan, nsteps, self.st))
