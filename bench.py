#!/usr/bin/env python
"""Benchmark of the rkstiff_b200 stepping engine (contract: see the task statement / DESIGN.md 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3]

Workloads (BASELINE.json configs):
  cfg2 (default, the configuration the metric is quoted on): NLS 1-D, n = 8192 complex128, ETD35
        adaptive (epsilon 1e-6), a batch of 4096 independent soliton trajectories PER GPU sharing
        one dt.  A "step" is one trial step (accepted or rejected) of the whole batch.
  cfg3: KS 1-D rfft n = 1024, ETD4 fixed step h = 0.05, 65536 trajectories per GPU.
Metric: real-space grid points x RK (trial) steps per second, whole job over all GPUs.
N > 1: one process per GPU under torchrun; the batch is sharded by rank (weak scaling, 4096 or
65536 trajectories per GPU) and, for the adaptive workload, the three error-norm scalars are
all-reduced (MAX, then SUM) over NCCL so that every rank takes the same accept/reject decision.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_NLS, B_NLS = 8192, 4096
N_KS, B_KS = 1024, 65536
METRIC = "rk_step_gridpoints_per_s"
UNIT = "gridpoint*steps/s"

# algorithmic bytes per complex element per trial (SURVEY.md 8d / DESIGN.md 4)
BYTES_PER_ELEM = {"IF4": 368, "ETD4": 400, "IF34": 416, "ETD34": 448, "ETD5": 688, "ETD35": 736, "IF45DP": 816}
# passes (reads + writes of a full state array) of each stage-combine kernel, from the reference formulas
STAGE_PASSES = {"IF4": [3, 3, 3, 6], "IF34": [3, 3, 3, 6], "ETD4": [3, 4, 4, 6], "ETD34": [3, 4, 4, 6],
                "ETD5": [3, 4, 4, 6, 7, 7], "ETD35": [3, 4, 4, 6, 7, 8], "IF45DP": [3, 4, 5, 6, 7, 7]}
NORM_PASSES = {"IF34": 3, "ETD34": 3, "ETD35": 2, "IF45DP": 7}
ADAPTIVE = ("IF34", "ETD34", "ETD35", "IF45DP")
# dram bytes / pipe utilisation of one nl_fast_pre_kernel<16> launch at 4096 x 8192 (profiles/, ncu --set full); None = not captured
NL_PRE_TRAFFIC = 536.95e6 + 480.89e6        # profiles/r01_v8_ncu_full_nl_fast_pre_8192.csv
NL_PRE_CO_BOUNDS = {"source": "profiles/r01_v8_ncu_full_nl_fast_pre_8192.csv", "lsu_wavefronts_pct": 62.3,
                    "fp64_pipe_pct": 53.3, "dram_pct": 39.9,
                    "note": "FP64 butterflies and shared-memory passes bound this kernel, not HBM; the first inverse "
                            "pass runs in the HBM-bound stage kernel instead"}


# ------------------------------------------------------------------------------------------
# synthetic inputs (generated with torch on the device; the oracle gets a slice of the same)
# ------------------------------------------------------------------------------------------
def nls_inputs(torch, batch, device, seed=2):
    n, w = N_NLS, 40.0 * math.pi
    dx = 2 * w / n
    x = torch.arange(n, dtype=torch.float64, device=device) * dx - w
    kx = 2 * math.pi * torch.fft.fftfreq(n, d=dx, dtype=torch.float64, device=device)
    g = torch.Generator(device="cpu").manual_seed(seed)
    eta = (0.5 + torch.rand(batch, 1, generator=g, dtype=torch.float64)).to(device)
    x0 = (-20.0 + 40.0 * torch.rand(batch, 1, generator=g, dtype=torch.float64)).to(device)
    c = (-0.5 + torch.rand(batch, 1, generator=g, dtype=torch.float64)).to(device)
    u0 = eta / torch.cosh(eta * (x[None, :] - x0)) * torch.exp(1j * c * x[None, :])
    return kx, torch.fft.fft(u0, dim=-1)


def ks_inputs(torch, batch, device, seed=0):
    n = N_KS
    dx = 32.0 * math.pi / n
    x = torch.arange(n, dtype=torch.float64, device=device) * dx
    kx = 2 * math.pi * torch.fft.rfftfreq(n, d=dx, dtype=torch.float64, device=device)
    g = torch.Generator(device="cpu").manual_seed(seed)
    phi = (2 * math.pi * torch.rand(batch, 1, generator=g, dtype=torch.float64)).to(device)
    u0 = torch.cos(x[None, :] / 16 + phi) * (1.0 + torch.sin(x[None, :] / 16))
    return kx, torch.fft.rfft(u0, dim=-1)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.tmp.read().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons)}
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle (NumPy port of the reference path), bounded sample
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    """`steps` trial steps of the oracle on `rows` trajectories; with min_secs > 0 it keeps stepping until that
    much time has passed (the cpu_baseline leg: a sample of 10-30 s whatever the host's speed)."""
    workload, rows, seed, steps, warmup, method, min_secs = args
    import numpy as np
    from oracle import problems
    from oracle.rk_oracle import Config, OracleSolver
    if workload == "cfg2":
        p = problems.nls(N_NLS, batch=rows, seed=seed)
        cfg, h, n = Config(epsilon=1e-6), (0.002 if method == "IF45DP" else 0.01), N_NLS
    else:
        p = problems.ks(N_KS, batch=rows, seed=seed)
        cfg, h, n = Config(epsilon=1e-4), 0.05, N_KS
    sol = OracleSolver(method, p.lin_op, p.nl_func, cfg)
    u = p.u0
    if method in ADAPTIVE:
        for _ in range(warmup):
            u, _, h = sol.step(u, h)
        sol.log.clear()
        t0 = time.perf_counter()
        done = 0
        while done < steps or time.perf_counter() - t0 < min_secs:
            u, _, h = sol.step(u, h)
            done += 1
        dt = time.perf_counter() - t0
        trials = len(sol.log)
    else:
        for _ in range(warmup):
            u = sol.step(u, h)
        t0 = time.perf_counter()
        trials = 0
        while trials < steps or time.perf_counter() - t0 < min_secs:
            u = sol.step(u, h)
            trials += 1
        dt = time.perf_counter() - t0
    assert np.isfinite(u).all()
    return trials * rows * n, dt


def cpu_baseline(workload, cores, steps, warmup, rows, method, min_secs=0.0):
    """gp*steps/s of the oracle on `cores` processes, each stepping its own `rows`-trajectory shard."""
    import multiprocessing as mp
    jobs = [(workload, rows, 100 + i, steps, warmup, method, min_secs) for i in range(cores)]
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    work = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return work / wall, wall, work


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port: the reference is
    pure Python/NumPy and cannot travel to the GPU box) on all host cores, same metric and config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    # a step = one trial step of a bounded sample of the workload: `rows` trajectories per process; K and W are
    # honoured as given (K = 40: ~4 s per process for cfg2, ~2 s for cfg3)
    rows = 64 if args.workload == "cfg2" else 1024
    steps, warm = max(1, args.steps), max(0, args.warmup)
    method = args.method or ("ETD35" if args.workload == "cfg2" else "ETD4")
    value, wall, _ = cpu_baseline(args.workload, cores, steps, warm, rows, method)
    sample = (f"{cores} processes x {rows} trajectories x {steps} steps of the oracle "
              f"({'NLS n=8192' if args.workload == 'cfg2' else 'KS n=1024'} {method})")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, args.gpus, method),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(workload, gpus, method=None):
    cfg = _workload_config(workload, gpus)
    if method and method != cfg["method"]:
        cfg["workload"] = cfg["workload"].replace(cfg["method"], method) + " [method overridden]"
        cfg["method"] = method
    return cfg


def _workload_config(workload, gpus):
    if workload == "cfg2":
        return {"workload": "cfg2: NLS 1-D n=8192 complex128, ETD35 adaptive eps=1e-6, 4096 soliton trajectories "
                            "per GPU, one shared dt", "method": "ETD35", "n": N_NLS, "batch_per_gpu": B_NLS,
                "parallelism": f"batch-sharded x{gpus}", "l2": "working set 5.4 GB >> 126 MB L2 (no flush needed)"}
    return {"workload": "cfg3: KS 1-D rfft n=1024, ETD4 fixed step h=0.05, 65536 trajectories per GPU",
            "method": "ETD4", "n": N_KS, "batch_per_gpu": B_KS, "parallelism": f"batch-sharded x{gpus}",
            "l2": "working set 3.2 GB >> 126 MB L2 (no flush needed)"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def parity_sample(torch, rk, workload, method, device):
    """Same small sample through the oracle and through the engine: relative state error after a few steps and,
    for adaptive methods, whether the accept/reject sequence and the dt sequence (1e-9) agree."""
    import numpy as np
    from oracle import problems
    from oracle.rk_oracle import Config, OracleSolver
    rows = 4
    if workload == "cfg2":
        p = problems.nls(N_NLS, batch=rows, seed=11)
        lin, nl = rk.models.nls_ops(torch.from_numpy(p.kx).to(device), 2.0)
        eps, tf, h = 1e-6, 0.1, 0.002
    else:
        p = problems.ks(N_KS, batch=rows, seed=11)
        lin, nl = rk.models.ks_ops(torch.from_numpy(p.kx).to(device))
        eps, tf, h = 1e-4, 1.0, 0.05
    u0 = torch.from_numpy(p.u0).to(device)
    ora = OracleSolver(method, p.lin_op, p.nl_func, Config(epsilon=eps))
    if method in ADAPTIVE:
        if method == "IF45DP":
            tf *= 0.05
        sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=eps))
        uf = sol.evolve(u0, 0.0, tf, store_data=False).cpu().numpy()
        uo = ora.evolve(p.u0, 0.0, tf, store_data=False)
        hs, acc = [r[0] for r in sol.trial_log], [r[2] for r in sol.trial_log]
        same = acc == [r.accepted for r in ora.log] and bool(np.allclose(hs, [r.h for r in ora.log], rtol=1e-9, atol=0))
        return {"rel_err_final": float(np.linalg.norm(uf - uo) / np.linalg.norm(uo)), "trials": len(hs),
                "dt_sequence_matches_oracle": same}
    sol = getattr(rk, method)(lin, nl)
    worst, u = 0.0, u0
    for _ in range(5):
        ref = OracleSolver(method, p.lin_op, p.nl_func).step(u.cpu().numpy(), h)
        sol.reset()
        u = sol.step(u, h)
        worst = max(worst, float(np.linalg.norm(u.cpu().numpy() - ref) / np.linalg.norm(ref)))
    return {"rel_err_per_step_max": worst, "steps": 5}


def time_kernel(torch, fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def run_cfg5(args):
    """3-D NLS, ETD35 adaptive, one grid slab-decomposed over all ranks (BASELINE cfg 5).  Strong scaling:
    the grid is fixed, the ranks split it.  NL = the engine's strided-axis and fused last-axis FFT kernels
    around two NCCL all-to-alls (dist_fft.SlabFFT.fused_nl)."""
    import torch
    import torch.distributed as dist

    import rkstiff_b200 as rk
    from rkstiff_b200.dist_fft import nls_slab_ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD
    n = args.size
    dx = 12.0 / n
    x = torch.arange(n, dtype=torch.float64, device=device) * dx - 6.0
    k = 2 * math.pi * torch.fft.fftfreq(n, d=dx, dtype=torch.float64, device=device)
    lin, nl, fft = nls_slab_ops([k, k, k], gamma=2.0, group=group)
    xs = fft.real_slice(x)
    f0 = torch.exp(-(xs[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2)).to(torch.complex128)
    u0 = fft.forward(f0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-5), group=group)
    sol.evolve(u0, 0.0, 0.02, store_data=False)              # warm-up (plans, NCCL)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sol.evolve(u0, 0.0, 0.2, store_data=False)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    if world > 1:
        t = torch.tensor([secs], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs = float(t.item())
    trials = len(sol.trial_log)
    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": n ** 3 * trials / secs, "unit": UNIT, "n_gpus": world,
                          "steps": trials, "warmup": 0, "ms_per_step": 1e3 * secs / trials, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": {"workload": f"cfg5: 3-D NLS {n}^3 complex128, ETD35 adaptive eps=1e-5, t 0->0.2, "
                                                 f"slab-decomposed FFT (hand-written axis/row FFT kernels + NCCL all-to-all) over {world} GPU(s)",
                                     "method": "ETD35", "n": n, "parallelism": f"slab x{world}"},
                          "accepted_steps": sum(1 for r in sol.trial_log if r[2]),
                          "gpu_launches": sol._engine.launches()}))
    if world > 1:
        dist.destroy_process_group()


def run_cfg4(args):
    """2-D periodic Allen-Cahn, rfft2 half spectrum, IF45DP adaptive with device-side dt control and on-device
    exp() coefficient recompute (BASELINE cfg 4).  Single GPU; the 2-D transform is the engine's own (column FFT
    kernel around the fused c2r-cube-r2c row kernel), K1/K2/K3 from this engine on the (n, n/2+1) 'lin_op shaped like u' layout."""
    import torch

    import rkstiff_b200 as rk

    torch.cuda.set_device(0)
    n = args.size if args.size != 256 else 4096
    lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01)
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.arange(n, dtype=torch.float64, device="cuda") * (2 * math.pi / n)
    u0 = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    for _ in range(16):
        amp = float(torch.randn(1, generator=g))
        m, q = int(torch.randint(-4, 5, (1,), generator=g)), int(torch.randint(-4, 5, (1,), generator=g))
        th = float(torch.rand(1, generator=g)) * 2 * math.pi
        u0 += 0.1 * amp * torch.cos(m * x[None, :] + q * x[:, None] + th)
    uf0 = torch.fft.rfft2(u0)
    sol = rk.IF45DP(lin, nl, config=rk.SolverConfig(epsilon=1e-4))
    sol.evolve(uf0, 0.0, 0.01, store_data=False)          # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sol.evolve(uf0, 0.0, 0.2, store_data=False)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    trials = len(sol.trial_log)
    n_c = n * (n // 2 + 1)
    alg = (16 * 87 + 8 * 29) * n_c * trials            # P = 87 passes + 29 real coefficient reads
    print(json.dumps({"metric": METRIC, "value": n * n * trials / secs, "unit": UNIT, "n_gpus": 1, "steps": trials,
                      "warmup": 0, "ms_per_step": 1e3 * secs / trials, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": f"cfg4: Allen-Cahn 2-D {n}^2 Fourier grid (rfft2 half spectrum), IF45DP adaptive "
                                             "eps=1e-4, t 0->0.2, NL = column FFT kernels around the fused c2r/cube/r2c row kernel", "method": "IF45DP", "n": n},
                      "accepted_steps": sum(1 for r in sol.trial_log if r[2]),
                      "roofline": {"bound": "hbm", "kernel": "whole trial (model of SURVEY 8d)", "achieved": alg / secs / 1e9,
                                   "peak": 6549.8, "unit": "GB/s", "frac": alg / secs / 1e9 / 6549.8, "traffic": None},
                      "gpu_launches": sol._engine.launches()}))


def run_cfg2b(args):
    """cfg 2b: the cfg-2 soliton ensemble with an INDEPENDENT dt per trajectory (one controller, coefficient
    set and role state per row; one set of launches with gridDim.z = trajectory).  Batch sharded over ranks,
    no collective.  value = sum over trajectories of their trial steps x n / time of the slowest rank."""
    import torch
    import torch.distributed as dist

    import rkstiff_b200 as rk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    n, batch = 8192, 4096
    kx, u0 = nls_inputs(torch, batch, device, seed=2 + rank)
    lin, nl = rk.models.nls_ops(kx, 2.0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6))
    sol.evolve_independent(u0, 0.0, 0.02, keep_log=False)          # warm-up: plan, graph capture
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sol.evolve_independent(u0, 0.0, 1.0, keep_log=False)
    e1.record()
    torch.cuda.synchronize()
    secs = e0.elapsed_time(e1) * 1e-3
    rows = sol._engine.read_rows()
    trials = sum(int(r.trial_count) for r in rows)
    steps_max = max(int(r.trial_count) for r in rows)
    updates = sum(int(r.coeff_updates) for r in rows)
    secs_local = secs
    if world > 1:
        t = torch.tensor([secs, float(trials), float(steps_max)], dtype=torch.float64, device=device)
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        secs, trials, steps_max = float(tmax[0]), int(t[1]), int(tmax[2])
    # byte model of a row-trial with per-row coefficients: the 736 B of the shared-dt trial
    # + 23 coefficient reads in K1 (16 B each) + 21 coefficient writes whenever the row's dt changed
    local_trials = sum(int(r.trial_count) for r in rows)
    alg = n * (local_trials * (736 + 23 * 16) + updates * 21 * 16)
    peak = 6549.8
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f).get("hbm_gbs", peak))
    except OSError:
        pass
    if rank == 0:
        print(json.dumps({"metric": METRIC, "value": n * trials / secs, "unit": UNIT, "n_gpus": world,
                          "steps": steps_max, "warmup": 0, "ms_per_step": 1e3 * secs / steps_max,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic",
                          "config": {"workload": "cfg2b: NLS 1-D n=8192 complex128, ETD35 adaptive eps=1e-6, 4096 soliton "
                                                 "trajectories per GPU, independent dt per trajectory, t 0->1",
                                     "method": "ETD35", "n": n, "batch_per_gpu": batch,
                                     "parallelism": f"batch x{world} (no collective)"},
                          "row_trials": trials, "launch_rounds": steps_max,
                          "roofline": {"bound": "hbm", "kernel": "whole run, rank 0 (per-row coefficient byte model)",
                                       "achieved": alg / secs_local / 1e9, "peak": peak, "unit": "GB/s",
                                       "frac": alg / secs_local / 1e9 / peak, "traffic": None},
                          "gpu_launches": sol._engine.launches()}))
    if world > 1:
        dist.destroy_process_group()


def run_ours(args):
    if args.workload == "cfg2b":
        return run_cfg2b(args)
    if args.workload == "cfg5":
        return run_cfg5(args)
    if args.workload == "cfg4":
        return run_cfg4(args)
    import torch
    import torch.distributed as dist

    import rkstiff_b200 as rk

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: rkstiff_b200 has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
        group = dist.group.WORLD

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    peak_gbs = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"

    K, W = args.steps, max(3, args.warmup)
    method = args.method or ("ETD35" if args.workload == "cfg2" else "ETD4")
    adaptive = method in ADAPTIVE
    if args.workload == "cfg2":
        n, batch, n_c = N_NLS, B_NLS, N_NLS
        kx, u0 = nls_inputs(torch, batch, device, seed=2 + rank)
        lin, nl = rk.models.nls_ops(kx, gamma=2.0)
        h0 = 0.002 if method == "IF45DP" else 0.01
    else:
        n, batch, n_c = N_KS, B_KS, N_KS // 2 + 1
        kx, u0 = ks_inputs(torch, batch, device, seed=rank)
        lin, nl = rk.models.ks_ops(kx)
        h0 = 0.05
    cls = getattr(rk, method)
    if adaptive:
        sol = cls(lin, nl, config=rk.SolverConfig(epsilon=1e-6 if args.workload == "cfg2" else 1e-4), group=group)
    else:
        sol = cls(lin, nl, group=group)
    eng = sol._get_engine(u0)

    # ---- device-resident throughput: K steps, inputs already in HBM --------------------------
    if adaptive:
        eng.begin(0.0, 1e9, h0, 0, False)
        eng.set_u(u0)
        run = eng.run_trials
    else:
        eng.begin(0.0, 0.0, h0, 0, True)
        eng.ensure_fixed_coeffs(h0)
        eng.set_u(u0)
        run = eng.run_fixed
    run(W)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(K)
    e1.record()
    barrier()
    secs = max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    launches = eng.launches() - l0
    clocks = sampler.stop() if sampler else None
    if adaptive:
        c = eng.read_ctrl()
        assert c.status == 0 and c.trial_count == K + W, (c.status, c.trial_count)
        accepted = int(c.step_count)
    else:
        accepted = K
        assert torch.isfinite(eng.get_u().real).all()
    value = world * batch * n * K / secs

    # ---- per-kernel roofline (each kernel timed alone on the launching stream) ---------------
    elems = batch * n_c
    reps = 10
    kern = {}
    S = eng.stages
    n_nl = S if (not adaptive or method in ("IF34", "ETD34", "IF45DP")) else S - 1
    if method == "ETD35":
        n_nl = S                                   # 5 stage NLs + N1 = N(u) after an accept
    # the kernels of the timed region, one at a time: rks_stage_nl_part launches exactly what rks_stage_nl(s) does
    # (for the NLS workload the intermediate stages use the pre-transforming pair; RKS_PT=0: plain)
    part = lambda s, which: rk._abi.check(rk._abi.lib.rks_stage_nl_part(eng.plan, s, which, eng.st))
    pt = args.workload == "cfg2" and os.environ.get("RKS_PT", "1")[:1] != "0"
    n_pre = S - 1 if pt else 0
    t_nl = time_kernel(torch, lambda: eng.nl(2), reps)
    kern["nl (K4 fused spectral nonlinearity)"] = (t_nl, 2 * 16 * elems, n_nl - n_pre)
    if pt:
        t_pre = time_kernel(torch, lambda: part(1, 2), reps)
        kern["nl-pre (K4 on rows K1 pre-transformed)"] = (t_pre, 2 * 16 * elems, n_pre)
    for s in range(1, S + 1):
        if not adaptive and s == S:
            continue            # in place: timing it alone would advance u repeatedly
        t = time_kernel(torch, lambda s=s: part(s, 1), reps)
        name = f"stage{s} (K1 combine + first inverse FFT pass)" if pt and s < S else f"stage{s} (K1 combine)"
        kern[name] = (t, STAGE_PASSES[method][s - 1] * 16 * elems, 1)
    if adaptive:
        t = time_kernel(torch, lambda: rk._abi.check(rk._abi.lib.rks_error_sums(eng.plan, eng.st)), reps)
        kern["norm (K3 masked norms)"] = (t, NORM_PASSES[method] * 16 * elems, 1)
    shares = {k: v[0] * v[2] for k, v in kern.items()}
    top = max(shares, key=shares.get)
    t_top, b_top, _ = kern[top]
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the committed
    # ncu --set full captures under profiles/ (same geometry); None when that kernel/geometry was not captured
    traffic = None
    co_bounds = None
    if top.startswith("nl-pre") and batch * n_c == 4096 * 8192:
        traffic = NL_PRE_TRAFFIC
        co_bounds = NL_PRE_CO_BOUNDS
    elif top.startswith("nl") and batch * n_c == 4096 * 8192:
        traffic = 537.16e6 + 480.97e6           # profiles/r01_final2_ncu_full_nl_fast_8192_tma.csv
        # the FFT pair is not HBM-bound: ncu counts, per SM and launch, 424 k LSU wavefront cycles and ~370 k FP64
        # pipe cycles against 234 k cycles of its HBM share (673 k elapsed) -- DESIGN.md section 4
        co_bounds = {"source": "profiles/r01_final2_ncu_full_nl_fast_8192_tma.csv", "lsu_wavefronts_pct": 63.1,
                     "fp64_pipe_pct": 55.0, "dram_pct": 34.8,
                     "note": "FP64 butterflies and shared-memory passes bound this kernel, not HBM"}
    elif top.startswith("nl") and batch * n_c == 65536 * 513:
        traffic = 540.44e6 + 480.53e6           # profiles/r01_v4_nl_fast_cfg3_ncu_raw.csv
        co_bounds = {"source": "profiles/r01_v4_nl_fast_cfg3_ncu_raw.csv", "lsu_wavefronts_pct": 93.0,
                     "note": "rfft models run full-length complex transforms: shared-memory (LSU) bound"}
    roofline = {"bound": "hbm", "kernel": top, "achieved": b_top / t_top / 1e9, "peak": peak_gbs, "unit": "GB/s",
                "frac": b_top / t_top / 1e9 / peak_gbs, "traffic": traffic, "peak_source": peak_src,
                "co_bounds": co_bounds,
                "share_of_step": shares[top] / sum(shares.values()),
                "whole_step": {"algorithmic_bytes_per_elem": BYTES_PER_ELEM[method],
                               "achieved": BYTES_PER_ELEM[method] * elems * K / secs / 1e9,
                               "frac": BYTES_PER_ELEM[method] * elems * K / secs / 1e9 / peak_gbs},
                "kernels": {k: {"us": v[0] * 1e6, "GBps": v[1] / v[0] / 1e9, "frac": v[1] / v[0] / 1e9 / peak_gbs,
                                "launches_per_step": v[2]} for k, v in kern.items()}}

    # ---- end to end through the public API: host buffers in, host buffers out ----------------
    u_host = torch.empty(u0.shape, dtype=u0.dtype, pin_memory=True)
    u_host.copy_(u0)
    out_host = torch.empty(u0.shape, dtype=u0.dtype, pin_memory=True)
    state_bytes = u0.numel() * 16

    def e2e_once():
        ud = u_host.to(device, non_blocking=True)
        if adaptive:
            # IF45DP: the reference's r4 weight makes the error estimate O(h) (~40x more trials): shorter horizon
            horizon = (1.0 if args.workload == "cfg2" else 2.0) * (0.02 if method == "IF45DP" else 1.0)
            uf = sol.evolve(ud, 0.0, horizon, store_data=False)
            steps_done = len(sol.trial_log)
        else:
            tf_e2e = 2.0 if args.workload == "cfg3" else 0.4
            uf = sol.evolve(ud, 0.0, tf_e2e, h0, store_data=False)
            steps_done, tc = 0, 0.0
            while tc < tf_e2e:                   # the reference's float-accumulated loop count
                tc += h0
                steps_done += 1
        out_host.copy_(uf, non_blocking=True)
        torch.cuda.synchronize()
        return steps_done

    e2e_once()
    barrier()
    t0 = time.perf_counter()
    reps_e2e, steps_e2e = 2, 0
    for _ in range(reps_e2e):
        steps_e2e += e2e_once()
    barrier()
    e2e_secs = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * batch * n * steps_e2e / e2e_secs, "unit": UNIT,
           "h2d_bytes_per_step": state_bytes * reps_e2e / steps_e2e, "d2h_bytes_per_step": state_bytes * reps_e2e / steps_e2e,
           "call": f"{method}.evolve(u0, ...) of the workload's horizon with u0 copied from pinned host memory and the "
                   "final state copied back, per call",
           "steps_per_call": steps_e2e / reps_e2e, "h2d_bytes_per_call": state_bytes, "d2h_bytes_per_call": state_bytes}

    # ---- CPU baseline on rank 0, N = 1 only (the oracle also serves as the checker of a small sample) --
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # 12 s of single-core NumPy work (cfg 2 at B=64, cfg 3 at B=1024), at least 24 / 60 steps
        rows, cs, cw = (64, 24, 1) if args.workload == "cfg2" else (1024, 60, 2)
        v, wall, work = cpu_baseline(args.workload, 1, cs, cw, rows, method, min_secs=12.0)
        cs = work // (rows * n)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"oracle (NumPy port of the reference path), 1 process, {rows} trajectories x {cs} steps, "
                         f"{wall:.1f} s",
               "parity": parity_sample(torch, rk, args.workload, method, device)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": workload_config(args.workload, world, method),
                "accepted_steps": accepted, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": launches, "clocks": clocks}
        line["config"]["kernel_pair"] = ("pre-transforming K1/K4 pair on the intermediate stages (DESIGN.md 4)" if pt
                                         else "plain K1 + K4")
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg2b", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--method", default=None, help="override the method of cfg2/cfg3 (IF4 ETD4 ETD5 IF34 ETD34 ETD35 IF45DP)")
    ap.add_argument("--size", type=int, default=256, help="cfg4/cfg5: points per axis of the 2-D/3-D grid (cfg4 default 4096)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
