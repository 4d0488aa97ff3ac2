#!/usr/bin/env python
"""Benchmark of the rkstiff_b200 stepping engine (contract: see the task statement / DESIGN.md 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4|cfg5|cfg2b]

Workloads (BASELINE.json configs):
  cfg2 (default, the configuration the metric is quoted on): NLS 1-D, n = 8192 complex128, ETD35
        adaptive (epsilon 1e-6), a batch of 4096 independent soliton trajectories PER GPU sharing
        one dt.  A "step" is one trial step (accepted or rejected) of the whole batch.
  cfg3: KS 1-D rfft n = 1024, ETD4 fixed step h = 0.05, 65536 trajectories per GPU.
  cfg4: Allen-Cahn 2-D 4096^2 (rfft2 half spectrum), IF45DP adaptive, one GPU.
  cfg5: NLS 3-D 512^3, ETD35 adaptive; N = 1: the engine's N-D model; N > 1: slab-decomposed, NCCL all-to-all.
  cfg2b: cfg2 with an independent dt per trajectory.
Metric: real-space grid points x RK (trial) steps per second, whole job over all GPUs.

Without --workload the ONE JSON line is the cfg2 line plus `secondary` (the same measurement of cfg3, cfg2b, cfg4
and cfg5 at N = 1; of cfg3, cfg2b and the slab-decomposed cfg5 at N > 1, each with roofline / clocks and -- except
cfg2b -- e2e and, at N = 1, cpu_baseline) and, at N > 1, `parity` (a small sharded shared-dt ensemble and a 16^3 slab run against the oracle:
the GPU test lease has one GPU, so this is where multi-GPU parity is checked on hardware).

N > 1: one process per GPU under torchrun; the batch is sharded by rank (weak scaling, 4096 or
65536 trajectories per GPU) and, for the adaptive workload, the three error-norm scalars are
all-reduced (MAX, then SUM) over NCCL so that every rank takes the same accept/reject decision.
"""
from __future__ import annotations

import argparse
import csv
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
# N-D evaluations cost one read + one write per extra FFT axis pass.  cfg 4 (2-D real): 3 kernels = 6
# passes per evaluation here (SURVEY counts 8 for separate c2r / r2c) -> IF45DP P = 26 + 6 + 7 + 7*6 = 81 of the
# survey's 87; cfg 5 (3-D): 5 kernels = 10 passes (survey: ~16) -> ETD35 P = 25 + 7 + 2 + 6*10 = 94 of ~130.
# Fractions are reported against the SURVEY figures (the contract) and, separately, against what this design moves.
NL_PASSES_ND = {"cfg4": (6, 8), "cfg5": (10, 16)}          # (this design)
NCU_INVENTORY = os.path.join(ROOT, "profiles", "r02_ncu_all_kernels.csv")


# ------------------------------------------------------------------------------------------
# peaks, ncu inventory
# ------------------------------------------------------------------------------------------
def measured_peak():
    """(GB/s, source) of the HBM roofline denominator: MEASURED_PEAKS.json (driver-written) or the recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
        if "hbm_gbs" in peaks:
            return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, ValueError):
        pass
    return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


_INVENTORY = None


def ncu_row(kernel_substr, grid_hint=None):
    """The row of profiles/r02_ncu_all_kernels.csv (one `ncu --set full` launch per kernel at the bench geometry,
    tools/ncu_summary.py) whose kernel name contains `kernel_substr`; None when that kernel was not captured."""
    global _INVENTORY
    if _INVENTORY is None:
        _INVENTORY = []
        try:
            with open(NCU_INVENTORY, newline="") as f:
                _INVENTORY = list(csv.DictReader(f))
        except OSError:
            pass
    for row in _INVENTORY:
        if kernel_substr in row.get("kernel", "") and (grid_hint is None or row.get("workload") == grid_hint):
            return row
    return None


def traffic_of(kernel_substr, workload):
    """(dram bytes per launch, co-bound pipe utilisations) of a kernel from the committed ncu inventory."""
    row = ncu_row(kernel_substr, workload)
    if row is None:
        return None, None
    try:
        traffic = float(row["dram_read_bytes"]) + float(row["dram_write_bytes"])
        co = {"source": os.path.relpath(NCU_INVENTORY, ROOT), "kernel": row["kernel"], "ncu_us": float(row["us"]),
              "dram_pct": float(row["dram_pct"]), "lsu_wavefronts_pct": float(row["lsu_pct"]),
              "fp64_pipe_pct": float(row["fp64_pct"])}
        return traffic, co
    except (KeyError, ValueError):
        return None, None


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


def allen_cahn_inputs(torch, n, device):
    """SURVEY 8d cfg 4: 16 random low modes, rfft2 half spectrum."""
    g = torch.Generator(device="cpu").manual_seed(1234)
    x = torch.arange(n, dtype=torch.float64, device=device) * (2 * math.pi / n)
    u0 = torch.zeros(n, n, dtype=torch.float64, device=device)
    for _ in range(16):
        amp = float(torch.randn(1, generator=g))
        m, q = int(torch.randint(-4, 5, (1,), generator=g)), int(torch.randint(-4, 5, (1,), generator=g))
        th = float(torch.rand(1, generator=g)) * 2 * math.pi
        u0 += 0.1 * amp * torch.cos(m * x[None, :] + q * x[:, None] + th)
    return torch.fft.rfft2(u0)


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
# CPU arm: the reference's own solver classes where its tree is present (dev container), else the oracle
# (NumPy restatement of the same path, pinned bit-for-bit to the reference: tests/test_oracle_golden.py)
# ------------------------------------------------------------------------------------------
def cpu_kind():
    from oracle.ref_loader import reference_root
    return "reference" if reference_root() else "port"


class _CpuSolver:
    """step()-level driver with a trial counter over either the unmodified reference class or the oracle."""

    def __init__(self, method, lin_op, nl_func, eps):
        from oracle.ref_loader import load_reference
        self.adaptive = method in ADAPTIVE
        self.trials = 0
        ref = load_reference()
        if ref is not None:
            from oracle.make_golden import make_solver
            self.kind = "reference"
            self.sol = make_solver(ref, method, lin_op, nl_func, eps if self.adaptive else None)
            if self.adaptive:
                inner = self.sol._update_stages            # bound-method wrapper: counts trials from the outside

                def counted(u, h):
                    self.trials += 1
                    return inner(u, h)
                self.sol._update_stages = counted
        else:
            from oracle.rk_oracle import Config, OracleSolver
            self.kind = "port"
            self.sol = OracleSolver(method, lin_op, nl_func, Config(epsilon=eps))

    def step(self, u, h):
        if self.adaptive:
            before = len(self.sol.log) if self.kind == "port" else 0
            u, _, h = self.sol.step(u, h)
            if self.kind == "port":
                self.trials += len(self.sol.log) - before
            return u, h
        self.trials += 1
        return self.sol.step(u, h), h


def cpu_problem(workload, rows, seed, size=None):
    """(Problem, epsilon, h0, grid points per trajectory) of the CPU arm's bounded sample."""
    from oracle import problems
    if workload == "cfg2":
        return problems.nls(N_NLS, batch=rows, seed=seed), 1e-6, 0.01, N_NLS
    if workload == "cfg3":
        return problems.ks(N_KS, batch=rows, seed=seed), 1e-4, 0.05, N_KS
    if workload == "cfg4":
        return problems.allen_cahn_2d(size), 1e-4, 0.002, size * size
    if workload == "cfg5":
        return problems.nls_3d(size), 1e-5, 0.002, size ** 3
    raise ValueError(workload)


def _cpu_worker(args):
    """`steps` trial steps on `rows` trajectories; with min_secs > 0 it keeps stepping until that much time has
    passed (the cpu_baseline leg: a sample of 10-30 s whatever the host's speed)."""
    workload, rows, seed, steps, warmup, method, min_secs, size = args
    import numpy as np
    p, eps, h, n = cpu_problem(workload, rows, seed, size)
    if method == "IF45DP" and workload == "cfg2":
        h = 0.002
    sol = _CpuSolver(method, p.lin_op, p.nl_func, eps)
    u = p.u0
    with np.errstate(all="ignore"):
        for _ in range(warmup):
            u, h = sol.step(u, h)
        sol.trials = 0
        t0 = time.perf_counter()
        while sol.trials < steps or time.perf_counter() - t0 < min_secs:
            u, h = sol.step(u, h)
        dt = time.perf_counter() - t0
    assert np.isfinite(u).all()
    return sol.trials * max(1, rows) * n, dt, sol.kind


def cpu_baseline(workload, cores, steps, warmup, rows, method, min_secs=0.0, size=None):
    """gp*steps/s of the CPU arm on `cores` processes, each stepping its own `rows`-trajectory shard."""
    import multiprocessing as mp
    jobs = [(workload, rows, 100 + i, steps, warmup, method, min_secs, size) for i in range(cores)]
    if cores == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            res = pool.map(_cpu_worker, jobs)
    work = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return work / wall, wall, work, res[0][2]


CPU_ROWS = {"cfg2": 64, "cfg3": 1024}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores, same metric and config:
    the unmodified reference classes when /root/reference (or $RKSTIFF_REF) exists, else the oracle port (the
    reference is pure Python/NumPy and does not travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    cores = os.cpu_count() or 1
    workload = args.workload or "cfg2"
    if workload not in CPU_ROWS:
        raise SystemExit("--impl reference times cfg2 or cfg3")
    # a step = one trial step of a bounded sample of the workload: `rows` trajectories per process; K and W are
    # honoured as given (K = 40: ~4 s per process for cfg2, ~2 s for cfg3)
    rows = CPU_ROWS[workload]
    steps, warm = max(1, args.steps), max(0, args.warmup)
    method = args.method or ("ETD35" if workload == "cfg2" else "ETD4")
    value, wall, _, kind = cpu_baseline(workload, cores, steps, warm, rows, method)
    what = "the unmodified reference classes (rkstiff.%s.%s)" % (method.lower(), method) if kind == "reference" \
        else "the oracle (NumPy port of the reference path)"
    sample = (f"{cores} processes x {rows} trajectories x {steps} steps of {what} "
              f"({'NLS n=8192' if workload == 'cfg2' else 'KS n=1024'} {method})")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(workload, args.gpus, method),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
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
    # identical in both arms.  The CPU arm (--impl reference, cpu_baseline) cannot run 4096 x 8192 x K in minutes:
    # it times the bounded sample named in `cpu_arm_sample` and reports the same metric.
    if workload == "cfg2":
        return {"workload": "cfg2: NLS 1-D n=8192 complex128, ETD35 adaptive eps=1e-6, 4096 soliton trajectories "
                            "per GPU, one shared dt", "method": "ETD35", "n": N_NLS, "batch_per_gpu": B_NLS,
                "parallelism": f"batch-sharded x{gpus}", "l2": "working set 5.4 GB >> 126 MB L2 (no flush needed)",
                "cpu_arm_sample": f"{CPU_ROWS['cfg2']} trajectories per host process, same n / method / tolerance"}
    return {"workload": "cfg3: KS 1-D rfft n=1024, ETD4 fixed step h=0.05, 65536 trajectories per GPU",
            "method": "ETD4", "n": N_KS, "batch_per_gpu": B_KS, "parallelism": f"batch-sharded x{gpus}",
            "l2": "working set 3.2 GB >> 126 MB L2 (no flush needed)",
            "cpu_arm_sample": f"{CPU_ROWS['cfg3']} trajectories per host process, same n / method / h"}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class Ctx:
    """One process per GPU; the NCCL group is created once and shared by every workload of the run."""

    def __init__(self):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: rkstiff_b200 has no CPU path")
        torch.cuda.set_device(self.local)
        self.device = torch.device("cuda", self.local)
        self.group = None
        self.cpu_affinity = None
        if self.world > 1 and os.environ.get("RKS_BENCH_NUMA", "1")[:1] != "0":
            # one process per GPU: run on the CPUs next to this rank's GPU, so that the pinned host buffers of the e2e
            # leg are NUMA-local and eight ranks do not pull their shards through one socket's memory
            try:
                import pynvml
                pynvml.nvmlInit()
                handle = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda._parse_visible_devices()[self.local]
                                                           if hasattr(torch.cuda, "_parse_visible_devices") else self.local)
                pynvml.nvmlDeviceSetCpuAffinity(handle)
                self.cpu_affinity = len(os.sched_getaffinity(0))
            except Exception:                                    # noqa: BLE001
                self.cpu_affinity = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.device)
            self.group = dist.group.WORLD
        self.peak, self.peak_src = measured_peak()

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        import torch.distributed as dist
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        if self.world == 1:
            return x
        import torch.distributed as dist
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def release(self):
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()

    def close(self, rc=0):
        if self.world > 1:
            import threading
            import torch.distributed as dist
            self.release()
            # the JSON line is out; a communicator that refuses to shut down must not hold the job hostage
            sys.stdout.flush()
            sys.stderr.flush()
            threading.Timer(60.0, lambda: os._exit(rc)).start()
            self.torch.cuda.synchronize()
            dist.destroy_process_group()
            os._exit(rc)
        if rc:
            raise SystemExit(rc)


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


def parity_grid_sample(torch, rk, workload, device):
    """cfg4 / cfg5 at the SURVEY 8d parity sizes (256^2 / 32^3) against the flattened oracle run."""
    import numpy as np
    from oracle import problems
    from oracle.rk_oracle import Config, OracleSolver
    if workload == "cfg4":
        n, method, eps, tf = 256, "IF45DP", 1e-4, 0.02
        p = problems.allen_cahn_2d(n)
        lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01, device=device)
        shape = (n, n // 2 + 1)
    else:
        n, method, eps, tf = 32, "ETD35", 1e-5, 0.2
        p = problems.nls_3d(n)
        k = torch.from_numpy(p.kx).to(device)
        lin, nl = rk.models.nls_nd_ops([k, k, k], gamma=2.0)
        shape = (n, n, n)
    sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=eps))
    uf = sol.evolve(torch.from_numpy(p.u0.reshape(shape)).to(device), 0.0, tf, store_data=False).cpu().numpy().ravel()
    ora = OracleSolver(method, p.lin_op, p.nl_func, Config(epsilon=eps))
    uo = ora.evolve(p.u0, 0.0, tf, store_data=False)
    hs, acc = [r[0] for r in sol.trial_log], [r[2] for r in sol.trial_log]
    same = acc == [r.accepted for r in ora.log] and bool(np.allclose(hs, [r.h for r in ora.log], rtol=1e-9, atol=0))
    return {"grid": "x".join(str(s) for s in shape), "rel_err_final": float(np.linalg.norm(uf - uo) / np.linalg.norm(uo)),
            "trials": len(hs), "dt_sequence_matches_oracle": same}


def _abi_mid(rk, method):
    return rk._abi.METHOD_IDS[method]


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


def timed_evolve(ctx, fn):
    """Device time of fn() (one evolve call), max over ranks.  Returns (result, seconds, clock sampler): the caller
    stops the sampler after the e2e repetitions of the same call, so that a timed region shorter than the 100 ms
    sampling period (cfg 4: 28 trials x 3 ms) still gets clock samples taken under the same load."""
    torch = ctx.torch
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if ctx.rank == 0 else None
    time.sleep(0.15 if sampler else 0.0)               # nvidia-smi is up before the timed region starts
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    ctx.barrier()
    secs = ctx.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
    return out, secs, sampler


def e2e_evolve(ctx, sol, u0, call, reps=2):
    """The public call with HOST buffers: u0 from pinned host memory, the final state back to pinned host memory,
    both copies inside the timed region (wall clock around barriers, max over ranks).  Returns (secs, trials, bytes)."""
    torch = ctx.torch
    u_host = torch.empty(u0.shape, dtype=u0.dtype, pin_memory=True)
    u_host.copy_(u0)
    out_host = torch.empty(u0.shape, dtype=u0.dtype, pin_memory=True)

    def once():
        ud = u_host.to(ctx.device, non_blocking=True)
        uf = call(ud)
        out_host.copy_(uf, non_blocking=True)
        torch.cuda.synchronize()
        return len(sol.trial_log)

    once()
    ctx.barrier()
    t0 = time.perf_counter()
    trials = 0
    for _ in range(reps):
        trials += once()
    ctx.barrier()
    secs = ctx.max_over_ranks(time.perf_counter() - t0)
    return secs, trials, u0.numel() * u0.element_size() * reps


def run_cfg5(ctx, args, cpu=True):
    """3-D NLS, ETD35 adaptive (BASELINE cfg 5).  One GPU: the engine's N-D model (axis / row FFT kernels launched by
    the engine, trials graph-replayed).  N > 1: one grid slab-decomposed over all ranks -- strong scaling: the grid
    is fixed, the ranks split it; NL = the engine's strided-axis and fused last-axis FFT kernels around two NCCL
    all-to-alls (dist_fft.SlabFFT.fused_nl)."""
    torch = ctx.torch
    import rkstiff_b200 as rk
    from rkstiff_b200.dist_fft import nls_slab_ops

    world, device = ctx.world, ctx.device
    n = args.size or 512
    dx = 12.0 / n
    x = torch.arange(n, dtype=torch.float64, device=device) * dx - 6.0
    k = 2 * math.pi * torch.fft.fftfreq(n, d=dx, dtype=torch.float64, device=device)
    if world > 1:
        lin, nl, fft = nls_slab_ops([k, k, k], gamma=2.0, group=ctx.group)
        xs = fft.real_slice(x)
        f0 = torch.exp(-(xs[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2)).to(torch.complex128)
        u0 = fft.forward(f0)
        del f0
    else:
        lin, nl = rk.models.nls_nd_ops([k, k, k], gamma=2.0)
        u0 = torch.exp(-(x[:, None, None] ** 2 + x[None, :, None] ** 2 + x[None, None, :] ** 2)).to(torch.complex128)
        u0 = torch.fft.fftn(u0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-5), group=ctx.group)
    sol.evolve(u0, 0.0, 0.02, store_data=False)              # warm-up (plans, NCCL, graph capture)
    l0 = sol._engine.launches()
    _, secs, sampler = timed_evolve(ctx, lambda: sol.evolve(u0, 0.0, 0.2, store_data=False))
    trials = len(sol.trial_log)
    launches = sol._engine.launches() - l0
    accepted = sum(1 for r in sol.trial_log if r[2])
    local_elems = u0.numel()
    ours, survey = NL_PASSES_ND["cfg5"]
    p_ours = 25 + 7 + 2 + 6 * ours                       # combine reads + writes (+err) + norm + 6 evaluations
    p_survey = 25 + 7 + 2 + 6 * survey
    alg = 16.0 * p_survey * local_elems * trials
    e_secs, e_trials, e_bytes = e2e_evolve(ctx, sol, u0, lambda ud: sol.evolve(ud, 0.0, 0.2, store_data=False), reps=1)
    clocks = sampler.stop() if sampler else None
    out = {"metric": METRIC, "value": n ** 3 * trials / secs, "unit": UNIT, "n_gpus": world, "steps": trials, "warmup": 0,
           "ms_per_step": 1e3 * secs / trials, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"cfg5: 3-D NLS {n}^3 complex128, ETD35 adaptive eps=1e-5, t 0->0.2, "
                                  + ("engine N-D model on one GPU" if world == 1 else
                                     f"slab-decomposed over {world} GPUs (axis/row FFT kernels + NCCL all-to-all)"),
                      "method": "ETD35", "n": n, "parallelism": f"slab x{world}",
                      "l2": f"working set {10 * 16 * n ** 3 / world / 1e9:.1f} GB per GPU >> 126 MB L2"},
           "accepted_steps": accepted,
           "roofline": {"bound": "hbm", "kernel": "whole trial, per GPU (byte model of SURVEY 8d: P = %d passes)" % p_survey,
                        "achieved": alg / secs / 1e9, "peak": ctx.peak, "unit": "GB/s", "frac": alg / secs / 1e9 / ctx.peak,
                        "traffic": None, "peak_source": ctx.peak_src,
                        "passes_this_design": p_ours,
                        "frac_of_this_designs_bytes": 16.0 * p_ours * local_elems * trials / secs / 1e9 / ctx.peak},
           "e2e": {"value": n ** 3 * e_trials / e_secs, "unit": UNIT, "h2d_bytes_per_step": e_bytes / e_trials,
                   "d2h_bytes_per_step": e_bytes / e_trials,
                   "call": "ETD35.evolve(u0, 0, 0.2) with this rank's spectral block copied from pinned host memory and "
                           "the final block copied back"},
           "cpu_baseline": None, "gpu_launches": launches, "clocks": clocks}
    del sol, lin, nl, u0
    ctx.release()
    if cpu and ctx.rank == 0 and world == 1:
        v, wall, work, kind = cpu_baseline("cfg5", 1, 4, 1, 0, "ETD35", min_secs=8.0, size=64)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                               "sample": f"64^3 grid (SURVEY 8d), flattened like demos/nls.ipynb:496-511, "
                                         f"{work // 64 ** 3} trials, {wall:.1f} s",
                               "parity": parity_grid_sample(torch, rk, "cfg5", ctx.device)}
    return out


def run_cfg4(ctx, args, cpu=True):
    """2-D periodic Allen-Cahn, rfft2 half spectrum, IF45DP adaptive with device-side dt control and on-device
    exp() coefficient recompute (BASELINE cfg 4).  Single GPU (N > 1: every rank runs its own replica of the grid);
    the 2-D transform is the engine's own (column FFT kernels around the fused c2r-cube-r2c row kernel)."""
    torch = ctx.torch
    import rkstiff_b200 as rk

    n = args.size or 4096
    lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01, device=ctx.device)
    uf0 = allen_cahn_inputs(torch, n, ctx.device)
    sol = rk.IF45DP(lin, nl, config=rk.SolverConfig(epsilon=1e-4))
    sol.evolve(uf0, 0.0, 0.01, store_data=False)          # warm-up
    l0 = sol._engine.launches()
    _, secs, sampler = timed_evolve(ctx, lambda: sol.evolve(uf0, 0.0, 0.2, store_data=False))
    trials = len(sol.trial_log)
    launches = sol._engine.launches() - l0
    n_c = n * (n // 2 + 1)
    ours, survey = NL_PASSES_ND["cfg4"]
    p_survey = 26 + 6 + 7 + 7 * survey                  # P = 87 (+ 29 real coefficient reads if full size)
    p_ours = 26 + 6 + 7 + 7 * ours
    coef_storage = sol._engine.coef_storage
    alg = (16.0 * p_survey + (8 * 29 if coef_storage == "arrays" else 0)) * n_c * trials
    e_secs, e_trials, e_bytes = e2e_evolve(ctx, sol, uf0, lambda ud: sol.evolve(ud, 0.0, 0.2, store_data=False))
    clocks = sampler.stop() if sampler else None
    out = {"metric": METRIC, "value": ctx.world * n * n * trials / secs, "unit": UNIT, "n_gpus": ctx.world, "steps": trials,
           "warmup": 0, "ms_per_step": 1e3 * secs / trials, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"cfg4: Allen-Cahn 2-D {n}^2 Fourier grid (rfft2 half spectrum), IF45DP adaptive "
                                  "eps=1e-4, t 0->0.2, NL = column FFT kernels around the fused c2r/cube/r2c row kernel",
                      "method": "IF45DP", "n": n, "coefficients": coef_storage,
                      "l2": f"working set {12 * 16 * n_c / 1e9:.1f} GB >> 126 MB L2"},
           "accepted_steps": sum(1 for r in sol.trial_log if r[2]),
           "roofline": {"bound": "hbm", "kernel": "whole trial (byte model of SURVEY 8d: P = %d passes)" % p_survey,
                        "achieved": alg / secs / 1e9, "peak": ctx.peak, "unit": "GB/s", "frac": alg / secs / 1e9 / ctx.peak,
                        "traffic": None, "peak_source": ctx.peak_src, "passes_this_design": p_ours,
                        "frac_of_this_designs_bytes": 16.0 * p_ours * n_c * trials / secs / 1e9 / ctx.peak},
           "e2e": {"value": ctx.world * n * n * e_trials / e_secs, "unit": UNIT, "h2d_bytes_per_step": e_bytes / e_trials,
                   "d2h_bytes_per_step": e_bytes / e_trials,
                   "call": "IF45DP.evolve(u0, 0, 0.2) with the half spectrum copied from pinned host memory and back"},
           "cpu_baseline": None, "gpu_launches": launches, "clocks": clocks}
    del sol, lin, nl, uf0
    ctx.release()
    if cpu and ctx.rank == 0 and ctx.world == 1:
        v, wall, work, kind = cpu_baseline("cfg4", 1, 6, 1, 0, "IF45DP", min_secs=8.0, size=512)
        out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                               "sample": f"512^2 grid (SURVEY 8d), flattened, {work // 512 ** 2} trials, {wall:.1f} s",
                               "parity": parity_grid_sample(torch, rk, "cfg4", ctx.device)}
    return out


def run_cfg2b(ctx, args):
    """cfg 2b: the cfg-2 soliton ensemble with an INDEPENDENT dt per trajectory (one controller, coefficient
    set and role state per row; one set of launches with gridDim.z = trajectory).  Batch sharded over ranks,
    no collective.  value = sum over trajectories of their trial steps x n / time of the slowest rank."""
    torch = ctx.torch
    import rkstiff_b200 as rk

    world, rank, device = ctx.world, ctx.rank, ctx.device
    n, batch = 8192, 4096
    kx, u0 = nls_inputs(torch, batch, device, seed=2 + rank)
    lin, nl = rk.models.nls_ops(kx, 2.0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6))
    sol.evolve_independent(u0, 0.0, 0.02, keep_log=False)          # warm-up: plan, graph capture
    _, secs_max, sampler = timed_evolve(ctx, lambda: sol.evolve_independent(u0, 0.0, 1.0, keep_log=False))
    clocks = sampler.stop() if sampler else None
    rows = sol._engine.read_rows()
    local_trials = sum(int(r.trial_count) for r in rows)
    steps_max = int(ctx.max_over_ranks(float(max(int(r.trial_count) for r in rows))))
    updates = sum(int(r.coeff_updates) for r in rows)
    trials = int(ctx.sum_over_ranks(float(local_trials)))
    # byte model of a row-trial with per-row coefficients: the 736 B of the shared-dt trial
    # + 23 coefficient reads in K1 (16 B each) + 21 coefficient writes whenever the row's dt changed
    alg = n * (local_trials * (736 + 23 * 16) + updates * 21 * 16)
    return {"metric": METRIC, "value": n * trials / secs_max, "unit": UNIT, "n_gpus": world,
            "steps": steps_max, "warmup": 0, "ms_per_step": 1e3 * secs_max / steps_max,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "cfg2b: NLS 1-D n=8192 complex128, ETD35 adaptive eps=1e-6, 4096 soliton "
                                   "trajectories per GPU, independent dt per trajectory, t 0->1",
                       "method": "ETD35", "n": n, "batch_per_gpu": batch, "parallelism": f"batch x{world} (no collective)"},
            "row_trials": trials, "launch_rounds": steps_max,
            "roofline": {"bound": "hbm", "kernel": "whole run, rank 0 (per-row coefficient byte model)",
                         "achieved": alg / secs_max / 1e9, "peak": ctx.peak, "unit": "GB/s",
                         "frac": alg / secs_max / 1e9 / ctx.peak, "traffic": None, "peak_source": ctx.peak_src},
            "gpu_launches": sol._engine.launches(), "clocks": clocks}


def _release_after(ctx, fn):
    try:
        return fn()
    finally:
        ctx.release()


def run_1d(ctx, args, workload, cpu=True):
    """cfg2 / cfg3: device-resident K timed steps, the per-kernel roofline, e2e through evolve(), CPU baseline."""
    torch = ctx.torch
    import rkstiff_b200 as rk

    world, rank, device, group = ctx.world, ctx.rank, ctx.device, ctx.group
    peak_gbs, peak_src = ctx.peak, ctx.peak_src
    K, W = args.steps, max(3, args.warmup)
    method = args.method or ("ETD35" if workload == "cfg2" else "ETD4")
    adaptive = method in ADAPTIVE
    if workload == "cfg2":
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
        sol = cls(lin, nl, config=rk.SolverConfig(epsilon=1e-6 if workload == "cfg2" else 1e-4), group=group)
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
    ctx.barrier()
    sampler = ClockSampler(ctx.local) if rank == 0 else None
    l0 = eng.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(K)
    e1.record()
    ctx.barrier()
    secs = ctx.max_over_ranks(e0.elapsed_time(e1) * 1e-3)
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
    pt = workload == "cfg2" and os.environ.get("RKS_PT", "1")[:1] != "0"
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
    # dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel and its pipe utilisations,
    # looked up by kernel name in the committed ncu --set full inventory of this workload's geometry
    if top.startswith("nl-pre"):
        sass_names = ["nl_fast_pre_kernel"]
    elif top.startswith("nl"):
        sass_names = ["nl_fast_real_kernel", "nl_fast_kernel<"] if workload == "cfg3" else ["nl_fast_kernel<"]
    elif top.startswith("norm"):
        sass_names = ["norm_kernel"]
    else:
        s_top = top[5]
        sass_names = [f"stage_pre_kernel<{_abi_mid(rk, method)}, {s_top}", f"stage_kernel<{_abi_mid(rk, method)}, {s_top}"]
    traffic, co_bounds = None, None
    for name in sass_names:
        traffic, co_bounds = traffic_of(name, workload)
        if traffic is not None:
            break
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
            horizon = (1.0 if workload == "cfg2" else 2.0) * (0.02 if method == "IF45DP" else 1.0)
            uf = sol.evolve(ud, 0.0, horizon, store_data=False)
            steps_done = len(sol.trial_log)
        else:
            tf_e2e = 2.0 if workload == "cfg3" else 0.4
            uf = sol.evolve(ud, 0.0, tf_e2e, h0, store_data=False)
            steps_done, tc = 0, 0.0
            while tc < tf_e2e:                   # the reference's float-accumulated loop count
                tc += h0
                steps_done += 1
        out_host.copy_(uf, non_blocking=True)
        torch.cuda.synchronize()
        return steps_done

    e2e_once()
    ctx.barrier()
    t0 = time.perf_counter()
    reps_e2e, steps_e2e = 2, 0
    for _ in range(reps_e2e):
        steps_e2e += e2e_once()
    ctx.barrier()
    e2e_secs = ctx.max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * batch * n * steps_e2e / e2e_secs, "unit": UNIT,
           "h2d_bytes_per_step": state_bytes * reps_e2e / steps_e2e, "d2h_bytes_per_step": state_bytes * reps_e2e / steps_e2e,
           "call": f"{method}.evolve(u0, ...) of the workload's horizon with u0 copied from pinned host memory and the "
                   "final state copied back, per call",
           "steps_per_call": steps_e2e / reps_e2e, "h2d_bytes_per_call": state_bytes, "d2h_bytes_per_call": state_bytes,
           "cpus_bound_to_this_gpu": ctx.cpu_affinity}
    del sol, eng, u0, u_host, out_host, lin, nl
    ctx.release()

    # ---- CPU baseline on rank 0, N = 1 only (the oracle also serves as the checker of a small sample) --
    cpu_line = None
    if cpu and rank == 0 and world == 1 and not args.no_cpu_baseline:
        # 12 s of single-core NumPy work (cfg 2 at B=64, cfg 3 at B=1024), at least 24 / 60 steps
        rows, cs, cw = (CPU_ROWS["cfg2"], 24, 1) if workload == "cfg2" else (CPU_ROWS["cfg3"], 60, 2)
        v, wall, work, kind = cpu_baseline(workload, 1, cs, cw, rows, method, min_secs=12.0)
        cs = work // (rows * n)
        what = "unmodified reference classes" if kind == "reference" else "oracle (NumPy port of the reference path)"
        cpu_line = {"value": v, "unit": UNIT, "cores": 1, "kind": kind,
                    "sample": f"{what}, 1 process, {rows} trajectories x {cs} steps, {wall:.1f} s",
                    "parity": parity_sample(torch, rk, workload, method, device)}

    return {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * secs / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(workload, world, method),
            "accepted_steps": accepted, "roofline": roofline, "cpu_baseline": cpu_line, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
            "kernel_pair": ("pre-transforming K1/K4 pair on the intermediate stages (DESIGN.md 4)" if pt
                            else "plain K1 + K4")}


def parity_multi(ctx):
    """N > 1: multi-GPU parity on hardware (tests/test_gpu_multi.py needs two GPUs and the GPU test lease has one).
    (i) an NLS ensemble of 2 x world rows, ETD35 eps 1e-6, sharded over the ranks with ONE shared dt (MAX / SUM
    all-reduce of the three error-norm scalars) against the oracle's run of the whole batch (the reference's global
    norms, solveras.py:451-454); (ii) the 16^3 3-D NLS slab-decomposed over all ranks against the oracle's flattened
    run.  Bars: accept flags equal, dt within 1e-9, final state within 1e-9, dt BIT-identical across ranks."""
    import numpy as np
    import torch.distributed as dist
    torch = ctx.torch
    import rkstiff_b200 as rk
    from oracle import problems
    from oracle.rk_oracle import Config, OracleSolver
    from rkstiff_b200.dist import shard_bounds
    from rkstiff_b200.dist_fft import nls_slab_ops

    world, rank, device = ctx.world, ctx.rank, ctx.device
    out = {}

    def gather(obj):
        res = [None] * world
        dist.all_gather_object(res, obj)
        return res

    # (i) sharded shared-dt ensemble
    batch = 2 * world
    p = problems.nls(512, batch=batch, seed=2, half_width=20.0)
    lo, hi = shard_bounds(batch, rank, world)
    lin, nl = rk.models.nls_ops(torch.from_numpy(p.kx).to(device), 2.0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6), group=ctx.group)
    uf = sol.evolve(torch.from_numpy(p.u0[lo:hi].copy()).to(device), 0.0, 0.2, store_data=False).cpu().numpy()
    ora = OracleSolver("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-6))
    uo = ora.evolve(p.u0, 0.0, 0.2, store_data=False)
    hs, acc = [r[0] for r in sol.trial_log], [r[2] for r in sol.trial_log]
    mine = {"accept_equal": acc == [r.accepted for r in ora.log],
            "dt_within_1e-9": len(hs) == len(ora.log) and bool(np.allclose(hs, [r.h for r in ora.log], rtol=1e-9, atol=0)),
            "rel_err_final": float(np.linalg.norm(uf - uo[lo:hi]) / np.linalg.norm(uo[lo:hi])),
            "dt_bits": np.asarray(hs, dtype=np.float64).tobytes().hex()}
    allr = gather(mine)
    out["sharded_shared_dt"] = {
        "sample": f"NLS n=512, {batch} rows over {world} ranks, ETD35 eps=1e-6, t 0->0.2, {len(hs)} trials",
        "accept_flags_equal": all(r["accept_equal"] for r in allr),
        "dt_within_1e-9": all(r["dt_within_1e-9"] for r in allr),
        "rel_err_final_max": max(r["rel_err_final"] for r in allr),
        "dt_bit_identical_across_ranks": all(r["dt_bits"] == allr[0]["dt_bits"] for r in allr)}
    out["sharded_shared_dt"]["ok"] = bool(out["sharded_shared_dt"]["accept_flags_equal"]
                                          and out["sharded_shared_dt"]["dt_within_1e-9"]
                                          and out["sharded_shared_dt"]["rel_err_final_max"] < 1e-9
                                          and out["sharded_shared_dt"]["dt_bit_identical_across_ranks"])
    del sol

    # (ii) slab-decomposed 3-D NLS
    n = 16
    p = problems.nls_3d(n)
    k = torch.from_numpy(p.kx).to(device)
    lin, nl, fft = nls_slab_ops([k, k, k], gamma=2.0, group=ctx.group)
    u0 = fft.spec_slice(torch.from_numpy(p.u0.reshape(n, n, n)).to(device))
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-5), group=ctx.group)
    uf = sol.evolve(u0, 0.0, 0.2, store_data=False).cpu().numpy()
    ora = OracleSolver("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-5))
    uo = ora.evolve(p.u0, 0.0, 0.2, store_data=False).reshape(n, n, n)
    m = n // world
    want = uo[:, rank * m:(rank + 1) * m]
    hs, acc = [r[0] for r in sol.trial_log], [r[2] for r in sol.trial_log]
    mine = {"accept_equal": acc == [r.accepted for r in ora.log],
            "dt_within_1e-9": len(hs) == len(ora.log) and bool(np.allclose(hs, [r.h for r in ora.log], rtol=1e-9, atol=0)),
            "rel_err_final": float(np.linalg.norm(uf - want) / np.linalg.norm(want)),
            "dt_bits": np.asarray(hs, dtype=np.float64).tobytes().hex()}
    allr = gather(mine)
    out["slab_16cubed"] = {
        "sample": f"3-D NLS 16^3 over {world} ranks (axis 1 sharded, NCCL all-to-all), ETD35 eps=1e-5, t 0->0.2, {len(hs)} trials",
        "accept_flags_equal": all(r["accept_equal"] for r in allr),
        "dt_within_1e-9": all(r["dt_within_1e-9"] for r in allr),
        "rel_err_final_max": max(r["rel_err_final"] for r in allr),
        "dt_bit_identical_across_ranks": all(r["dt_bits"] == allr[0]["dt_bits"] for r in allr)}
    out["slab_16cubed"]["ok"] = bool(out["slab_16cubed"]["accept_flags_equal"] and out["slab_16cubed"]["dt_within_1e-9"]
                                     and out["slab_16cubed"]["rel_err_final_max"] < 1e-9
                                     and out["slab_16cubed"]["dt_bit_identical_across_ranks"])
    del sol
    ctx.release()
    return out


#: seconds a secondary workload may take before the line is emitted without it
SECONDARY_LIMIT = 150.0


def guarded(ctx, name, fn, line=None, slot=None):
    """A secondary workload must never take the primary line down with it: an exception becomes an `error` entry, and
    a workload that does not come back within SECONDARY_LIMIT (a rank stuck in a collective cannot be interrupted)
    makes every rank leave -- rank 0 prints the line with what has been measured so far first."""
    import threading

    def bail():
        if line is not None and ctx.rank == 0:
            if slot is not None:
                slot[name] = {"error": f"{name}: no result within {SECONDARY_LIMIT:.0f} s"}
            print(json.dumps(line), flush=True)
        sys.stderr.write(f"bench.py: secondary workload {name} timed out on rank {ctx.rank}\n")
        sys.stderr.flush()
        os._exit(0)

    timer = threading.Timer(SECONDARY_LIMIT, bail) if line is not None else None
    if timer:
        timer.daemon = True
        timer.start()
    try:
        return fn()
    except Exception as exc:                                    # noqa: BLE001
        import traceback
        traceback.print_exc(file=sys.stderr)
        ctx.release()
        return {"error": f"{name}: {type(exc).__name__}: {exc}"}
    finally:
        if timer:
            timer.cancel()


def run_ours(args):
    ctx = Ctx()
    rc = 1
    try:
        if args.workload == "cfg2b":
            line = run_cfg2b(ctx, args)
        elif args.workload == "cfg5":
            line = run_cfg5(ctx, args, cpu=not args.no_cpu_baseline)
        elif args.workload == "cfg4":
            line = run_cfg4(ctx, args, cpu=not args.no_cpu_baseline)
        elif args.workload in ("cfg2", "cfg3"):
            line = run_1d(ctx, args, args.workload)
        else:
            # the default run: the headline workload plus the other BASELINE configs in the same line
            line = run_1d(ctx, args, "cfg2")
            if not args.no_secondary:
                cpu = not args.no_cpu_baseline
                sub = argparse.Namespace(**vars(args))
                sub.method, sub.size = None, None
                sec = {}
                line["secondary"] = sec
                sec["cfg3"] = guarded(ctx, "cfg3", lambda: run_1d(ctx, sub, "cfg3", cpu=cpu), line, sec)
                sec["cfg2b"] = guarded(ctx, "cfg2b", lambda: _release_after(ctx, lambda: run_cfg2b(ctx, sub)), line, sec)
                if ctx.world == 1:
                    sec["cfg4"] = guarded(ctx, "cfg4", lambda: run_cfg4(ctx, sub, cpu=cpu), line, sec)
                    sec["cfg5"] = guarded(ctx, "cfg5", lambda: run_cfg5(ctx, sub, cpu=cpu), line, sec)
                else:
                    line["parity"] = guarded(ctx, "parity", lambda: parity_multi(ctx), line, line)
                    sec["cfg5"] = guarded(ctx, "cfg5", lambda: run_cfg5(ctx, sub, cpu=False), line, sec)
        if ctx.rank == 0:
            print(json.dumps(line), flush=True)
        rc = 0
    except Exception:                                           # noqa: BLE001
        import traceback
        traceback.print_exc()
    finally:
        ctx.close(rc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["cfg2", "cfg2b", "cfg3", "cfg4", "cfg5"],
                    help="one workload only (default: cfg2 with cfg3/cfg4/cfg5 under `secondary`)")
    ap.add_argument("--method", default=None, help="override the method of cfg2/cfg3 (IF4 ETD4 ETD5 IF34 ETD34 ETD35 IF45DP)")
    ap.add_argument("--size", type=int, default=None, help="cfg4/cfg5: points per axis of the 2-D/3-D grid (defaults 4096 / 512)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="default run: the cfg2 line only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
