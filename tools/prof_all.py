"""Launch every kernel of a workload's trial ONCE at the bench geometry inside a cudaProfilerStart/Stop range
(for the ncu --set full inventory; not a benchmark).

    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:rks \
        -f -o /tmp/inv_cfg2 python tools/prof_all.py cfg2
    ncu -i /tmp/inv_cfg2.ncu-rep --page raw --csv > gpurun_out/inv_cfg2_raw.csv
    python tools/ncu_summary.py cfg2=gpurun_out/inv_cfg2_raw.csv ... > profiles/r02_ncu_all_kernels.csv
"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import rkstiff_b200 as rk  # noqa: E402
from rkstiff_b200._abi import check, lib  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
size = int(sys.argv[2]) if len(sys.argv) > 2 else None
dev = torch.device("cuda", 0)
rt = torch.cuda.cudart()


def one_trial_by_parts(eng, adaptive, h):
    """coefficients (forced by a new h), N1, every stage kernel and every NL kernel once, the norm kernel."""
    eng.set_h(h)
    eng.update_coeffs()
    eng.nl(1)                     # (adaptive methods: predicated off unless N1 is stale)
    eng.nl(2)                     # the plain evaluation kernel on the stage-value array
    for s in range(1, eng.stages + 1):
        check(lib.rks_stage_nl(eng.plan, s, eng.st))
    if adaptive:
        check(lib.rks_error_sums(eng.plan, eng.st))


if workload in ("cfg2", "cfg2b"):
    # (cfg2b: a quarter of the bench batch -- ncu saves and restores device memory around every replay pass)
    kx, u0 = bench.nls_inputs(torch, bench.B_NLS if workload == "cfg2" else bench.B_NLS // 4, dev)
    lin, nl = rk.models.nls_ops(kx, 2.0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6))
    if workload == "cfg2b":
        sol.evolve_independent(u0, 0.0, 0.02, keep_log=False)
        eng = sol._engine
        eng.begin(0.0, 1e9, 0.01, 0, False)          # every row running again (the warm-up ended at tf)
        eng.set_u(u0)
        eng.run_trials(1)
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        eng.run_trials(1)
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
    else:
        eng = sol._get_engine(u0)
        eng.begin(0.0, 1e9, 0.01, 0, False)
        eng.set_u(u0)
        eng.run_trials(2)
        torch.cuda.synchronize()
        rt.cudaProfilerStart()
        one_trial_by_parts(eng, True, 0.009)
        torch.cuda.synchronize()
        rt.cudaProfilerStop()
elif workload == "cfg3":
    kx, u0 = bench.ks_inputs(torch, bench.B_KS, dev)
    lin, nl = rk.models.ks_ops(kx)
    sol = rk.ETD4(lin, nl)
    eng = sol._get_engine(u0)
    eng.begin(0.0, 0.0, 0.05, 0, True)
    eng.ensure_fixed_coeffs(0.05)
    eng.set_u(u0)
    eng.run_fixed(2)
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    one_trial_by_parts(eng, False, 0.04)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
elif workload == "cfg4":
    n = size or 4096
    lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01, device=dev)
    u0 = bench.allen_cahn_inputs(torch, n, dev)
    sol = rk.IF45DP(lin, nl, config=rk.SolverConfig(epsilon=1e-4))
    sol.evolve(u0, 0.0, 0.01, store_data=False)
    eng = sol._engine
    eng.begin(0.0, 1e9, 0.002, 0, False)
    eng.set_u(u0)
    eng.run_trials(1)
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    one_trial_by_parts(eng, True, 0.0019)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
elif workload == "cfg5":
    # the bench grid is 512^3 (26 GB of plan arrays, minutes of ncu replay per kernel); 512 x 512 x 64 runs the same
    # 512-point strided-axis kernels, indexed-coefficient stage kernels and norm kernel on 1/8 of the memory
    dims = (size or 512, size or 512, 64)
    ks = [2 * math.pi * torch.fft.fftfreq(n, d=12.0 / n, dtype=torch.float64, device=dev) for n in dims]
    xs = [torch.arange(n, dtype=torch.float64, device=dev) * (12.0 / n) - 6.0 for n in dims]
    lin, nl = rk.models.nls_nd_ops(ks, gamma=2.0)
    u0 = torch.fft.fftn(torch.exp(-(xs[0][:, None, None] ** 2 + xs[1][None, :, None] ** 2 + xs[2][None, None, :] ** 2)).to(torch.complex128))
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-5))
    sol.evolve(u0, 0.0, 0.004, store_data=False)
    eng = sol._engine
    eng.begin(0.0, 1e9, 0.002, 0, False)
    eng.set_u(u0)
    eng.run_trials(1)
    torch.cuda.synchronize()
    rt.cudaProfilerStart()
    one_trial_by_parts(eng, True, 0.0019)
    torch.cuda.synchronize()
    rt.cudaProfilerStop()
else:
    raise SystemExit(f"unknown workload {workload}")
print("done", workload, eng.launches())
