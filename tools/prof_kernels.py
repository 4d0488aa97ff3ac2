"""Launch each hot kernel a few times at the bench geometry (for ncu captures; not a benchmark).

    ncu --set full --clock-control none --import-source on -k regex:nl_fast -c 2 -o gpurun_out/prof \
        python tools/prof_kernels.py cfg2
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import rkstiff_b200 as rk  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
if workload == "cfg2":
    kx, u0 = bench.nls_inputs(torch, bench.B_NLS, dev)
    lin, nl = rk.models.nls_ops(kx, 2.0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6))
    eng = sol._get_engine(u0)
    eng.begin(0.0, 1e9, 0.01, 0, False)
    eng.set_u(u0)
    eng.run_trials(2)
else:
    kx, u0 = bench.ks_inputs(torch, bench.B_KS, dev)
    lin, nl = rk.models.ks_ops(kx)
    sol = rk.ETD4(lin, nl)
    eng = sol._get_engine(u0)
    eng.begin(0.0, 0.0, 0.05, 0, True)
    eng.ensure_fixed_coeffs(0.05)
    eng.set_u(u0)
    eng.run_fixed(2)
torch.cuda.synchronize()
from rkstiff_b200._abi import check, lib  # noqa: E402
for _ in range(reps):
    for s in range(1, eng.stages):
        check(lib.rks_stage_nl(eng.plan, s, eng.st))      # fused K1+K4
    eng.nl(2)
    for s in range(1, eng.stages + (0 if workload == "cfg3" else 1)):
        eng.stage(s)
torch.cuda.synchronize()
print("done", eng.launches())
