"""Time the kernels rks_stage_nl launches, one at a time, at the cfg-2 geometry (tuning aid, not a benchmark).
Usage: [RKS_LIB=path/to/lib.so] [RKS_PT=0] python tools/time_stage_parts.py [method=ETD35] [reps=20]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import rkstiff_b200 as rk  # noqa: E402
from rkstiff_b200._abi import check, lib  # noqa: E402

method = sys.argv[1] if len(sys.argv) > 1 else "ETD35"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
kx, u0 = bench.nls_inputs(torch, bench.B_NLS, dev)
lin, nl = rk.models.nls_ops(kx, 2.0)
adaptive = method in bench.ADAPTIVE
sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=1e-6)) if adaptive else getattr(rk, method)(lin, nl)
eng = sol._get_engine(u0)
if adaptive:
    eng.begin(0.0, 1e9, 0.01, 0, False)
    eng.set_u(u0)
    eng.run_trials(3)
else:
    eng.begin(0.0, 0.0, 0.01, 0, True)
    eng.ensure_fixed_coeffs(0.01)
    eng.set_u(u0)
    eng.run_fixed(3)
elems = u0.numel()
out = [os.environ.get("RKS_LIB", "default") + f" {method} PT={os.environ.get('RKS_PT', '1')}"]
for s in range(1, eng.stages):
    t = bench.time_kernel(torch, lambda: check(lib.rks_stage_nl_part(eng.plan, s, 1, eng.st)), reps)
    passes = bench.STAGE_PASSES[method][s - 1]
    out.append(f"  stage{s}: {t*1e6:7.1f} us {passes*16*elems/t/1e9:6.0f} GB/s")
t = bench.time_kernel(torch, lambda: check(lib.rks_stage_nl_part(eng.plan, 1, 2, eng.st)), reps)
out.append(f"  nl    : {t*1e6:7.1f} us {2*16*elems/t/1e9:6.0f} GB/s")
print("\n".join(out))
