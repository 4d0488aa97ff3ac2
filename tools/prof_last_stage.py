"""Launch the last stage kernel of a method at the cfg-2 geometry a few times (for ncu captures).
Usage: ncu ... python tools/prof_last_stage.py [method=IF45DP]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import rkstiff_b200 as rk  # noqa: E402

method = sys.argv[1] if len(sys.argv) > 1 else "IF45DP"
dev = torch.device("cuda", 0)
kx, u0 = bench.nls_inputs(torch, bench.B_NLS, dev)
lin, nl = rk.models.nls_ops(kx, 2.0)
sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=1e-6))
eng = sol._get_engine(u0)
eng.begin(0.0, 1e9, 0.002, 0, False)
eng.set_u(u0)
eng.run_trials(2)
for _ in range(3):
    eng.stage(eng.stages)
torch.cuda.synchronize()
t = bench.time_kernel(torch, lambda: eng.stage(eng.stages), 20)
passes = bench.STAGE_PASSES[method][-1]
print(f"{os.environ.get('RKS_LIB', 'default')} {method} last stage: {t * 1e6:.1f} us  {passes * 16 * u0.numel() / t / 1e9:.0f} GB/s")
