"""ms per trial of a batched NLS ensemble (2^25 modes in total) for every fast row length, with the pre-transforming
K1/K4 pair on and off (tuning aid).  Usage: python tools/pt_sweep.py [method=ETD35] [trials=12]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rkstiff_b200 as rk  # noqa: E402

method = sys.argv[1] if len(sys.argv) > 1 else "ETD35"
trials = int(sys.argv[2]) if len(sys.argv) > 2 else 12
dev = torch.device("cuda", 0)
adaptive = method in ("IF34", "ETD34", "ETD35", "IF45DP")
for n in [int(v) for v in os.environ.get("PT_SWEEP_N", "512,1024,2048,4096,8192").split(",")]:
    batch = (1 << 25) // n
    w = 40.0 * math.pi
    dx = 2 * w / n
    x = torch.arange(n, dtype=torch.float64, device=dev) * dx - w
    kx = 2 * math.pi * torch.fft.fftfreq(n, d=dx, dtype=torch.float64, device=dev)
    g = torch.Generator(device="cpu").manual_seed(n)
    eta = (0.5 + torch.rand(batch, 1, generator=g, dtype=torch.float64)).to(dev)
    x0 = (-20.0 + 40.0 * torch.rand(batch, 1, generator=g, dtype=torch.float64)).to(dev)
    u0 = torch.fft.fft((eta / torch.cosh(eta * (x[None, :] - x0))).to(torch.complex128), dim=-1)
    res = {}
    for pt in ("1", "0"):
        os.environ["RKS_PT"] = pt
        lin, nl = rk.models.nls_ops(kx, 2.0)
        sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=1e-6)) if adaptive else getattr(rk, method)(lin, nl)
        eng = sol._get_engine(u0)
        if adaptive:
            eng.begin(0.0, 1e9, 0.005, 0, False)
            eng.set_u(u0)
            run = eng.run_trials
        else:
            eng.begin(0.0, 0.0, 0.005, 0, True)
            eng.ensure_fixed_coeffs(0.005)
            eng.set_u(u0)
            run = eng.run_fixed
        run(4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(trials)
        e1.record()
        torch.cuda.synchronize()
        res[pt] = e0.elapsed_time(e1) / trials
        del sol, eng
        torch.cuda.empty_cache()
    print(f"{method} n={n:5d} B={batch:6d}  pair on {res['1']:.3f} ms  off {res['0']:.3f} ms  ({100 * (res['1'] / res['0'] - 1):+.1f} %)")
