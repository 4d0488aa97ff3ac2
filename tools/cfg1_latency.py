"""BASELINE cfg 1 (README quick-start): KS n=1024, IF34 adaptive, t 0->50, store_freq=20, one trajectory.
Wall time of evolve() on the GPU (fused and torch-callable paths) next to the oracle on one host core."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import rkstiff_b200 as rk  # noqa: E402
from oracle import problems  # noqa: E402
from oracle.rk_oracle import OracleSolver  # noqa: E402

p = problems.ks(1024)
kx = torch.from_numpy(p.kx).cuda()
u0 = torch.from_numpy(p.u0).cuda()
lin, nl = rk.models.ks_ops(kx)
for label, f in (("fused", nl), ("callable", lambda v: nl(v))):
    for method in ("IF34", "ETD35"):
        sol = getattr(rk, method)(lin, f)
        sol.evolve(u0, 0.0, 5.0, store_freq=20)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        uf = sol.evolve(u0, 0.0, 50.0, store_freq=20)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        acc = sum(1 for r in sol.trial_log if r[2])
        print(f"gpu {label:8s} {method}: {dt*1e3:8.1f} ms  {len(sol.trial_log)} trials ({acc} accepted), {len(sol.u)} snapshots, "
              f"{dt/len(sol.trial_log)*1e6:.1f} us/trial, {1024*len(sol.trial_log)/dt:.3e} gp*trials/s")
for method in ("IF34", "ETD35"):
    ora = OracleSolver(method, p.lin_op, p.nl_func)
    t0 = time.perf_counter()
    ora.evolve(p.u0, 0.0, 50.0, store_freq=20)
    dt = time.perf_counter() - t0
    print(f"cpu oracle   {method}: {dt*1e3:8.1f} ms  {len(ora.log)} trials, {dt/len(ora.log)*1e6:.1f} us/trial, "
          f"{1024*len(ora.log)/dt:.3e} gp*trials/s")
