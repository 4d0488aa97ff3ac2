# quick timing of batched independent-dt vs shared-dt on cfg2-like sizes
import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
import rkstiff_b200 as rk
from oracle import problems
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
TF = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
for n, B in [(8192, 4096), (1024, 16384)]:
    p = problems.nls(n, batch=B, seed=2)
    lin, nl = rk.models.nls_ops(dev(p.kx), 2.0)
    u0 = dev(p.u0)
    for mode in ("shared", "independent"):
        sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6))
        for rep in range(2):
            torch.cuda.synchronize(); t = time.time()
            if mode == "shared":
                uf = sol.evolve(u0, 0.0, TF, store_data=False); trials = len(sol.trial_log) * B
            else:
                uf, logs = sol.evolve_independent(u0, 0.0, TF, keep_log=(rep == 0))
                if rep == 0: trials = sum(len(l) for l in logs); maxt = max(len(l) for l in logs)
            torch.cuda.synchronize(); dt = time.time() - t
        print(n, B, mode, f"{dt*1e3:.1f} ms", f"row-trials={trials}", f"{trials*n/dt:.3e} gp.trials/s", flush=True)
    print("max trials/row", maxt)
