"""ncu target: one ETD35 trial on a (512, 512, 64) NLS grid (the 512-point strided-axis kernels of cfg 5 at 1/8 of its memory)."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rkstiff_b200 as rk
from rkstiff_b200._abi import check, lib
dev = torch.device("cuda", 0)
dims = (512, 512, 64)
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
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
eng.set_h(0.0019)
eng.update_coeffs()
for s in range(1, eng.stages + 1):
    check(lib.rks_stage_nl(eng.plan, s, eng.st))
check(lib.rks_error_sums(eng.plan, eng.st))
torch.cuda.synchronize()
rt.cudaProfilerStop()
print("done", eng.coef_storage)
