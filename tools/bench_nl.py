"""Time the fused NL kernel (K4) alone over row lengths and models: GB/s against the 2-pass byte model.
Usage: [RKS_LIB=path/to/lib.so] python tools/bench_nl.py [total_elems_log2=26]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rkstiff_b200 as rk  # noqa: E402

tot = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 25)
dev = torch.device("cuda", 0)
out = []
for model in ("nls", "uux", "cubic"):
    for n in [int(v) for v in os.environ.get("BENCH_NL_N", "512,1024,2048,4096,8192").split(",")]:
        n_c = n if model == "nls" else n // 2 + 1
        batch = tot // n
        kx = torch.linspace(0, 10, n_c, dtype=torch.float64, device=dev)
        if model == "nls":
            lin, nl = rk.models.nls_ops(kx, 2.0)
        elif model == "cubic":
            lin, nl = rk.models.allen_cahn_1d_ops(kx)
        else:
            lin, nl = rk.models.ks_ops(kx)
        sol = rk.ETD4(lin, nl)
        u = torch.randn(batch, n_c, dtype=torch.complex128, device=dev)
        eng = sol._get_engine(u)
        eng.set_u(u)
        eng.nl(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record()
        for _ in range(reps):
            eng.nl(1)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        gbs = 2 * 16 * batch * n_c / t / 1e9
        out.append(f"{model} n={n:5d} B={batch:6d}  {t*1e6:8.1f} us  {gbs:7.0f} GB/s  {batch*n/t/1e9:6.2f} Ggp/s")
        del sol, eng, u
        torch.cuda.empty_cache()
print(os.environ.get("RKS_LIB", "default"), {k: v for k, v in os.environ.items() if k.startswith("RKS_") and k != "RKS_LIB"})
print("\n".join(out))
