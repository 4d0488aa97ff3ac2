import sys, os, torch
sys.path.insert(0, "/root/repo")
import rkstiff_b200 as rk
dev = torch.device("cuda", 0)
tot = 1 << 25
for model in ("nls", "uux"):
    for n in (64, 128, 256, 512):
        n_c = n if model == "nls" else n // 2 + 1
        batch = tot // n
        kx = torch.linspace(0, 10, n_c, dtype=torch.float64, device=dev)
        lin, nl = rk.models.nls_ops(kx, 2.0) if model == "nls" else rk.models.ks_ops(kx)
        sol = rk.ETD4(lin, nl)
        u = torch.randn(batch, n_c, dtype=torch.complex128, device=dev)
        eng = sol._get_engine(u); eng.set_u(u); eng.nl(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): eng.nl(1)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-4
        print(f"{model} n={n:5d} B={batch:7d} {t*1e6:8.1f} us {2*16*batch*n_c/t/1e9:7.0f} GB/s", flush=True)
        del sol, eng, u
