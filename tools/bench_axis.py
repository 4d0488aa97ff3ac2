"""Time the strided-axis transforms alone (fft_axis.cuh): GB/s against the 2-pass byte model.
Usage: [RKS_LIB=...] python tools/bench_axis.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import rkstiff_b200 as rk
dev = torch.device("cuda", 0)
print(os.environ.get("RKS_LIB", "default"))
for shape, dim in (((512, 512, 128), 0), ((512, 512, 128), 1), ((128, 512, 512), 1), ((4096, 2049), 0), ((256, 256, 256), 0), ((1024, 1024, 16), 0)):
    n = shape[dim]
    ax = rk.models.AxisFFT(n, dev)
    x = torch.randn(shape, dtype=torch.float64, device=dev).to(torch.complex128)
    y = torch.empty_like(x)
    for name, fn in (("inv", lambda: ax.inverse_(x, dim, out=y)), ("fwd", lambda: ax.forward_(x, dim, out=y))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / 10
        print(f"{str(shape):18s} dim {dim} n={n:5d} {name}  {t*1e6:8.1f} us  {2*16*x.numel()/t/1e9:7.0f} GB/s")
    del ax, x, y
    torch.cuda.empty_cache()
