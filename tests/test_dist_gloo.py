"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch sharding and the MAX-then-SUM
exchange that makes a sharded ensemble take the reference's global accept/reject decision."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import problems
from oracle.rk_oracle import Config, OracleSolver


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, out_queue):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rkstiff_b200.dist import allreduce_error_scalars, shard_bounds
    p = problems.nls(128, batch=batch, seed=2, half_width=20.0)
    cfg = Config(epsilon=1e-5)
    sol = OracleSolver("ETD35", p.lin_op, p.nl_func, cfg)
    unew, err = sol.trial(p.u0, 0.01)                    # every rank can form the global answer ...
    lo, hi = shard_bounds(batch, rank, world)            # ... but only reduces its own rows
    u_loc, e_loc = unew[lo:hi], err[lo:hi]
    red = torch.zeros(3, dtype=torch.float64)
    red[0] = float((np.abs(u_loc) ** 2).max()) if hi > lo else 0.0

    def masked_sums():
        m = float(np.sqrt(red[0].item()))
        mask = np.sqrt(np.abs(u_loc) ** 2) / m > cfg.adapt_cutoff
        red[1] = float((np.abs(u_loc[mask]) ** 2).sum())
        red[2] = float((np.abs(e_loc[mask]) ** 2).sum())

    allreduce_error_scalars(red, None, between=masked_sums)
    s = cfg.safety_f * (cfg.epsilon * np.sqrt(red[1].item()) / np.sqrt(red[2].item())) ** 0.25
    out_queue.put((rank, lo, hi, s, float(sol.compute_s(unew, err))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 5])        # even and ragged split
def test_sharded_error_control_matches_global_decision(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    (r0, lo0, hi0, s0, g0), (r1, lo1, hi1, s1, g1) = res
    assert (lo0, hi1) == (0, batch) and hi0 == lo1            # shards tile the batch
    assert s0 == s1                                            # every rank takes the same decision
    assert s0 == pytest.approx(g0, rel=1e-12)                  # and it is the reference's global one


def test_shard_bounds_cover_every_row_once():
    from rkstiff_b200.dist import shard_bounds
    for batch in (1, 7, 64, 4096, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _fft_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rkstiff_b200.dist_fft import SlabFFT
    shape = (8, 4, 6)
    g = torch.Generator().manual_seed(0)
    full = torch.randn(shape, dtype=torch.float64, generator=g) + 1j * torch.randn(shape, dtype=torch.float64, generator=g)
    fft = SlabFFT(shape)
    spec = fft.forward(fft.real_slice(full))
    want = fft.spec_slice(torch.fft.fftn(full))
    back = fft.inverse(spec)
    q.put((rank, float((spec - want).abs().max()), float((back - fft.real_slice(full)).abs().max())))
    dist.barrier()
    dist.destroy_process_group()


def test_slab_fft_matches_fftn_and_round_trips():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_fft_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for _, e_fwd, e_back in res:
        assert e_fwd < 1e-12 and e_back < 1e-13


class _TorchAxis:
    """CPU stand-in for models.AxisFFT (natural order instead of digit-reversed: invisible to a pointwise N)."""

    def inverse_(self, x, dim, out=None):
        res = torch.fft.ifft(x, dim=dim)
        if out is None:
            x.copy_(res)
            return x
        out.copy_(res)
        return out

    def forward_(self, x, dim, out=None):
        res = torch.fft.fft(x, dim=dim)
        if out is None:
            x.copy_(res)
            return x
        out.copy_(res)
        return out

    def chunked_(self, x, inverse):
        # x: (G, outer, n/G, inner...) block-major along the transformed axis
        G, outer, nb = x.shape[:3]
        y = x.permute((1, 0, 2) + tuple(range(3, x.dim()))).reshape((outer, G * nb) + tuple(x.shape[3:]))
        y = torch.fft.ifft(y, dim=1) if inverse else torch.fft.fft(y, dim=1)
        x.copy_(y.reshape((outer, G, nb) + tuple(x.shape[3:])).permute((1, 0, 2) + tuple(range(3, x.dim()))))
        return x


def _rows_nls(a, out=None):
    f = torch.fft.ifft(a, dim=-1)
    res = 2j * torch.fft.fft((f.real ** 2 + f.imag ** 2) * f, dim=-1)
    out.copy_(res)
    return out


def _pipeline_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rkstiff_b200.dist_fft import SlabFFT
    shape = (8, 4, 6)
    g = torch.Generator().manual_seed(1)
    full = torch.randn(shape, dtype=torch.float64, generator=g) + 1j * torch.randn(shape, dtype=torch.float64, generator=g)
    f = torch.fft.ifftn(full)
    want_full = 2j * torch.fft.fftn((f.real ** 2 + f.imag ** 2) * f)
    fft = SlabFFT(shape)
    s = fft.spec_slice(full)
    axes = [_TorchAxis(), _TorchAxis()]
    errs = []
    for chunks in ("1", "2", "4"):
        os.environ["RKS_SLAB_CHUNKS"] = chunks
        keep = s.clone()
        got = fft.fused_nl(s, _rows_nls, axes)
        errs.append((chunks, fft._pipeline_chunks(), float((got - fft.spec_slice(want_full)).abs().max()),
                     float((s - keep).abs().max())))
    q.put((rank, errs))
    dist.barrier()
    dist.destroy_process_group()


def test_pipelined_slab_exchange_equals_single_exchange():
    """The x-plane-chunked all-to-all pipeline of SlabFFT.fused_nl (dist_fft.py) against fftn of the whole grid."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pipeline_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for _, errs in res:
        assert [e[1] for e in errs] == [1, 2, 4]          # 8 / 2 = 4 local planes: 1, 2 and 4 groups
        for _, _, err, touched in errs:
            assert err < 1e-10 and touched == 0.0           # the input spectrum stays intact
