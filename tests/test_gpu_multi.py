"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): a batch-sharded adaptive ensemble that
shares one dt must reproduce the single-process dt sequence of the whole batch (= the reference's
global-norm semantics), rank by rank, with NCCL carrying only three doubles per trial."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import problems  # noqa: E402
from oracle.rk_oracle import Config, OracleSolver  # noqa: E402

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _teardown(sol, q):
    """The sharded trial is a captured CUDA graph holding NCCL kernels: release it before the communicator goes.
    The worker then leaves without running interpreter shutdown (symmetric-memory mappings of the slab exchange
    and NCCL's own threads make that order-dependent): results are flushed to the queue first."""
    import torch.distributed as dist
    sol.close()
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    q.close()
    q.join_thread()
    os._exit(0)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import rkstiff_b200 as rk
    from rkstiff_b200.dist import shard_bounds
    p = problems.nls(512, batch=6, seed=2, half_width=20.0)
    lo, hi = shard_bounds(6, rank, world)
    kx = torch.from_numpy(p.kx).cuda()
    lin, nl = rk.models.nls_ops(kx, 2.0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6), group=dist.group.WORLD)
    uf = sol.evolve(torch.from_numpy(p.u0[lo:hi].copy()).cuda(), 0.0, 0.2, store_freq=3)
    q.put((rank, lo, hi, [r[0] for r in sol.trial_log], [r[2] for r in sol.trial_log], list(sol.t),
           uf.cpu().numpy()))
    _teardown(sol, q)


def test_sharded_shared_dt_ensemble_matches_whole_batch():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda r: r[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = problems.nls(512, batch=6, seed=2, half_width=20.0)
    ora = OracleSolver("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-6))
    uo = ora.evolve(p.u0, 0.0, 0.2, store_freq=3)
    for rank, lo, hi, hs, acc, ts, uf in res:
        assert acc == [r.accepted for r in ora.log]
        np.testing.assert_allclose(hs, [r.h for r in ora.log], rtol=1e-9)
        np.testing.assert_allclose(ts, ora.t, rtol=1e-9)
        assert np.linalg.norm(uf - uo[lo:hi]) / np.linalg.norm(uo[lo:hi]) < 1e-9
    assert res[0][3] == res[1][3]            # bit-identical dt sequence on every rank


def _slab_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import rkstiff_b200 as rk
    from rkstiff_b200.dist_fft import nls_slab_ops
    n = 16
    p = problems.nls_3d(n)
    k = torch.from_numpy(p.kx).cuda()
    lin, nl, fft = nls_slab_ops([k, k, k], gamma=2.0, group=dist.group.WORLD)
    u0 = fft.spec_slice(torch.from_numpy(p.u0.reshape(n, n, n)).cuda())
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-5), group=dist.group.WORLD)
    uf = sol.evolve(u0, 0.0, 0.2)
    q.put((rank, [r[0] for r in sol.trial_log], [r[2] for r in sol.trial_log], uf.cpu().numpy()))
    _teardown(sol, q)


def test_cfg5_slab_decomposed_nls3d_matches_flattened_oracle():
    """BASELINE cfg 5 at reduced size: 3-D NLS, ETD35, grid slab-decomposed over 2 GPUs with the
    FFT transpose as an NCCL all-to-all, against the reference's flattened single-process run."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_slab_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = sorted((q.get(timeout=300) for _ in range(2)), key=lambda r: r[0])
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    n = 16
    p = problems.nls_3d(n)
    ora = OracleSolver("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-5))
    uo = ora.evolve(p.u0, 0.0, 0.2).reshape(n, n, n)
    for rank, hs, acc, uf in res:
        assert acc == [r.accepted for r in ora.log]
        np.testing.assert_allclose(hs, [r.h for r in ora.log], rtol=1e-9)
        want = uo[:, rank * (n // 2):(rank + 1) * (n // 2)]          # spectral layout: axis 1 sharded
        assert np.linalg.norm(uf - want) / np.linalg.norm(want) < 1e-9
