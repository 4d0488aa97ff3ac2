"""The reference's own accuracy tests, run through the B200 classes:
  * KdV soliton against the exact solution, evolve() and step() (tests/testing_util.py:42-124) with the
    reference's tolerances (test_etd4.py:12,19; test_etd5.py:14,21; test_if4.py:28,35; test_if34.py:141,148;
    test_etd34.py:32,40; test_etd35.py:27,33);
  * Burgers norm invariant (test_if34.py:15-21, test_if45dp.py:10-16);
  * convergence order by log-log least squares (test_order_convergence.py:221-295);
  * structural checks: snapshot cadence, h caching, tf < t0, explosive nonlinearity."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rk():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import rkstiff_b200
    return rkstiff_b200


def kdv_setup(rk):
    x, kx = rk.grids.construct_x_kx_rfft(256, -30.0, 30.0)
    h, steps = 0.025, 200
    u0 = torch.fft.rfft(rk.models.kdv_soliton(x, ampl=1.0, x0=-5.0, t=0.0))
    exact = torch.fft.rfft(rk.models.kdv_soliton(x, ampl=1.0, x0=-5.0, t=h * steps))
    lin, nl = rk.models.kdv_ops(kx)
    return u0, lin, nl, exact, h, steps


def relerr(a, b):
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


FIXED_TOL = {"ETD4": 1e-6, "ETD5": 1e-6, "IF4": 1e-5}
ADAPT = {"IF34": 1e-5, "ETD34": 1e-4, "ETD35": 1e-4}      # epsilon used by the reference tests; bound 1e-4


@pytest.mark.parametrize("method", list(FIXED_TOL))
def test_kdv_soliton_fixed_step_evolve_and_step(rk, method):
    u0, lin, nl, exact, h, steps = kdv_setup(rk)
    sol = getattr(rk, method)(lin, nl)
    assert relerr(sol.evolve(u0, 0.0, h * steps, h, store_data=False), exact) < FIXED_TOL[method]
    sol = getattr(rk, method)(lin, nl)
    u = u0.clone()
    for _ in range(steps):
        u = sol.step(u, h)
    assert relerr(u, exact) < FIXED_TOL[method]


@pytest.mark.parametrize("method", list(ADAPT))
def test_kdv_soliton_adaptive_evolve_and_step(rk, method):
    u0, lin, nl, exact, h, steps = kdv_setup(rk)
    sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=ADAPT[method]))
    assert relerr(sol.evolve(u0, 0.0, h * steps, h_init=h, store_data=False), exact) < 1e-4
    # step(): with a loose tolerance the controller must not change h (testing_util.py:110-117)
    sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=0.1))
    u = u0.clone()
    for _ in range(steps):
        u, h_actual, _ = sol.step(u, h)
        assert abs(h_actual - h) < 1e-10
    assert relerr(u, exact) < 1e-4


@pytest.mark.parametrize("method", ["IF34", "IF45DP", "ETD35"])
def test_burgers_norm_invariant(rk, method):
    x, kx = rk.grids.construct_x_kx_rfft(1024, -np.pi, np.pi)
    lin, nl = rk.models.burgers_ops(kx, 0.0005)
    u0 = torch.fft.rfft(torch.exp(-10 * torch.sin(x / 2) ** 2))
    sol = getattr(rk, method)(lin, nl)
    uf = sol.evolve(u0, 0.0, 0.85, store_data=False)
    assert abs(float(torch.linalg.norm(uf) / torch.linalg.norm(u0)) - 1.0) < 1e-2


@pytest.mark.parametrize("method,expected", [("ETD4", 4), ("IF4", 4), ("ETD5", 5)])
def test_convergence_order_on_kdv(rk, method, expected):
    u0, lin, nl, _, _, _ = kdv_setup(rk)
    hs = [0.05, 0.025, 0.0125] if expected == 4 else [0.1, 0.05, 0.025]
    tf = 1.0

    def run(h):
        # fixed number of step() calls like the reference's order tests (evolve() counts steps by float
        # accumulation and may take one more, solvercs.py:258-261)
        sol = getattr(rk, method)(lin, nl)
        u = u0.clone()
        for _ in range(int(round(tf / h))):
            u = sol.step(u, h)
        return u

    ref = run(hs[-1] / 4)
    errs = [relerr(run(h), ref) for h in hs]
    order = np.polyfit(np.log(hs), np.log(errs), 1)[0]
    assert order > expected - 0.5, (order, errs)


def test_snapshot_cadence_and_reset(rk):
    u0, lin, nl, _, h, _ = kdv_setup(rk)
    sol = rk.ETD4(lin, nl)
    sol.evolve(u0, 0.0, 10 * h, h, store_data=True, store_freq=2)
    assert len(sol.t) == len(sol.u) == 6 and sol.t[0] == 0.0
    assert sol.u[0] is u0                                  # the reference stores the caller's array
    sol.reset()
    assert sol.t == [] and sol.u == []
    sol.evolve(u0, 0.0, 10 * h, h, store_data=False)
    assert sol.t == []


def test_step_size_cache_key_and_live_config(rk):
    u0, lin, nl, _, h, _ = kdv_setup(rk)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-3))
    sol.evolve(u0, 0.0, 0.5, store_data=False)
    updates_loose = sol._engine.read_ctrl().coeff_updates
    trials_loose = len(sol.trial_log)
    assert 0 < updates_loose <= trials_loose              # coefficients rebuilt only when h changes (etd35.py:851)
    sol.config.epsilon = 1e-7                              # config is read live (solveras.py:452-454)
    sol.evolve(u0, 0.0, 0.5, store_data=False)
    assert len(sol.trial_log) > trials_loose


def test_host_snapshot_pipeline_matches_device_snapshots(rk):
    """snapshot_device="cpu": pinned-host snapshots through an asynchronous side-stream copy."""
    u0, lin, nl, _, h, _ = kdv_setup(rk)
    for cls, kw in ((rk.ETD4, dict(h=h)), (rk.ETD35, dict(h_init=h))):
        dev_sol, host_sol = cls(lin, nl), cls(lin, nl)
        host_sol.snapshot_device = "cpu"
        dev_sol.evolve(u0, 0.0, 20 * h, store_freq=3, **kw)
        host_sol.evolve(u0, 0.0, 20 * h, store_freq=3, **kw)
        assert host_sol.t == dev_sol.t and len(host_sol.u) == len(dev_sol.u) > 2
        for a, b in zip(host_sol.u[1:], dev_sol.u[1:]):
            assert not a.is_cuda and a.is_pinned()
            assert torch.equal(a, b.cpu())


def test_device_derivatives_on_the_gpu():
    """derivatives.dx_rfft / dx_fft on CUDA tensors against NumPy (reference derivatives.py:47-179), batched: every
    kernel family behind rks_rows_* (generic 16 / 32, packed 64 ... 256, register path 512 ... 8192), odd row
    counts (real rows go in pairs), and lengths the engine does not transform (16384, 96: torch.fft on the device)."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rkstiff_b200 import _abi
    from rkstiff_b200 import derivatives as d
    assert _abi.lib is not None
    rng = np.random.default_rng(11)
    for n in (16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 96):
        x = np.arange(n) * (2 * np.pi / n)
        kr = 2 * np.pi * np.fft.rfftfreq(n, d=2 * np.pi / n)
        kc = 2 * np.pi * np.fft.fftfreq(n, d=2 * np.pi / n)
        u = np.stack([np.sin(3 * x) + 0.5 * np.cos(5 * x), np.cos(x) ** 3, rng.standard_normal(n)])      # 3 rows: odd
        for order in (1, 2, 3):
            got = d.dx_rfft(torch.from_numpy(kr).cuda(), torch.from_numpy(u).cuda(), order).cpu().numpy()
            ref = np.fft.irfft((1j * kr) ** order * np.fft.rfft(u, axis=-1), n=n, axis=-1)
            assert got.shape == ref.shape
            np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
        z = np.stack([u[0] + 1j * u[1], u[2] - 0.5j * u[0]]).reshape(2, 1, n)
        got = d.dx_fft(torch.from_numpy(kc).cuda(), torch.from_numpy(z).cuda(), 2).cpu().numpy()
        ref = np.fft.ifft((1j * kc) ** 2 * np.fft.fft(z, axis=-1), axis=-1)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    # a reusable handle on a larger batch, single row included
    n = 1024
    kr = torch.from_numpy(2 * np.pi * np.fft.rfftfreq(n, d=0.05)).cuda()
    dd = d.SpectralDerivative(kr, n, 1, real=True)
    big = torch.from_numpy(rng.standard_normal((257, n))).cuda()
    ref = np.fft.irfft((1j * kr.cpu().numpy()) * np.fft.rfft(big.cpu().numpy(), axis=-1), n=n, axis=-1)
    np.testing.assert_allclose(dd(big).cpu().numpy(), ref, rtol=0, atol=1e-12 * np.abs(ref).max())
    np.testing.assert_allclose(dd(big[0]).cpu().numpy(), ref[0], rtol=0, atol=1e-12 * np.abs(ref).max())


def test_gemv_kernel_matches_numpy():
    """rks_gemv (the S / S^-1 products of diagonalize=True, etd35.py:463, 495) against NumPy: vector and batch of
    row vectors, sizes that are not multiples of the warp or of the rows-per-CTA."""
    import numpy as np
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from rkstiff_b200.solver import BaseSolver
    rng = np.random.default_rng(3)
    for n, batch in ((1, 1), (5, 1), (33, 3), (130, 1), (257, 4)):
        a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        x = rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))
        xs = x[0] if batch == 1 else x
        got = BaseSolver._gemv(torch.from_numpy(a).cuda(), torch.from_numpy(np.ascontiguousarray(xs)).cuda()).cpu().numpy()
        ref = xs @ a.T
        assert got.shape == ref.shape
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-13 * np.abs(ref).max() * n)
    # a real state vector is widened (the reference's nl_func may return float arrays)
    a = rng.standard_normal((7, 7)) + 0j
    v = rng.standard_normal(7)
    got = BaseSolver._gemv(torch.from_numpy(a).cuda(), torch.from_numpy(v).cuda()).cpu().numpy()
    np.testing.assert_allclose(got, a @ v, rtol=0, atol=1e-13)
    with pytest.raises(ValueError):
        BaseSolver._gemv(torch.from_numpy(a).cuda(), torch.zeros(6, dtype=torch.complex128, device="cuda"))
