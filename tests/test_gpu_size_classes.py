"""Parity at the size classes bench.py measures: the kernels the benchmark launches -- not their
small-size siblings -- against the oracle.

  cfg 4  IF45DP step on the 4096^2 Allen-Cahn grid (axis_fft_plan_kernel<4096>, nl_fast_kernel<8,3,0>, separable
         coefficient tables) vs OracleSolver.step on the flattened arrays (demos/nls.ipynb:496-511 formulation)
  cfg 5  ETD35 step on a (512, 16, 64) and on a 64^3 NLS grid (axis_fft kernels of length 512 / 64 / 16, indexed
         coefficient records) vs the flattened oracle
  cfg 3  KS n=1024 ETD4, 64 trajectories x 50 steps, <= 1e-12 relative per step
  cfg 2a NLS n=8192 ETD35, 64 trajectories sharing one dt, accept/reject and dt sequence + final state
Bars: state <= 1e-12 relative per step (same input to both sides), h and h_suggest within 1e-9.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import problems  # noqa: E402
from oracle.rk_oracle import Config, OracleSolver  # noqa: E402

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-12
DT_TOL = 1e-9
FINAL_TOL = 1e-9


@pytest.fixture(scope="module")
def rk():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import rkstiff_b200
    return rkstiff_b200


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def test_cfg4_if45dp_step_at_4096_squared(rk):
    n = 4096
    p = problems.allen_cahn_2d(n)
    shape = p.params["shape"]
    lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01)
    assert isinstance(nl, rk.models.FusedGridNL)
    sol = rk.IF45DP(lin, nl, config=rk.SolverConfig(epsilon=1e-4))
    ora = OracleSolver("IF45DP", p.lin_op, p.nl_func, Config(epsilon=1e-4))
    u, h, h_next = sol.step(dev(p.u0.reshape(shape)), 0.002)
    uo, ho, ho_next = ora.step(p.u0, 0.002)
    assert sol._engine.coef_storage == "separable"
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    assert h == pytest.approx(ho, rel=DT_TOL) and h_next == pytest.approx(ho_next, rel=DT_TOL)
    assert rel(host(u).ravel(), uo) < STEP_TOL       # (one oracle step of this size is ~18 s of NumPy)


def _nls_grid_problem(dims, half_width=6.0, gamma=2.0):
    """problems.nls_3d on a non-cubic grid (same formulation, flattened like the reference's demo)."""
    axes = [problems.x_kx_fft(n, -half_width, half_width) for n in dims]
    K = np.meshgrid(*[a[1] for a in axes], indexing="ij")
    X = np.meshgrid(*[a[0] for a in axes], indexing="ij")
    lin = (-1j * sum(k ** 2 for k in K)).ravel()
    u0 = np.exp(-sum(x ** 2 for x in X)).astype(np.complex128)

    def nl(uf):
        f = np.fft.ifftn(uf.reshape(dims))
        f2 = f.real ** 2 + f.imag ** 2
        return (1j * gamma * np.fft.fftn(f2 * f)).ravel()

    return lin, nl, np.fft.fftn(u0).ravel(), [a[1] for a in axes]


@pytest.mark.parametrize("dims", [(512, 16, 64), (64, 64, 64)], ids=["512x16x64", "64cubed"])
def test_cfg5_etd35_step_on_large_axes(rk, dims):
    lin_o, nl_o, u0, ks = _nls_grid_problem(dims)
    lin, nl = rk.models.nls_nd_ops([dev(k) for k in ks], gamma=2.0)
    assert isinstance(nl, rk.models.FusedGridNL)
    np.testing.assert_allclose(host(lin).ravel(), lin_o, rtol=1e-14)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-5))
    ora = OracleSolver("ETD35", lin_o, nl_o, Config(epsilon=1e-5))
    u, h, h_next = sol.step(dev(u0.reshape(dims)), 0.002)
    uo, ho, ho_next = ora.step(u0, 0.002)
    assert sol._engine.coef_storage == "indexed"
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    assert h == pytest.approx(ho, rel=DT_TOL) and h_next == pytest.approx(ho_next, rel=DT_TOL)
    assert rel(host(u).ravel(), uo) < STEP_TOL
    u2, h2, _ = sol.step(u, h_next)
    uo2, ho2, _ = ora.step(host(u).ravel(), h_next)
    assert h2 == pytest.approx(ho2, rel=DT_TOL)
    assert rel(host(u2).ravel(), uo2) < STEP_TOL


def test_cfg3_ks_etd4_64_trajectories_50_steps(rk):
    """SURVEY 8d cfg 3: n = 1024, 64 trajectories x 50 steps, <= 1e-12 per step (same input both sides each step)."""
    p = problems.ks(1024, batch=64, seed=0)
    lin, nl = rk.models.ks_ops(dev(p.kx))
    sol = rk.ETD4(lin, nl)
    ora = OracleSolver("ETD4", p.lin_op, p.nl_func)
    u = dev(p.u0)
    worst = 0.0
    for _ in range(50):
        ora.reset()
        ref = ora.step(host(u), 0.05)
        sol.reset()
        u = sol.step(u, 0.05)
        worst = max(worst, rel(host(u), ref))
    assert worst < STEP_TOL
    # and the zero-sync fixed-step evolve of the same 50 steps against the oracle's own 50-step run
    uf = rk.ETD4(lin, nl).evolve(dev(p.u0), 0.0, 2.5, 0.05, store_data=False)
    uo = OracleSolver("ETD4", p.lin_op, p.nl_func).evolve(p.u0, 0.0, 2.5, 0.05, store_data=False)
    assert rel(host(uf), uo) < 1e-10          # 50 chaotic steps: per-step 1e-12 compounded


def test_cfg2a_nls_etd35_64_trajectories_shared_dt(rk):
    """SURVEY 8d cfg 2a: B = 64 solitons at n = 8192 sharing one dt (global norms, solveras.py:451-454)."""
    p = problems.nls(8192, batch=64, seed=2)
    lin, nl = rk.models.nls_ops(dev(p.kx), 2.0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6))
    ora = OracleSolver("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-6))
    uf = sol.evolve(dev(p.u0), 0.0, 0.06, store_data=False)
    uo = ora.evolve(p.u0, 0.0, 0.06, store_data=False)
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    np.testing.assert_allclose([r[0] for r in sol.trial_log], [r.h for r in ora.log], rtol=DT_TOL)
    assert rel(host(uf), uo) < FINAL_TOL
