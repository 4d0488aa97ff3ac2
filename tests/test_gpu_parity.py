"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference goldens.

Bars (BASELINE.json north_star):
  * fixed-step states within 1e-12 relative per step in FP64;
  * adaptive runs reproduce the reference's accepted/rejected dt sequence over the compared
    horizon (flags identical, dt within 1e-9 relative) with the final state within 1e-9 relative.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import problems  # noqa: E402
from oracle.rk_oracle import ADAPTIVE, FIXED, METHODS, Config, OracleSolver, coefficients  # noqa: E402

pytestmark = pytest.mark.gpu

STEP_TOL = 1e-12        # fixed-step relative tolerance per step
DT_TOL = 1e-9           # adaptive dt-sequence relative tolerance
FINAL_TOL = 1e-9        # adaptive final-state relative tolerance


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


def fused_for(rk, prob):
    kx = dev(prob.kx)
    if prob.model == "cubic":
        return rk.models.FusedNL(3, prob.n, None, prob.params["c"], "allen_cahn")
    if prob.model == "sine_gordon":
        return rk.models.FusedNL(4, prob.n, kx, 0.0, "sine_gordon")
    if prob.model == "nls":
        return rk.models.FusedNL(2, prob.n, kx, prob.params["gamma"], "nls")
    return rk.models.FusedNL(1, prob.n, kx, prob.params["c"], prob.model)


def torch_callable(rk, prob):
    f = fused_for(rk, prob)
    return lambda v: f(v)          # plain callable => torch.fft path


def make(rk, method, prob, nl, epsilon=None):
    cls = getattr(rk, method)
    lin = dev(prob.lin_op)
    if method in ADAPTIVE:
        cfg = rk.SolverConfig() if epsilon is None else rk.SolverConfig(epsilon=epsilon)
        return cls(lin, nl, config=cfg)
    return cls(lin, nl)


# --------------------------------------------------------------------------------------------
# K4: fused nonlinearity
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [16, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("batch", [1, 3, 40, 301])
def test_fused_uux_matches_numpy(rk, n, batch):
    p = problems.ks(n, batch=batch, seed=n) if n >= 64 else problems.kdv(n, batch=batch, seed=n)
    sol = rk.ETD4(dev(p.lin_op), fused_for(rk, p))
    u = dev(p.u0)
    eng = sol._get_engine(u)
    eng.set_u(u)
    eng.nl(1)
    got = host(eng.state_view("N1"))
    assert rel(got, p.nl_func(p.u0)) < 2e-14 * np.log2(n)


@pytest.mark.parametrize("n", [16, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("batch", [1, 5, 150])
def test_fused_nls_matches_numpy(rk, n, batch):
    p = problems.nls(n, batch=batch, seed=n, half_width=20.0)
    sol = rk.ETD4(dev(p.lin_op), fused_for(rk, p))
    u = dev(p.u0)
    eng = sol._get_engine(u)
    eng.set_u(u)
    eng.nl(1)
    assert rel(host(eng.state_view("N1")), p.nl_func(p.u0)) < 2e-14 * np.log2(n)


# --------------------------------------------------------------------------------------------
# K2: coefficient arrays
# --------------------------------------------------------------------------------------------
NAMES = {
    "ETD4": ["E", "E2", "a21", "a31", "a32", "a41", "a43", "a51", "a52", "a54"],
    "ETD5": ["E14", "E12", "E34", "E", "a21", "a31", "a32", "a41", "a43", "a51", "a52", "a54", "a61", "a62", "a63",
             "a65", "a71", "a73", "a74", "a75", "a76"],
    "IF4": ["E", "E2"],
    "IF45DP": ["E15", "E310", "E45", "E89", "E", "a21", "a31", "a32", "a41", "a42", "a43", "a51", "a52", "a53", "a54",
               "a61", "a62", "a63", "a64", "a65", "a71", "a73", "a74", "a75", "r1", "r3", "r4", "r5"],
}
NAMES["ETD34"] = NAMES["ETD4"]
NAMES["ETD35"] = NAMES["ETD5"]
NAMES["IF34"] = NAMES["IF4"]


@pytest.mark.parametrize("probname,h", [("ks", 0.05), ("nls", 0.013), ("kdv", 0.025)])
@pytest.mark.parametrize("method", METHODS)
def test_coefficient_kernel_matches_oracle(rk, probname, h, method):
    p = {"ks": lambda: problems.ks(256), "nls": lambda: problems.nls(256, half_width=20.0),
         "kdv": lambda: problems.kdv(256)}[probname]()
    sol = make(rk, method, p, torch_callable(rk, p))
    u = dev(p.u0)
    eng = sol._get_engine(u)
    eng.set_h(h)
    from rkstiff_b200._abi import check, lib
    check(lib.rks_update_coeffs(eng.plan, eng.st))
    ref = coefficients(method, p.lin_op, h, Config())
    z = np.abs(h * p.lin_op)
    band = (z >= 0.01) & (z < 0.5)        # cancellation band of the closed forms
    for name in NAMES[method]:
        got = host(eng.coef_view(name)).astype(np.complex128)
        want = np.asarray(ref[name], dtype=np.complex128)
        np.testing.assert_allclose(got[~band], want[~band], rtol=5e-13, atol=1e-15 * h, err_msg=f"{method}.{name}")
        np.testing.assert_allclose(got[band], want[band], rtol=1e-7, atol=1e-15 * h, err_msg=f"{method}.{name}")


# --------------------------------------------------------------------------------------------
# fixed step: per-step parity and reference goldens
# --------------------------------------------------------------------------------------------
FIXED_CASES = {"kdv": lambda: problems.kdv(256), "ks": lambda: problems.ks(256),
               "burgers": lambda: problems.burgers(256, mu=0.01), "nls": lambda: problems.nls(256, half_width=20.0),
               "ksb": lambda: problems.ks(128, batch=3, seed=0)}


@pytest.mark.parametrize("path", ["fused", "callable"])
@pytest.mark.parametrize("tag", list(FIXED_CASES))
@pytest.mark.parametrize("method", FIXED)
def test_fixed_step_parity_per_step(rk, golden, tag, method, path):
    """Feed the SAME input state to both sides every step; <= 1e-12 relative per step."""
    g = golden("fixed_runs.npz")
    p = FIXED_CASES[tag]()
    pre = f"{tag}_{method}_"
    h = float(g[pre + "h"])
    nl = fused_for(rk, p) if path == "fused" else torch_callable(rk, p)
    sol = make(rk, method, p, nl)
    got1 = host(sol.step(dev(p.u0), h))
    assert rel(got1, g[pre + "u_step1"]) < STEP_TOL
    # 20 more steps, oracle advanced from the engine's state each time
    u = dev(p.u0)
    for _ in range(20):
        ref = OracleSolver(method, p.lin_op, p.nl_func).step(host(u), h)
        sol.reset()
        u = sol.step(u, h)
        assert rel(host(u), ref) < STEP_TOL


@pytest.mark.parametrize("path", ["fused", "callable"])
@pytest.mark.parametrize("tag", list(FIXED_CASES))
@pytest.mark.parametrize("method", FIXED)
def test_fixed_evolve_matches_reference_golden(rk, golden, tag, method, path):
    g = golden("fixed_runs.npz")
    p = FIXED_CASES[tag]()
    pre = f"{tag}_{method}_"
    h, tf, steps = float(g[pre + "h"]), float(g[pre + "tf"]), int(g[pre + "steps"])
    nl = fused_for(rk, p) if path == "fused" else torch_callable(rk, p)
    sol = make(rk, method, p, nl)
    uf = sol.evolve(dev(p.u0), 0.0, tf, h, store_freq=max(1, steps // 4))
    np.testing.assert_array_equal(np.array(sol.t), g[pre + "t"])          # float-accumulated loop count
    assert len(sol.u) == int(g[pre + "n_snap"])
    tol = 1e-9 if tag.startswith("ks") else steps * STEP_TOL               # KS is chaotic: looser over 60 steps
    assert rel(host(uf), g[pre + "u_final"]) < tol
    assert rel(host(sol.u[1]), g[pre + "u_snap_1"]) < tol


def test_fixed_step_h_larger_than_interval_raises(rk):
    p = problems.kdv(64)
    sol = rk.ETD4(dev(p.lin_op), fused_for(rk, p))
    with pytest.raises(ValueError):
        sol.evolve(dev(p.u0), 0.0, 0.1, 0.2)
    with pytest.raises(AssertionError):
        sol.step(dev(p.u0), -0.1)


# --------------------------------------------------------------------------------------------
# adaptive: dt sequence + final state against the reference goldens
# --------------------------------------------------------------------------------------------
AD_CASES = {"kdv": lambda: problems.kdv(256), "ks": lambda: problems.ks(256),
            "burgers": lambda: problems.burgers(256, mu=0.01), "nls": lambda: problems.nls(256, half_width=20.0),
            "nlsb": lambda: problems.nls(128, batch=3, seed=2, half_width=20.0),
            "ksb": lambda: problems.ks(128, batch=4, seed=0)}


@pytest.mark.parametrize("path", ["fused", "callable"])
@pytest.mark.parametrize("tag", list(AD_CASES))
@pytest.mark.parametrize("method", ADAPTIVE)
def test_adaptive_dt_sequence_and_final_state(rk, golden, tag, method, path):
    g = golden("adaptive_runs.npz")
    p = AD_CASES[tag]()
    pre = f"{tag}_{method}_"
    h0 = float(g[pre + "h_init"])
    nl = fused_for(rk, p) if path == "fused" else torch_callable(rk, p)
    sol = make(rk, method, p, nl, float(g[pre + "epsilon"]))
    uf = sol.evolve(dev(p.u0), 0.0, float(g[pre + "tf"]), None if np.isnan(h0) else h0,
                    store_freq=int(g[pre + "store_freq"]))
    hs = np.array([r[0] for r in sol.trial_log])
    acc = np.array([r[2] for r in sol.trial_log])
    ref_h, ref_acc = g[pre + "trial_h"], g[pre + "trial_accepted"]
    assert len(hs) == len(ref_h), f"{len(hs)} trials vs {len(ref_h)} in the reference"
    np.testing.assert_array_equal(acc, ref_acc)
    np.testing.assert_allclose(hs[:-1], ref_h[:-1], rtol=DT_TOL, atol=0)
    # the last dt is the clamp tf - t (solveras.py:637): it carries the summed absolute deviation of every dt
    # before it, each within DT_TOL relative, so its own bound is DT_TOL * tf absolute
    np.testing.assert_allclose(hs[-1], ref_h[-1], rtol=0, atol=DT_TOL * float(g[pre + "tf"]))
    np.testing.assert_allclose(np.array(sol.t), g[pre + "t"], rtol=DT_TOL, atol=0)
    assert len(sol.u) == int(g[pre + "n_snap"])
    assert rel(host(uf), g[pre + "u_final"]) < FINAL_TOL
    assert rel(host(sol.u[-1]), g[pre + "u_snap_last"]) < FINAL_TOL


def test_cfg1_readme_quickstart(rk, golden):
    """BASELINE cfg 1: KS n=1024, IF34, t 0->50, store_freq=20.  The run is chaotic, so the dt
    sequence is compared over the short horizon the north_star allows and the totals loosely."""
    g = golden("ks1024_if34_cfg1.npz")
    p = problems.ks(1024)
    sol = rk.IF34(dev(p.lin_op), fused_for(rk, p))
    uf = sol.evolve(dev(p.u0), 0.0, 50.0, store_freq=20)
    hs = np.array([r[0] for r in sol.trial_log])
    acc = np.array([r[2] for r in sol.trial_log])
    k = 150                                         # t <~ 12
    np.testing.assert_array_equal(acc[:k], g["trial_accepted"][:k])
    np.testing.assert_allclose(hs[:k], g["trial_h"][:k], rtol=1e-7)
    assert abs(len(hs) - len(g["trial_h"])) <= 30
    assert abs(len(sol.u) - int(g["n_snap"])) <= 2
    assert torch.isfinite(uf.real).all()


@pytest.mark.parametrize("method", ADAPTIVE)
def test_step_api_matches_oracle(rk, method):
    """User-driven step() loop (demos/ks.ipynb cell 14): u, h, h_suggest = solver.step(u, h)."""
    p = problems.kdv(256)
    sol = make(rk, method, p, fused_for(rk, p), 1e-5)
    ora = OracleSolver(method, p.lin_op, p.nl_func, Config(epsilon=1e-5))
    u, uo = dev(p.u0), p.u0
    h = ho = 0.025
    for _ in range(12):
        u, hu, h = sol.step(u, h)
        uo, huo, ho = ora.step(uo, ho)
        assert hu == pytest.approx(huo, rel=DT_TOL)
        assert h == pytest.approx(ho, rel=DT_TOL)
        assert rel(host(u), uo) < FINAL_TOL


def test_explosive_nonlinearity_raises_minimum_step(rk):
    """tests/test_etd35.py:111-121: an explosive N drives h below minh."""
    p = problems.kdv(64)
    sol = rk.ETD35(dev(p.lin_op), lambda v: 1e20 * v * v.abs() ** 2 + 1e30,
                   config=rk.SolverConfig(minh=1e-4))
    with pytest.raises(rk.solveras.BaseSolverAS.SolverError):
        sol.evolve(dev(p.u0), 0.0, 1.0)


def test_tf_before_t0_returns_input(rk):
    p = problems.kdv(64)
    sol = rk.ETD35(dev(p.lin_op), fused_for(rk, p))
    u0 = dev(p.u0)
    out = sol.evolve(u0, 1.0, 0.5)
    assert out is u0 and len(sol.t) == 1


def test_if45dp_r4_fix_changes_step_count(rk):
    p = problems.kdv(256)
    counts = {}
    for fix in (False, True):
        sol = rk.IF45DP(dev(p.lin_op), fused_for(rk, p), r4_fix=fix)
        sol.evolve(dev(p.u0), 0.0, 1.0, store_data=False)
        counts[fix] = len(sol.trial_log)
    assert counts[False] > 5 * counts[True]        # the shipped 17/1920 weight makes the estimate O(h)


def test_real_linop_if_coefficients_stay_real(rk):
    """if34.py:71: IF strategies do not cast lin_op; real L gives real exponentials."""
    p = problems.ks(64)
    sol = rk.IF34(dev(p.lin_op), fused_for(rk, p))
    eng = sol._get_engine(dev(p.u0))
    assert eng.coef_view("E").dtype == torch.float64
    p2 = problems.kdv(64)
    sol2 = rk.IF34(dev(p2.lin_op), fused_for(rk, p2))
    assert sol2._get_engine(dev(p2.u0)).coef_view("E").dtype == torch.complex128


# --------------------------------------------------------------------------------------------
# N-D grids (BASELINE cfg 4 / cfg 5 at reduced size): "lin_op shaped like u" against the
# reference's own formulation, which flattens lin_op/u to 1-D (demos/nls.ipynb:496-511)
# --------------------------------------------------------------------------------------------
def test_cfg4_allen_cahn_2d_if45dp_matches_flattened_oracle(rk):
    n = 64
    p = problems.allen_cahn_2d(n)
    shape = p.params["shape"]
    lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01)
    np.testing.assert_allclose(host(lin).ravel(), p.lin_op, rtol=1e-14)
    sol = rk.IF45DP(lin, nl, config=rk.SolverConfig(epsilon=1e-4))
    uf = sol.evolve(dev(p.u0.reshape(shape)), 0.0, 0.1, store_freq=5)
    ora = OracleSolver("IF45DP", p.lin_op, p.nl_func, Config(epsilon=1e-4))
    uo = ora.evolve(p.u0, 0.0, 0.1, store_freq=5)
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    np.testing.assert_allclose([r[0] for r in sol.trial_log], [r.h for r in ora.log], rtol=DT_TOL)
    np.testing.assert_allclose(sol.t, ora.t, rtol=DT_TOL)
    assert rel(host(uf).ravel(), uo) < FINAL_TOL
    assert sol.u[-1].shape == tuple(shape)


def test_cfg5_nls_3d_etd35_matches_flattened_oracle(rk):
    n = 16
    p = problems.nls_3d(n)
    k = dev(p.kx)
    lin, nl = rk.models.nls_nd_ops([k, k, k], gamma=2.0)
    np.testing.assert_allclose(host(lin).ravel(), p.lin_op, rtol=1e-14)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-5))
    uf = sol.evolve(dev(p.u0.reshape(n, n, n)), 0.0, 0.2)
    ora = OracleSolver("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-5))
    uo = ora.evolve(p.u0, 0.0, 0.2)
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    np.testing.assert_allclose([r[0] for r in sol.trial_log], [r.h for r in ora.log], rtol=DT_TOL)
    assert rel(host(uf).ravel(), uo) < FINAL_TOL


def test_batched_2d_grid_shares_one_dt(rk):
    """leading batch dim over a 2-D spectral grid: lin_op (n, n/2+1), u (B, n, n/2+1)."""
    n = 32
    p = problems.allen_cahn_2d(n)
    shape = p.params["shape"]
    lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01)
    u0 = np.stack([p.u0.reshape(shape), 0.5 * p.u0.reshape(shape)])
    sol = rk.IF34(lin, nl)
    uf = sol.evolve(dev(u0), 0.0, 0.5)
    # oracle: reference broadcast semantics with the batch flattened alongside (lin_op tiled)
    def nl_b(v):
        return np.concatenate([p.nl_func(v[: p.lin_op.size]), p.nl_func(v[p.lin_op.size:])])
    ora = OracleSolver("IF34", np.tile(p.lin_op, 2), nl_b)
    uo = ora.evolve(u0.ravel(), 0.0, 0.5)
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    assert rel(host(uf).ravel(), uo) < FINAL_TOL


# --------------------------------------------------------------------------------------------
# fused Allen-Cahn (cubic) and sine-Gordon models: oracle = the reference's solver
# logic driven by a NumPy nl_func of the same model
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [64, 128, 256, 512, 1024, 2048, 4096, 8192])      # packed, fast and TMA-staged kernels
@pytest.mark.parametrize("name", ["allen_cahn_1d", "sine_gordon"])
def test_fused_new_models_match_numpy(rk, name, n):
    p = getattr(problems, name)(n, batch=5 if n < 8192 else 300)              # 8192: several rows per persistent CTA
    sol = rk.ETD4(dev(p.lin_op), fused_for(rk, p))
    u = dev(p.u0)
    eng = sol._get_engine(u)
    eng.set_u(u)
    eng.nl(1)
    assert rel(host(eng.state_view("N1")), p.nl_func(p.u0)) < 2e-14 * np.log2(n)
    f = fused_for(rk, p)
    assert rel(host(f(u)), p.nl_func(p.u0)) < 1e-13          # the torch.fft restatement of the same model


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("batch", [1, 2, 7, 300])
def test_paired_rows_cubic_model(rk, n, batch):
    """fft_pair.cuh: the cubic model runs two rows per complex transform (n = 512 ... 4096): even / odd batches, a
    single row, several pair groups per persistent CTA, and the in-place evaluation the N-D grids use (RowNL)."""
    p = problems.allen_cahn_1d(n, batch=batch, seed=n + batch)
    u0 = p.u0.reshape(batch, -1) * (1.0 + np.arange(batch))[:, None]       # rows of different size share a transform
    ref = p.nl_func(u0)
    sol = rk.ETD4(dev(p.lin_op), fused_for(rk, p))
    u = dev(u0)
    eng = sol._get_engine(u)
    eng.set_u(u)
    eng.nl(1)
    got = host(eng.state_view("N1"))
    scale = np.linalg.norm(ref, axis=1).max()
    for b in range(batch):                      # per row: the cross-talk is relative to the larger row of a pair
        assert np.linalg.norm(got[b] - ref[b]) < 2e-14 * np.log2(n) * scale
    assert rel(got, ref) < 2e-14 * np.log2(n)
    rows = rk.models.RowNL(3, n, None, -1.0, "cuda")
    work = u.clone()
    rows(work, out=work)
    np.testing.assert_array_equal(host(work), got)


@pytest.mark.parametrize("path", ["fused", "callable"])
@pytest.mark.parametrize("name,method,tf,eps", [("allen_cahn_1d", "IF45DP", 0.5, 1e-4), ("allen_cahn_1d", "ETD35", 3.0, 1e-5),
                                                ("sine_gordon", "ETD35", 2.0, 1e-6), ("sine_gordon", "IF34", 2.0, 1e-5)])
def test_new_models_adaptive_parity(rk, name, method, tf, eps, path):
    p = getattr(problems, name)(512, batch=3)
    nl = fused_for(rk, p) if path == "fused" else torch_callable(rk, p)
    sol = make(rk, method, p, nl, eps)
    uf = sol.evolve(dev(p.u0), 0.0, tf, store_freq=4)
    ora = OracleSolver(method, p.lin_op, p.nl_func, Config(epsilon=eps))
    uo = ora.evolve(p.u0, 0.0, tf, store_freq=4)
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    np.testing.assert_allclose([r[0] for r in sol.trial_log], [r.h for r in ora.log], rtol=DT_TOL)
    np.testing.assert_allclose(sol.t, ora.t, rtol=DT_TOL)
    assert rel(host(uf), uo) < FINAL_TOL


def test_sine_gordon_breather_returns_after_one_period(rk):
    """physics check of the first-order restatement: a breather of frequency w is 2 pi / w periodic."""
    n, w = 1024, 0.5
    p = problems.sine_gordon(n)
    lin, nl = rk.models.sine_gordon_ops(dev(np.sqrt(p.kx ** 2 - 1.0) * np.sign(np.fft.fftfreq(n))))
    np.testing.assert_allclose(host(lin), p.lin_op, rtol=1e-13)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-8))
    uf = sol.evolve(dev(p.u0), 0.0, 2 * np.pi / w, store_data=False)
    assert rel(host(uf), p.u0) < 1e-5


def test_cfg2b_independent_dt_matches_per_trajectory_reference_runs(rk):
    """cfg 2b: every trajectory keeps its own dt sequence == B separate reference runs (subset of 8)."""
    from rkstiff_b200.ensemble import evolve_independent
    p = problems.nls(512, batch=8, seed=2, half_width=20.0)
    lin, nl = rk.models.nls_ops(dev(p.kx), 2.0)
    uf, logs = evolve_independent(lambda: rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-6)), dev(p.u0), 0.0, 0.2)
    lengths = set()
    for b in range(8):
        ora = OracleSolver("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-6))
        uo = ora.evolve(p.u0[b], 0.0, 0.2)
        assert [r[2] for r in logs[b]] == [r.accepted for r in ora.log]
        np.testing.assert_allclose([r[0] for r in logs[b]], [r.h for r in ora.log], rtol=DT_TOL)
        assert rel(host(uf[b]), uo) < FINAL_TOL
        lengths.add(tuple(round(r.h, 12) for r in ora.log))
    assert len(lengths) > 1                     # the trajectories really take different dt sequences


INDEP_CASES = [
    ("ETD35", "nls", 512), ("ETD34", "ks", 1024), ("IF34", "kdv", 256), ("IF45DP", "nls", 128),
    ("ETD35", "burgers", 2048), ("ETD34", "nls", 64),        # n < 512: generic shared-memory NL kernel
]


@pytest.mark.parametrize("method,prob,n", INDEP_CASES, ids=[f"{m}-{q}{n}" for m, q, n in INDEP_CASES])
def test_cfg2b_batched_independent_dt_matches_per_trajectory_oracle(rk, method, prob, n):
    """cfg 2b in one plan: per-trajectory control blocks, coefficient arrays and buffer roles, stepped by
    one set of launches, reproduce B separate reference runs (accept/reject sequence, dt, final state)."""
    B = 6
    if prob == "nls":
        p = problems.nls(n, batch=B, seed=3, half_width=20.0)
        lin, nl = rk.models.nls_ops(dev(p.kx), 2.0)
        tf = 0.15
    elif prob == "ks":
        p = problems.ks(n, batch=B, seed=3)
        lin, nl = rk.models.ks_ops(dev(p.kx))
        tf = 2.0
    elif prob == "kdv":
        p = problems.kdv(n, batch=B, seed=3)
        lin, nl = rk.models.kdv_ops(dev(p.kx))
        tf = 0.2
    else:
        p = problems.burgers(n, mu=0.01, batch=B, seed=3)
        lin, nl = rk.models.burgers_ops(dev(p.kx), 0.01)
        tf = 0.1
    # rows of different amplitude so that the controllers really diverge
    scale = np.linspace(0.6, 1.4, B).reshape(B, 1)
    u0 = p.u0 * scale
    eps = 1e-7 if prob == "kdv" else 1e-6
    sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=eps))
    uf, logs = sol.evolve_independent(dev(u0), 0.0, tf)
    seqs = set()
    for b in range(B):
        ora = OracleSolver(method, p.lin_op, p.nl_func, Config(epsilon=eps))
        uo = ora.evolve(u0[b], 0.0, tf)
        assert [r[2] for r in logs[b]] == [r.accepted for r in ora.log], f"row {b}"
        hs, ho = np.array([r[0] for r in logs[b]]), np.array([r.h for r in ora.log])
        np.testing.assert_allclose(hs[:-1], ho[:-1], rtol=DT_TOL)
        # the last dt is tf - t: a difference that carries the summed absolute deviation of all dts before it
        np.testing.assert_allclose(hs[-1], ho[-1], rtol=0, atol=DT_TOL * tf)
        assert rel(host(uf[b]), uo) < FINAL_TOL, f"row {b}"
        seqs.add(tuple(round(r.h, 12) for r in ora.log))
    assert len(seqs) > 1
    # a second evolve on the same plan restarts every row cleanly
    uf2, logs2 = sol.evolve_independent(dev(u0), 0.0, tf)
    assert torch.equal(uf, uf2) and logs == logs2


def test_cfg2b_batched_independent_failure_status_surfaces(rk):
    """One trajectory hitting the minimum step makes the whole call raise the reference's exception."""
    p = problems.nls(256, batch=3, seed=1, half_width=20.0)
    lin, nl = rk.models.nls_ops(dev(p.kx), 2.0)
    sol = rk.ETD35(lin, nl, config=rk.SolverConfig(epsilon=1e-12, minh=1e-3))
    with pytest.raises(sol.MinimumStepReached):
        sol.evolve_independent(dev(p.u0), 0.0, 1.0, h_init=0.5)


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_axis_fft_kernels_match_torch_fft(rk, n):
    """rks_axis_*: hand-written transform along a strided axis == torch.fft.ifft / fft along that axis
    (the inverse leaves the axis digit-reversed; the permutation is read off a ramp)."""
    ax = rk.models.AxisFFT(n, "cuda")
    outer, inner = 3, (13 if n <= 512 else 6)
    g = torch.Generator(device="cpu").manual_seed(n)
    x = torch.randn(outer, n, inner, 2, generator=g, dtype=torch.float64)
    x = torch.view_as_complex(x).cuda().contiguous()
    ramp = torch.fft.fft(torch.arange(n, dtype=torch.float64, device="cuda")).to(torch.complex128)
    ramp = ramp[None, :, None].expand(1, n, 8).contiguous()
    perm = torch.round(ax.inverse_(ramp, 1)[0, :, 0].real).long()
    assert sorted(perm.tolist()) == list(range(n))
    y = ax.inverse_(x.clone(), 1)
    ref = torch.fft.ifft(x, dim=1)
    tol = 1e-13 * float(ref.abs().max()) * np.log2(n)
    assert float((y - ref[:, perm, :]).abs().max()) < tol
    out = torch.empty_like(x)
    assert ax.inverse_(x, 1, out=out) is out and torch.equal(out, y)          # out-of-place variant, input untouched
    z = ax.forward_(y.clone(), 1)
    assert float((z - x).abs().max()) < 1e-13 * float(x.abs().max()) * np.log2(n)
    # the same axis delivered in G row blocks, block-major (slab-decomposition layout)
    for G in (2, 8):
        blocked = x.reshape(outer, G, n // G, inner).permute(1, 0, 2, 3).contiguous()
        got = ax.chunked_(blocked.clone(), True)
        want = y.reshape(outer, G, n // G, inner).permute(1, 0, 2, 3)
        assert torch.equal(got, want.contiguous())
        assert float((ax.chunked_(got, False) - blocked).abs().max()) < 1e-13 * float(x.abs().max()) * np.log2(n)
    # 3-D array, middle axis and leading axis
    if n <= 256:
        w = torch.view_as_complex(torch.randn(n, n, 24, 2, generator=g, dtype=torch.float64)).cuda().contiguous()
        v = ax.inverse_(ax.inverse_(w.clone(), 0), 1)
        refw = torch.fft.ifftn(w, dim=(0, 1))
        assert float((v - refw[perm][:, perm]).abs().max()) < 1e-13 * float(refw.abs().max()) * 2 * np.log2(n)
        back = ax.forward_(ax.forward_(v, 1), 0)
        assert float((back - w).abs().max()) < 1e-12 * float(w.abs().max())



# --------------------------------------------------------------------------------------------
# diagonalize=True: dense lin_op, eigenbasis stepping against the reference goldens
# --------------------------------------------------------------------------------------------
DIAG_CASES = [("IF34", 1e-6), ("ETD34", 1e-6), ("ETD35", 1e-6), ("ETD35", 1e-9)]


@pytest.mark.parametrize("method,eps", DIAG_CASES)
def test_diagonalized_dense_operator_matches_reference(rk, golden, method, eps):
    g = golden("diagonalized_runs.npz")
    p = problems.dense_advection_diffusion()
    pre = f"{method}_{eps:g}_"
    tf = float(g[pre + "tf"])
    sol = getattr(rk, method)(dev(p.lin_op), lambda u: u - u ** 3, config=rk.SolverConfig(epsilon=eps), diagonalize=True)
    uf = sol.evolve(dev(p.u0), 0.0, tf, store_freq=3)
    hs = np.array([r[0] for r in sol.trial_log])
    acc = np.array([r[2] for r in sol.trial_log])
    np.testing.assert_array_equal(acc, g[pre + "trial_accepted"])
    # eps = 1e-9 puts the error estimate nine digits below u while the eigenvector matrix (cond ~ 1e2) amplifies
    # the rounding of every S / S^-1 product: the estimate itself only carries ~5 digits there, so dt (a fourth
    # root of it) is compared at 1e-7; the accept/reject sequence and the states keep the usual bars
    dt_tol = DT_TOL if eps >= 1e-6 else 1e-7
    np.testing.assert_allclose(hs[:-1], g[pre + "trial_h"][:-1], rtol=dt_tol, atol=0)
    np.testing.assert_allclose(hs[-1], g[pre + "trial_h"][-1], rtol=0, atol=dt_tol * tf)      # the clamp tf - t
    np.testing.assert_allclose(np.array(sol.t), g[pre + "t"], rtol=dt_tol, atol=0)
    assert len(sol.u) == int(g[pre + "n_snap"])
    assert rel(host(uf), g[pre + "u_final"]) < FINAL_TOL
    # a stored snapshot sits at a time that itself moved by dt_tol: allow for that shift in the tight case
    assert rel(host(sol.u[-1]), g[pre + "u_snap_last"]) < (FINAL_TOL if eps >= 1e-6 else 100 * dt_tol)
    assert sol.u[0].shape == (p.u0.shape[0],)


def test_diagonalized_step_api_and_errors(rk, caplog):
    from oracle.rk_oracle import OracleDiagonalized
    p = problems.dense_advection_diffusion()
    sol = rk.ETD35(dev(p.lin_op), lambda u: u - u ** 3, config=rk.SolverConfig(epsilon=1e-6), diagonalize=True)
    ora = OracleDiagonalized("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-6))
    u, uo, h, ho = dev(p.u0), p.u0.copy(), 0.01, 0.01
    for _ in range(6):
        u, hu, h = sol.step(u, h)
        uo, huo, ho = ora.step(uo, ho)
        assert hu == pytest.approx(huo, rel=DT_TOL) and h == pytest.approx(ho, rel=DT_TOL)
        assert rel(host(u), uo) < FINAL_TOL
    with pytest.raises(ValueError):
        rk.ETD34(dev(np.array([[1.0, 2.0], [2.0, 4.0]])), lambda v: v, diagonalize=True)      # singular
    with pytest.raises(ValueError):
        rk.IF34(dev(np.ones((3, 4))), lambda v: v, diagonalize=True)                            # not square
    import logging
    with caplog.at_level(logging.WARNING):
        ill = np.array([[1.0, 1e4], [0.0, 2.0]])
        rk.ETD35(dev(ill), lambda v: v, diagonalize=True)
    assert any("condition number" in r.getMessage() for r in caplog.records)


def test_reference_allen_cahn_chebyshev_test_through_diagonalize(rk):
    """The reference's dense-operator test (tests/test_etd35.py:36-48, its matrix-exponential path) run through
    diagonalize=True: same physical assertions, and the run equals the oracle's diagonalized strategy."""
    from oracle.rk_oracle import OracleDiagonalized
    p = problems.allen_cahn_cheb(20)
    x, d = rk.grids.construct_x_dx_cheb(20, -1.0, 1.0)
    lin_dev, nl = rk.models.allen_cahn_ops(x, d, 0.01)
    assert float((lin_dev - dev(p.lin_op)).abs().max()) < 1e-10
    sol = rk.ETD35(dev(p.lin_op), nl, config=rk.SolverConfig(epsilon=1e-4),
                   etd_config=rk.ETDConfig(contour_points=32, contour_radius=10), diagonalize=True)
    wf = sol.evolve(dev(p.u0), 0.0, 60.0, store_data=False)
    uf = host(wf).real + p.kx
    u0int = p.params["u0int"]
    assert abs(u0int[0] - uf[0]) < 0.01 and abs(u0int[7] - uf[7]) > 1
    ora = OracleDiagonalized("ETD35", p.lin_op, p.nl_func, Config(epsilon=1e-4, contour_points=32, contour_radius=10.0))
    wo = ora.evolve(p.u0.copy(), 0.0, 60.0, store_data=False)
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    # t = 60 of metastable front dynamics amplifies roundoff along the way (the reference's own test only checks
    # two grid values): dt is compared tightly over the first 40 trials, loosely over the whole run
    hs, ho = np.array([r[0] for r in sol.trial_log]), np.array([r.h for r in ora.log])
    np.testing.assert_allclose(hs[:40], ho[:40], rtol=1e-8)
    np.testing.assert_allclose(hs[:-1], ho[:-1], rtol=1e-4)
    assert rel(host(wf), wo) < 1e-5



# --------------------------------------------------------------------------------------------
# pre-transformed intermediate stages (complex-field fused model, n = 512 ... 8192): K1 applies the first inverse
# FFT pass (stage_pre_kernel), K4 starts one pass later (nl_fast_pre_kernel).  DESIGN.md section 4.
# --------------------------------------------------------------------------------------------
PT_SIZES = [512, 1024, 2048, 4096, 8192]


def _stage_outputs(rk, method, p, monkeypatch, pt):
    """N_2 ... N_S of one trial at h = 0.004 through rks_stage_nl, with the pre-transforming pair on or off."""
    monkeypatch.setenv("RKS_PT", "1" if pt else "0")
    sol = make(rk, method, p, fused_for(rk, p))
    u = dev(p.u0)
    eng = sol._get_engine(u)
    adaptive = method in ADAPTIVE
    eng.begin(0.0, 1.0, 0.004, 0, not adaptive)
    if not adaptive:
        eng.ensure_fixed_coeffs(0.004)
    eng.set_u(u)
    eng.update_coeffs()
    eng.nl(1)
    outs = []
    for s in range(1, eng.stages):
        rk._abi.check(rk._abi.lib.rks_stage_nl(eng.plan, s, eng.st))
        outs.append(host(eng.state_view(f"N{s + 1}")).copy())
    return outs, sol


@pytest.mark.parametrize("n", PT_SIZES)
@pytest.mark.parametrize("method", METHODS)
def test_pretransformed_pair_equals_plain_pair(rk, method, n, monkeypatch):
    """Every intermediate N of every method, batch 5 (a ragged last CTA for every rows-per-CTA): same operations
    in the same order, so only the compiler's FMA contraction may differ between the two routes."""
    p = problems.nls(n, batch=5, seed=n, half_width=20.0)
    plain, _ = _stage_outputs(rk, method, p, monkeypatch, pt=False)
    pre, sol = _stage_outputs(rk, method, p, monkeypatch, pt=True)
    assert len(pre) == len(plain) > 0
    for a, b in zip(pre, plain):
        assert rel(a, b) < 1e-14
    # the two kernels rks_stage_nl launches, separately
    eng = sol._engine
    for which in (1, 2):
        rk._abi.check(rk._abi.lib.rks_stage_nl_part(eng.plan, 1, which, eng.st))
    assert rel(host(eng.state_view("N2")), plain[0]) < 1e-14
    assert rk._abi.lib.rks_stage_nl_part(eng.plan, 1, 0, eng.st) != 0          # part must be 1 or 2


@pytest.mark.parametrize("method,n,batch", [("ETD35", 8192, 301), ("ETD4", 8192, 301), ("IF45DP", 4096, 700),
                                            ("ETD35", 2048, 1), ("IF34", 1024, 2)])
def test_pretransformed_pair_ragged_batches(rk, method, n, batch, monkeypatch):
    """Batches the persistent stage kernel walks through several rows per warp (its row counters are re-armed by the
    last warp of every launch: stages 1..S-1 run back to back here), and batches smaller than one CTA."""
    p = problems.nls(n, batch=batch, seed=batch, half_width=20.0)
    plain, _ = _stage_outputs(rk, method, p, monkeypatch, pt=False)
    pre, sol = _stage_outputs(rk, method, p, monkeypatch, pt=True)
    for a, b in zip(pre, plain):
        assert rel(a, b) < 1e-14
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()          # no row skipped or taken twice
    eng = sol._engine
    for _ in range(3):                                                  # the same launch again: counters back at zero
        rk._abi.check(rk._abi.lib.rks_stage_nl(eng.plan, 1, eng.st))
    assert rel(host(eng.state_view("N2")), plain[0]) < 1e-14


@pytest.mark.parametrize("n", PT_SIZES)
@pytest.mark.parametrize("method", FIXED)
def test_pretransformed_fixed_step_parity(rk, method, n):
    p = problems.nls(n, batch=5, seed=n + 1, half_width=20.0)
    sol = make(rk, method, p, fused_for(rk, p))
    u, h = dev(p.u0), 0.004
    for _ in range(3):
        ref = OracleSolver(method, p.lin_op, p.nl_func).step(host(u), h)
        sol.reset()
        u = sol.step(u, h)
        assert rel(host(u), ref) < STEP_TOL
    # and a run of consecutive steps through the captured graph (FSAL / N1 carried over)
    uf = sol.evolve(dev(p.u0), 0.0, 0.02, h, store_data=False)
    ora = OracleSolver(method, p.lin_op, p.nl_func)
    assert rel(host(uf), ora.evolve(p.u0, 0.0, 0.02, h, store_data=False)) < 10 * STEP_TOL


@pytest.mark.parametrize("n", [512, 2048, 8192])
@pytest.mark.parametrize("method", ADAPTIVE)
def test_pretransformed_adaptive_parity(rk, method, n):
    p = problems.nls(n, batch=3, seed=n + 2, half_width=20.0)
    eps, tf = 1e-6, (0.004 if method == "IF45DP" else 0.08)
    sol = make(rk, method, p, fused_for(rk, p), eps)
    uf = sol.evolve(dev(p.u0), 0.0, tf, store_data=False)
    ora = OracleSolver(method, p.lin_op, p.nl_func, Config(epsilon=eps))
    uo = ora.evolve(p.u0, 0.0, tf, store_data=False)
    hs, acc = np.array([r[0] for r in sol.trial_log]), [r[2] for r in sol.trial_log]
    ho = np.array([r.h for r in ora.log])
    assert acc == [r.accepted for r in ora.log]
    np.testing.assert_allclose(hs[:-1], ho[:-1], rtol=DT_TOL, atol=0)
    np.testing.assert_allclose(hs[-1], ho[-1], rtol=0, atol=DT_TOL * tf)
    assert rel(host(uf), uo) < FINAL_TOL


# --------------------------------------------------------------------------------------------
# csrc/fft_real.cuh: half-length forward transform for the real-field models.  Default for n = 512 and 1024 (measured
# 7 % faster there, slower for longer rows: profiles/r02_k4_ab.md); RKS_RFFT_HALF=1 / 0 forces it on (n <= 4096) / off.
# Both settings are checked for every row length the kernel exists for.
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("setting", ["1", "0"])
@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
@pytest.mark.parametrize("batch", [1, 5, 150])
@pytest.mark.parametrize("name", ["ks", "allen_cahn_1d"])
def test_rfft_half_length_forward_matches_full_length(rk, name, n, batch, setting, monkeypatch):
    monkeypatch.setenv("RKS_RFFT_HALF", setting)
    p = problems.ks(n, batch=batch, seed=n) if name == "ks" else problems.allen_cahn_1d(n, batch=batch, seed=n)
    sol = rk.ETD4(dev(p.lin_op), fused_for(rk, p))
    u = dev(p.u0)
    eng = sol._get_engine(u)
    eng.set_u(u)
    eng.nl(1)
    assert rel(host(eng.state_view("N1")), p.nl_func(p.u0)) < 2e-14 * np.log2(n)


# --------------------------------------------------------------------------------------------
# coefficient storage of large grids: records per distinct lin_op value ("indexed") and per-axis
# exponential tables ("separable") against full-size arrays and against the flattened oracle
# --------------------------------------------------------------------------------------------
def _grid_case(rk, kind, n):
    if kind == "allen_cahn_2d":
        p = problems.allen_cahn_2d(n)
        lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01)
        return p, lin, nl, p.params["shape"]
    p = problems.nls_3d(n)
    k = dev(p.kx)
    lin, nl = rk.models.nls_nd_ops([k, k, k], gamma=2.0)
    return p, lin, nl, (n, n, n)


GRID_CASES = [("IF45DP", "allen_cahn_2d", 64, "separable", 0.05, 1e-4), ("IF34", "allen_cahn_2d", 64, "separable", 0.5, 1e-5),
              ("IF45DP", "allen_cahn_2d", 32, "indexed", 0.05, 1e-4), ("ETD34", "allen_cahn_2d", 64, "indexed", 0.5, 1e-5),
              ("ETD35", "nls_3d", 16, "indexed", 0.2, 1e-5), ("IF34", "nls_3d", 16, "separable", 0.1, 1e-5),
              ("IF45DP", "nls_3d", 16, "separable", 0.02, 1e-5), ("IF34", "nls_3d", 16, "indexed", 0.1, 1e-5)]


@pytest.mark.parametrize("method,kind,n,storage,tf,eps", GRID_CASES, ids=[f"{c[0]}-{c[1]}{c[2]}-{c[3]}" for c in GRID_CASES])
def test_grid_coefficient_storage_matches_flattened_oracle(rk, method, kind, n, storage, tf, eps):
    p, lin, nl, shape = _grid_case(rk, kind, n)
    sol = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=eps))
    sol.coef_storage = storage
    uf = sol.evolve(dev(p.u0.reshape(shape)), 0.0, tf, store_freq=3)
    assert sol._engine.coef_storage == storage
    ora = OracleSolver(method, p.lin_op, p.nl_func, Config(epsilon=eps))
    uo = ora.evolve(p.u0, 0.0, tf, store_freq=3)
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    np.testing.assert_allclose([r[0] for r in sol.trial_log], [r.h for r in ora.log], rtol=DT_TOL)
    np.testing.assert_allclose(sol.t, ora.t, rtol=DT_TOL)
    assert rel(host(uf).ravel(), uo) < FINAL_TOL


@pytest.mark.parametrize("method", ["ETD4", "ETD5", "IF4", "ETD34", "ETD35", "IF34", "IF45DP"])
def test_indexed_records_are_bit_identical_to_full_arrays(rk, method):
    """K2 evaluates the same formulas on the same z: gathering a record must give the bits of the full-size arrays."""
    p, lin, nl, shape = _grid_case(rk, "allen_cahn_2d", 32)
    lin = lin.to(torch.complex128) * (1.0 + 0.3j) if method.startswith("ETD") else lin
    u0 = dev(np.stack([p.u0.reshape(shape), 0.7 * p.u0.reshape(shape), -0.4 * p.u0.reshape(shape)]))
    outs = {}
    for storage in ("arrays", "indexed"):
        sol = getattr(rk, method)(lin, nl)
        sol.coef_storage = storage
        if method in ADAPTIVE:
            outs[storage] = (host(sol.evolve(u0, 0.0, 0.05, store_data=False)), [r[:3] for r in sol.trial_log])
        else:
            u = u0
            for _ in range(3):
                u = sol.step(u, 0.01)
            outs[storage] = (host(u), None)
        assert sol._engine.coef_storage == storage
    np.testing.assert_array_equal(outs["arrays"][0], outs["indexed"][0])
    assert outs["arrays"][1] == outs["indexed"][1]


@pytest.mark.parametrize("method", ["IF4", "IF34", "IF45DP"])
@pytest.mark.parametrize("kind,n", [("allen_cahn_2d", 32), ("nls_3d", 16)])
def test_separable_tables_agree_with_full_arrays(rk, method, kind, n):
    """products of per-axis exponentials vs exp of the summed exponent: one rounding per factor"""
    p, lin, nl, shape = _grid_case(rk, kind, n)
    u0 = dev(np.stack([p.u0.reshape(shape), 0.5 * p.u0.reshape(shape)]))
    outs = {}
    for storage in ("arrays", "separable"):
        sol = getattr(rk, method)(lin, nl)
        sol.coef_storage = storage
        if method in ADAPTIVE:
            outs[storage] = host(sol.evolve(u0, 0.0, 0.02, store_data=False))
            outs[storage + "_log"] = [r[2] for r in sol.trial_log]
        else:
            outs[storage] = host(sol.step(u0, 0.01))
        assert sol._engine.coef_storage == storage
    assert rel(outs["separable"], outs["arrays"]) < 1e-13
    assert outs.get("separable_log") == outs.get("arrays_log")


def test_non_separable_operator_is_detected(rk):
    from rkstiff_b200._engine import distinct_values, separable_terms
    a = torch.linspace(0, 1, 7, dtype=torch.float64, device="cuda")
    sep = 1.0 - a[:, None] ** 2 - 3 * a[None, :5]
    assert separable_terms(sep) is not None
    assert separable_terms(sep * (1 + a[:, None] * a[None, :5])) is None
    vals, idx = distinct_values((a[:, None] ** 2 + a[None, :] ** 2).to(torch.complex128) * 1j)
    assert vals.numel() < 49 and idx.dtype == torch.int32
    np.testing.assert_array_equal(host(vals)[host(idx)].reshape(7, 7), host((a[:, None] ** 2 + a[None, :] ** 2) * 1j))
    sol = rk.IF34(sep * (1 + a[:, None] * a[None, :5]), lambda v: v)
    sol.coef_storage = "separable"
    with pytest.raises(ValueError):
        sol.step(torch.zeros(7, 5, dtype=torch.complex128, device="cuda"), 0.1)


def test_step_recognises_only_the_tensor_it_returned(rk):
    """a new tensor (even at a recycled address) must be copied in, not mistaken for the previous output.
    Same call sequence on both sides: like the reference, step() keeps N1 = N(previous output) (etd4.py:174)."""
    p = problems.ks(256, batch=2)
    sol = rk.ETD4(dev(p.lin_op), fused_for(rk, p))
    ora = OracleSolver("ETD4", p.lin_op, p.nl_func)
    u1 = sol.step(dev(p.u0), 0.05)
    o1 = ora.step(p.u0, 0.05)
    ref = ora.step(0.5 * o1, 0.05)
    u_new = (0.5 * u1).clone()
    del u1
    again = torch.empty_like(u_new)          # the caching allocator may hand the freed block out again
    again.copy_(u_new)
    got = sol.step(again, 0.05)
    assert rel(host(got), ref) < 1e-11
    # and the tensor step() returned is recognised (no copy, same result as feeding it back)
    nxt = sol.step(got, 0.05)
    assert rel(host(nxt), ora.step(ref, 0.05)) < 1e-11


# --------------------------------------------------------------------------------------------
# N-D grid models as engine models (rks_set_model_nd): the trial loop runs without a host sync per trial
# --------------------------------------------------------------------------------------------
def _grid_handles(rk, kind, n):
    if kind == "nls_2d":
        p = problems.nls(n, half_width=6.0)
        k = dev(p.kx)
        lin, nl = rk.models.nls_nd_ops([k, k], gamma=2.0)
        x = p.x
        u0 = np.fft.fft2(np.exp(-(x[:, None] ** 2 + x[None, :] ** 2)) * (1.0 + 0.0j))
        return lin, nl, dev(u0)
    p, lin, nl, shape = _grid_case(rk, kind, n)
    return lin, nl, dev(p.u0.reshape(shape))


@pytest.mark.parametrize("method,kind,n,tf", [("IF45DP", "allen_cahn_2d", 64, 0.05), ("ETD35", "nls_3d", 16, 0.2),
                                              ("ETD35", "nls_2d", 64, 0.1), ("IF34", "allen_cahn_2d", 128, 0.5),
                                              ("ETD34", "nls_3d", 32, 0.05)])
def test_grid_model_in_engine_equals_python_composition(rk, method, kind, n, tf):
    """the engine launching the axis / row kernels itself (32-trial graph-replayed chunks) == the same kernels
    composed from Python with one control-block read per trial: identical decisions, identical bits"""
    lin, nl, u0 = _grid_handles(rk, kind, n)
    assert isinstance(nl, rk.models.FusedGridNL)
    a = getattr(rk, method)(lin, nl, config=rk.SolverConfig(epsilon=1e-5))
    ua = a.evolve(u0, 0.0, tf, store_freq=2)
    b = getattr(rk, method)(lin, lambda v, out=None: nl(v), config=rk.SolverConfig(epsilon=1e-5))
    ub = b.evolve(u0, 0.0, tf, store_freq=2)
    assert len(a.trial_log) > 3 and a.trial_log == b.trial_log
    assert a.t == b.t
    np.testing.assert_array_equal(host(ua), host(ub))
    for x, y in zip(a.u[1:], b.u[1:]):
        np.testing.assert_array_equal(host(x), host(y))
    # far fewer host syncs: launches per trial are the same, but the fused run reads the control block per chunk
    assert a._engine.fused is not None and b._engine.fused is None


def test_grid_model_fixed_step_and_batch(rk):
    lin, nl, u0 = _grid_handles(rk, "allen_cahn_2d", 32)
    u0 = torch.stack([u0, 0.5 * u0])
    a = rk.ETD4(lin, nl)
    ua = a.evolve(u0, 0.0, 0.1, 0.01, store_freq=5)
    b = rk.ETD4(lin, lambda v, out=None: nl(v))
    ub = b.evolve(u0, 0.0, 0.1, 0.01, store_freq=5)
    np.testing.assert_array_equal(host(ua), host(ub))
    assert a.t == b.t and len(a.u) == len(b.u)


def test_non_power_of_two_grid_falls_back_to_a_callable(rk):
    n = 24
    p = problems.allen_cahn_2d(n)
    lin, nl = rk.models.allen_cahn_fourier_ops(n, eps=0.01)
    assert not isinstance(nl, rk.models.FusedGridNL)
    sol = rk.IF34(lin, nl)
    uf = sol.evolve(dev(p.u0.reshape(p.params["shape"])), 0.0, 0.3)
    ora = OracleSolver("IF34", p.lin_op, p.nl_func)
    uo = ora.evolve(p.u0, 0.0, 0.3)
    assert [r[2] for r in sol.trial_log] == [r.accepted for r in ora.log]
    assert rel(host(uf).ravel(), uo) < FINAL_TOL
