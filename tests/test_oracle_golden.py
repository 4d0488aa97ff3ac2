"""Pin the NumPy oracle (oracle/rk_oracle.py) to the reference's own outputs.

The fixtures under tests/golden/ were produced by oracle/make_golden.py running the
unmodified reference; the literals in test_survey_fingerprints are the values SURVEY.md 8c
recorded from the reference.  CPU only.
"""
import numpy as np
import pytest

from oracle import problems
from oracle.rk_oracle import ADAPTIVE, FIXED, Config, OracleSolver, coefficients


def rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(np.asarray(b))


def test_survey_fingerprints(golden):
    g = golden("ks1024_one_trial.npz")
    p = problems.ks(1024)
    assert np.linalg.norm(p.u0) == pytest.approx(5.724334022399462e+02, rel=1e-14)
    expect = {"IF34": (5.742325855605143e+02, 4.525060844081852e-04, 4.458496176118095),
              "ETD34": (5.742325852568775e+02, 4.404481948578522e-04, 4.402598668872002),
              "ETD35": (5.742325864471438e+02, 1.466407638290303e-03, 3.459158196158707),
              "IF45DP": (5.742325863716673e+02, 4.422257954306392e-01, 0.5373225147366475)}
    for m, (nk, ne, s) in expect.items():
        sol = OracleSolver(m, p.lin_op, p.nl_func)
        k, err = sol.trial(p.u0, 0.5)
        assert np.linalg.norm(k) == pytest.approx(nk, rel=1e-13)
        assert np.linalg.norm(err) == pytest.approx(ne, rel=1e-9)
        assert sol.compute_s(k, err) == pytest.approx(s, rel=1e-9)
        assert rel(k, g[f"{m}_k"]) < 1e-15
        assert rel(err, g[f"{m}_err"]) < 1e-12
    for m, nk in {"ETD4": 5.726108053426919e+02, "ETD5": 5.726108053426983e+02, "IF4": 5.726108053426951e+02}.items():
        k = OracleSolver(m, p.lin_op, p.nl_func).step(p.u0, 0.05)
        assert np.linalg.norm(k) == pytest.approx(nk, rel=1e-14)
        assert rel(k, g[f"{m}_k"]) < 1e-15


@pytest.mark.parametrize("tag", ["ks", "nls", "kdv"])
@pytest.mark.parametrize("method", ADAPTIVE + FIXED)
def test_coefficients_match_reference(golden, tag, method):
    g = golden("coefficients.npz")
    lin, h = g[f"{tag}_lin_op"], float(g[f"{tag}_h"])
    c = coefficients(method, lin, h, Config())
    alias = {"EL": "E", "EL2": "E2", "EL14": "E14", "EL12": "E12", "EL34": "E34", "EL15": "E15",
             "EL310": "E310", "EL45": "E45", "EL89": "E89", "b1": None, "b2": None, "b3": None,
             "b4": None, "b5": None, "b6": None}
    bmap = {"ETD4": {"b1": "a51", "b2": "a52", "b4": "a54"},
            "ETD5": {"b1": "a71", "b3": "a73", "b4": "a74", "b5": "a75", "b6": "a76"}}
    checked = 0
    pre = f"{tag}_{method}_"
    for key, ref in g.items():
        if not key.startswith(pre):
            continue
        name = key[len(pre):]
        mine = bmap.get(method, {}).get(name) or alias.get(name, name) or name
        if method in ("IF4", "IF34") and mine not in ("E", "E2"):
            continue
        assert mine in c, (key, mine)
        np.testing.assert_array_equal(np.asarray(c[mine]), ref, err_msg=key)   # same NumPy expressions => bitwise
        checked += 1
    assert checked >= 2


@pytest.mark.parametrize("tag,builder", [("kdv", lambda: problems.kdv(256)), ("ks", lambda: problems.ks(256)),
                                         ("burgers", lambda: problems.burgers(256, mu=0.01)),
                                         ("nls", lambda: problems.nls(256, half_width=20.0)),
                                         ("ksb", lambda: problems.ks(128, batch=3, seed=0))])
@pytest.mark.parametrize("method", FIXED)
def test_fixed_runs_match_reference(golden, tag, builder, method):
    g = golden("fixed_runs.npz")
    p = builder()
    pre = f"{tag}_{method}_"
    np.testing.assert_array_equal(p.u0, g[pre + "u0"])
    h, tf = float(g[pre + "h"]), float(g[pre + "tf"])
    steps = int(g[pre + "steps"])
    sol = OracleSolver(method, p.lin_op, p.nl_func)
    np.testing.assert_array_equal(sol.step(p.u0, h), g[pre + "u_step1"])
    sol.reset()
    uf = sol.evolve(p.u0, 0.0, tf, h, store_freq=max(1, steps // 4))
    np.testing.assert_array_equal(uf, g[pre + "u_final"])
    np.testing.assert_array_equal(np.array(sol.t), g[pre + "t"])
    assert len(sol.u) == int(g[pre + "n_snap"])


CASES = {"kdv": lambda: problems.kdv(256), "ks": lambda: problems.ks(256),
         "burgers": lambda: problems.burgers(256, mu=0.01), "nls": lambda: problems.nls(256, half_width=20.0),
         "nlsb": lambda: problems.nls(128, batch=3, seed=2, half_width=20.0),
         "ksb": lambda: problems.ks(128, batch=4, seed=0)}


@pytest.mark.parametrize("tag", list(CASES))
@pytest.mark.parametrize("method", ADAPTIVE)
def test_adaptive_runs_match_reference(golden, tag, method):
    g = golden("adaptive_runs.npz")
    p = CASES[tag]()
    pre = f"{tag}_{method}_"
    np.testing.assert_array_equal(p.u0, g[pre + "u0"])
    h0 = float(g[pre + "h_init"])
    sol = OracleSolver(method, p.lin_op, p.nl_func, Config(epsilon=float(g[pre + "epsilon"])))
    uf = sol.evolve(p.u0, 0.0, float(g[pre + "tf"]), None if np.isnan(h0) else h0,
                    store_freq=int(g[pre + "store_freq"]))
    np.testing.assert_array_equal(np.array([r.h for r in sol.log]), g[pre + "trial_h"])
    np.testing.assert_array_equal(np.array([r.accepted for r in sol.log]), g[pre + "trial_accepted"])
    np.testing.assert_array_equal(uf, g[pre + "u_final"])
    np.testing.assert_array_equal(np.array(sol.t), g[pre + "t"])
    assert len(sol.u) == int(g[pre + "n_snap"])


def test_cfg1_readme_quickstart(golden):
    """BASELINE cfg 1: 546 accepted + 21 rejected trials, 28 snapshots, last t = 49.6579."""
    g = golden("ks1024_if34_cfg1.npz")
    p = problems.ks(1024)
    sol = OracleSolver("IF34", p.lin_op, p.nl_func)
    uf = sol.evolve(p.u0, 0.0, 50.0, store_freq=20)
    acc = np.array([r.accepted for r in sol.log])
    assert acc.sum() == 546 and (~acc).sum() == 21 and len(sol.u) == 28
    assert sol.t[-1] == pytest.approx(49.65786485353072, rel=1e-15)
    np.testing.assert_array_equal(np.array([r.h for r in sol.log]), g["trial_h"])
    assert np.linalg.norm(uf) == pytest.approx(float(g["u_final_norm"]), rel=1e-12)


DIAG_CASES = [("IF34", 1e-6), ("ETD34", 1e-6), ("ETD35", 1e-6), ("ETD35", 1e-9)]


@pytest.mark.parametrize("method,eps", DIAG_CASES)
def test_diagonalized_runs_match_reference(golden, method, eps):
    """diagonalize=True (dense lin_op, etd35.py:348-495): the oracle's eigenbasis strategy reproduces the
    reference's trial sequence and states bit for bit (same LAPACK eig / inv on the host)."""
    from oracle.rk_oracle import OracleDiagonalized
    g = golden("diagonalized_runs.npz")
    p = problems.dense_advection_diffusion()
    np.testing.assert_array_equal(p.lin_op, g["lin_op"])
    np.testing.assert_array_equal(p.u0, g["u0"])
    pre = f"{method}_{eps:g}_"
    sol = OracleDiagonalized(method, p.lin_op, p.nl_func, Config(epsilon=eps))
    uf = sol.evolve(p.u0.copy(), 0.0, float(g[pre + "tf"]), store_freq=3)
    np.testing.assert_array_equal(np.array([r.h for r in sol.log]), g[pre + "trial_h"])
    np.testing.assert_array_equal(np.array([r.accepted for r in sol.log]), g[pre + "trial_accepted"])
    np.testing.assert_array_equal(uf, g[pre + "u_final"])
    np.testing.assert_array_equal(np.array(sol.t), g[pre + "t"])
    assert len(sol.u) == int(g[pre + "n_snap"])

