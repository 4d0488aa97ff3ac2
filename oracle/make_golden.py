"""Generate tests/golden/*.npz by running the UNMODIFIED reference (dev container only).

TEST INFRASTRUCTURE ONLY.  Usage:  python -m oracle.make_golden
The reference package is imported from /root/reference through oracle/ref_loader.py; its
solver classes are driven through their public API (evolve/step) and instrumented from the
outside (bound-method wrappers record every trial's h and s) -- no reference code is copied.
Each case stores its inputs as well, so the fixtures are self-contained on the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import problems  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

ADAPTIVE = ("IF34", "ETD34", "ETD35", "IF45DP")
FIXED = ("IF4", "ETD4", "ETD5")


def make_solver(ref, method, lin_op, nl, epsilon=None):
    mod = getattr(ref, method.lower())
    cls = getattr(mod, method)
    if method in ADAPTIVE:
        cfg = ref.solveras.SolverConfig() if epsilon is None else ref.solveras.SolverConfig(epsilon=epsilon)
        if method in ("ETD34", "ETD35"):
            return cls(lin_op, nl, config=cfg, etd_config=ref.etd.ETDConfig())
        return cls(lin_op, nl, config=cfg)
    if method in ("ETD4", "ETD5"):
        return cls(lin_op, nl, etd_config=ref.etd.ETDConfig())
    return cls(lin_op, nl)


def instrument(solver):
    """Record (h, s) of every trial of an adaptive reference solver from the outside."""
    log = {"h": [], "s": []}
    upd, cs = solver._update_stages, solver._compute_s

    def upd_wrapped(u, h):
        log["h"].append(float(h))
        return upd(u, h)

    def cs_wrapped(u, err):
        s = cs(u, err)
        log["s"].append(float(s))
        return s

    solver._update_stages = upd_wrapped
    solver._compute_s = cs_wrapped
    return log


def run_adaptive(ref, method, prob, tf, epsilon, store_freq=1, h_init=None):
    sol = make_solver(ref, method, prob.lin_op, prob.nl_func, epsilon)
    log = instrument(sol)
    with np.errstate(all="ignore"):
        uf = sol.evolve(prob.u0, 0.0, tf, h_init=h_init, store_data=True, store_freq=store_freq)
    h = np.array(log["h"])
    s = np.array(log["s"])
    acc = ~(np.isinf(s) | np.isnan(s) | (s < 1.0))
    return dict(lin_op=prob.lin_op, u0=prob.u0, kx=prob.kx, tf=tf, epsilon=epsilon, store_freq=store_freq,
                h_init=np.nan if h_init is None else h_init,
                u_final=uf, trial_h=h, trial_s=s, trial_accepted=acc,
                t=np.array(sol.t), u_snap_last=np.asarray(sol.u[-1]), n_snap=len(sol.u))


def run_fixed(ref, method, prob, h, steps):
    sol = make_solver(ref, method, prob.lin_op, prob.nl_func)
    u1 = sol.step(prob.u0, h)                       # one step from a fresh solver
    sol.reset()
    tf = h * steps
    uf = sol.evolve(prob.u0, 0.0, tf, h, store_data=True, store_freq=max(1, steps // 4))
    return dict(lin_op=prob.lin_op, u0=prob.u0, kx=prob.kx, h=h, steps=steps, tf=tf,
                u_step1=u1, u_final=uf, t=np.array(sol.t), n_snap=len(sol.u),
                u_snap_1=np.asarray(sol.u[1]))


def coefficient_arrays(ref, method, lin_op, h):
    """Pull the coefficient arrays the reference builds for step size h."""
    nl = lambda v: np.zeros_like(v)  # noqa: E731
    sol = make_solver(ref, method, lin_op, nl)
    out = {}
    if method == "IF45DP":
        sol._update_coeffs(h)
        src = sol
    else:
        sol._method.update_coeffs(h)
        src = sol._method
    for k, v in vars(src).items():
        if not k.startswith("_"):
            continue
        name = k.lstrip("_")
        if name.startswith(("EL", "a", "b", "r")) and name not in ("accept",) and isinstance(v, (np.ndarray, float)):
            if name == "a64" and method in ("ETD5", "ETD35"):
                continue                                 # allocated, never written (etd35.py:143)
            out[name] = np.asarray(v)
    return out


def main():
    ref = load_reference()
    if ref is None:
        raise SystemExit("reference not found; goldens can only be generated in the dev container")
    os.makedirs(GOLD, exist_ok=True)

    # ---- 1. SURVEY 8c fingerprints: one trial on KS cfg-1 inputs -----------------------
    p = problems.ks(1024)
    fp = dict(u0_norm=np.linalg.norm(p.u0))
    for m in ADAPTIVE:
        sol = make_solver(ref, m, p.lin_op, p.nl_func)
        k, err = sol._update_stages(p.u0, 0.5)
        fp[f"{m}_k"] = k
        fp[f"{m}_err"] = err
        fp[f"{m}_s"] = sol._compute_s(k, err)
    for m in FIXED:
        sol = make_solver(ref, m, p.lin_op, p.nl_func)
        fp[f"{m}_k"] = sol.step(p.u0, 0.05)
    np.savez_compressed(os.path.join(GOLD, "ks1024_one_trial.npz"), **fp)

    # ---- 2. coefficient arrays (real L: KS; complex L: NLS, KdV) -------------------------
    co = {}
    for tag, prob, h in (("ks", problems.ks(64), 0.05), ("nls", problems.nls(64, half_width=20.0), 0.013),
                         ("kdv", problems.kdv(64), 0.025)):
        co[f"{tag}_lin_op"] = prob.lin_op
        co[f"{tag}_h"] = h
        for m in ADAPTIVE + FIXED:
            for name, arr in coefficient_arrays(ref, m, prob.lin_op, h).items():
                co[f"{tag}_{m}_{name}"] = arr
    np.savez_compressed(os.path.join(GOLD, "coefficients.npz"), **co)

    # ---- 3. fixed-step runs -------------------------------------------------------------
    fx = {}
    cases = (("kdv", problems.kdv(256), 0.025, 200), ("ks", problems.ks(256), 0.05, 60),
             ("burgers", problems.burgers(256, mu=0.01), 0.005, 60),
             ("nls", problems.nls(256, half_width=20.0), 0.002, 100),
             ("ksb", problems.ks(128, batch=3, seed=0), 0.05, 40))
    for tag, prob, h, steps in cases:
        for m in FIXED:
            for k, v in run_fixed(ref, m, prob, h, steps).items():
                fx[f"{tag}_{m}_{k}"] = v
    np.savez_compressed(os.path.join(GOLD, "fixed_runs.npz"), **fx)

    # ---- 4. adaptive runs ---------------------------------------------------------------
    ad = {}
    cases = (("kdv", problems.kdv(256), 5.0, 1e-4, 1, 0.025),
             ("ks", problems.ks(256), 8.0, 1e-4, 5, None),
             ("burgers", problems.burgers(256, mu=0.01), 0.4, 1e-4, 3, None),
             ("nls", problems.nls(256, half_width=20.0), 0.5, 1e-6, 2, None),
             ("nlsb", problems.nls(128, batch=3, seed=2, half_width=20.0), 0.25, 1e-5, 1, None),
             ("ksb", problems.ks(128, batch=4, seed=0), 4.0, 1e-4, 2, None))
    for tag, prob, tf, eps, sf, h0 in cases:
        for m in ADAPTIVE:
            tfm = tf
            if m == "IF45DP":                      # r4 quirk => ~40x more trials; shorten
                tfm = tf / 8
            for k, v in run_adaptive(ref, m, prob, tfm, eps, sf, h0).items():
                ad[f"{tag}_{m}_{k}"] = v
    np.savez_compressed(os.path.join(GOLD, "adaptive_runs.npz"), **ad)

    # ---- 4b. diagonalize=True on a dense operator --------------------------
    dg = {}
    prob = problems.dense_advection_diffusion()
    for m, eps, tf in (("IF34", 1e-6, 1.0), ("ETD34", 1e-6, 1.0), ("ETD35", 1e-6, 1.0), ("ETD35", 1e-9, 0.5)):
        cls = getattr(getattr(ref, m.lower()), m)
        sol = cls(prob.lin_op, prob.nl_func, config=ref.solveras.SolverConfig(epsilon=eps), diagonalize=True)
        log = instrument(sol)
        uf = sol.evolve(prob.u0.copy(), 0.0, tf, store_data=True, store_freq=3)
        s = np.array(log["s"])
        pre = f"{m}_{eps:g}_"
        dg.update({pre + "u_final": uf, pre + "trial_h": np.array(log["h"]), pre + "trial_s": s,
                   pre + "trial_accepted": ~(np.isinf(s) | np.isnan(s) | (s < 1.0)), pre + "t": np.array(sol.t),
                   pre + "u_snap_last": np.asarray(sol.u[-1]), pre + "n_snap": len(sol.u), pre + "tf": tf})
    dg["lin_op"], dg["u0"] = prob.lin_op, prob.u0
    np.savez_compressed(os.path.join(GOLD, "diagonalized_runs.npz"), **dg)

    # ---- 5. README quickstart (cfg 1) summary -------------------------------------------
    p = problems.ks(1024)
    r = run_adaptive(ref, "IF34", p, 50.0, None if False else 1e-4, 20, None)
    np.savez_compressed(os.path.join(GOLD, "ks1024_if34_cfg1.npz"),
                        trial_h=r["trial_h"], trial_accepted=r["trial_accepted"], t=r["t"],
                        n_snap=r["n_snap"], u_final_norm=np.linalg.norm(r["u_final"]))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
