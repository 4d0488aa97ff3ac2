"""CPU checks of the device math: the __host__ __device__ headers of rkstiff_b200/csrc are
compiled with g++ (tests/host_check) and compared with the oracle.  No GPU, no CUDA runtime.
These tests pin the algorithms (FFT pass structure, psi/tableau formulas, stage formulas,
controller state machine); the `-m gpu` tests then check the kernels that wrap them.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import problems
from oracle.rk_oracle import (ADAPTIVE, FIXED, METHODS, Config, MaxLoopsExceeded, MinimumStepReached,
                              OracleSolver, coefficients)

HERE = os.path.dirname(os.path.abspath(__file__))
MID = {"IF4": 0, "ETD4": 1, "ETD5": 2, "IF34": 3, "ETD34": 4, "ETD35": 5, "IF45DP": 6}
SLOTS = {
    "kro": ["E", "E2", "a21", "a31", "a32", "a41", "a43", "a51", "a52", "a54"],
    "e5": ["E14", "E12", "E34", "E", "a21", "a31", "a32", "a41", "a43", "a51", "a52", "a54", "a61", "a62", "a63",
           "a65", "a71", "a73", "a74", "a75", "a76"],
    "if": ["E", "E2"],
    "dp": ["E15", "E310", "E45", "E89", "E", "a21", "a31", "a32", "a41", "a42", "a43", "a51", "a52", "a53", "a54",
           "a61", "a62", "a63", "a64", "a65", "a71", "a73", "a74", "a75", "r1", "r3", "r4", "r5"],
}
FAMILY = {"IF4": "if", "IF34": "if", "ETD4": "kro", "ETD34": "kro", "ETD5": "e5", "ETD35": "e5", "IF45DP": "dp"}
STAGES = {"IF4": 4, "IF34": 4, "ETD4": 4, "ETD34": 4, "ETD5": 6, "ETD35": 6, "IF45DP": 6}


@pytest.fixture(scope="module")
def hc():
    src = os.path.join(HERE, "host_check", "host_check.cpp")
    lib = os.path.join(HERE, "host_check", "libhostcheck.so")
    deps = [src] + [os.path.join(HERE, "..", "rkstiff_b200", "csrc", f)
                    for f in ("common.cuh", "coeffs.cuh", "stages.cuh", "errctl.cuh", "fft.cuh", "fft_fast.cuh", "fft_axis.cuh")]
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", lib, src])
    return ctypes.CDLL(lib)


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("n", [16, 32, 64, 128, 512, 1024, 2048, 8192])
def test_fft_dif_dit_roundtrip(hc, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y = np.empty_like(x)
    hc.hc_fft_roundtrip(n, ptr(x), ptr(y))
    assert rel(y / n, x) < 5e-16 * np.log2(n)


@pytest.mark.parametrize("n", [16, 64, 256, 1024, 8192])
@pytest.mark.parametrize("nthreads", [1, 7, 64])
def test_nl_uux_matches_numpy(hc, n, nthreads):
    p = problems.ks(n) if n >= 64 else problems.kdv(n)
    rng = np.random.default_rng(n)
    uf = p.u0 + 1e-3 * (rng.standard_normal(p.u0.shape) + 1j * rng.standard_normal(p.u0.shape))
    uf[-1] += 0.3j          # Nyquist with an imaginary part: irfft must ignore it
    for c in (1.0, 6.0):
        ref = -c * np.fft.rfft(np.fft.irfft(uf) * np.fft.irfft(1j * p.kx * uf))
        out = np.empty_like(uf)
        hc.hc_nl(1, n, ptr(uf), ptr(p.kx), ctypes.c_double(c), ptr(out), nthreads)
        assert rel(out, ref) < 1e-14 * np.log2(n)


@pytest.mark.parametrize("n", [16, 128, 2048, 8192])
def test_nl_nls_matches_numpy(hc, n):
    p = problems.nls(n, batch=2, half_width=20.0)
    for row in p.u0:
        out = np.empty_like(row)
        hc.hc_nl(2, n, ptr(np.ascontiguousarray(row)), None, ctypes.c_double(2.0), ptr(out), 32)
        assert rel(out, p.nl_func(row)) < 1e-14 * np.log2(n)


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192])
def test_pretransformed_route_is_bit_identical(hc, n):
    """fft_fast.cuh pre_butterfly/phase_pre: K1 applying the first inverse pass and K4 starting one pass later
    perform the same operations in the same order as the plain K4, so the two routes agree to the last bit."""
    p = problems.nls(n, batch=3, half_width=20.0, seed=n)
    rng = np.random.default_rng(n)
    for row in p.u0:
        row = np.ascontiguousarray(row + 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n)))
        plain, pre = np.empty_like(row), np.empty_like(row)
        assert hc.hc_nl_fast(2, n, ptr(row), None, ctypes.c_double(2.0), ptr(plain)) == 0
        assert hc.hc_nl_pre(n, ptr(row), ctypes.c_double(2.0), ptr(pre)) == 0
        np.testing.assert_array_equal(pre, plain)
        assert rel(pre, p.nl_func(row)) < 2e-15 * np.log2(n)


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096, 8192])
def test_fast_register_fft_nl_matches_numpy(hc, n):
    """fft_fast.cuh: the register-resident W x 8 x 8 x 8 pipeline, emulated thread by thread."""
    p = problems.nls(n, batch=2, half_width=20.0)
    row = np.ascontiguousarray(p.u0[1])
    out = np.empty_like(row)
    assert hc.hc_nl_fast(2, n, ptr(row), None, ctypes.c_double(2.0), ptr(out)) == 0
    assert rel(out, p.nl_func(row)) < 2e-15 * np.log2(n)
    p = problems.ks(n)
    rng = np.random.default_rng(n)
    uf = p.u0 + 1e-3 * (rng.standard_normal(p.u0.shape) + 1j * rng.standard_normal(p.u0.shape))
    uf[-1] += 0.3j
    out = np.empty_like(uf)
    assert hc.hc_nl_fast(1, n, ptr(uf), ptr(p.kx), ctypes.c_double(6.0), ptr(out)) == 0
    ref = -6 * np.fft.rfft(np.fft.irfft(uf) * np.fft.irfft(1j * p.kx * uf))
    assert rel(out, ref) < 1e-14 * np.log2(n)


@pytest.mark.parametrize("n", [64, 256, 1024, 4096])
def test_nl_cubic_and_sine_gordon_match_numpy(hc, n):
    """models 3 (Allen-Cahn cubic, rfft) and 4 (sine-Gordon, first-order complex form): generic and fast paths."""
    p = problems.allen_cahn_1d(n)
    uf = p.u0.copy()
    uf[-1] += 0.2j
    ref = p.nl_func(uf)
    out = np.empty_like(uf)
    hc.hc_nl(3, n, ptr(uf), None, ctypes.c_double(-1.0), ptr(out), 32)
    assert rel(out, ref) < 2e-14 * np.log2(n)
    if n >= 512:
        assert hc.hc_nl_fast(3, n, ptr(uf), None, ctypes.c_double(-1.0), ptr(out)) == 0
        assert rel(out, ref) < 2e-14 * np.log2(n)
    p = problems.sine_gordon(n)
    rng = np.random.default_rng(n)
    pf = p.u0 + 0.01 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    ref = p.nl_func(pf)
    out = np.empty_like(pf)
    hc.hc_nl(4, n, ptr(pf), ptr(p.kx), ctypes.c_double(0.0), ptr(out), 32)
    assert rel(out, ref) < 2e-14 * np.log2(n)
    if n >= 512:
        assert hc.hc_nl_fast(4, n, ptr(pf), ptr(p.kx), ctypes.c_double(0.0), ptr(out)) == 0
        assert rel(out, ref) < 2e-14 * np.log2(n)


def device_coeffs(hc, method, lin, h, cfg=Config()):
    lin = np.ascontiguousarray(lin)
    n = lin.shape[0]
    names = SLOTS[FAMILY[method]]
    out = np.zeros((len(names), n), dtype=np.complex128)
    nc = hc.hc_coeffs(MID[method], n, ptr(lin), int(np.iscomplexobj(lin)), ctypes.c_double(h),
                      ctypes.c_double(cfg.modecutoff), cfg.contour_points, ctypes.c_double(cfg.contour_radius),
                      int(cfg.if45dp_r4_fix), ptr(out))
    assert nc == len(names)
    return dict(zip(names, out)), out


@pytest.mark.parametrize("prob,h", [(problems.ks(256), 0.05), (problems.nls(256, half_width=20.0), 0.013),
                                    (problems.kdv(256), 0.025), (problems.burgers(256, mu=0.01), 0.005)])
@pytest.mark.parametrize("method", METHODS)
def test_coefficients_match_oracle(hc, prob, h, method):
    mine, _ = device_coeffs(hc, method, prob.lin_op, h)
    ref = coefficients(method, prob.lin_op, h, Config())
    z = np.abs(h * prob.lin_op)
    for name, arr in mine.items():
        r = np.asarray(ref[name], dtype=np.complex128)
        # cancellation band just above modecutoff: the reference itself carries ~1e-9 relative
        # rounding noise there; elsewhere the two agree to rounding
        band = (z >= 0.01) & (z < 0.5)
        np.testing.assert_allclose(arr[~band], r[~band], rtol=2e-13, atol=1e-15 * h, err_msg=f"{method}.{name}")
        np.testing.assert_allclose(arr[band], r[band], rtol=5e-8, atol=1e-15 * h, err_msg=f"{method}.{name} band")


def test_if45dp_r4_quirk_and_fix(hc):
    p = problems.ks(64)
    quirk, _ = device_coeffs(hc, "IF45DP", p.lin_op, 0.1)
    fixed, _ = device_coeffs(hc, "IF45DP", p.lin_op, 0.1, Config(if45dp_r4_fix=True))
    np.testing.assert_allclose(fixed["r4"] / quirk["r4"], 71.0 / 17.0, rtol=1e-15)


def device_trial(hc, method, prob, u, h, N1=None):
    """One pass over the stages with device formulas + NumPy nl_func; returns (u_new, err, N dict)."""
    lin = prob.lin_op
    mine, coef = device_coeffs(hc, method, lin, h)
    n = u.shape[0]
    N = {1: prob.nl_func(u) if N1 is None else N1}
    S = STAGES[method]
    k = None
    err = np.zeros(n, dtype=np.complex128)
    for s in range(1, S + 1):
        arr = (ctypes.c_void_p * 8)()
        for j in range(1, 8):
            arr[j] = ptr(N[j]) if j in N else None
        out = np.empty(n, dtype=np.complex128)
        rc = hc.hc_stage(MID[method], s, n, ptr(u), arr, ptr(coef), ctypes.c_double(h), ptr(out), ptr(err))
        assert rc == 0
        k = out
        if s < S:
            N[s + 1] = prob.nl_func(k)
    if method in ("IF34", "ETD34", "IF45DP"):
        N[S + 1] = prob.nl_func(k)
        arr = (ctypes.c_void_p * 8)()
        for j in range(1, 8):
            arr[j] = ptr(N[j]) if j in N else None
        assert hc.hc_embedded_err(MID[method], n, arr, ptr(coef), ctypes.c_double(h), ptr(err)) == 0
    return k, err, N


@pytest.mark.parametrize("prob,h", [(problems.ks(256), 0.05), (problems.nls(256, half_width=20.0), 0.004),
                                    (problems.kdv(256), 0.025)])
@pytest.mark.parametrize("method", METHODS)
def test_one_trial_state_matches_oracle(hc, prob, h, method):
    """Fixed-step parity bar: <= 1e-12 relative per step (BASELINE north_star)."""
    sol = OracleSolver(method, prob.lin_op, prob.nl_func)
    ref = sol.trial(prob.u0, h)
    k, err, _ = device_trial(hc, method, prob, prob.u0, h)
    if method in ADAPTIVE:
        assert rel(k, ref[0]) < 1e-13
        # err is a difference of O(|N|) terms, so the two agree absolutely (rounding of N), not relatively
        assert np.linalg.norm(err - ref[1]) < 1e-14 * np.linalg.norm(k)
    else:
        assert rel(k, ref) < 1e-13


class _Canned(OracleSolver):
    """Oracle controller fed with canned (u, err) pairs: checks the state machine alone."""

    def __init__(self, method, pairs, cfg):
        super().__init__(method, np.zeros(1), lambda v: v, cfg)
        self.pairs = list(pairs)
        self.i = 0

    def trial(self, u, h):
        x, y = self.pairs[self.i % len(self.pairs)]
        self.i += 1
        return np.array([x + 0j]), np.array([y + 0j])


@pytest.mark.parametrize("method", ADAPTIVE)
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_controller_matches_oracle(hc, method, seed):
    rng = np.random.default_rng(seed)
    cfg = Config(epsilon=1e-4)
    q = 5 if method == "IF45DP" else 4
    # error/tolerance ratios spread around the accept threshold, with occasional NaN/inf/zero
    pairs = []
    for _ in range(400):
        x = rng.uniform(0.5, 2.0)
        ratio = 10 ** rng.uniform(-1.2, 1.2)
        y = cfg.epsilon * x / ratio
        r = rng.uniform()
        if r < 0.02:
            y = 0.0                 # ||err|| = 0 -> s = inf -> rejected
        elif r < 0.04:
            x = float("nan")
        pairs.append((x, y))
    sol = _Canned(method, pairs, cfg)
    t0, tf, store_freq = 0.0, 3.0, 3
    expect_status = 1
    with np.errstate(all="ignore"):
        try:
            sol.evolve(np.zeros(1, dtype=complex), t0, tf, None, store_freq=store_freq)
        except MinimumStepReached:
            expect_status = 3
        except MaxLoopsExceeded:
            expect_status = 2
    size = hc.hc_ctrl_size()
    blob = ctypes.create_string_buffer(size)
    h0 = (tf - t0) / 100.0
    hc.hc_ctrl_init(blob, ctypes.c_double(t0), ctypes.c_double(tf), ctypes.c_double(h0), ctypes.c_longlong(store_freq),
                    0, int(method == "ETD35"), ctypes.c_double(cfg.epsilon), ctypes.c_double(cfg.incr_f),
                    ctypes.c_double(cfg.decr_f), ctypes.c_double(cfg.safety_f), ctypes.c_double(cfg.minh), q)
    out = np.zeros(13)
    hs, accs, ts, snaps = [], [], [], [t0]
    h = h0
    i = 0
    status = 0
    while status == 0:
        x, y = pairs[i % len(pairs)]
        i += 1
        hs.append(h)
        hc.hc_ctrl_advance(blob, ctypes.c_double(x * x), ctypes.c_double(y * y), ptr(out))
        h, status = out[0], int(out[4])
        accs.append(bool(out[5]))
        if out[5]:
            ts.append(out[2])
        if out[11]:
            snaps.append(out[2])
    assert status == expect_status
    # numpy's pow is its own SIMD kernel (not libm's), so h agrees to rounding, not bitwise
    np.testing.assert_allclose(np.array(hs), np.array([r.h for r in sol.log]), rtol=1e-13, atol=0)
    np.testing.assert_array_equal(np.array(accs), np.array([r.accepted for r in sol.log]))
    ref_ts = np.array([r.t_after for r in sol.log if r.accepted])
    np.testing.assert_allclose(np.array(ts)[:len(ref_ts)], ref_ts, rtol=1e-13, atol=0)
    np.testing.assert_allclose(np.array(snaps), np.array(sol.t), rtol=1e-13, atol=0)


def test_controller_failure_paths(hc):
    """||err|| never small enough: MinimumStepReached / MaxLoopsExceeded (solveras.py:399-410)."""
    size = hc.hc_ctrl_size()
    out = np.zeros(13)
    for minh, expect in ((1e-3, 3), (1e-300, 2)):
        blob = ctypes.create_string_buffer(size)
        hc.hc_ctrl_init(blob, ctypes.c_double(0.0), ctypes.c_double(1.0), ctypes.c_double(0.01), ctypes.c_longlong(1),
                        0, 0, ctypes.c_double(1e-4), ctypes.c_double(1.25), ctypes.c_double(0.85),
                        ctypes.c_double(0.8), ctypes.c_double(minh), 4)
        n = 0
        while True:
            hc.hc_ctrl_advance(blob, ctypes.c_double(1.0), ctypes.c_double(1.0), ptr(out))
            n += 1
            if out[4] != 0:
                break
        assert int(out[4]) == expect
        if expect == 2:
            assert n == 51      # numloops > MAX_LOOPS on the 51st rejection


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096])
def test_axis_fft_levels_match_numpy(hc, n):
    """fft_axis.cuh on the host: the inverse (DIF) transform gives ifft in digit-reversed row order, the
    forward (DIT) transform takes that order back to fft's natural order; inner need not be a tile multiple."""
    rng = np.random.default_rng(n)
    inner = 5 if n > 512 else 11
    x = rng.standard_normal((n, inner)) + 1j * rng.standard_normal((n, inner))
    # the row permutation, read off a ramp: physical value m sits in the row whose output equals m
    ramp = np.fft.fft(np.arange(n, dtype=float))[:, None] * np.ones((1, inner))
    ramp = np.ascontiguousarray(ramp.astype(np.complex128))
    out = np.empty_like(ramp)
    assert hc.hc_axis_fft(n, 1, ptr(ramp), ptr(out), ctypes.c_longlong(inner)) == 0
    perm = np.rint(out[:, 0].real).astype(int)
    assert sorted(perm) == list(range(n))
    np.testing.assert_allclose(out, perm[:, None] * np.ones((1, inner)), atol=1e-9 * n)
    y = np.empty_like(x)
    assert hc.hc_axis_fft(n, 1, ptr(x), ptr(y), ctypes.c_longlong(inner)) == 0
    ref = np.fft.ifft(x, axis=0)
    np.testing.assert_allclose(y, ref[perm], rtol=0, atol=1e-13 * np.abs(ref).max() * np.log2(n))
    z = np.empty_like(x)
    assert hc.hc_axis_fft(n, 0, ptr(y), ptr(z), ctypes.c_longlong(inner)) == 0
    np.testing.assert_allclose(z, x, rtol=0, atol=1e-13 * np.abs(x).max() * np.log2(n))
    # forward of digit-reversed physical data == numpy fft of the natural-order data
    phys = rng.standard_normal((n, inner)) + 1j * rng.standard_normal((n, inner))
    w = np.empty_like(phys)
    assert hc.hc_axis_fft(n, 0, ptr(np.ascontiguousarray(phys[perm])), ptr(w), ctypes.c_longlong(inner)) == 0
    reff = np.fft.fft(phys, axis=0)
    np.testing.assert_allclose(w, reff, rtol=0, atol=1e-13 * np.abs(reff).max() * np.log2(n))


@pytest.mark.parametrize("n", [64, 128, 256])
@pytest.mark.parametrize("model", [1, 2, 3, 4])
def test_packed_short_row_pipeline_matches_generic_rows(hc, n, model):
    """nl_small_kernel on the host: 512/n rows packed per slab, ragged last slab, all four models == the
    generic per-row pipeline (which the other tests pin to NumPy)."""
    rng = np.random.default_rng(100 * n + model)
    rows = 2 * (512 // n) + 1                       # two full slabs and a ragged one
    n_c = n // 2 + 1 if model in (1, 3) else n
    x = rng.standard_normal((rows, n_c)) + 1j * rng.standard_normal((rows, n_c))
    if model in (1, 3):
        x[:, 0] = x[:, 0].real
        x[:, -1] = x[:, -1].real
    x = np.ascontiguousarray(x)
    kx = np.ascontiguousarray(np.sqrt(1.0 + np.arange(n, dtype=float) ** 2) if model == 4 else np.arange(n_c, dtype=float) * 0.37)
    p0 = {1: 6.0, 2: 2.0, 3: -1.0, 4: 0.0}[model]
    got = np.full_like(x, np.nan)
    assert hc.hc_nl_packed(model, n, rows, ptr(x), ptr(kx), ctypes.c_double(p0), ptr(got)) == 0
    want = np.empty_like(x)
    for r in range(rows):
        row_out = np.empty(n_c, dtype=np.complex128)
        hc.hc_nl(model, n, ptr(np.ascontiguousarray(x[r])), ptr(kx), ctypes.c_double(p0), ptr(row_out), 32)
        want[r] = row_out
    assert np.isfinite(got).all()
    scale = np.abs(want).max()
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-13 * scale)


@pytest.mark.parametrize("d", [3, 6, 40, 84, 525])
def test_div_const_is_the_correctly_rounded_quotient(hc, d):
    """common.cuh div_const<D>: q = x RN(1/D), r = fma(-D, q, x), q + r RN(1/D) must equal x / D bit for bit
    (NumPy's division, which the reference's `/ 6`, `/ 3`, `/ 84` ... are) over the normal range; inf and NaN pass through."""
    rng = np.random.default_rng(d)
    bits = rng.integers(0, 1 << 52, size=2_000_000, dtype=np.uint64)
    expo = rng.integers(1023 - 900, 1023 + 900, size=bits.size, dtype=np.uint64)
    sign = rng.integers(0, 2, size=bits.size, dtype=np.uint64) << np.uint64(63)
    x = (bits | (expo << np.uint64(52)) | sign).view(np.float64)
    # adversarial mantissas: small odd integers, all-ones tails, exact multiples of D
    extra = np.concatenate([np.arange(1, 200001, 2, dtype=np.float64), np.arange(0, 100000, dtype=np.float64) * d,
                            (np.uint64(0x3ff0000000000000) | (np.uint64(0xfffffffffffff) - np.arange(100000, dtype=np.uint64))).view(np.float64),
                            np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-300, -1e300, 5e-324])])
    x = np.ascontiguousarray(np.concatenate([x, extra]))
    out = np.empty_like(x)
    assert hc.hc_div_const(d, x.size, ptr(x), ptr(out)) == 0
    with np.errstate(invalid="ignore"):
        ref = x / float(d)
    normal = ~(np.abs(ref) < 2.3e-308) | (ref == 0)          # subnormal quotients may differ by one subnormal ulp
    np.testing.assert_array_equal(out[normal], ref[normal])
    assert np.all(np.abs(out[~normal] - ref[~normal]) <= 5e-324)


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
def test_real_field_half_length_forward_matches_numpy(hc, n):
    """fft_real.cuh (used by the opt-in nl_fast_real_kernel only): the forward transform of the real pointwise product as a half-length
    complex transform on the digit-reversed in-place layout -- paired core pass, even-block middle pass, radix-R1/2 last
    pass with the C[k] / conj C[n/2-k] exchange -- for the u u_x model (incl. a Nyquist mode with an imaginary part)
    and the cubic model, against NumPy and against the full-length route."""
    p = problems.ks(n)
    rng = np.random.default_rng(n)
    uf = p.u0 + 1e-3 * (rng.standard_normal(p.u0.shape) + 1j * rng.standard_normal(p.u0.shape))
    uf[-1] += 0.3j
    out, full = np.empty_like(uf), np.empty_like(uf)
    assert hc.hc_nl_fast_real(1, n, ptr(uf), ptr(p.kx), ctypes.c_double(6.0), ptr(out)) == 0
    assert hc.hc_nl_fast(1, n, ptr(uf), ptr(p.kx), ctypes.c_double(6.0), ptr(full)) == 0
    ref = -6 * np.fft.rfft(np.fft.irfft(uf) * np.fft.irfft(1j * p.kx * uf))
    assert rel(out, ref) < 1e-14 * np.log2(n)
    assert rel(out, full) < 1e-14 * np.log2(n)
    # cubic model: N = -rfft(irfft(u)^3)
    uf = uf / np.abs(uf).max()
    out = np.empty_like(uf)
    assert hc.hc_nl_fast_real(3, n, ptr(uf), None, ctypes.c_double(-1.0), ptr(out)) == 0
    ref = -np.fft.rfft(np.fft.irfft(uf) ** 3)
    assert rel(out, ref) < 1e-14 * np.log2(n)
    assert hc.hc_nl_fast_real(2, n, ptr(uf), None, ctypes.c_double(1.0), ptr(out)) != 0     # complex-field model: not applicable


@pytest.mark.parametrize("n", [512, 1024, 2048, 4096])
def test_paired_rows_cubic_model_matches_numpy(hc, n):
    """fft_pair.cuh: two real rows through ONE complex transform pair == -rfft(irfft(u)^3) of each row; rows of very
    different size (the rounding cross-talk is relative to the larger row) and the odd tail (second row absent)."""
    rng = np.random.default_rng(n + 1)
    nc = n // 2 + 1
    decay = np.exp(-0.02 * np.arange(nc))

    def row(scale):
        v = scale * decay * (rng.standard_normal(nc) + 1j * rng.standard_normal(nc))
        v[0] += 0.7j * scale                     # c2r drops the imaginary parts of DC and Nyquist
        v[-1] -= 0.4j * scale
        return np.ascontiguousarray(v)

    a, b = row(1.0), row(3.0)
    oa, ob = np.empty_like(a), np.empty_like(b)
    assert hc.hc_nl_fast_pair(n, ptr(a), ptr(b), ctypes.c_double(-1.0), ptr(oa), ptr(ob)) == 0
    ra = -np.fft.rfft(np.fft.irfft(a, n) ** 3)
    rb = -np.fft.rfft(np.fft.irfft(b, n) ** 3)
    tol = 1e-14 * np.log2(n)
    assert rel(oa, ra) < tol * (np.linalg.norm(rb) / np.linalg.norm(ra)) and rel(ob, rb) < tol
    assert oa[0].imag == 0.0 and oa[-1].imag == 0.0
    # odd tail: one row alone, nothing written for the absent partner
    oa2, sentinel = np.empty_like(a), np.full_like(b, 123.0)
    assert hc.hc_nl_fast_pair(n, ptr(a), None, ctypes.c_double(-1.0), ptr(oa2), ptr(sentinel)) == 0
    assert rel(oa2, ra) < tol and np.all(sentinel == 123.0)
    # the full-length single-row pipeline computes the same thing
    full = np.empty_like(a)
    assert hc.hc_nl_fast(3, n, ptr(a), None, ctypes.c_double(-1.0), ptr(full)) == 0
    assert rel(oa2, full) < tol


# ------------------------------------------------------------------------------------------------
# coefficient storage of large grids: per-axis exponential tables, grouped records
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("method", ["IF4", "IF34", "IF45DP"])
@pytest.mark.parametrize("cx", [False, True])
def test_separable_if_coefficients_match_the_direct_ones(hc, method, cx):
    """exp(q h (a_i + b_j)) = exp(q h a_i) exp(q h b_j): the per-axis tables reproduce every coefficient slot of the
    reference formulas (if4.py:72-83, if45dp.py:204-237) on the flattened grid to rounding."""
    rng = np.random.default_rng(5)
    n0, n1, h = 12, 9, 0.037
    ky, kx = np.fft.fftfreq(n0, 1.0 / n0), np.fft.rfftfreq(2 * (n1 - 1), 1.0 / (2 * (n1 - 1)))
    if cx:
        a0, a1 = (0.3 - 1j * ky ** 2).astype(np.complex128), (-1j * kx ** 2 + 0.01 * rng.standard_normal(n1)).astype(np.complex128)
    else:
        a0, a1 = 1.0 - 0.05 * ky ** 2, -0.05 * kx ** 2
    lin = np.ascontiguousarray((a0[:, None] + a1[None, :]).ravel())
    direct, _ = device_coeffs(hc, method, lin, h)
    names = SLOTS[FAMILY[method]]
    out = np.zeros((len(names), n0 * n1), dtype=np.complex128)
    nc = hc.hc_sep_coeffs(MID[method], n0, n1, ptr(np.ascontiguousarray(a0)), ptr(np.ascontiguousarray(a1)), int(cx),
                          ctypes.c_double(h), 0, ptr(out))
    assert nc == len(names)
    for name, arr in zip(names, out):
        np.testing.assert_allclose(arr, direct[name], rtol=2e-15 * (1 + np.abs(h * lin).max()), atol=0, err_msg=f"{method}.{name}")


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("real", [False, True])
def test_grouped_record_layout(hc, method, real):
    """every stage's slots are contiguous, start on a 32-byte sector and do not overlap another group"""
    if real and method not in ("IF4", "IF34", "IF45DP"):
        pytest.skip("real coefficient arrays exist for the IF methods only")
    out = (ctypes.c_int * 9)()
    masks = (ctypes.c_uint * 9)()
    ng = hc.hc_record_layout(MID[method], int(real), out, masks)
    assert ng == {"IF4": 5, "ETD4": 5, "IF34": 5, "ETD34": 5}.get(method, 7)
    elem = 8 if real else 16
    end = 0
    for g in range(1, ng + 1):
        assert out[g] * elem % 32 == 0 and out[g] >= end
        end = out[g] + bin(masks[g]).count("1")
    assert out[0] >= end and out[0] * elem % 32 == 0


# ---- spectral derivative rows (fft_fast.cuh DerivModel / DerivPairModel, rkstiff/derivatives.py:47-179) ----
def _deriv_table(kx_full, order, n):
    """what rkstiff_b200/derivatives.py hands to rks_rows_create: conj((i kx)^order) / n in FFT order"""
    return np.ascontiguousarray(np.conj((1j * kx_full) ** order) / n)


def _hc_deriv(hc, model, n, x, table):
    """one row (model 5: complex; model 6: a row PAIR of reals) through the kernel family the engine uses for n"""
    out = np.empty_like(x)
    if n >= 512:
        assert hc.hc_nl_fast(model, n, ptr(x), ptr(table), ctypes.c_double(0.0), ptr(out)) == 0
    elif n >= 64:
        assert hc.hc_nl_packed(model, n, 1, ptr(x), ptr(table), ctypes.c_double(0.0), ptr(out)) == 0
    else:
        hc.hc_nl(model, n, ptr(x), ptr(table), ctypes.c_double(0.0), ptr(out), 8)
    return out


@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_derivative_rows_complex_match_numpy(hc, n, order):
    rng = np.random.default_rng(100 * n + order)
    kx = 2 * np.pi * np.fft.fftfreq(n, d=0.37)
    z = np.ascontiguousarray(rng.standard_normal(n) + 1j * rng.standard_normal(n))
    got = _hc_deriv(hc, 5, n, z, _deriv_table(kx, order, n))
    ref = np.fft.ifft((1j * kx) ** order * np.fft.fft(z))
    assert rel(got, ref) < 4e-16 * np.log2(n) * 4


@pytest.mark.parametrize("n", [16, 64, 256, 512, 1024, 8192])
@pytest.mark.parametrize("order", [1, 2, 3, 4])
def test_derivative_rows_real_pairs_match_numpy(hc, n, order):
    """two real rows per complex transform; the Nyquist multiplier keeps its real part only (what irfft does)"""
    rng = np.random.default_rng(7 * n + order)
    kr = 2 * np.pi * np.fft.rfftfreq(n, d=0.11)
    half = (1j * kr) ** order
    half[-1] = half[-1].real
    full = np.concatenate([half, np.conj(half[1:-1][::-1])])
    table = np.ascontiguousarray(np.conj(full) / n)
    pair = np.ascontiguousarray(rng.standard_normal((2, n)))
    got = _hc_deriv(hc, 6, n, pair, table)
    ref = np.fft.irfft((1j * kr) ** order * np.fft.rfft(pair, axis=-1), n=n, axis=-1)
    assert rel(got, ref) < 4e-16 * np.log2(n) * 4


def test_packed_derivative_rows_keep_rows_apart(hc):
    """n = 64: eight rows share one 512-point slab; positions map to frequencies row by row (PackedModel::pointwise_at)"""
    n, rows, order = 64, 11, 1
    rng = np.random.default_rng(5)
    kx = 2 * np.pi * np.fft.fftfreq(n, d=0.5)
    z = np.ascontiguousarray(rng.standard_normal((rows, n)) + 1j * rng.standard_normal((rows, n)))
    out = np.empty_like(z)
    assert hc.hc_nl_packed(5, n, rows, ptr(z), ptr(_deriv_table(kx, order, n)), ctypes.c_double(0.0), ptr(out)) == 0
    ref = np.fft.ifft((1j * kx) ** order * np.fft.fft(z, axis=-1), axis=-1)
    assert rel(out, ref) < 2e-14
