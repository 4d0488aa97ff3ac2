"""Host-side API checks that mirror the reference's own unit tests (no GPU needed):
SolverConfig / ETDConfig validation (tests/test_solveras.py:43-85, tests/test_etd.py:68-104),
SolverType (tests/test_solver_type.py), logger naming (tests/test_loghelper.py), module surface."""
import logging

import pytest

import rkstiff_b200 as rk
from rkstiff_b200.etd import ETDConfig
from rkstiff_b200.solveras import BaseSolverAS, SolverConfig
from rkstiff_b200.util.loghelper import get_solver_logger, set_log_level
from rkstiff_b200.util.solver_type import SolverType


def test_solverconfig_valid_defaults():
    cfg = SolverConfig()
    assert (cfg.epsilon, cfg.incr_f, cfg.decr_f, cfg.safety_f, cfg.adapt_cutoff, cfg.minh) == \
        (1e-4, 1.25, 0.85, 0.8, 0.01, 1e-16)


@pytest.mark.parametrize("kw", [dict(epsilon=0.0), dict(epsilon=-1e-3), dict(incr_f=1.0), dict(incr_f=0.5),
                                dict(decr_f=1.0), dict(decr_f=2.0), dict(safety_f=1.01), dict(adapt_cutoff=1.0),
                                dict(minh=0), dict(minh=-1e-5)])
def test_solverconfig_rejects_out_of_range(kw):
    with pytest.raises(ValueError):
        SolverConfig(**kw)


def test_solverconfig_setters_validate_and_fresh_default_per_solver():
    cfg = SolverConfig()
    cfg.epsilon = 1e-6
    assert cfg.epsilon == 1e-6
    with pytest.raises(ValueError):
        cfg.decr_f = 1.5
    assert SolverConfig().epsilon == 1e-4          # no shared mutable default (SURVEY.md 4 hazard)


def test_etdconfig_validation():
    assert ETDConfig(modecutoff=0.5).modecutoff == 0.5
    assert ETDConfig(contour_points=8).contour_points == 8
    assert ETDConfig(contour_radius=2.5).contour_radius == 2.5
    for bad in (dict(modecutoff=1.5), dict(modecutoff=0.0), dict(contour_points=1), dict(contour_radius=0.0),
                dict(contour_radius=-1.0)):
        with pytest.raises(ValueError):
            ETDConfig(**bad)
    with pytest.raises(TypeError):
        ETDConfig(contour_points=3.14)


def test_solver_type_enum():
    assert SolverType.CS is SolverType.CONSTANT_STEP and SolverType.AS is SolverType.ADAPTIVE_STEP
    assert str(SolverType.ADAPTIVE_STEP) == "Adaptive Step"
    with pytest.raises(TypeError):
        SolverType.from_solver(object())


def test_logger_names_and_levels():
    lg = get_solver_logger(rk.ETD35, "INFO")
    assert lg.name == "rkstiff.ETD35" and lg.level == logging.INFO
    set_log_level(lg, "debug")
    assert lg.level == logging.DEBUG
    with pytest.raises(ValueError):
        set_log_level(lg, "NOT_A_LEVEL")


def test_controller_limits_and_exceptions_exposed():
    assert (BaseSolverAS.MAX_LOOPS, BaseSolverAS.MAX_S, BaseSolverAS.MIN_S) == (50, 4.0, 0.25)
    assert issubclass(BaseSolverAS.MaxLoopsExceeded, BaseSolverAS.SolverError)
    assert issubclass(BaseSolverAS.MinimumStepReached, BaseSolverAS.SolverError)
    assert issubclass(BaseSolverAS.SolverError, RuntimeError)


def test_module_surface_matches_reference_imports():
    """import paths the reference's demos and tests use."""
    from rkstiff_b200.etd import SolverConfig as a  # noqa: F401
    from rkstiff_b200.etd34 import ETD34, ETDConfig as b, SolverConfig as c  # noqa: F401
    from rkstiff_b200.etd35 import ETD35, ETDConfig as d, SolverConfig as e  # noqa: F401
    from rkstiff_b200.etd4 import ETD4  # noqa: F401
    from rkstiff_b200.etd5 import ETD5  # noqa: F401
    from rkstiff_b200.if4 import IF4  # noqa: F401
    from rkstiff_b200.if34 import IF34  # noqa: F401
    from rkstiff_b200.if45dp import IF45DP  # noqa: F401
    from rkstiff_b200 import grids, models  # noqa: F401
    for cls in (rk.IF34, rk.ETD34, rk.ETD35, rk.IF45DP):
        assert issubclass(cls, BaseSolverAS)


def test_matrix_operators_are_rejected_not_emulated():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        # a 1-D operator is already diagonal: the flag is ignored (reference tests/test_etd35.py:202-205)
        lin = torch.zeros(8, dtype=torch.float64, device="cuda")
        assert rk.ETD35(lin, lambda v: v, diagonalize=True)._S is None
        # a dense matrix is diagonalised; a singular one is refused (tests/test_etd35.py:255-262)
        with pytest.raises(ValueError):
            rk.IF34(torch.tensor([[1.0, 2.0], [2.0, 4.0]], dtype=torch.float64, device="cuda"), lambda v: v, diagonalize=True)
    with pytest.raises(TypeError):
        rk.ETD4([1.0, 2.0], lambda v: v)


def test_device_derivatives_match_reference_semantics():
    """derivatives.dx_rfft / dx_fft: same results and the same errors as rkstiff/derivatives.py:47-179
    (its doctests: d/dx sin = cos, d/dx e^{ix} = i e^{ix}); torch tensors, leading batch allowed."""
    import math

    import numpy as np
    import torch
    from rkstiff_b200 import derivatives as d
    n, length = 128, 2 * math.pi
    x = torch.arange(n, dtype=torch.float64) * (length / n)
    kr = torch.from_numpy(np.fft.rfftfreq(n, d=length / n) * 2 * np.pi)
    kc = torch.from_numpy(np.fft.fftfreq(n, d=length / n) * 2 * np.pi)
    assert torch.allclose(d.dx_rfft(kr, torch.sin(x)), torch.cos(x), atol=1e-10)
    assert torch.allclose(d.dx_rfft(kr, torch.sin(x), 2), -torch.sin(x), atol=1e-10)
    u = torch.exp(1j * x)
    assert torch.allclose(d.dx_fft(kc, u), 1j * u, atol=1e-10)
    batch = torch.stack([torch.sin(x), torch.cos(2 * x)])
    ref = np.fft.irfft((1j * kr.numpy()) ** 3 * np.fft.rfft(batch.numpy(), axis=-1), n=n, axis=-1)
    np.testing.assert_allclose(d.dx_rfft(kr, batch, 3).numpy(), ref, rtol=0, atol=1e-9)
    assert d.dx_rfft(kr, batch, 0) is batch and d.dx_fft(kc, u, 0) is u
    with pytest.raises(TypeError):
        d.dx_rfft(kr, torch.sin(x), 1.5)
    with pytest.raises(ValueError):
        d.dx_rfft(kr, torch.sin(x), -1)
    with pytest.raises(TypeError):
        d.dx_rfft(kr, u)
    with pytest.raises(ValueError):
        d.dx_rfft(kc, torch.sin(x))
    with pytest.raises(ValueError):
        d.dx_fft(kr, u)
    assert d.dx_rfft(kr, torch.empty(0, dtype=torch.float64)).numel() == 0


def test_chebyshev_grid_and_dense_allen_cahn_operator():
    """grids.construct_x_dx_cheb / models.allen_cahn_ops (rkstiff/grids.py:134-220, models.py:202-262) on the
    CPU device: points of chebpts2, D 1 = 0, D x = 1, D x^2 = 2x, and the interior operator eps D^2 + I."""
    import numpy as np
    import torch
    from oracle import problems
    from rkstiff_b200 import grids, models
    n = 20
    x, d = grids.construct_x_dx_cheb(n, -1.0, 1.0, device="cpu")
    np.testing.assert_allclose(x.numpy(), np.polynomial.chebyshev.chebpts2(n + 1), rtol=0, atol=1e-15)
    assert float((d @ torch.ones_like(x)).abs().max()) < 1e-12
    assert float((d @ x - 1.0).abs().max()) < 1e-12
    assert float((d @ x ** 2 - 2 * x).abs().max()) < 1e-11
    lin, nl = models.allen_cahn_ops(x, d, 0.01)
    p = problems.allen_cahn_cheb(n)
    np.testing.assert_allclose(lin.numpy(), p.lin_op, rtol=0, atol=1e-11)
    w = torch.from_numpy(p.u0)
    np.testing.assert_allclose(nl(w).numpy(), p.nl_func(p.u0), rtol=0, atol=1e-15)
    with pytest.raises(ValueError):
        grids.construct_x_cheb(1, device="cpu")
    with pytest.raises(TypeError):
        grids.construct_x_cheb(4.0, device="cpu")



def test_derivative_multiplier_matches_numpy_powers():
    """derivatives._ik_power: (i kx)^n with exact zeros in the vanishing part, as NumPy's repeated complex products
    give (rkstiff/derivatives.py:119,176); magnitudes within one rounding of NumPy's."""
    import numpy as np
    import torch
    from rkstiff_b200.derivatives import _engine_length, _ik_power
    kx = 2 * np.pi * np.fft.fftfreq(64, d=0.3)
    for order in range(0, 9):
        got = _ik_power(torch.from_numpy(kx), order).numpy()
        ref = (1j * kx) ** order
        assert np.array_equal(got.real == 0, ref.real == 0) and np.array_equal(got.imag == 0, ref.imag == 0)
        np.testing.assert_allclose(got, ref, rtol=4e-16, atol=0)
    assert [n for n in (8, 16, 96, 512, 8192, 16384) if _engine_length(n)] == [16, 512, 8192]
