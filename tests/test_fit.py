"""``Param`` / ``ScaledSolution`` / ``fit``.

The two ``ScaledSolution`` tests restate /root/reference/tests/test_inverse.py:65-101 with the
reference's own tolerances; ``fit`` is checked on synthetic data generated from a known model
(the reference has no test of ``fit`` itself).
"""

from __future__ import annotations

import numpy as np
import pytest
from conftest import B_PAPER, I_PAPER

import frontx_b200 as fx
from frontx_b200 import _fit
from frontx_b200.models import LETd, VanGenuchten


def letd():
    return LETd(L=0.004569, E=12930, T=1.505, Dwt=4.660e-4, theta_range=(0.019852, 0.7))


# ---- host logic (no GPU) -------------------------------------------------------------------
def test_param_defaults_and_substitution():
    assert fx.Param(min=0, max=2).value == 1.0 and fx.Param(min=3).value == 4.0  # param.py:21-32
    assert fx.Param(max=3).value == 2.0 and fx.Param().value == 0.0 and fx.Param(5, min=0, max=9).value == 5.0
    m = LETd(L=fx.Param(0.5, min=0, max=2), E=fx.Param(1e4, min=1, max=1e5), T=1.5, Dwt=4.66e-4,
             theta_range=(fx.Param(0.01, min=0, max=0.02), 0.7))
    ps = _fit.get_params(m)
    assert [p.value for p in ps] == [0.5, 1e4, 0.01]
    fixed = _fit.set_param_values(m, [0.1, 2e4, 0.015])
    assert fixed == LETd(L=0.1, E=2e4, T=1.5, Dwt=4.66e-4, theta_range=(0.015, 0.7))
    X = np.array([[0.1, 0.2], [2e4, 3e4], [0.015, 0.012]])
    mb = _fit._candidates(m, X)
    np.testing.assert_array_equal(mb.params[0], fixed.params())
    np.testing.assert_array_equal(mb.params[1], LETd(L=0.2, E=3e4, T=1.5, Dwt=4.66e-4, theta_range=(0.012, 0.7)).params())
    with pytest.raises(ValueError):
        fx.fit(letd(), [0, 1], [0.5, 0.1], i=0.025, b=0.7)  # nothing to fit
    with pytest.raises(TypeError):
        _fit.get_params(lambda th: th)


# ---- reference tests (GPU) -----------------------------------------------------------------
@pytest.mark.gpu
def test_scaled_solution_sorptivity():
    """/root/reference/tests/test_inverse.py:65-80"""
    D = letd()
    sol = fx.solve(D, b=B_PAPER, i=I_PAPER)
    unscaled = fx.ScaledSolution(sol, D0=1 / D.Dwt)
    scaled = fx.ScaledSolution.with_sorptivity(unscaled, sol.sorptivity())
    assert scaled.sorptivity() == pytest.approx(sol.sorptivity())
    assert scaled.D0 == pytest.approx(D.Dwt)
    assert scaled.oi == pytest.approx(sol.oi)
    assert scaled.result == fx.RESULTS.successful


@pytest.mark.gpu
def test_scaled_solution_data():
    """/root/reference/tests/test_inverse.py:83-101"""
    D = letd()
    sol = fx.solve(D, b=B_PAPER, i=I_PAPER)
    unscaled = fx.ScaledSolution(sol, D0=1 / D.Dwt)
    o = np.linspace(0, sol.oi, 500)
    scaled = fx.ScaledSolution.fitting_data(unscaled, o, sol(o))
    assert scaled(o) == pytest.approx(sol(o), abs=1e-6)
    assert scaled.d_do(o) == pytest.approx(sol.d_do(o), abs=1e-2)
    assert scaled.D0 == pytest.approx(D.Dwt)
    assert scaled.sorptivity() == pytest.approx(sol.sorptivity())
    assert scaled.oi == pytest.approx(sol.oi)
    assert scaled.result == fx.RESULTS.successful


@pytest.mark.gpu
def test_scaled_fit_batch_recovers_known_factors():
    """Every candidate of a batch gets its own D0: data generated from candidate j rescaled by a
    known factor f_j is fitted back to f_j (and to cost ~ 0); a failed solve costs inf (fit.py:145-151)."""
    import torch

    models = [letd(), VanGenuchten(n=8.093, l=2.344, Ks=2.079e-6, theta_range=(0.004943, 0.7))]
    sols = fx.solve_batch(models[:2], b=B_PAPER, i=I_PAPER, k_max=512)
    factors = [2.5, 0.3]
    for j, f in enumerate(factors):
        o = np.linspace(0, float(sols.oi[j]) * np.sqrt(f), 300)
        theta = sols[j](o / np.sqrt(f))
        D0, cost, status = _fit._scaled_fit(sols, o, theta, 1.0, mode=1)
        assert int(status[j]) == 0
        assert float(D0[j]) == pytest.approx(f, rel=1e-9)
        assert float(cost[j]) < 1e-20
        # the other candidate is a different curve: positive cost, finite D0
        assert float(cost[1 - j]) > 1e-8 and np.isfinite(float(D0[1 - j]))
    # mode 0 / mode 2 agree with the definition evaluated through the Solution surface
    o = np.linspace(0, 0.002, 200)
    theta = np.full_like(o, 0.3)
    sig = np.linspace(1.0, 2.0, 200)
    _, cost0, _ = _fit._scaled_fit(sols, o, theta, sig, mode=0)
    _, cost2, _ = _fit._scaled_fit(sols, o, theta, sig, mode=2, D0=torch.tensor([2.0, 0.5], dtype=torch.float64))
    for j in range(2):
        assert float(cost0[j]) == pytest.approx(np.mean(((sols[j](o) - theta) / sig) ** 2), rel=1e-12)
        d0 = (2.0, 0.5)[j]
        assert float(cost2[j]) == pytest.approx(np.mean(((sols[j](o / np.sqrt(d0)) - theta) / sig) ** 2), rel=1e-12)
    # an unsolvable candidate (tests/test_solve.py:148-156: D = theta, b = 1, i = -1) costs inf
    bad = fx.solve_batch([fx.D.power_law(k=1)], b=1.0, i=-1.0, k_max=512, throw=False)
    _, cost, status = _fit._scaled_fit(bad, o, theta, 1.0, mode=1)
    assert np.isinf(float(cost[0])) and int(status[0]) == 1


@pytest.mark.gpu
def test_fit_recovers_a_known_model():
    """Data from a LETd model with a different overall scale: fit() finds L, E and the scale."""
    truth = LETd(L=0.3, E=5000.0, T=1.5, Dwt=4.66e-4, theta_range=(0.019852, 0.7))
    sol = fx.solve(truth, b=B_PAPER, i=I_PAPER)
    o = np.linspace(0, sol.oi, 120)
    theta = sol(o)
    family = LETd(L=fx.Param(1.0, min=0.01, max=2.0), E=fx.Param(1e4, min=1e3, max=2e4), T=1.5, Dwt=1.0,
                  theta_range=(0.019852, 0.7))
    res = fx.fit(family, o, theta, i=I_PAPER, b=B_PAPER, max_steps=25, seed=1)
    assert isinstance(res, fx.ScaledSolution) and res.result == fx.RESULTS.successful
    assert res.cost < 1e-7
    assert np.max(np.abs(res(o) - theta)) < 2e-3
    assert res.sorptivity() == pytest.approx(sol.sorptivity(), rel=5e-3)
    # the product Dwt*D0 is what the data determine
    assert res.D0 * res.fitted_model.Dwt == pytest.approx(truth.Dwt, rel=0.1)
    # fit_D0="sorptivity" and None run the same generations through the other two cost modes
    res_s = fx.fit(family, o, theta, i=I_PAPER, b=B_PAPER, fit_D0="sorptivity", max_steps=6, seed=1, polish=False)
    assert isinstance(res_s, fx.ScaledSolution) and np.isfinite(res_s.cost)
    fam2 = LETd(L=fx.Param(0.5, min=0.01, max=2.0), E=5000.0, T=1.5, Dwt=4.66e-4, theta_range=(0.019852, 0.7))
    res_n = fx.fit(fam2, o, theta, i=I_PAPER, b=B_PAPER, fit_D0=None, max_steps=10, seed=2)
    assert isinstance(res_n, fx.Solution) and res_n.fitted_model.L == pytest.approx(0.3, abs=0.05)
