"""Pin the inverse-path oracle (oracle/fx_oracle_inverse.c).

interpax's PchipInterpolator/PPoly (what the reference calls,
/root/reference/src/frontx/_inverse/interpolated.py:6,30,47-49) are ports of
scipy's; scipy is installed, so the restatement is checked against a literal
scipy twin of interpolated.py:18-73 and against the reference's analytic tests
(/root/reference/tests/test_inverse.py:29-62).
"""

from __future__ import annotations

import numpy as np
import pytest
from scipy.interpolate import PchipInterpolator


def scipy_twin(o, theta):
    """interpolated.py:30-50,59-73 with scipy in place of interpax (b=None, i=None)."""
    sol = PchipInterpolator(o, theta)
    i = theta[-1]
    oi = o[-1]
    th_u, idx = np.unique(theta, return_index=True)
    o_u = o[idx]
    inverse = PchipInterpolator(th_u, o_u, extrapolate=False)
    do_dtheta = inverse.derivative()
    Iodtheta = inverse.antiderivative()
    c = Iodtheta(i)

    def D(th):
        return -(do_dtheta(th) * (Iodtheta(th) - c)) / 2

    anti = sol.antiderivative()
    sorp = (anti(oi) - anti(0)) - sol(oi) * (oi - 0)
    return D, float(sorp)


class ScipyInterpolated:
    """interpolated.py:18-73 line for line, scipy's PchipInterpolator/PPoly in place of interpax's port of them
    (b=, i= insertions, jnp.unique, extrapolate=False on the inverse, sorptivity at any o)."""

    def __init__(self, o, theta, b=None, i=None):
        o = np.asarray(o, dtype=np.float64)
        theta = np.asarray(theta, dtype=np.float64)
        self._sol = PchipInterpolator(o, theta)
        if b is not None:
            o = np.insert(o, 0, 0)
            theta = np.insert(theta, 0, b)
        if i is not None:
            o = np.append(o, o[-1] + 1)
            theta = np.append(theta, i)
        else:
            i = theta[-1]
        self.oi = o[-1]
        theta, indices = np.unique(theta, return_index=True)
        o = o[indices]
        inverse = PchipInterpolator(theta, o, extrapolate=False)
        self._do_dtheta = inverse.derivative()
        self._Iodtheta = inverse.antiderivative()
        self._c = self._Iodtheta(i)

    def __call__(self, o):
        return self._sol(o)

    def d_do(self, o):
        return self._sol.derivative()(o)

    @property
    def i(self):
        return self(self.oi)

    def D(self, theta):  # noqa: N802
        return -(self._do_dtheta(theta) * (self._Iodtheta(theta) - self._c)) / 2

    def sorptivity(self, o=0):
        anti = self._sol.antiderivative()
        return (anti(self.oi) - anti(o)) - self.i * (self.oi - o)


def profiles():
    rng = np.random.default_rng(7)
    o = np.linspace(0, 20, 100)
    yield "exp", o, np.exp(-o)
    a = 1.7
    o2 = np.sort(rng.uniform(0, 12, 257))
    o2[0] = 0.0
    yield "ragged", o2, 0.1 + 0.8 * np.exp(-a * o2)
    o3 = np.linspace(0.5, 9, 40)  # does not start at 0: sorptivity extrapolates the first cubic
    yield "offset", o3, 1.0 / (1.0 + o3) ** 2
    o4 = np.linspace(0, 3, 33)
    yield "increasing", o4, 0.2 + 0.5 * np.tanh(o4)  # i > b
    yield "three", np.array([0.0, 1.0, 2.5]), np.array([1.0, 0.4, 0.1])
    yield "two", np.array([0.0, 2.0]), np.array([1.0, 0.25])


@pytest.mark.parametrize("case", list(profiles()), ids=lambda c: c[0])
def test_oracle_matches_scipy_twin(fxo, case):
    _, o, th = case
    rc, Dk, s_pchip, s_trapz = fxo.inverse(o, th)
    assert rc == 0
    D, sorp = scipy_twin(o, th)
    want = D(th)
    # D at the far-field knot is 0 * slope; compare absolutely there
    np.testing.assert_allclose(Dk, want, rtol=1e-12, atol=1e-15 * np.abs(want).max())
    assert s_pchip == pytest.approx(sorp, rel=1e-12)
    assert s_trapz == pytest.approx(np.trapezoid(th, o), rel=1e-13)


def test_reference_analytic_cases(fxo):
    """tests/test_inverse.py:29-39,51-62 of the reference, through the oracle."""
    o = np.linspace(0, 20, 100)
    th = np.exp(-o)
    rc, Dk, s_pchip, _ = fxo.inverse(o, th)
    assert rc == 0
    # D(theta) = (1 - ln theta)/2 ; the reference checks abs 5e-2 on theta in [1e-6, 1]
    sel = th >= 1e-6
    assert np.abs(Dk[sel] - 0.5 * (1 - np.log(th[sel]))).max() < 5e-2
    assert s_pchip == pytest.approx(1.0, abs=1e-4)
    o5 = np.linspace(0, 20, 500)
    _, _, _, tz = fxo.inverse(o5, np.exp(-o5))
    # frontx.sorptivity(o, theta, b=1, i=0): prepend (0, b) and subtract i*oi -- both no-ops here
    assert tz == pytest.approx(1.0, abs=1e-3)


def test_duplicates_are_dropped_like_unique(fxo):
    """jnp.unique keeps the first occurrence of a repeated theta (flat far-field tail)."""
    o = np.linspace(0, 10, 30)
    th = np.maximum(np.exp(-o), np.exp(-6.0))  # flat tail
    rc, Dk, _, _ = fxo.inverse(o, th)
    assert rc == 0
    flat = np.where(th == th[-1])[0]
    assert np.isfinite(Dk[flat[0]]) and np.all(np.isnan(Dk[flat[1:]]))


def test_scipy_interpolated_twin_passes_the_reference_tests():
    """The literal scipy twin used by the GPU tests of duplicates / insertions / d_do / sorptivity(o) reproduces
    /root/reference/tests/test_inverse.py:29-39,51-62 and agrees with the simpler twin above."""
    o = np.linspace(0, 20, 100)
    sol = ScipyInterpolated(o, np.exp(-o))
    assert sol(o) == pytest.approx(np.exp(-o))
    theta = np.linspace(1e-6, 1, 100)
    assert sol.D(theta) == pytest.approx(0.5 * (1 - np.log(theta)), abs=5e-2)
    assert sol.sorptivity() == pytest.approx(1, abs=1e-4)
    D, sorp = scipy_twin(o, np.exp(-o))
    np.testing.assert_array_equal(sol.D(theta), D(theta))
    assert float(sol.sorptivity()) == sorp
    # jax.grad of the PCHIP = its derivative; of exp(-o): -exp(-o)
    assert sol.d_do(o[5:60]) == pytest.approx(-np.exp(-o[5:60]), rel=2e-2)
