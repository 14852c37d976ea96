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


def test_tabulated_diffusivity_oracle(fxo):
    """FXO_TABULATED = InterpolatedSolution.D as a model of the forward oracle (so that the reference's
    solve(D=sol.D) test has a checker): its D equals the scipy twin's, its D' and D'' the derivatives of that, and
    the kernel's own arithmetic (twin) agrees on the same table."""
    o = np.linspace(0, 20, 100)
    th = np.exp(-o)
    p, table = fxo.tabulated(o, th)
    want = ScipyInterpolated(o, th)
    tq = np.concatenate([np.linspace(th.min(), th.max(), 400), th[3:40:5]])
    got = fxo.D(fxo.TABULATED, p, tq)
    np.testing.assert_allclose(got[:, 0], want.D(tq), rtol=1e-12, atol=1e-300)
    eps = 1e-9
    away = np.min(np.abs(tq[:, None] - th[None, :]), axis=1) > 3 * eps  # D' jumps at the knots
    inner = tq[away & (tq > th.min() + 3 * eps) & (tq < th.max() - 3 * eps)]
    d = fxo.D(fxo.TABULATED, p, inner)
    num1 = (want.D(inner + eps) - want.D(inner - eps)) / (2 * eps)
    np.testing.assert_allclose(d[:, 1], num1, rtol=1e-5, atol=1e-6 * np.abs(num1).max())
    np.testing.assert_allclose(fxo.twin_D(fxo.TABULATED, p, tq), got, rtol=1e-12, atol=1e-300)
    assert np.isnan(fxo.D(fxo.TABULATED, p, [th.min() - 1e-9, th.max() + 1e-9])).all()  # extrapolate=False
    del table


def test_reference_interpolated_exact_solve_through_the_oracle(fxo):
    """/root/reference/tests/test_inverse.py:42-48: solve(D=sol.D, b=1, i=1e-3, itol=5e-3) vs exp(-o) to 1e-2.

    NOT REPRODUCED by the restatement, and this test records why.  The PCHIP-derived D has a kink at every knot
    and is NaN below the smallest theta of the data; the event's second clause (theta < i - itol = -4e-3) can never
    fire, so a shot ends where theta' rounds to >= 0 on a plateau reached in ~25 huge steps, or crawls into the NaN
    region (-inf).  The residual as a function of d_dob is then a SAWTOOTH: on the shallow side of the jump to -inf
    only ~5 % of the d_dob values give |residual| < itol.  Whether a bisection succeeds depends on which midpoints
    it happens to visit, i.e. on the last bits of the arithmetic; the reference's run hits a window, this
    restatement (under every bracket-expansion rule, U6) walks to the jump, where the residual is 8e-3.  Had it
    visited a window, the reference's own criterion holds -- checked below."""
    import ctypes as C

    o = np.linspace(0, 20, 100)
    p, table = fxo.tabulated(o, np.exp(-o))
    for rule in (0, 1, 2):
        r = fxo.solve(fxo.TABULATED, p, b=1.0, i=1e-3, itol=5e-3, options=fxo.default_options(u6_expand_rule=rule))
        th, _ = fxo.evaluate(r["knots"], o)
        assert r["counters"][0] == 1 and r["counters"][2] == 1  # max_steps_reached after one expansion
        assert np.abs(th - np.exp(-o)).max() < 4e-2 and abs(r["summary"][0] + 1.06) < 1e-3

    def shoot(d):
        cnt = np.zeros(8, dtype=np.int32)
        nk = C.c_int()
        kn = np.zeros((512, 5))
        res = fxo.lib().fxo_shoot(fxo.TABULATED, fxo._dp(p), 1.0, 1e-3, 5e-3, float(d), None, 512, fxo._dp(kn),
                                  C.byref(nk), fxo._ip(cnt))
        return res, kn[: min(nk.value, 512)]

    ds = np.linspace(-1.0599, -0.85, 120)
    res = np.array([shoot(d)[0] for d in ds])
    frac = float((np.abs(res) < 5e-3).mean())
    assert 0.0 < frac < 0.2, frac  # a few narrow windows
    hit = ds[np.abs(res) < 5e-3]
    hit = hit[np.argmin(np.abs(hit + 1.0))]
    _, kn = shoot(hit)
    th, _ = fxo.evaluate(kn, o)
    assert np.abs(th - np.exp(-o)).max() < 1e-2  # the reference's criterion, in a window
    # the kernel's arithmetic (twin) on the same table behaves the same way
    tw = fxo.twin_solve_batch([fxo.TABULATED], p[None, :], 1.0, 1e-3, itol=5e-3, k_max=512)
    assert tw["counters"][0, 0] == 1 and abs(tw["summary"][0, 0] + 1.06) < 1e-3
    del table
