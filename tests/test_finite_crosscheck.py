"""Independent cross-check of the shooting solve against a method-of-lines PDE integration
(the reference's own validation of ``solve``:
/root/reference/tests/test_finite.py:39-94 with /root/reference/src/frontx/finite.py:127-158).

Nothing here shares code or algorithm with the product or with the oracle: the PDE
``d theta/dt = d/dr (D(theta) d theta/dr)`` is discretised with the reference's three-point stencil
(face diffusivity = mean of the two cells, first cell clamped to b, zero-flux outer end) and
integrated in time by scipy's BDF; D(theta) is the LETd definition (src/frontx/models.py:62-66)
written in numpy.  Tolerances are the reference's.
"""

from __future__ import annotations

import numpy as np
import pytest
from conftest import B_PAPER, I_PAPER
from scipy.integrate import solve_ivp
from scipy.sparse import diags

import frontx_b200 as fx
from frontx_b200.models import LETd

L_, E_, T_, DWT, TR, TS = 0.004569, 12930.0, 1.505, 4.660e-4, 0.019852, 0.7


def letd_numpy(theta):
    Se = (theta - TR) / (TS - TR)
    return DWT * Se**L_ / (Se**L_ + E_ * (1 - Se) ** T_)


def finite_solve(r1, t_eval, y0, clamp_first):
    """finite.py:127-158 with scipy BDF in place of diffrax Kvaerno5."""
    n = y0.size
    dr2 = (r1 / (n - 1)) ** 2

    def rhs(_, theta):
        D = letd_numpy(theta)
        Df = (D[1:] + D[:-1]) / 2
        out = np.empty_like(theta)
        out[0] = 0.0 if clamp_first else Df[0] / dr2 * (theta[1] - theta[0])
        out[1:-1] = Df[:-1] / dr2 * theta[:-2] - (Df[1:] + Df[:-1]) / dr2 * theta[1:-1] + Df[1:] / dr2 * theta[2:]
        out[-1] = Df[-1] / dr2 * (theta[-2] - theta[-1])
        return out

    sparsity = diags([np.ones(n - 1), np.ones(n), np.ones(n - 1)], [-1, 0, 1])
    sol = solve_ivp(rhs, (0.0, float(t_eval[-1])), y0, method="BDF", t_eval=t_eval, rtol=1e-6, atol=1e-9,
                    jac_sparsity=sparsity)
    assert sol.success
    return sol.y


@pytest.mark.gpu
def test_validity_against_the_pde():
    """test_finite.py:39-69 (LETd case): max |theta_PDE(r, 1) - theta_shooting(r, 1)| < 2e-2."""
    D = LETd(L=L_, E=E_, T=T_, Dwt=DWT, theta_range=(TR, TS))
    ref = fx.solve(D, b=B_PAPER, i=I_PAPER)
    r = np.linspace(0, 0.0025, 500)
    y0 = np.full(r.size, I_PAPER)
    y0[0] = B_PAPER
    y = finite_solve(r[-1], np.array([1.0]), y0, clamp_first=True)[:, -1]
    assert y == pytest.approx(ref(r, 1.0), abs=2e-2)
    # much tighter away from the front, where the 5e-6 grid resolves the profile
    assert np.median(np.abs(y - ref(r, 1.0))) < 2e-3


@pytest.mark.gpu
def test_mass_conservation_of_the_shooting_profile():
    """test_finite.py:72-94: started from the shooting profile at t = 1, the PDE keeps the uptake
    Int (theta - i) dr equal to the shooting solution's sorptivity (the profile is self-similar, the
    uptake at time t is S sqrt(t); here the outer end is closed, so what is in stays in)."""
    D = LETd(L=L_, E=E_, T=T_, Dwt=DWT, theta_range=(TR, TS))
    sol = fx.solve(D, b=B_PAPER, i=I_PAPER)
    r = np.linspace(0, 0.0025, 500)
    t = np.linspace(0, 1000.0, 26)
    total = sol.sorptivity()
    y = finite_solve(r[-1], t, np.asarray(sol(r, 1.0), dtype=np.float64), clamp_first=False)
    for k in range(t.size):
        assert np.trapezoid(y[:, k] - I_PAPER, r) == pytest.approx(total, abs=1e-4)
