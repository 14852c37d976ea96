"""CPU restatement of the fit path's two numerical pieces.

TEST INFRASTRUCTURE ONLY (see fx_oracle.h): the checker of fx_scaled_fit_batch and of the batched
objective of frontx_b200.fit; nothing under frontx_b200/ may import it.

Follows /root/reference/src/frontx/_inverse/fit.py:
  :17-107   ScaledSolution: theta(o) = original(o / sqrt(D0)), oi = original.oi sqrt(D0)
  :45-82    ScaledSolution.fitting_data: least squares on the single unknown D0 of the residuals
            (scaled(o) - theta) / sigma, started from (o[-1] / original.oi)**2
  :145-151  the cost of fit(): mean(((sol(o) - theta) / sigma)**2), inf for a failed solve
The reference minimises with optimistix.LevenbergMarquardt(rtol=1e-3, atol=1e-6) (not installable here);
this restatement runs MINPACK's Levenberg-Marquardt (scipy.optimize.least_squares(method="lm")) on the same
residual function of D0 from the same start, to full precision: the minimiser both converge to.  The
reference's iterate stops within its own tolerances of it, so D0 is pinned here at the MINIMISER, not at the
reference's last iterate ("parity unpinned" for the iterate itself).
"""

from __future__ import annotations

import numpy as np
from scipy.optimize import brentq, least_squares

from . import fxo


def _k5(knots: np.ndarray) -> np.ndarray:
    k = np.asarray(knots, dtype=np.float64)
    return k if k.shape[1] == 5 else np.concatenate([k[:, :3], k[:, 2:3], k[:, 3:4]], axis=1)


def scaled(knots: np.ndarray, D0: float, o) -> np.ndarray:
    """ScaledSolution.__call__ (fit.py:84-89) on a dense-output table: original(o / sqrt(D0)), where
    original clips to [0, oi] (src/frontx/_forward.py:30)."""
    th, _ = fxo.evaluate(_k5(knots), np.asarray(o, dtype=np.float64) / np.sqrt(D0))
    return th


def cost(knots: np.ndarray, D0: float, o, theta, sigma=1.0, solved: bool = True) -> float:
    """fit.py:145-151."""
    if not solved:
        return float("inf")
    r = (scaled(knots, D0, o) - np.asarray(theta, dtype=np.float64)) / np.asarray(sigma, dtype=np.float64)
    return float(np.mean(r ** 2))


def fitting_data(knots: np.ndarray, oi: float, o, theta, sigma=1.0):
    """fit.py:45-82: returns (D0, cost at D0)."""
    o = np.asarray(o, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), o.shape)

    def residuals(x):
        return (scaled(knots, float(x[0]), o) - theta) / sigma

    y0 = (o[-1] / oi) ** 2  # fit.py:66
    opt = least_squares(residuals, x0=[y0], method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15, x_scale=[y0],
                        diff_step=1e-7)
    D0 = float(opt.x[0])
    # MINPACK's finite-difference Jacobian leaves D0 good to ~1e-8; the stationary point itself is the root of the
    # ANALYTIC gradient sum r dr/dD0 (the table's cubics differentiate exactly), bracketed around LM's answer
    def gradient(D):
        x = o / np.sqrt(D)
        th, dth = hermite(knots, x)
        return float(np.sum(((th - theta) / sigma) * (dth * x * (-0.5 / D) / sigma)))

    lo, hi = D0 * (1 - 1e-6), D0 * (1 + 1e-6)
    if gradient(lo) * gradient(hi) < 0:
        D0 = float(brentq(gradient, lo, hi, xtol=1e-300, rtol=1e-15))
    return D0, cost(knots, D0, o, theta, sigma)


def hermite(knots: np.ndarray, x):
    """theta and d theta/dx of the dense-output cubics at x (clipped to the table like Solution.__call__;
    the derivative of the clipped function is 0 outside)."""
    k = _k5(knots)
    x = np.asarray(x, dtype=np.float64)
    inside = (x > k[0, 0]) & (x < k[-1, 0])
    xc = np.clip(x, k[0, 0], k[-1, 0])
    idx = np.clip(np.searchsorted(k[:, 0], xc, side="left") - 1, 0, len(k) - 2)
    a, b = k[idx], k[idx + 1]
    dt = b[:, 0] - a[:, 0]
    s = (xc - a[:, 0]) / dt
    k0, k1 = dt * a[:, 3], dt * b[:, 3]
    ca = k0 + k1 + 2 * a[:, 1] - 2 * b[:, 1]
    cb = -2 * k0 - k1 - 3 * a[:, 1] + 3 * b[:, 1]
    th = ((ca * s + cb) * s + k0) * s + a[:, 1]
    dth = np.where(inside, ((3 * ca * s + 2 * cb) * s + k0) / dt, 0.0)
    return th, dth
