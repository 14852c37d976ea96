"""Profile -> D(theta) inverse: ``InterpolatedSolution``, ``sorptivity`` and the
batched ``inverse`` entry point.

Reference: src/frontx/_inverse/interpolated.py:18-73 (PCHIP of theta(o); PCHIP of
o(theta) on ``jnp.unique``-sorted theta; ``D = -(do/dtheta * Int_i^theta o dtheta)/2``)
and src/frontx/_inverse/sorptivity.py:8-19.  The PCHIP slopes, the per-interval cubic
integrals and their running sum are one scan kernel (``fx_inverse_batch``); queries at
arbitrary theta are ``fx_inverse_query_batch``; value, derivative and antiderivative
of the forward interpolant are ``fx_pchip_query_batch`` on what the same scan leaves
behind when o and theta swap roles.

A profile whose theta is strictly monotone in o -- the case the scan is built for -- needs
no sort: ``jnp.unique`` is then a reversal.  Anything else (plateaus at theta = i, repeated
or noisy values, an appended ``i`` equal to the last value) goes through
``fx_unique_batch``: a stable radix sort by theta and a first-occurrence dedupe, exactly
what ``jnp.unique(theta, return_index=True)`` returns (interpolated.py:44-45).
"""

from __future__ import annotations

from typing import Any

import numpy as np

from . import _lib
from ._boltzmann import AbstractSolution, boltzmannmethod


def _stream(torch, dev):
    return torch.cuda.current_stream(dev).cuda_stream


def unique_profile(o: Any, theta: Any, *, device: Any = None):
    """``theta, idx = jnp.unique(theta, return_index=True); o = o[idx]`` for ONE profile, on the GPU.
    Returns (o_u, theta_u, idx) device tensors of the deduplicated length, theta ascending."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    o_t = torch.as_tensor(o, dtype=torch.float64, device=dev).reshape(-1).contiguous()
    th_t = torch.as_tensor(theta, dtype=torch.float64, device=dev).reshape(-1).contiguous()
    n = int(o_t.numel())
    o_u, th_u = torch.empty_like(o_t), torch.empty_like(th_t)
    idx = torch.empty((n,), dtype=torch.int32, device=dev)
    count = torch.zeros((1,), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fx_unique_batch(1, n, o_t.data_ptr(), th_t.data_ptr(), o_u.data_ptr(), th_u.data_ptr(),
                                              idx.data_ptr(), count.data_ptr(), _stream(torch, dev)))
    m = int(count.item())
    return o_u[:m], th_u[:m], idx[:m]


class BatchInverse:
    """Result of ``inverse(o, theta)`` for [batch, npts] profiles; tensors stay on the GPU."""

    def __init__(self, o, theta, D, sorp, status, slope, anti) -> None:
        self.o, self.theta = o, theta
        self.D_at_knots = D
        self._sorp = sorp
        self.status = status
        self._slope, self._anti = slope, anti
        # profiles that needed the unique-sort: their own scan of the sorted, deduplicated rows + the theta
        # the integral starts from (the caller's last value, interpolated.py:40,50)
        self._deduped: dict[int, tuple["BatchInverse", float]] = {}

    def D(self, theta_query: Any):
        """InterpolatedSolution.D at theta_query[batch, q] (NaN outside each profile's range)."""
        if self._slope is None:
            raise ValueError("this inverse was computed with query=False: D at the knots only")
        torch = _lib.require_cuda()
        dev = self.o.device
        tq = torch.as_tensor(theta_query, dtype=torch.float64, device=dev)
        batch, npts = self.o.shape
        if tq.ndim == 1:
            tq = tq.unsqueeze(0).expand(batch, -1)
        tq = tq.contiguous()
        q = int(tq.shape[1])
        out = torch.empty((batch, q), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fx_inverse_query_batch(batch, npts, self.o.data_ptr(), self.theta.data_ptr(),
                                                         self._slope.data_ptr(), self._anti.data_ptr(), q,
                                                         tq.data_ptr(), out.data_ptr(), _stream(torch, dev)))
        for p, (sub, i_ref) in self._deduped.items():
            out[p] = _deduped_D(sub, i_ref, tq[p:p + 1])[0]
        return out

    def _need_sorp(self):
        if self._sorp is None:
            raise ValueError("this inverse was computed with sorptivity=False: D only")

    def sorptivity(self):
        """InterpolatedSolution.sorptivity(o=0) per profile (interpolated.py:69-73)."""
        self._need_sorp()
        oi = self.o[:, -1]
        i = self.theta[:, -1]
        return (self._sorp[:, 0] + self._sorp[:, 2]) - i * oi

    def sorptivity_trapezoid(self, b: Any, i: Any):
        """frontx.sorptivity(o, theta, b=, i=) per profile (_inverse/sorptivity.py:16-19)."""
        self._need_sorp()
        torch = _lib.require_cuda()
        dev = self.o.device
        b_ = torch.as_tensor(b, dtype=torch.float64, device=dev)
        i_ = torch.as_tensor(i, dtype=torch.float64, device=dev)
        o0, th0, oi = self.o[:, 0], self.theta[:, 0], self.o[:, -1]
        # prepend (0, b): one more trapezoid; subtract i over [0, oi]
        return self._sorp[:, 1] + 0.5 * (th0 + b_) * o0 - i_ * oi


def _deduped_D(sub: BatchInverse, i_ref: float, tq):
    """D(theta) of a profile that went through the unique-sort.  Its rows are in ascending theta, so the scan
    measured Int o dtheta from the LARGEST theta; the reference measures it from theta = i
    (``c = Iodtheta(i)``, interpolated.py:50,65):  D = -(o'(theta) (I(theta) - I(i)))/2 = D_scan + o'(theta) I(i)/2."""
    torch = _lib.require_cuda()
    dev = sub.o.device
    npts, q = int(sub.o.shape[1]), int(tq.shape[1])
    tq = tq.contiguous()
    at_i = torch.full((1, 1), float(i_ref), dtype=torch.float64, device=dev)
    ant_i = torch.empty((1, 1), dtype=torch.float64, device=dev)
    der = torch.empty((1, q), dtype=torch.float64, device=dev)
    L = _lib.lib()
    args = (1, npts, sub.theta.data_ptr(), sub.o.data_ptr(), sub._slope.data_ptr(), sub._anti.data_ptr())
    with torch.cuda.device(dev):  # the inverse interpolant o(theta): x = theta, y = o, extrapolate=False
        _lib.check(L.fx_pchip_query_batch(*args, 1, at_i.data_ptr(), 1, 0, None, None, ant_i.data_ptr(),
                                          _stream(torch, dev)))
        _lib.check(L.fx_pchip_query_batch(*args, q, tq.data_ptr(), 1, 0, None, der.data_ptr(), None,
                                          _stream(torch, dev)))
    return BatchInverse.D(sub, tq) + der * ant_i[0, 0] / 2


def _scan(o_t, th_t, sorptivity: bool, query: bool):
    torch = _lib.require_cuda()
    dev = o_t.device
    batch, npts = o_t.shape
    D = torch.empty_like(o_t)
    slope = torch.empty_like(o_t) if query else None
    anti = torch.empty_like(o_t) if query else None
    sorp = torch.empty((batch, 4), dtype=torch.float64, device=dev) if sorptivity else None
    status = torch.empty((batch,), dtype=torch.int32, device=dev)
    ptr = lambda x: None if x is None else x.data_ptr()  # noqa: E731
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fx_inverse_batch(batch, npts, o_t.data_ptr(), th_t.data_ptr(), D.data_ptr(),
                                               ptr(sorp), status.data_ptr(), ptr(slope), ptr(anti),
                                               _stream(torch, dev)))
    return BatchInverse(o_t, th_t, D, sorp, status, slope, anti)


def inverse(o: Any, theta: Any, *, device: Any = None, throw: bool = True, sorptivity: bool = True,
            query: bool = True, unique: bool = True) -> BatchInverse:
    """``InterpolatedSolution(o, theta).D`` for a batch of profiles, on the GPU.

    ``o`` and ``theta`` are [batch, npts] (or [npts]); numpy arrays or CUDA tensors.
    ``sorptivity=False`` skips the sorptivity integrals of the forward interpolant (a second PCHIP
    slope per point: a quarter of the kernel's time), ``query=False`` the two knot arrays that
    ``BatchInverse.D(theta_query)`` needs (16 B/point more written): D at the knots only.

    Profiles whose theta is strictly monotone in o take the single-pass scan as they are.  With
    ``unique=True`` (default, the reference's behaviour) every other profile is sorted by theta and
    deduplicated on the device first (``jnp.unique``, interpolated.py:44-45) and its D evaluated at the
    caller's points; ``unique=False`` leaves such profiles flagged (``status`` 1).  ``throw`` raises if a
    profile is left without a valid inverse (fewer than two distinct theta values).
    """
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    o_t = torch.as_tensor(o, dtype=torch.float64, device=dev)
    th_t = torch.as_tensor(theta, dtype=torch.float64, device=dev)
    if o_t.ndim == 1:
        o_t, th_t = o_t.unsqueeze(0), th_t.unsqueeze(0)
    if o_t.shape != th_t.shape or o_t.ndim != 2:
        raise ValueError("o and theta must have the same [batch, npts] shape")
    o_t, th_t = o_t.contiguous(), th_t.contiguous()
    res = _scan(o_t, th_t, sorptivity, query)
    flagged = torch.nonzero(res.status != 0).reshape(-1).tolist() if unique else []
    for p in flagged:
        o_u, th_u, _ = unique_profile(o_t[p], th_t[p], device=dev)
        if o_u.numel() < 2:
            continue  # a constant profile has no inverse: stays flagged
        sub = _scan(o_u.unsqueeze(0).contiguous(), th_u.unsqueeze(0).contiguous(), False, True)
        if int(sub.status[0].item()) != 0:
            continue
        res._deduped[p] = (sub, float(th_t[p, -1].item()))
        res.D_at_knots[p] = _deduped_D(sub, res._deduped[p][1], th_t[p:p + 1])[0]  # at the caller's points
        res.status[p] = 0
    if throw and int(res.status.sum().item()) != 0:
        raise ValueError("theta must take at least two distinct values in every profile" if unique else
                         "theta must be strictly monotone in o for every profile (unique=False)")
    return res


class _ForwardInterpolant:
    """PCHIP theta(o) of one profile: knot slopes and knot antiderivative values from the scan kernel run
    with the roles of o and theta swapped; value / derivative / antiderivative by fx_pchip_query_batch."""

    def __init__(self, o: np.ndarray, theta: np.ndarray) -> None:
        torch = _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device())
        self.x = torch.as_tensor(o, dtype=torch.float64, device=dev).reshape(1, -1).contiguous()
        self.y = torch.as_tensor(theta, dtype=torch.float64, device=dev).reshape(1, -1).contiguous()
        swapped = _scan(self.y, self.x, False, True)  # the kernel's "theta" argument is the abscissa
        if int(swapped.status[0].item()) != 0:
            raise ValueError("o must be strictly increasing")
        self.slope, self.anti = swapped._slope, swapped._anti

    def query(self, o: Any, which: int) -> np.ndarray:
        """which = 0 value, 1 derivative, 2 antiderivative; the end cubics extrapolate (scipy/interpax default)."""
        torch = _lib.require_cuda()
        dev = self.x.device
        oo = np.asarray(o, dtype=np.float64)
        xq = torch.as_tensor(np.ascontiguousarray(oo.reshape(-1)), device=dev)
        q = int(xq.numel())
        out = torch.empty((1, q), dtype=torch.float64, device=dev)
        ptrs = [None, None, None]
        ptrs[which] = out.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fx_pchip_query_batch(1, int(self.x.shape[1]), self.x.data_ptr(), self.y.data_ptr(),
                                                       self.slope.data_ptr(), self.anti.data_ptr(), q, xq.data_ptr(),
                                                       0, 1, ptrs[0], ptrs[1], ptrs[2], _stream(torch, dev)))
        res = out[0].cpu().numpy().reshape(oo.shape)
        return float(res) if oo.ndim == 0 else res


class InterpolatedSolution(AbstractSolution):
    """``frontx.InterpolatedSolution`` (src/frontx/_inverse/interpolated.py:11-73).

    ``sol(o)`` interpolates the data with PCHIP, ``sol.d_do(o)`` is that interpolant's derivative
    (_boltzmann.py:130-135), ``sol.D(theta)`` the diffusivity the profile implies, ``sol.sorptivity(o)``
    its sorptivity.  ``b`` / ``i`` add the boundary point (0, b) and the far-field point (o[-1]+1, i) to
    the inverse interpolant as the reference does (:32-40); repeated theta values are collapsed onto their
    first occurrence (:44-45).
    """

    def __init__(self, o: Any, theta: Any, /, *, b: float | None = None, i: float | None = None) -> None:
        o = np.asarray(o, dtype=np.float64)
        theta = np.asarray(theta, dtype=np.float64)
        if o.ndim != 1 or o.shape != theta.shape:
            raise ValueError("o and theta must be 1-D arrays of the same length")
        self._fwd = _ForwardInterpolant(o, theta)  # interpolated.py:30
        o_inv, th_inv = o, theta
        if b is not None:
            o_inv = np.insert(o_inv, 0, 0.0)
            th_inv = np.insert(th_inv, 0, b)
        if i is not None:
            o_inv = np.append(o_inv, o_inv[-1] + 1)
            th_inv = np.append(th_inv, i)
        self.oi = float(o_inv[-1])  # :42
        # jnp.unique + PchipInterpolator(theta, o, extrapolate=False), its derivative and antiderivative, and
        # c = Iodtheta(i) with i the last value of the (extended) data (:40,44-50)
        self._inv = inverse(o_inv, th_inv, sorptivity=False)
        self._o, self._theta = o, theta

    @boltzmannmethod
    def __call__(self, o: Any) -> Any:  # interpolated.py:52-57: the forward PCHIP theta(o)
        return self._fwd.query(o, 0)

    @boltzmannmethod
    def d_do(self, o: Any) -> Any:  # _boltzmann.py:130-135: jax.grad of __call__ = the PCHIP's derivative
        return self._fwd.query(o, 1)

    def D(self, theta: Any, /) -> Any:  # noqa: N802  interpolated.py:59-67
        torch = _lib.require_cuda()
        th = np.asarray(theta, dtype=np.float64)
        tq = torch.as_tensor(np.ascontiguousarray(th.reshape(1, -1)), device=self._inv.o.device)
        out = self._inv.D(tq)[0].cpu().numpy().reshape(th.shape)
        return float(out) if th.ndim == 0 else out

    def tabulated_model(self):
        """``self.D`` as a model the solve kernels evaluate (``frontx_b200.models.Tabulated``): the knots of the
        inverse interpolant in ascending theta with its slopes and ``Iodtheta - c``.  ``solve(D=sol.D, ...)`` calls
        this (the reference passes the bound method itself, tests/test_inverse.py:42-48)."""
        torch = _lib.require_cuda()
        from .models import Tabulated

        if 0 in self._inv._deduped:  # rows already ascending; the scan measured the integral from the largest theta
            sub, i_ref = self._inv._deduped[0]
            dev = sub.o.device
            at_i = torch.full((1, 1), float(i_ref), dtype=torch.float64, device=dev)
            ant_i = torch.empty((1, 1), dtype=torch.float64, device=dev)
            with torch.cuda.device(dev):
                _lib.check(_lib.lib().fx_pchip_query_batch(1, int(sub.o.shape[1]), sub.theta.data_ptr(), sub.o.data_ptr(),
                                                           sub._slope.data_ptr(), sub._anti.data_ptr(), 1, at_i.data_ptr(),
                                                           1, 0, None, None, ant_i.data_ptr(), _stream(torch, dev)))
            rows = [sub.theta[0], sub.o[0], sub._slope[0], sub._anti[0] - ant_i[0, 0]]
        else:  # strictly monotone data in the caller's order: the integral already starts at the last point, theta = i
            inv = self._inv
            rows = [inv.theta[0], inv.o[0], inv._slope[0], inv._anti[0]]
            if bool(rows[0][0] > rows[0][-1]):
                rows = [r.flip(0) for r in rows]
        return Tabulated(torch.stack(rows).contiguous())

    @property
    def i(self) -> float:
        return float(self(self.oi))

    def sorptivity(self, o: Any = 0) -> Any:  # interpolated.py:69-73
        A = self._fwd.query  # noqa: N806  antiderivative of the forward PCHIP, end cubics extrapolated
        oo = np.asarray(o, dtype=np.float64)
        out = (A(self.oi, 2) - A(oo, 2)) - self.i * (self.oi - oo)
        return float(out) if oo.ndim == 0 else out


def sorptivity(o: Any, theta: Any, /, *, b: float, i: float) -> float:
    """``frontx.sorptivity`` (src/frontx/_inverse/sorptivity.py:8-19): trapezoid of theta - i
    over o with the boundary point (0, b) prepended; the sum runs on the GPU."""
    res = inverse(np.asarray(o, dtype=np.float64), np.asarray(theta, dtype=np.float64), throw=False, unique=False)
    return float(res.sorptivity_trapezoid(b, i)[0].item())
