"""Profile -> D(theta) inverse: ``InterpolatedSolution``, ``sorptivity`` and the
batched ``inverse`` entry point.

Reference: src/frontx/_inverse/interpolated.py:18-73 (PCHIP of theta(o); PCHIP of
o(theta) on sorted theta; ``D = -(do/dtheta * Int_i^theta o dtheta)/2``) and
src/frontx/_inverse/sorptivity.py:8-19.  The PCHIP slopes, the per-interval cubic
integrals and their running sum are one scan kernel (``fx_inverse_batch``);
queries at arbitrary theta are ``fx_inverse_query_batch``.  theta must be
strictly monotone in o (the reference's ``jnp.unique`` sort is then a reversal).
"""

from __future__ import annotations

from typing import Any

import numpy as np

from . import _lib
from ._boltzmann import AbstractSolution, boltzmannmethod


class BatchInverse:
    """Result of ``inverse(o, theta)`` for [batch, npts] profiles; tensors stay on the GPU."""

    def __init__(self, o, theta, D, sorp, status, slope, anti) -> None:
        self.o, self.theta = o, theta
        self.D_at_knots = D
        self._sorp = sorp
        self.status = status
        self._slope, self._anti = slope, anti

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
                                                         tq.data_ptr(), out.data_ptr(),
                                                         torch.cuda.current_stream(dev).cuda_stream))
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


def inverse(o: Any, theta: Any, *, device: Any = None, throw: bool = True, sorptivity: bool = True,
            query: bool = True) -> BatchInverse:
    """``InterpolatedSolution(o, theta).D`` for a batch of profiles, on the GPU.

    ``o`` and ``theta`` are [batch, npts] (or [npts]); numpy arrays or CUDA tensors.
    ``sorptivity=False`` skips the sorptivity integrals of the forward interpolant (a second PCHIP
    slope per point: a quarter of the kernel's time), ``query=False`` the two knot arrays that
    ``BatchInverse.D(theta_query)`` needs (16 B/point more written): D at the knots only.
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
                                               torch.cuda.current_stream(dev).cuda_stream))
    res = BatchInverse(o_t, th_t, D, sorp, status, slope, anti)
    if throw and int(status.sum().item()) != 0:
        raise ValueError("theta must be strictly monotone in o for every profile")
    return res


class InterpolatedSolution(AbstractSolution):
    """``frontx.InterpolatedSolution`` (src/frontx/_inverse/interpolated.py:11-73).

    ``sol(o)`` interpolates the data with PCHIP, ``sol.D(theta)`` is the diffusivity
    the profile implies, ``sol.sorptivity()`` its sorptivity.  ``b`` / ``i`` add the
    boundary point (0, b) and the far-field point (o[-1]+1, i) to the inverse
    interpolant as the reference does (:32-40).
    """

    def __init__(self, o: Any, theta: Any, /, *, b: float | None = None, i: float | None = None) -> None:
        o = np.asarray(o, dtype=np.float64)
        theta = np.asarray(theta, dtype=np.float64)
        if o.ndim != 1 or o.shape != theta.shape:
            raise ValueError("o and theta must be 1-D arrays of the same length")
        self._fwd = inverse(o, theta)  # forward interpolant data + sorptivity integrals
        o_inv, th_inv = o, theta
        if b is not None:
            o_inv = np.insert(o_inv, 0, 0.0)
            th_inv = np.insert(th_inv, 0, b)
        if i is not None:
            o_inv = np.append(o_inv, o_inv[-1] + 1)
            th_inv = np.append(th_inv, i)
        self.oi = float(o_inv[-1])
        self._inv = self._fwd if (b is None and i is None) else inverse(o_inv, th_inv)
        self._o, self._theta = o, theta

    @boltzmannmethod
    def __call__(self, o: Any) -> Any:  # interpolated.py:52-57: the forward PCHIP theta(o)
        torch = _lib.require_cuda()
        oo = np.asarray(o, dtype=np.float64)
        f = self._fwd
        # theta(o) = value of the forward PCHIP: evaluate it as the inverse of (theta, o) swapped
        fwd = _forward_interpolant(f)
        out = fwd(torch.as_tensor(oo.reshape(1, -1), device=f.o.device))[0].cpu().numpy().reshape(oo.shape)
        return float(out) if oo.ndim == 0 else out

    @boltzmannmethod
    def d_do(self, o: Any) -> Any:
        raise NotImplementedError("InterpolatedSolution.d_do is not part of the accelerated path")

    def D(self, theta: Any, /) -> Any:  # interpolated.py:59-67
        torch = _lib.require_cuda()
        th = np.asarray(theta, dtype=np.float64)
        out = self._inv.D(torch.as_tensor(th.reshape(1, -1), device=self._inv.o.device))[0]
        out = out.cpu().numpy().reshape(th.shape)
        return float(out) if th.ndim == 0 else out

    @property
    def i(self) -> float:
        return float(self._theta[-1]) if self.oi == self._o[-1] else float(self(self.oi))

    def sorptivity(self, o: Any = 0) -> Any:  # interpolated.py:69-73
        if np.ndim(o) != 0 or float(o) != 0.0:
            raise NotImplementedError("InterpolatedSolution.sorptivity is accelerated for o=0 only")
        f = self._fwd
        raw = float((f._sorp[0, 0] + f._sorp[0, 2]).item())  # Int_0^{o_last} theta do (PCHIP)
        o_last = float(self._o[-1])
        th_oi = self(self.oi)
        # (I(oi) - I(0)) - i*(oi - 0); beyond the data the forward PCHIP extrapolates (scipy default)
        extra = 0.0
        if self.oi != o_last:
            raise NotImplementedError("sorptivity with an appended far-field point is not accelerated")
        return (raw + extra) - th_oi * (self.oi - 0.0)


def _forward_interpolant(f: BatchInverse):
    """theta(o) of the forward PCHIP, through the query kernel on the swapped data.

    The query kernel evaluates a PCHIP's derivative and antiderivative; the VALUE of
    the forward interpolant is recovered here from its antiderivative's derivative,
    i.e. by a dedicated value kernel -- see fx_pchip_eval_batch.
    """
    torch = _lib.require_cuda()

    def run(oq):
        oq = oq.contiguous()
        batch, npts = f.o.shape
        q = int(oq.shape[1])
        out = torch.empty((batch, q), dtype=torch.float64, device=oq.device)
        with torch.cuda.device(oq.device):
            _lib.check(_lib.lib().fx_pchip_eval_batch(batch, npts, f.o.data_ptr(), f.theta.data_ptr(), q,
                                                      oq.data_ptr(), out.data_ptr(),
                                                      torch.cuda.current_stream(oq.device).cuda_stream))
        return out

    return run


def sorptivity(o: Any, theta: Any, /, *, b: float, i: float) -> float:
    """``frontx.sorptivity`` (src/frontx/_inverse/sorptivity.py:8-19): trapezoid of theta - i
    over o with the boundary point (0, b) prepended; the sum runs on the GPU."""
    res = inverse(np.asarray(o, dtype=np.float64), np.asarray(theta, dtype=np.float64), throw=False)
    return float(res.sorptivity_trapezoid(b, i)[0].item())
