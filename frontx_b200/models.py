"""Moisture diffusivity models, tagged for the CUDA kernels.

Same constructors, defaults and error behaviour as the reference's
``frontx.models`` (src/frontx/models.py:35-315): ``LETd``, ``BrooksAndCorey``,
``VanGenuchten``, ``LETxs``.  Where the reference hands an arbitrary
JAX-traceable callable to diffrax, a C ABI cannot: each model here packs itself
into ``(model_id, params[12])`` (layouts in include/frontx_b200.h) and the
kernels evaluate D, D', D'' in closed form.  ``Constant`` / ``PowerLaw`` cover
the ``fronts.D`` helpers the north star names (no class in the reference;
lambdas in tests/test_solve.py:40,151), ``LogDiffusivity`` the Philip (1960)
exact case of tests/test_solve.py:38-42.

Calling a model, ``D(theta)``, evaluates it ON THE GPU (fx_diffusivity_batch);
there is no CPU implementation in this package.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Sequence

import numpy as np

from . import _lib


class _TaggedModel:
    """Base: a model the kernels know, as (model_id, 12 doubles)."""

    model_id: int = -1

    def _pack(self) -> Sequence[float]:
        raise NotImplementedError

    def params(self) -> np.ndarray:
        out = np.zeros(_lib.NPARAM, dtype=np.float64)
        p = np.asarray(self._pack(), dtype=np.float64)
        out[: p.size] = p
        return out

    # --- evaluation on the device -----------------------------------------
    def derivatives(self, theta: Any) -> np.ndarray:
        """(D, dD/dtheta, d2D/dtheta2) stacked on a trailing axis, as numpy."""
        return ModelBatch.from_models([self]).derivatives(np.asarray(theta, dtype=np.float64)[None, ...])[0]

    def __call__(self, theta: Any) -> Any:
        th = np.asarray(theta, dtype=np.float64)
        out = self.derivatives(th)[..., 0]
        return float(out) if th.ndim == 0 else out

    @classmethod
    def batch(cls, **kwargs: Any) -> "ModelBatch":
        """Vectorised constructor: every keyword may be an array of the batch size.

        ``VanGenuchten.batch(n=n_arr, Ks=Ks_arr, alpha=a_arr, l=0.5, theta_range=(0.005, 0.7))``
        """
        conv = {}
        for key, v in kwargs.items():
            if v is None:
                conv[key] = None
            elif key == "theta_range":
                conv[key] = (np.asarray(v[0], dtype=np.float64), np.asarray(v[1], dtype=np.float64))
            elif isinstance(v, str):
                conv[key] = v
            else:
                conv[key] = np.asarray(v, dtype=np.float64)
        cols = np.broadcast_arrays(*[np.asarray(c, dtype=np.float64) for c in cls(**conv)._pack()])
        if cols[0].ndim > 1:
            raise ValueError("batch parameters must be scalars or 1-D arrays")
        n = cols[0].size
        rows = np.zeros((n, _lib.NPARAM), dtype=np.float64)
        for c, col in enumerate(cols):
            rows[:, c] = col.reshape(n)
        return ModelBatch(np.full(n, cls.model_id, dtype=np.int32), rows)


def _resolve_Ks(Ks, k, g, rho, mu):
    # src/frontx/models.py:76-86
    if Ks is None:
        if k is None:
            return 1.0
        return rho * g * k / mu
    if k is not None:
        msg = "Cannot set both Ks and k"
        raise ValueError(msg)
    return Ks


@dataclass(frozen=True)
class Constant(_TaggedModel):
    """``D(theta) = D0`` (fronts.D.constant)."""

    D0: float
    model_id = _lib.CONSTANT

    def _pack(self):
        return [self.D0]


@dataclass(frozen=True)
class PowerLaw(_TaggedModel):
    """``D(theta) = a*theta**k + epsilon`` (fronts.D.power_law)."""

    k: float
    a: float = 1.0
    epsilon: float = 0.0
    model_id = _lib.POWER_LAW

    def _pack(self):
        return [self.a, self.k, self.epsilon]


@dataclass(frozen=True)
class LogDiffusivity(_TaggedModel):
    """``D(theta) = c0 + c1*ln(theta)``; Philip's exact case is c0=1/2, c1=-1/2."""

    c0: float = 0.5
    c1: float = -0.5
    model_id = _lib.LOG

    def _pack(self):
        return [self.c0, self.c1]


@dataclass(frozen=True)
class LETd(_TaggedModel):
    """LET diffusivity, ``Dwt*Se**L/(Se**L + E*(1-Se)**T)`` (src/frontx/models.py:35-66)."""

    L: float
    E: float
    T: float
    Dwt: float = 1
    theta_range: tuple = (0, 1)
    model_id = _lib.LETD

    def _pack(self):
        return [self.L, self.E, self.T, self.Dwt, self.theta_range[0], self.theta_range[1]]


@dataclass(frozen=True)
class BrooksAndCorey(_TaggedModel):
    """Brooks and Corey, ``D = K/C`` (src/frontx/models.py:126-174)."""

    n: float
    l: float = 1  # noqa: E741
    Ks: float | None = None
    k: float | None = None
    g: float = 9.81
    rho: float = 1e3
    mu: float = 1e-3
    alpha: float = 1
    theta_range: tuple = (0, 1)
    model_id = _lib.BROOKS_COREY

    def _pack(self):
        Ks = _resolve_Ks(self.Ks, self.k, self.g, self.rho, self.mu)
        return [self.n, self.l, Ks, self.alpha, self.theta_range[0], self.theta_range[1]]


@dataclass(frozen=True)
class VanGenuchten(_TaggedModel):
    """van Genuchten-Mualem, ``D = K/C`` (src/frontx/models.py:177-253).

    Either ``n`` or ``m`` must be given; the other follows from ``m = 1 - 1/n``.  The kernels use the
    Mualem constraint ``m = 1 - 1/n`` in closed form, so a pair that violates it is refused (the
    reference would evaluate the unconstrained expression).

    ``form`` selects how ``1 - Se**(1/m)`` is formed: ``"literal"`` (default) is the subtraction the
    reference writes (src/frontx/models.py:245,253), which cancels ~7 digits at ``b = theta_s - 1e-7``;
    ``"expm1"`` is the same value without the cancellation.
    """

    n: float | None = None
    m: float | None = None
    l: float = 0.5  # noqa: E741
    Ks: float | None = None
    k: float | None = None
    g: float = 9.81
    rho: float = 1e3
    mu: float = 1e-3
    alpha: float = 1
    theta_range: tuple = (0, 1)
    form: str = "literal"
    model_id = _lib.VAN_GENUCHTEN

    @property
    def _n(self):  # src/frontx/models.py:217-227
        if self.n is not None:
            return self.n
        if self.m is None:
            msg = "Either n or m must be set"
            raise ValueError(msg)
        return 1 / (1 - self.m)

    @property
    def _m(self):  # src/frontx/models.py:229-239
        if self.m is not None:
            return self.m
        if self.n is None:
            msg = "Either n or m must be set"
            raise ValueError(msg)
        return 1 - 1 / self.n

    def _pack(self):
        Ks = _resolve_Ks(self.Ks, self.k, self.g, self.rho, self.mu)
        if self.form not in ("literal", "expm1"):
            msg = "form must be 'literal' or 'expm1'"
            raise ValueError(msg)
        n, m = self._n, self._m
        if self.n is not None and self.m is not None and np.any(np.abs(np.asarray(m) - (1 - 1 / np.asarray(n))) > 1e-12):
            msg = "the kernels need m = 1 - 1/n; give n or m, not an inconsistent pair"
            raise ValueError(msg)
        return [n, m, self.l, Ks, self.alpha, self.theta_range[0], self.theta_range[1],
                0.0 if self.form == "literal" else 1.0]


@dataclass(frozen=True)
class LETxs(_TaggedModel):
    """LET relative conductivity and capillary head (src/frontx/models.py:256-315)."""

    Lw: float
    Ew: float
    Tw: float
    Ls: float
    Es: float
    Ts: float
    Ks: float | None = None
    k: float | None = None
    g: float = 9.81
    rho: float = 1e3
    mu: float = 1e-3
    alpha: float = 1
    theta_range: tuple = (0, 1)
    model_id = _lib.LETXS

    def _pack(self):
        Ks = _resolve_Ks(self.Ks, self.k, self.g, self.rho, self.mu)
        return [self.Lw, self.Ew, self.Tw, self.Ls, self.Es, self.Ts, Ks, self.alpha,
                self.theta_range[0], self.theta_range[1]]


class Tabulated(_TaggedModel):
    """The diffusivity a measured profile implies, ``InterpolatedSolution.D``
    (src/frontx/_inverse/interpolated.py:59-67), as a model the kernels know -- so that it can be fed back into
    ``solve`` like the reference's ``solve(D=sol.D, ...)`` (tests/test_inverse.py:42-48).

    ``table`` is a contiguous CUDA float64 tensor ``[4, n]``: theta ascending, o, do/dtheta at the knots and
    ``Iodtheta - c`` at the knots.  The parameter block carries its device address; this object (and any
    ``ModelBatch`` made from it) keeps the tensor alive.  ``InterpolatedSolution.tabulated_model()`` builds one.
    """

    model_id = _lib.TABULATED

    def __init__(self, table: Any) -> None:
        if table.dim() != 2 or table.shape[0] != 4 or table.shape[1] < 2 or not table.is_cuda or not table.is_contiguous():
            raise ValueError("table must be a contiguous CUDA float64 tensor of shape [4, n], n >= 2")
        self.table = table

    def _pack(self):
        address = np.array([self.table.data_ptr()], dtype=np.uint64).view(np.float64)[0]
        return [address, float(self.table.shape[1])]


class ModelBatch:
    """n tagged models as two arrays: ``model_id[n]`` (int32) and ``params[n, 12]`` (float64)."""

    def __init__(self, model_id: Any, params: Any, keepalive: Any = None) -> None:
        self.model_id = np.ascontiguousarray(model_id, dtype=np.int32)
        self.params = np.ascontiguousarray(params, dtype=np.float64).reshape(self.model_id.size, _lib.NPARAM)
        self._dev = None
        self._keepalive = keepalive  # device tables the parameter blocks point into (Tabulated)

    def __len__(self) -> int:
        return int(self.model_id.size)

    @staticmethod
    def from_models(models: Sequence[_TaggedModel]) -> "ModelBatch":
        for m in models:
            if not isinstance(m, _TaggedModel):
                raise TypeError(
                    "frontx_b200 solves tagged models only (frontx_b200.models / frontx_b200.D); "
                    f"got {type(m).__name__}. An arbitrary callable cannot cross the C ABI and there is no CPU fallback."
                )
        ids = np.array([m.model_id for m in models], dtype=np.int32)
        par = np.stack([m.params() for m in models]) if models else np.zeros((0, _lib.NPARAM))
        return ModelBatch(ids, par, [m.table for m in models if isinstance(m, Tabulated)] or None)

    @staticmethod
    def concat(batches: Sequence["ModelBatch"]) -> "ModelBatch":
        return ModelBatch(np.concatenate([b.model_id for b in batches]),
                          np.concatenate([b.params for b in batches]),
                          [k for b in batches for k in (b._keepalive or [])] or None)

    def __getitem__(self, idx) -> "ModelBatch":
        ids = np.atleast_1d(self.model_id[idx])
        return ModelBatch(ids, self.params[idx].reshape(ids.size, _lib.NPARAM), self._keepalive)

    def mask(self) -> int:
        """Bit m set = model id m occurs in the batch (``fx_solve_options.model_mask``: only those models are
        launched; ids the library does not know contribute nothing and are flagged by the kernels)."""
        ids = np.unique(self.model_id)
        return int(sum(1 << int(m) for m in ids if 0 <= m < len(_lib.MODEL_NAMES)))

    def saturation(self, margin: float = 1e-7) -> np.ndarray:
        """Largest boundary value worth bracketing per problem: ``theta_s - margin*(theta_s - theta_r)``
        for the models with a saturated value, ``inf`` for the unbounded closed forms."""
        col = {_lib.BROOKS_COREY: (4, 5), _lib.VAN_GENUCHTEN: (5, 6), _lib.LETXS: (8, 9), _lib.LETD: (4, 5)}
        out = np.full(len(self), np.inf)
        for mid, (r, s_) in col.items():
            sel = self.model_id == mid
            tr, ts = self.params[sel, r], self.params[sel, s_]
            out[sel] = ts - margin * (ts - tr)
        return out

    def to_device(self, device=None):
        torch = _lib.require_cuda()
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self._dev is None or self._dev[0].device != dev:
            self._dev = (torch.from_numpy(self.model_id).to(dev), torch.from_numpy(self.params).to(dev))
        return self._dev

    def derivatives(self, theta: Any):
        """D, D', D'' at theta[n, ...] for each model; numpy in, numpy out (computed on the GPU)."""
        torch = _lib.require_cuda()
        ids, par = self.to_device()
        th = torch.as_tensor(np.ascontiguousarray(theta, dtype=np.float64), device=ids.device)
        n = len(self)
        if th.shape[0] != n:
            raise ValueError("theta must have one leading row per model")
        q = int(th[0].numel()) if n else 0
        th2 = th.reshape(n, q).contiguous()
        out = torch.empty((n, q, 3), dtype=torch.float64, device=ids.device)
        stream = torch.cuda.current_stream(ids.device).cuda_stream
        with torch.cuda.device(ids.device):
            _lib.check(_lib.lib().fx_diffusivity_batch(n, ids.data_ptr(), par.data_ptr(), q, th2.data_ptr(),
                                                       out.data_ptr(), stream))
        return out.reshape(tuple(th.shape) + (3,)).cpu().numpy()
