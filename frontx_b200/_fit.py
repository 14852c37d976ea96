"""Parametric fitting on top of the batched solve: ``Param``, ``ScaledSolution``, ``fit``.

Reference: src/frontx/_inverse/param.py:11-112 and src/frontx/_inverse/fit.py:17-153.  The
reference's ``fit`` hands scipy's ``differential_evolution`` a jitted objective that solves ONE
candidate per call (param.py:105-110), so a generation of 15 x n_params candidates is that many
sequential solves.  Here a whole generation is one ``fx_solve_batch`` launch
(``vectorized=True, updating="deferred"``), the one-parameter D0 fit of every candidate
(``ScaledSolution.fitting_data``, fit.py:45-82) and its cost (fit.py:145-151) one
``fx_scaled_fit_batch`` launch on the candidates' dense-output tables.  The evolution itself stays
in scipy on the host: it is control logic over a few hundred numbers.

A model is made fittable the way the reference does it: put ``Param(value, min=, max=)`` where a
number would go, e.g. ``LETd(L=Param(min=0, max=2), E=Param(1e3, min=1, max=1e5), T=1.5, ...)``.
"""

from __future__ import annotations

from dataclasses import fields, is_dataclass, replace
from typing import Any, Sequence

import numpy as np

from . import _lib
from ._boltzmann import AbstractSolution, boltzmannmethod
from ._forward import RESULTS, Solution, solve, solve_batch
from .models import ModelBatch, _TaggedModel


class Param:
    """A fittable number with optional bounds (src/frontx/_inverse/param.py:11-48).

    ``value`` defaults as in the reference: the middle of ``[min, max]``, ``min + 1``, ``max - 1`` or 0.
    (The reference stores a logit/log ``raw_value`` for its gradient-based fitters; the evolution
    used by ``fit`` only ever reads ``value``, ``min`` and ``max``.)
    """

    def __init__(self, value: float | None = None, min: float | None = None, max: float | None = None) -> None:  # noqa: A002
        if value is None:
            if min is not None and max is not None:
                value = (min + max) / 2
            elif min is not None:
                value = min + 1.0
            elif max is not None:
                value = max - 1.0
            else:
                value = 0.0
        self.value = float(value)
        self.min = min
        self.max = max

    def __repr__(self) -> str:
        return f"Param({self.value!r}, min={self.min!r}, max={self.max!r})"


def _slots(model: Any) -> list[tuple[str, int | None]]:
    """Where the Params of a model sit: (field name, index inside a tuple field or None)."""
    if not (is_dataclass(model) and isinstance(model, _TaggedModel)):
        raise TypeError("fit() needs a frontx_b200 model whose fittable fields are Param objects")
    out: list[tuple[str, int | None]] = []
    for f in fields(model):
        v = getattr(model, f.name)
        if isinstance(v, Param):
            out.append((f.name, None))
        elif isinstance(v, tuple):
            out.extend((f.name, k) for k, e in enumerate(v) if isinstance(e, Param))
    return out


def get_params(model: Any) -> list[Param]:
    """The ``Param`` objects of a model, in field order (param.py:51-63)."""
    out = []
    for name, k in _slots(model):
        v = getattr(model, name)
        out.append(v if k is None else v[k])
    return out


def set_param_values(model: Any, values: Sequence[float]) -> Any:
    """The model with its Params replaced by plain numbers (param.py:70-90)."""
    changes: dict[str, Any] = {}
    for (name, k), x in zip(_slots(model), values):
        if k is None:
            changes[name] = float(x)
        else:
            t = list(changes.get(name, getattr(model, name)))
            t[k] = float(x)
            changes[name] = tuple(t)
    return replace(model, **changes)


def _candidates(model: Any, X: np.ndarray) -> ModelBatch:
    """One ModelBatch row per column of X[n_params, S], without a Python loop over candidates."""
    S = X.shape[1]
    kwargs: dict[str, Any] = {}
    slots = _slots(model)
    for f in fields(model):
        v = getattr(model, f.name)
        if isinstance(v, tuple):
            kwargs[f.name] = tuple(np.full(S, float(e.value if isinstance(e, Param) else e)) for e in v)
        elif isinstance(v, Param):
            kwargs[f.name] = np.full(S, v.value)
        else:
            kwargs[f.name] = v
    for (name, k), row in zip(slots, X):
        if k is None:
            kwargs[name] = np.asarray(row, dtype=np.float64)
        else:
            t = list(kwargs[name])
            t[k] = np.asarray(row, dtype=np.float64)
            kwargs[name] = tuple(t)
    return type(model).batch(**kwargs)


class ScaledSolution(AbstractSolution):
    """A solution rescaled by a constant diffusivity factor D0 (src/frontx/_inverse/fit.py:17-107):
    ``theta(o) = original(o/sqrt(D0))``, ``D = D0 * original.D``, ``oi = original.oi * sqrt(D0)``."""

    def __init__(self, original: AbstractSolution, /, D0: float, *, _result: RESULTS = RESULTS.successful) -> None:  # noqa: N803
        self.original = original
        self.D0 = float(D0)
        self.result = _result

    @staticmethod
    def with_sorptivity(original: AbstractSolution, S: float, /) -> "ScaledSolution":  # noqa: N803
        """fit.py:35-43: the D0 that gives the rescaled solution sorptivity S."""
        return ScaledSolution(original, D0=(float(S) / float(original.sorptivity())) ** 2)

    @staticmethod
    def fitting_data(original: AbstractSolution, o: Any, theta: Any, /, sigma: Any = 1, *,
                     throw: bool = True) -> "ScaledSolution":
        """fit.py:45-82: least-squares D0 for data (o, theta) with standard deviations sigma."""
        base, extra = _unscale(original)
        if not isinstance(base, Solution) or base._batch.knots is None:
            raise TypeError("fitting_data needs a frontx_b200 Solution with dense output")
        o_a = np.ascontiguousarray(o, dtype=np.float64).reshape(-1)
        # original(o) = base(o / sqrt(extra)): fit the base solution, then divide the factor out
        D0, _, status = _scaled_fit(base._batch, o_a, theta, sigma, mode=1)
        d0 = float(D0[0].item()) / extra
        st = int(status[0].item())
        result = RESULTS.successful if st == 0 else (RESULTS.max_steps_reached if st == 1 else RESULTS.internal_error)
        if throw and result != RESULTS.successful:
            raise RuntimeError(f"the D0 fit did not converge ({result.name})")
        return ScaledSolution(original, d0, _result=result)

    @boltzmannmethod
    def __call__(self, o: Any) -> Any:  # fit.py:84-89
        return self.original(np.asarray(o, dtype=np.float64) / np.sqrt(self.D0))

    @boltzmannmethod
    def d_do(self, o: Any) -> Any:  # fit.py:91-96
        return self.original.d_do(np.asarray(o, dtype=np.float64) / np.sqrt(self.D0)) / np.sqrt(self.D0)

    def D(self, theta: Any, /) -> Any:  # noqa: N802  fit.py:98-103
        return np.asarray(self.original.D(theta)) * self.D0

    @property
    def oi(self) -> float:  # fit.py:105-107
        return float(self.original.oi) * float(np.sqrt(self.D0))


def _unscale(sol: AbstractSolution) -> tuple[AbstractSolution, float]:
    """Peel nested ScaledSolutions: sol(o) = base(o / sqrt(factor))."""
    factor = 1.0
    while isinstance(sol, ScaledSolution):
        factor *= sol.D0
        sol = sol.original
    return sol, factor


def _scaled_fit(batch, o: np.ndarray, theta: Any, sigma: Any, *, mode: int, D0=None):
    """fx_scaled_fit_batch on a BatchSolution: returns (D0, cost, status) device tensors."""
    torch = _lib.require_cuda()
    dev = batch.summary.device
    n, m = len(batch), int(o.size)
    th = np.ascontiguousarray(np.broadcast_to(np.asarray(theta, dtype=np.float64).reshape(-1), (m,)))
    w = np.ascontiguousarray(np.broadcast_to(1.0 / np.asarray(sigma, dtype=np.float64).reshape(-1), (m,)))
    o_t, th_t, w_t = (torch.as_tensor(a.copy(), device=dev) for a in (o, th, w))
    d0_t = torch.ones((n,), dtype=torch.float64, device=dev) if D0 is None else D0.to(dev).contiguous()
    cost = torch.empty((n,), dtype=torch.float64, device=dev)
    status = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().fx_scaled_fit_batch(n, batch.counters.data_ptr(), batch.summary.data_ptr(),
                                                  batch.knots.data_ptr(), batch.k_max, m, o_t.data_ptr(),
                                                  th_t.data_ptr(), w_t.data_ptr(), mode, d0_t.data_ptr(),
                                                  cost.data_ptr(), status.data_ptr(),
                                                  torch.cuda.current_stream(dev).cuda_stream))
    return d0_t, cost, status


def generation_cost(D: Any, X: Any, o: Any, theta: Any, sigma: Any = 1, *, i: float, b: float,  # noqa: N803, PLR0913
                    fit_D0: str | None = "data", k_max: int = 512, S_data: float | None = None) -> np.ndarray:  # noqa: N803
    """``cost(candidate(x))`` of the reference's ``fit`` (fit.py:125-151) for a whole population at once:
    ``X[n_params, S]`` holds S candidate parameter vectors of the model family ``D``; returns their S costs.
    One ``fx_solve_batch`` launch (dense output) + one ``fx_scaled_fit_batch`` launch."""
    torch = _lib.require_cuda()
    o_a = np.ascontiguousarray(o, dtype=np.float64).reshape(-1)
    X = np.asarray(X, dtype=np.float64).reshape(len(get_params(D)), -1)
    cand = _candidates(D, X)
    sols = solve_batch(cand, b=b, i=i, throw=False, k_max=k_max)
    if fit_D0 == "data":
        _, cost, _ = _scaled_fit(sols, o_a, theta, sigma, mode=1)
    elif fit_D0 == "sorptivity":
        d0 = (S_data / sols.sorptivity()) ** 2
        _, cost, _ = _scaled_fit(sols, o_a, theta, sigma, mode=2, D0=d0)
    else:
        _, cost, _ = _scaled_fit(sols, o_a, theta, sigma, mode=0)
    trunc = sols.counters[:, _lib.C_KNOTS] > sols.k_max
    cost = torch.where(trunc, torch.full_like(cost, float("inf")), cost)
    return cost.cpu().numpy()


def fit(D: Any, o: Any, theta: Any, /, sigma: Any = 1, *, i: float, b: float,  # noqa: N803, PLR0913
        fit_D0: str | None = "data", max_steps: int = 15, k_max: int = 512, seed: Any = None,  # noqa: N803
        popsize: int = 15, polish: bool = True):
    """``frontx.fit`` (src/frontx/_inverse/fit.py:110-153): the model of the family ``D`` (a tagged
    model with ``Param`` fields) whose solution best matches the profile data ``(o, theta)``.

    ``fit_D0``: "data" fits an overall diffusivity factor to the data for every candidate,
    "sorptivity" fixes it from the data's sorptivity, ``None`` compares the bare solutions.
    Returns the best candidate's ``ScaledSolution`` (or ``Solution`` for ``fit_D0=None``); the fitted
    model is ``result.fitted_model``.  Each generation of the differential evolution is one batched
    solve on the GPU.
    """
    from scipy.optimize import differential_evolution

    from ._inverse import sorptivity as data_sorptivity

    if fit_D0 not in ("data", "sorptivity", None):
        raise ValueError("fit_D0 must be 'data', 'sorptivity' or None")
    params = get_params(D)
    if not params:
        raise ValueError("the model has no Param to fit")
    if any(p.min is None or p.max is None for p in params):
        raise ValueError("differential evolution needs min and max on every Param")
    o_a = np.ascontiguousarray(o, dtype=np.float64).reshape(-1)
    S_data = float(data_sorptivity(o_a, theta, b=b, i=i)) if fit_D0 == "sorptivity" else None

    def generation(X: np.ndarray) -> np.ndarray:
        return generation_cost(D, X, o_a, theta, sigma, i=i, b=b, fit_D0=fit_D0, k_max=k_max, S_data=S_data)

    bounds = [(p.min, p.max) for p in params]
    x0 = np.array([p.value for p in params])
    opt = differential_evolution(generation, bounds=bounds, x0=x0, maxiter=max_steps, vectorized=True,
                                 updating="deferred", popsize=popsize, polish=polish, seed=seed)
    best = set_param_values(D, opt.x)
    sol = solve(best, b=b, i=i, throw=False)
    if fit_D0 == "data":
        out: Any = ScaledSolution.fitting_data(sol, o_a, theta, sigma=sigma, throw=False)
    elif fit_D0 == "sorptivity":
        out = ScaledSolution.with_sorptivity(sol, S_data)
    else:
        out = sol
    out.fitted_model = best
    out.cost = float(opt.fun)
    out.generations = int(opt.nit)
    return out


__all__ = ("Param", "ScaledSolution", "fit", "generation_cost", "get_params", "set_param_values")
