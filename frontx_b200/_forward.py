"""Forward shooting solve: ``solve`` / ``Solution`` with the reference's surface,
plus the batched entry point ``solve_batch`` / ``BatchSolution``.

Reference: src/frontx/_forward.py:17-122.  The reference integrates one problem
per call through diffrax + optimistix under jit; here every call -- single or
batched -- goes through ``fx_solve_batch`` (include/frontx_b200.h), one GPU
thread per problem.  torch is used for device buffers and streams only.
"""

from __future__ import annotations

import enum
from typing import Any, Sequence

import numpy as np

from . import _lib
from ._boltzmann import AbstractSolution, boltzmannmethod
from .models import ModelBatch, _TaggedModel


class RESULTS(enum.IntEnum):
    """Status of a solve; the members the reference uses of ``diffrax.RESULTS``
    (src/frontx/_forward.py:14,98,116-120; _inverse/fit.py:72-77)."""

    successful = 0
    max_steps_reached = 1
    internal_error = 2
    event_occurred = 3


class SolveError(RuntimeError):
    """Raised by ``throw=True`` when the bisection did not converge; the analogue of the
    ``eqx.EquinoxRuntimeError`` the reference raises (tests/test_solve.py:150)."""


def _as_batch(D: Any) -> ModelBatch:
    if isinstance(D, ModelBatch):
        return D
    owner = getattr(D, "__self__", None)  # solve(D=sol.D) with sol an InterpolatedSolution (tests/test_inverse.py:42-48)
    if owner is not None and hasattr(owner, "tabulated_model") and getattr(D, "__name__", "") == "D":
        D = owner.tabulated_model()
    if isinstance(D, _TaggedModel):
        return ModelBatch.from_models([D])
    if isinstance(D, (list, tuple)):
        return ModelBatch.from_models(list(D))
    raise TypeError(
        "D must be a frontx_b200 model (frontx_b200.models.* / frontx_b200.D.*), a sequence of them or a "
        f"ModelBatch; got {type(D).__name__}. Arbitrary callables cannot cross the C ABI and there is no "
        "CPU fallback."
    )


class BatchSolution:
    """Solutions of n independent problems, resident on one GPU.

    ``summary[n, 8]``, ``counters[n, 8]`` and (if dense output was requested)
    ``knots[n, k_max, 4]`` are torch CUDA tensors laid out as the C ABI documents.
    """

    def __init__(self, models: ModelBatch, summary, counters, knots, k_max: int, itol: float) -> None:
        self.models = models
        self.summary = summary
        self.counters = counters
        self.knots = knots
        self.k_max = int(k_max)
        self.itol = float(itol)
        self.b = None  # boundary values found by solve_flowrate_batch (None for solve_batch: b is an input)

    def __len__(self) -> int:
        return int(self.summary.shape[0])

    # --- per-problem scalars (torch tensors on the device) -------------------
    @property
    def d_dob(self):
        return self.summary[:, _lib.S_D_DOB]

    @property
    def oi(self):
        return self.summary[:, _lib.S_OI]

    @property
    def i(self):
        return self.summary[:, _lib.S_THETA_OI]

    @property
    def D_b(self):
        return self.summary[:, _lib.S_D_B]

    @property
    def residual(self):
        return self.summary[:, _lib.S_RESIDUAL]

    @property
    def n_shots(self):
        """Shooting iterations per problem (initial bracket evaluations included)."""
        return self.counters[:, _lib.C_SHOTS]

    @property
    def result(self):
        """0 where successful, 1 (max_steps_reached) elsewhere -- _forward.py:116-120."""
        return (self.counters[:, _lib.C_RESULT] != 0).to(self.counters.dtype)

    def sorptivity(self, o: Any = None):
        """``-2 D(theta(o)) dtheta/do(o)`` (_boltzmann.py:158-161); at o=0 it is in the summary."""
        if o is None:
            return self.summary[:, _lib.S_S0]
        torch = _lib.require_cuda()
        th, dth = self.evaluate(o)
        ids, par = self.models.to_device(self.summary.device)
        n, q = th.shape
        out = torch.empty((n, q, 3), dtype=torch.float64, device=th.device)
        with torch.cuda.device(th.device):
            _lib.check(_lib.lib().fx_diffusivity_batch(n, ids.data_ptr(), par.data_ptr(), q, th.data_ptr(),
                                                       out.data_ptr(),
                                                       torch.cuda.current_stream(th.device).cuda_stream))
        return -2.0 * out[:, :, 0] * dth

    # --- dense output ------------------------------------------------------
    def evaluate(self, o: Any):
        """theta(o) and dtheta/do(o): ``o`` is [q] (shared) or [n, q]; returns two [n, q] tensors."""
        torch = _lib.require_cuda()
        if self.knots is None:
            raise _lib.FrontxB200Error("this batch was solved without dense output (k_max=0)")
        dev = self.summary.device
        oq = torch.as_tensor(o, dtype=torch.float64, device=dev)
        n = len(self)
        if oq.ndim == 0:
            oq = oq.reshape(1)
        per_problem = 1 if oq.ndim == 2 else 0
        if per_problem and oq.shape[0] != n:
            raise ValueError("per-problem queries need shape [n, q]")
        oq = oq.contiguous()
        q = int(oq.shape[-1])
        th = torch.empty((n, q), dtype=torch.float64, device=dev)
        dth = torch.empty((n, q), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().fx_eval_batch(n, self.counters.data_ptr(), self.knots.data_ptr(), self.k_max, q,
                                                oq.data_ptr(), per_problem, th.data_ptr(), dth.data_ptr(),
                                                torch.cuda.current_stream(dev).cuda_stream))
        return th, dth

    def theta(self, o: Any):
        return self.evaluate(o)[0]

    __call__ = theta

    def d_do(self, o: Any):
        return self.evaluate(o)[1]

    def flux(self, r: Any, t: Any):
        """``-D(theta) dtheta/dr`` at (r, t) (_boltzmann.py:151-156)."""
        torch = _lib.require_cuda()
        dev = self.summary.device
        r_ = torch.as_tensor(r, dtype=torch.float64, device=dev)
        t_ = torch.as_tensor(t, dtype=torch.float64, device=dev)
        o = r_ / torch.sqrt(t_)
        return self.sorptivity(o) / (2.0 * torch.sqrt(t_))

    # --- status ------------------------------------------------------------
    def raise_for_status(self) -> None:
        bad = int((self.counters[:, _lib.C_RESULT] != 0).sum().item())
        if bad:
            raise SolveError(f"{bad} of {len(self)} problems did not converge (max_steps_reached)")
        trunc = int((self.counters[:, _lib.C_KNOTS] > self.k_max).sum().item()) if self.knots is not None else 0
        if trunc:
            raise SolveError(f"{trunc} solutions have more accepted steps than k_max={self.k_max}; "
                             "re-solve with a larger k_max")

    def __getitem__(self, j: int) -> "Solution":
        j = int(j)
        if j < 0:
            j += len(self)
        sub = BatchSolution(self.models[j], self.summary[j:j + 1], self.counters[j:j + 1],
                            None if self.knots is None else self.knots[j:j + 1], self.k_max, self.itol)
        sub.b = None if self.b is None else self.b[j:j + 1]
        return Solution(sub)

    def cpu(self) -> dict:
        """Summary and counters as numpy arrays (named columns)."""
        s = self.summary.cpu().numpy()
        c = self.counters.cpu().numpy()
        out = {"d_dob": s[:, 0], "oi": s[:, 1], "i": s[:, 2], "sorptivity": s[:, 3], "D_b": s[:, 4],
               "residual": s[:, 5], "lower": s[:, 6], "upper": s[:, 7], "result": c[:, 0], "n_shots": c[:, 1],
               "n_expansions": c[:, 2], "n_attempts": c[:, 3], "n_rejected": c[:, 4], "n_rhs": c[:, 5],
               "n_jac": c[:, 6], "n_knots": c[:, 7]}
        return out


class Solution(AbstractSolution):
    """One solved problem, with the reference's ``Solution`` surface
    (src/frontx/_forward.py:17-57 + src/frontx/_boltzmann.py:101-161).

    Evaluation (``sol(o)``, ``sol(r, t)``, ``d_do``, ``flux``, ``sorptivity``) runs the
    dense-output kernel on the GPU and returns numpy arrays / floats.
    """

    def __init__(self, batch: BatchSolution) -> None:
        self._batch = batch
        s = batch.summary[0].cpu().numpy()
        c = batch.counters[0].cpu().numpy()
        self._summary = s
        self._counters = c
        self.result = RESULTS.successful if c[_lib.C_RESULT] == 0 else RESULTS.max_steps_reached
        model = batch.models
        self.D = _SingleModel(model)

    def _eval(self, o: Any, which: int) -> Any:
        oo = np.asarray(o, dtype=np.float64)
        flat = np.ascontiguousarray(oo.reshape(-1))
        th, dth = self._batch.evaluate(flat)
        out = (th if which == 0 else dth)[0].cpu().numpy().reshape(oo.shape)
        return float(out) if oo.ndim == 0 else out

    @boltzmannmethod
    def __call__(self, o: Any) -> Any:  # _forward.py:25-30
        return self._eval(o, 0)

    @boltzmannmethod
    def d_do(self, o: Any) -> Any:  # _forward.py:32-37
        return self._eval(o, 1)

    @property
    def oi(self) -> float:  # _forward.py:39-42
        return float(self._summary[_lib.S_OI])

    @property
    def i(self) -> float:  # _forward.py:44-47: the ACHIEVED far-field value
        return float(self._summary[_lib.S_THETA_OI])

    @property
    def b(self) -> float:  # _forward.py:49-52
        if getattr(self._batch, "b", None) is not None:  # flow-rate boundary: b is a result
            return float(self._batch.b[0].item())
        return float(self._batch.knots[0, 0, 1].item()) if self._batch.knots is not None else float("nan")

    @property
    def d_dob(self) -> float:  # _forward.py:54-57
        return float(self._summary[_lib.S_D_DOB])

    @property
    def n_shots(self) -> int:
        """Shooting iterations used (the two bracket evaluations included)."""
        return int(self._counters[_lib.C_SHOTS])

    @property
    def counters(self) -> dict:
        names = ("result", "n_shots", "n_expansions", "n_attempts", "n_rejected", "n_rhs", "n_jac", "n_knots")
        return {k: int(v) for k, v in zip(names, self._counters)}


class _SingleModel:
    """``Solution.D``: the model of one solved problem as a callable (evaluated on the GPU)."""

    def __init__(self, models: ModelBatch) -> None:
        self._models = models

    def __call__(self, theta: Any) -> Any:
        th = np.asarray(theta, dtype=np.float64)
        out = self._models.derivatives(th.reshape(1, -1))[0, :, 0].reshape(th.shape)
        return float(out) if th.ndim == 0 else out


def solve_batch(  # noqa: PLR0913
    D: Any,
    *,
    b: Any,
    i: Any,
    itol: float = 1e-3,
    max_steps: int = 100,
    throw: bool = True,
    k_max: int = 256,
    device: Any = None,
    options: dict | None = None,
) -> BatchSolution:
    """Solve n independent problems on one GPU.

    ``D`` is a ``ModelBatch`` (or a sequence of models), ``b`` and ``i`` scalars or
    arrays of length n.  Keyword meaning and defaults are those of
    ``frontx.solve`` (src/frontx/_forward.py:60-72); ``k_max`` is the dense-output
    capacity per problem (0 = summaries only).
    """
    models = _as_batch(D)
    torch = _lib.require_cuda()
    n = len(models)
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ids, par = models.to_device(dev)
    b_t = torch.as_tensor(np.broadcast_to(np.asarray(b, dtype=np.float64), (n,)).copy(), device=dev)
    i_t = torch.as_tensor(np.broadcast_to(np.asarray(i, dtype=np.float64), (n,)).copy(), device=dev)
    summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
    counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
    knots = torch.zeros((n, k_max, _lib.KNOT), dtype=torch.float64, device=dev) if k_max > 0 else None
    opt = _lib.default_options(itol=float(itol), max_bisect=int(max_steps),
                               **{"model_mask": models.mask(), **(options or {})})
    with torch.cuda.device(dev):
        rc = _lib.lib().fx_solve_batch(n, ids.data_ptr(), par.data_ptr(), b_t.data_ptr(), i_t.data_ptr(),
                                       opt, summary.data_ptr(), counters.data_ptr(), k_max,
                                       knots.data_ptr() if knots is not None else None,
                                       torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc)
    sol = BatchSolution(models, summary, counters, knots, k_max, itol)
    if throw:
        sol.raise_for_status()
    return sol


def solve(  # noqa: PLR0913
    D: Any,
    *,
    b: float,
    i: float,
    itol: float = 1e-3,
    max_steps: int = 100,
    throw: bool = True,
) -> Solution:
    """``frontx.solve`` (src/frontx/_forward.py:60-122) for one problem, on the GPU."""
    k_max = 512
    while True:
        batch = solve_batch(D, b=b, i=i, itol=itol, max_steps=max_steps, throw=False, k_max=k_max)
        if len(batch) != 1:
            raise ValueError("solve() takes one model; use solve_batch() for several")
        nk = int(batch.counters[0, _lib.C_KNOTS].item())
        if nk <= k_max:
            break
        k_max = 1 << (nk - 1).bit_length()
    sol = Solution(batch)
    if throw and sol.result != RESULTS.successful:
        raise SolveError("the shooting bisection did not converge (max_steps_reached)")
    return sol


_RADIAL_K = {"cylindrical": 1, "polar": 1, "spherical": 2}


def _flux_density(Qb: Any, radial: str, ob: float, angle: float | None, height: float | None) -> tuple[np.ndarray, int]:
    """Inlet flux density q with d_dob = -q/D(b) (fronts.solve_flowrate)."""
    if radial not in _RADIAL_K:
        raise ValueError(f"radial must be one of {sorted(_RADIAL_K)}, got {radial!r}")
    Qb = np.asarray(Qb, dtype=np.float64)
    if radial == "cylindrical":
        if height is None:
            raise TypeError("cylindrical flow needs a height")
        area = (2 * np.pi if angle is None else angle) * ob * height
    elif radial == "polar":
        if height is not None:
            raise TypeError("polar flow takes no height")
        area = (2 * np.pi if angle is None else angle) * ob
    else:
        if height is not None:
            raise TypeError("spherical flow takes no height")
        area = (4 * np.pi if angle is None else angle) * ob * ob
    return Qb / area, _RADIAL_K[radial]


def solve_flowrate_batch(  # noqa: PLR0913
    D: Any,
    *,
    Qb: Any,
    i: Any,
    radial: str = "cylindrical",
    ob: float = 1e-6,
    angle: float | None = None,
    height: float | None = None,
    itol: float = 1e-3,
    max_steps: int = 100,
    throw: bool = True,
    k_max: int = 256,
    b_bracket: Any = None,
    device: Any = None,
    options: dict | None = None,
) -> BatchSolution:
    """Radial problems with a FLOW-RATE boundary, n at a time (north-star config C4).

    Not in the reference; semantics are those of ``fronts.solve_flowrate``:
    the ODE gains the radial term ``k/o`` (k = 1 cylindrical/polar, 2 spherical), the boundary sits
    at ``ob > 0``, the inlet value ``b`` is unknown and the inlet gradient follows from the imposed
    flow rate, ``d_dob = -Qb / (D(b) * angle * ob * height)``.  The shooting bisection runs on ``b``
    inside ``b_bracket`` (default: from ``i`` to just below the model's saturated value; models with
    no upper bound need an explicit bracket) until the far field meets ``i`` within ``itol``.
    The returned ``BatchSolution`` has the usual layout; ``.b`` holds the boundary values found.
    """
    models = _as_batch(D)
    torch = _lib.require_cuda()
    n = len(models)
    q, radial_k = _flux_density(Qb, radial, ob, angle, height)
    i_np = np.broadcast_to(np.asarray(i, dtype=np.float64), (n,)).copy()
    if b_bracket is None:
        hi = models.saturation()
        if not np.all(np.isfinite(hi)):
            raise ValueError("models without a saturated value need an explicit b_bracket=(lo, hi)")
        lo_np, hi_np = i_np.copy(), hi
    else:
        lo_np = np.broadcast_to(np.asarray(b_bracket[0], dtype=np.float64), (n,)).copy()
        hi_np = np.broadcast_to(np.asarray(b_bracket[1], dtype=np.float64), (n,)).copy()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    ids, par = models.to_device(dev)
    q_t = torch.as_tensor(np.broadcast_to(q, (n,)).copy(), device=dev)
    i_t = torch.as_tensor(i_np, device=dev)
    lo_t = torch.as_tensor(lo_np, device=dev)
    hi_t = torch.as_tensor(hi_np, device=dev)
    summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
    counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
    knots = torch.zeros((n, k_max, _lib.KNOT), dtype=torch.float64, device=dev) if k_max > 0 else None
    opt = _lib.default_options(itol=float(itol), max_bisect=int(max_steps), radial_k=radial_k, ob=float(ob),
                               **{"model_mask": models.mask(), **(options or {})})
    with torch.cuda.device(dev):
        rc = _lib.lib().fx_solve_flowrate_batch(n, ids.data_ptr(), par.data_ptr(), q_t.data_ptr(), i_t.data_ptr(),
                                                lo_t.data_ptr(), hi_t.data_ptr(), opt, summary.data_ptr(),
                                                counters.data_ptr(), k_max,
                                                knots.data_ptr() if knots is not None else None,
                                                torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc)
    # C ABI layout {b, oi, theta(oi), d_dob, D(b), residual, b_lower, b_upper} -> the usual one
    b_found = summary[:, 0].clone()
    d_dob = summary[:, 3].clone()
    summary[:, _lib.S_D_DOB] = d_dob
    summary[:, _lib.S_S0] = -2.0 * summary[:, _lib.S_D_B] * d_dob
    sol = BatchSolution(models, summary, counters, knots, k_max, itol)
    sol.b = b_found
    sol.ob = float(ob)
    if throw:
        sol.raise_for_status()
    return sol


def solve_flowrate(D: Any, *, Qb: float, i: float, radial: str = "cylindrical", ob: float = 1e-6,  # noqa: PLR0913
                   angle: float | None = None, height: float | None = None, itol: float = 1e-3,
                   max_steps: int = 100, throw: bool = True, b_bracket: Any = None) -> Solution:
    """One radial problem with a flow-rate boundary (``fronts.solve_flowrate`` semantics), on the GPU."""
    k_max = 512
    while True:
        batch = solve_flowrate_batch(D, Qb=Qb, i=i, radial=radial, ob=ob, angle=angle, height=height, itol=itol,
                                     max_steps=max_steps, throw=False, k_max=k_max, b_bracket=b_bracket)
        if len(batch) != 1:
            raise ValueError("solve_flowrate() takes one model; use solve_flowrate_batch() for several")
        nk = int(batch.counters[0, _lib.C_KNOTS].item())
        if nk <= k_max:
            break
        k_max = 1 << (nk - 1).bit_length()
    sol = Solution(batch)
    if throw and sol.result != RESULTS.successful:
        raise SolveError("the shooting bisection on b did not converge (max_steps_reached or no bracket)")
    return sol


def solve_batch_host(models: ModelBatch, b: np.ndarray, i: np.ndarray, *, itol: float = 1e-3,
                     max_steps: int = 100, k_max: int = 0, summary: np.ndarray | None = None,
                     counters: np.ndarray | None = None, options: dict | None = None):
    """Host-buffer entry point (``fx_solve_batch_host``): numpy in, numpy out, copies inside."""
    _lib.require_cuda()
    n = len(models)
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64), (n,)))
    i = np.ascontiguousarray(np.broadcast_to(np.asarray(i, dtype=np.float64), (n,)))
    if summary is None:
        summary = np.empty((n, _lib.NSUM), dtype=np.float64)
    if counters is None:
        counters = np.zeros((n, _lib.NCNT), dtype=np.int32)
    knots = np.zeros((n, k_max, _lib.KNOT)) if k_max > 0 else None
    opt = _lib.default_options(itol=float(itol), max_bisect=int(max_steps),
                               **{"model_mask": models.mask(), **(options or {})})
    rc = _lib.lib().fx_solve_batch_host(n, models.model_id.ctypes.data, models.params.ctypes.data,
                                        b.ctypes.data, i.ctypes.data, opt, summary.ctypes.data,
                                        counters.ctypes.data, k_max,
                                        knots.ctypes.data if knots is not None else None)
    _lib.check(rc)
    return summary, counters, knots


__all__: Sequence[str] = ("RESULTS", "Solution", "BatchSolution", "SolveError", "solve", "solve_batch",
                          "solve_batch_host", "solve_flowrate", "solve_flowrate_batch")
