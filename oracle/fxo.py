"""ctypes binding of the CPU oracle (oracle/libfx_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py.  Nothing under
``frontx_b200/`` may import this module.  PARITY UNPINNED at 1e-6 (see
oracle/fx_oracle.h).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libfx_oracle.so"

NPARAM, NSUM, NCNT, KNOT = 12, 8, 8, 5

CONSTANT, POWER_LAW, BROOKS_COREY, VAN_GENUCHTEN, LETXS, LETD, LOG, TABULATED = range(8)


class Options(C.Structure):
    _fields_ = [
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("max_ode_steps", C.c_int),
        ("max_expansions", C.c_int),
        ("radial_k", C.c_int),
        ("ob", C.c_double),
        ("u4_failure_aborts", C.c_int),
        ("u5_or_termination", C.c_int),
        ("u6_expand_rule", C.c_int),
    ]


def build(force: bool = False) -> Path:
    csrc = HERE.parent / "frontx_b200" / "csrc"
    srcs = [HERE / "fx_oracle.c", HERE / "fx_oracle_inverse.c", HERE / "fx_oracle.h", HERE / "fx_twin.cpp",
            HERE / "Makefile", csrc / "fx_math.cuh", csrc / "fx_tables.inc", csrc / "fx_models.cuh", csrc / "fx_arith.cuh"]
    libs = [LIB, HERE / "libfx_twin.so", HERE / "libfx_twin_fma.so"]
    stale = any(not p.exists() for p in libs) or any(
        s.stat().st_mtime > min(p.stat().st_mtime for p in libs) for s in srcs if s.exists())
    missing_sources = not all(s.exists() for s in srcs)
    if (force or stale) and not missing_sources:
        env = dict(os.environ)
        env.pop("CC", None)
        env.pop("CXX", None)
        proc = subprocess.run(["make", "-C", str(HERE), "-B"], env=env, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("building the oracle failed:\n" + proc.stdout + proc.stderr)
    return LIB


def use_native() -> str:
    """For the CPU TIMING arm of bench.py: rebuild both checkers with -O3 -march=native ON THIS BOX (into
    oracle/_native/, git-ignored, never shipped: the instruction set is the box's own) and load those instead
    of the portable builds.  Same sources, -ffp-contract=off: same results, only faster.  Returns the flags."""
    global _lib, _twin, LIB, _TWIN_PATH
    proc = subprocess.run(["make", "-C", str(HERE), "-B", "native"], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building the native oracle failed:\n" + proc.stdout + proc.stderr)
    LIB = HERE / "_native" / "libfx_oracle.so"
    _TWIN_PATH = HERE / "_native" / "libfx_twin.so"
    _lib = _twin = None
    return "-O3 -march=native -ffp-contract=off"


_twin = None
_TWIN_PATH = None


def twin() -> C.CDLL:
    """Host build of the kernel's arithmetic under plain loops (oracle/fx_twin.cpp)."""
    global _twin
    if _twin is None:
        build()
        has_fma = False
        try:
            has_fma = " fma " in open("/proc/cpuinfo").read().replace("\n", " ")
        except OSError:
            pass
        L = C.CDLL(str(_TWIN_PATH or HERE / ("libfx_twin_fma.so" if has_fma else "libfx_twin.so")))
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.fxt_D.argtypes = [C.c_int, dp, C.c_double, dp]
        for name in ("fxt_log", "fxt_exp", "fxt_expm1"):
            getattr(L, name).argtypes = [C.c_double]
            getattr(L, name).restype = C.c_double
        L.fxt_solve_batch.argtypes = [C.c_int64, ip, dp, dp, dp, C.c_double, C.c_int, C.c_int, C.c_double,
                                      C.c_double, C.c_int, C.c_int, C.c_double, dp, ip, C.c_int, dp, C.c_int]
        L.fxt_solve_batch.restype = C.c_int
        L.fxt_solve_flowrate_batch.argtypes = [C.c_int64, ip, dp, dp, dp, dp, dp, C.c_double, C.c_int, C.c_int,
                                               C.c_double, C.c_double, C.c_int, C.c_double, dp, ip, C.c_int, dp,
                                               C.c_int]
        L.fxt_solve_flowrate_batch.restype = C.c_int
        _twin = L
    return _twin


def twin_D(model: int, params, theta) -> np.ndarray:
    p = pad_params(params)
    th = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    out = np.empty((th.size, 3))
    tmp = np.empty(3)
    for k, t in enumerate(th.ravel()):
        twin().fxt_D(model, _dp(p), float(t), _dp(tmp))
        out[k] = tmp
    return out.reshape(th.shape + (3,))


def twin_solve_batch(model, params, b, i, itol: float = 1e-3, max_steps: int = 100, k_max: int = 0,
                     nthreads: int = 0, max_ode_steps: int = 4096, rtol: float = 1e-3, atol: float = 1e-6,
                     max_expansions: int = 64, radial_k: int = 0, ob: float = 0.0):
    """Same outputs as the GPU entry point: summary[n,8], counters[n,8], knots[n,k_max,4]."""
    model = np.ascontiguousarray(model, dtype=np.int32)
    n = model.size
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(n, NPARAM)
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64), (n,)))
    i = np.ascontiguousarray(np.broadcast_to(np.asarray(i, dtype=np.float64), (n,)))
    summary = np.empty((n, NSUM))
    counters = np.zeros((n, NCNT), dtype=np.int32)
    knots = np.zeros((n, k_max, 4)) if k_max > 0 else None
    twin().fxt_solve_batch(n, _ip(model), _dp(params), _dp(b), _dp(i), float(itol), int(max_steps),
                           int(max_ode_steps), float(rtol), float(atol), int(max_expansions), int(radial_k),
                           float(ob), _dp(summary), _ip(counters), k_max,
                           _dp(knots) if knots is not None else None, int(nthreads))
    return {"summary": summary, "counters": counters, "knots": knots}


def twin_solve_flowrate_batch(model, params, q, i, b_lo, b_hi, itol: float = 1e-3, max_steps: int = 100,
                              k_max: int = 0, nthreads: int = 0, max_ode_steps: int = 4096, rtol: float = 1e-3,
                              atol: float = 1e-6, radial_k: int = 1, ob: float = 1e-6):
    """Flow-rate boundary (bisection on b): summary = {b, oi, theta(oi), d_dob, D(b), residual, b_lower, b_upper}."""
    model = np.ascontiguousarray(model, dtype=np.int32)
    n = model.size
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(n, NPARAM)
    full = lambda a: np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float64), (n,)))  # noqa: E731
    q, i, b_lo, b_hi = full(q), full(i), full(b_lo), full(b_hi)
    summary = np.empty((n, NSUM))
    counters = np.zeros((n, NCNT), dtype=np.int32)
    knots = np.zeros((n, k_max, 4)) if k_max > 0 else None
    twin().fxt_solve_flowrate_batch(n, _ip(model), _dp(params), _dp(q), _dp(i), _dp(b_lo), _dp(b_hi), float(itol),
                                    int(max_steps), int(max_ode_steps), float(rtol), float(atol), int(radial_k),
                                    float(ob), _dp(summary), _ip(counters), k_max,
                                    _dp(knots) if knots is not None else None, int(nthreads))
    return {"summary": summary, "counters": counters, "knots": knots}


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int32)
        L.fxo_default_options.argtypes = [C.POINTER(Options)]
        L.fxo_D.argtypes = [C.c_int, dp, C.c_double, dp]
        L.fxo_rhs.argtypes = [C.c_int, dp, C.c_int, C.c_double, dp, dp]
        L.fxo_shoot.argtypes = [C.c_int, dp, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.POINTER(Options), C.c_int, dp, C.POINTER(C.c_int), ip]
        L.fxo_shoot.restype = C.c_double
        L.fxo_solve.argtypes = [C.c_int, dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                C.POINTER(Options), dp, ip, C.c_int, dp]
        L.fxo_solve.restype = C.c_int
        L.fxo_solve_batch.argtypes = [C.c_int64, ip, dp, dp, dp, C.c_double, C.c_int,
                                      C.POINTER(Options), dp, ip, C.c_int, dp, C.c_int]
        L.fxo_solve_batch.restype = C.c_int
        L.fxo_eval.argtypes = [C.c_int, dp, C.c_int64, dp, dp, dp]
        L.fxo_inverse.argtypes = [C.c_int64, dp, dp, dp, dp, dp]
        L.fxo_inverse.restype = C.c_int
        L.fxo_num_threads.restype = C.c_int
        L.fxo_diag_get.argtypes = [C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def default_options(**kw) -> Options:
    o = Options()
    lib().fxo_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def pad_params(p) -> np.ndarray:
    out = np.zeros(NPARAM, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    out[: p.size] = p
    return out


def D(model: int, params, theta) -> np.ndarray:
    """(D, D', D'') at each theta; shape [..., 3]."""
    p = pad_params(params)
    th = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    out = np.empty((th.size, 3))
    tmp = np.empty(3)
    for k, t in enumerate(th.ravel()):
        lib().fxo_D(model, _dp(p), float(t), _dp(tmp))
        out[k] = tmp
    return out.reshape(th.shape + (3,))


def rhs(model: int, params, o: float, y, radial_k: int = 0) -> np.ndarray:
    p = pad_params(params)
    yy = np.asarray(y, dtype=np.float64).copy()
    f = np.empty(2)
    lib().fxo_rhs(model, _dp(p), radial_k, float(o), _dp(yy), _dp(f))
    return f


def solve(model: int, params, b: float, i: float, itol: float = 1e-3, max_steps: int = 100,
          k_max: int = 512, options: Options | None = None):
    """One problem. Returns dict(summary, counters, knots[n_knots,5])."""
    p = pad_params(params)
    summary = np.empty(NSUM)
    counters = np.zeros(NCNT, dtype=np.int32)
    knots = np.zeros((k_max, KNOT))
    opt = options if options is not None else default_options()
    lib().fxo_solve(model, _dp(p), float(b), float(i), float(itol), int(max_steps),
                    C.byref(opt), _dp(summary), _ip(counters), k_max, _dp(knots))
    nk = min(int(counters[7]), k_max)
    return {"summary": summary, "counters": counters, "knots": knots[:nk].copy()}


def solve_batch(model, params, b, i, itol: float = 1e-3, max_steps: int = 100, k_max: int = 0,
                options: Options | None = None, nthreads: int = 0):
    model = np.ascontiguousarray(model, dtype=np.int32)
    n = model.size
    params = np.ascontiguousarray(params, dtype=np.float64).reshape(n, NPARAM)
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(b, dtype=np.float64), (n,)))
    i = np.ascontiguousarray(np.broadcast_to(np.asarray(i, dtype=np.float64), (n,)))
    summary = np.empty((n, NSUM))
    counters = np.zeros((n, NCNT), dtype=np.int32)
    knots = np.zeros((n, k_max, KNOT)) if k_max > 0 else None
    opt = options if options is not None else default_options()
    lib().fxo_solve_batch(n, _ip(model), _dp(params), _dp(b), _dp(i), float(itol), int(max_steps),
                          C.byref(opt), _dp(summary), _ip(counters), k_max,
                          _dp(knots) if knots is not None else None, int(nthreads))
    return {"summary": summary, "counters": counters, "knots": knots}


def evaluate(knots: np.ndarray, o) -> tuple[np.ndarray, np.ndarray]:
    """theta(o), dtheta/do(o) from a [n_knots,5] table (Solution.__call__/d_do)."""
    k = np.ascontiguousarray(knots, dtype=np.float64)
    oo = np.ascontiguousarray(np.atleast_1d(o), dtype=np.float64)
    th = np.empty(oo.size)
    dth = np.empty(oo.size)
    lib().fxo_eval(k.shape[0], _dp(k), oo.size, _dp(oo.ravel()), _dp(th), _dp(dth))
    return th.reshape(oo.shape), dth.reshape(oo.shape)


def inverse(o, theta):
    oo = np.ascontiguousarray(o, dtype=np.float64)
    th = np.ascontiguousarray(theta, dtype=np.float64)
    Dk = np.empty(oo.size)
    s1 = C.c_double()
    s2 = C.c_double()
    rc = lib().fxo_inverse(oo.size, _dp(oo), _dp(th), _dp(Dk), C.byref(s1), C.byref(s2))
    return rc, Dk, s1.value, s2.value


def tabulated(o, theta, b=None, i=None):
    """The table of a FXO_TABULATED model from scipy's own PchipInterpolator / PPoly (what interpax ports), following
    interpolated.py:30-50 line for line: returns (params[12], table) -- keep `table` alive while the model is in use."""
    from scipy.interpolate import PchipInterpolator

    o = np.asarray(o, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    if b is not None:
        o, theta = np.insert(o, 0, 0), np.insert(theta, 0, b)
    if i is not None:
        o, theta = np.append(o, o[-1] + 1), np.append(theta, i)
    else:
        i = theta[-1]
    theta, idx = np.unique(theta, return_index=True)
    o = o[idx]
    inv = PchipInterpolator(theta, o, extrapolate=False)
    anti = inv.antiderivative()
    table = np.ascontiguousarray(np.stack([theta, o, inv.derivative()(theta), anti(theta) - anti(i)]))
    p = np.zeros(NPARAM)
    p[0] = np.array([table.ctypes.data], dtype=np.uint64).view(np.float64)[0]
    p[1] = theta.size
    return p, table


def vg_form(params, form: str) -> np.ndarray:
    """Copy of a [n,12] parameter block with every van Genuchten row's form flag (column 7) set:
    'literal' = the reference's cancelling subtraction (the default of the models), 'expm1' = the
    same value without cancellation."""
    out = np.array(params, dtype=np.float64, copy=True)
    out[..., 7] = {"literal": 0.0, "expm1": 1.0}[form]
    return out


DIAG = ("stage_fail", "error_reject", "nan_error", "expansion")


def diag_reset() -> None:
    lib().fxo_diag_reset()


def diag() -> dict:
    out = np.zeros(len(DIAG), dtype=np.int64)
    lib().fxo_diag_get(out.ctypes.data_as(C.POINTER(C.c_int64)))
    return dict(zip(DIAG, (int(v) for v in out)))


def num_threads() -> int:
    return int(lib().fxo_num_threads())
