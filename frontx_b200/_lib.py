"""ctypes binding of libfrontx_b200.so (the C ABI in include/frontx_b200.h).

There is no CPU fallback: every compute call goes to the CUDA library and
fails loudly when the library or a CUDA device is missing.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

NPARAM, NSUM, NCNT, KNOT = 12, 8, 8, 4
LIB_PATH = Path(__file__).resolve().parent / "libfrontx_b200.so"

# model ids (enum fx_model)
CONSTANT, POWER_LAW, BROOKS_COREY, VAN_GENUCHTEN, LETXS, LETD, LOG, TABULATED = range(8)
MODEL_NAMES = ("constant", "power_law", "brooks_and_corey", "van_genuchten", "letxs", "letd", "log", "tabulated")

# summary / counter columns
S_D_DOB, S_OI, S_THETA_OI, S_S0, S_D_B, S_RESIDUAL, S_LOWER, S_UPPER = range(8)
C_RESULT, C_SHOTS, C_EXPANSIONS, C_ATTEMPTS, C_REJECTED, C_RHS, C_JAC, C_KNOTS = range(8)


class SolveOptions(C.Structure):
    """struct fx_solve_options"""

    _fields_ = [
        ("itol", C.c_double),
        ("max_bisect", C.c_int32),
        ("max_ode_steps", C.c_int32),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("max_expansions", C.c_int32),
        ("radial_k", C.c_int32),
        ("ob", C.c_double),
        ("model_mask", C.c_uint32),
    ]


class FrontxB200Error(RuntimeError):
    """The CUDA library reported an error (argument, CUDA runtime, model id)."""


_lib = None

# every symbol include/frontx_b200.h declares
EXPORTS = (
    "fx_default_solve_options",
    "fx_solve_batch",
    "fx_solve_batch_host",
    "fx_solve_flowrate_batch",
    "fx_eval_batch",
    "fx_scaled_fit_batch",
    "fx_diffusivity_batch",
    "fx_inverse_batch",
    "fx_inverse_query_batch",
    "fx_pchip_eval_batch",
    "fx_pchip_query_batch",
    "fx_unique_batch",
    "fx_measure_fp64_peak",
    "fx_set_kernel_timing",
    "fx_last_solve_kernel_ms",
    "fx_launch_count",
    "fx_last_error",
    "fx_version",
    "fx_device_count",
)


def lib() -> C.CDLL:
    """Load the CUDA extension; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FrontxB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(frontx_b200 has no CPU fallback)"
        )
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    L.fx_default_solve_options.argtypes = [C.POINTER(SolveOptions)]
    L.fx_default_solve_options.restype = None
    L.fx_solve_batch.argtypes = [i64, vp, vp, vp, vp, C.POINTER(SolveOptions), vp, vp, i32, vp, vp]
    L.fx_solve_batch_host.argtypes = [i64, vp, vp, vp, vp, C.POINTER(SolveOptions), vp, vp, i32, vp]
    L.fx_solve_flowrate_batch.argtypes = [i64, vp, vp, vp, vp, vp, vp, C.POINTER(SolveOptions), vp, vp, i32, vp, vp]
    L.fx_eval_batch.argtypes = [i64, vp, vp, i32, i64, vp, i32, vp, vp, vp]
    L.fx_diffusivity_batch.argtypes = [i64, vp, vp, i64, vp, vp, vp]
    L.fx_scaled_fit_batch.argtypes = [i64, vp, vp, vp, i32, i64, vp, vp, vp, i32, vp, vp, vp, vp]
    L.fx_inverse_batch.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.fx_inverse_query_batch.argtypes = [i64, i64, vp, vp, vp, vp, i64, vp, vp, vp]
    L.fx_pchip_eval_batch.argtypes = [i64, i64, vp, vp, i64, vp, vp, vp]
    L.fx_pchip_query_batch.argtypes = [i64, i64, vp, vp, vp, vp, i64, vp, i32, i32, vp, vp, vp, vp]
    L.fx_unique_batch.argtypes = [i64, i64, vp, vp, vp, vp, vp, vp, vp]
    L.fx_measure_fp64_peak.argtypes = [C.POINTER(dbl), vp]
    L.fx_set_kernel_timing.argtypes = [i32]
    L.fx_set_kernel_timing.restype = None
    L.fx_last_solve_kernel_ms.argtypes = [i32, C.POINTER(C.c_float)]
    L.fx_last_solve_kernel_ms.restype = C.c_int
    L.fx_launch_count.restype = i64
    L.fx_last_error.restype = C.c_char_p
    for name in ("fx_solve_batch", "fx_solve_batch_host", "fx_solve_flowrate_batch", "fx_eval_batch", "fx_scaled_fit_batch", "fx_diffusivity_batch",
                 "fx_inverse_batch", "fx_inverse_query_batch", "fx_pchip_eval_batch", "fx_pchip_query_batch",
                 "fx_unique_batch", "fx_measure_fp64_peak",
                 "fx_version",
                 "fx_device_count"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().fx_last_error().decode("utf-8", "replace")
        raise FrontxB200Error(f"libfrontx_b200 error {rc}: {msg}")


def require_cuda():
    """Return torch after checking a CUDA device is usable (no CPU fallback)."""
    import torch

    lib()
    if not torch.cuda.is_available():
        raise FrontxB200Error("frontx_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def default_options(**kw) -> SolveOptions:
    o = SolveOptions()
    lib().fx_default_solve_options(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown solve option {k!r}")
        setattr(o, k, v)
    return o


def launch_count() -> int:
    return int(lib().fx_launch_count())
