"""frontx_b200 -- B200-native batched Boltzmann-transformed shooting solve.

Drop-in for the hot path of gerlero/frontx: ``solve`` / ``Solution`` (plus the
batched ``solve_batch``), the tagged diffusivity models, and the profile -> D
inverse (``InterpolatedSolution``, ``sorptivity``, batched ``inverse``), and the
parametric ``fit`` / ``ScaledSolution`` that call the solve a generation at a time.  All
compute runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/frontx_b200.h``; there is no CPU fallback.
"""

from . import D, models
from ._forward import (RESULTS, BatchSolution, Solution, SolveError, solve, solve_batch, solve_batch_host,
                       solve_flowrate, solve_flowrate_batch)
from ._fit import Param, ScaledSolution, fit
from ._inverse import BatchInverse, InterpolatedSolution, inverse, sorptivity, unique_profile
from ._lib import FrontxB200Error
from .models import ModelBatch

__version__ = "0.1.0"

__all__ = [
    "RESULTS",
    "BatchInverse",
    "BatchSolution",
    "D",
    "FrontxB200Error",
    "InterpolatedSolution",
    "ModelBatch",
    "Param",
    "ScaledSolution",
    "Solution",
    "SolveError",
    "fit",
    "inverse",
    "models",
    "solve",
    "solve_batch",
    "solve_batch_host",
    "solve_flowrate",
    "solve_flowrate_batch",
    "sorptivity",
    "unique_profile",
]
