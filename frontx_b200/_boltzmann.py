"""Boltzmann-variable plumbing shared by every solution type.

Mirrors src/frontx/_boltzmann.py:55-161 of the reference: methods accept either
the Boltzmann variable ``o`` or physical ``(r, t)`` with ``o = r/sqrt(t)``, and
``d_dr``, ``d_dt``, ``flux``, ``sorptivity`` are derived from ``__call__`` and
``d_do`` exactly as there.  The arithmetic here is on a handful of host arrays
after the device kernels have run; it is API glue, not the hot path.
"""

from __future__ import annotations

from functools import wraps
from typing import Any, Callable

import numpy as np


def boltzmannmethod(meth: Callable) -> Callable:
    """``meth(self, o)`` also callable as ``(r, t)``, ``o=``, or ``r=, t=`` (_boltzmann.py:55-98)."""

    @wraps(meth)
    def boltzmann_wrapper(self, *args: Any, **kwargs: Any) -> Any:
        if len(args) == 1 and not kwargs:
            o = args[0]
        elif len(args) == 2 and not kwargs:  # noqa: PLR2004
            r, t = args
            o = np.asarray(r, dtype=np.float64) / np.sqrt(np.asarray(t, dtype=np.float64))
        elif not args and set(kwargs) == {"o"}:
            o = kwargs["o"]
        elif not args and set(kwargs) == {"r", "t"}:
            o = np.asarray(kwargs["r"], dtype=np.float64) / np.sqrt(np.asarray(kwargs["t"], dtype=np.float64))
        else:
            meth_name = getattr(meth, "__name__", "method")
            msg = f"{meth_name} takes (r, t) or (o,) as arguments"
            raise TypeError(msg)
        return meth(self, o)

    return boltzmann_wrapper


class AbstractSolution:
    """Solution algebra of src/frontx/_boltzmann.py:101-161.

    Subclasses provide ``__call__(o)``, ``d_do(o)``, ``oi`` and ``D``.
    """

    D: Callable[[Any], Any]
    oi: float

    @property
    def b(self) -> float:
        return float(self(0.0))

    @property
    def d_dob(self) -> float:
        return float(self.d_do(0.0))

    @property
    def i(self) -> float:
        return float(self(self.oi))

    def d_dr(self, r: Any, t: Any) -> Any:
        return self.d_do(r, t) / np.sqrt(t)

    def d_dt(self, r: Any, t: Any) -> Any:
        return -np.asarray(r) / (np.sqrt(t) * 2 * np.asarray(t)) * self.d_do(r, t)

    def flux(self, r: Any, t: Any) -> Any:
        return -np.asarray(self.D(self(r, t))) * self.d_dr(r, t)

    def sorptivity(self, o: Any = 0.0) -> Any:
        return -2 * np.asarray(self.D(self(o))) * self.d_do(o)
