"""Diffusivity functions under the names the north star uses (the ``fronts.D`` API):
``constant``, ``power_law``, ``brooks_and_corey``, ``van_genuchten``, ``letxs``, ``letd``.

Each returns the corresponding tagged model of :mod:`frontx_b200.models`
(reference classes: src/frontx/models.py:35,126,177,256; ``constant`` and
``power_law`` have no class in the reference -- tests/test_solve.py:40,151 use lambdas).
"""

from __future__ import annotations

from .models import BrooksAndCorey, Constant, LETd, LETxs, LogDiffusivity, PowerLaw, VanGenuchten


def constant(D0: float) -> Constant:
    return Constant(D0)


def power_law(k: float, a: float = 1.0, epsilon: float = 0.0) -> PowerLaw:
    return PowerLaw(k=k, a=a, epsilon=epsilon)


def brooks_and_corey(**kwargs) -> BrooksAndCorey:
    return BrooksAndCorey(**kwargs)


def van_genuchten(**kwargs) -> VanGenuchten:
    return VanGenuchten(**kwargs)


def letxs(**kwargs) -> LETxs:
    return LETxs(**kwargs)


def letd(**kwargs) -> LETd:
    return LETd(**kwargs)


def philip_exact() -> LogDiffusivity:
    """D = (1 - ln theta)/2, whose solution for b=1, i=0 is exp(-o) (Philip 1960)."""
    return LogDiffusivity(0.5, -0.5)
