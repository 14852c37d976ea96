"""Shared fixtures.  Tests marked ``gpu`` need a B200; everything else runs on CPU.

The CPU oracle (oracle/) is test infrastructure: it is the checker here, never
the thing under test in the ``gpu`` tests, and the product package never
imports it.
"""

from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


# boundary values of the reference's forward tests (tests/test_solve.py:127-128)
B_PAPER = 0.7 - 1e-7
I_PAPER = 0.025


def paper_models():
    """The six parameter sets of /root/reference/tests/test_solve.py:47-113."""
    from frontx_b200 import models as M

    return {
        "LETd#1": M.LETd(L=0.004569, E=12930, T=1.505, Dwt=4.660e-4, theta_range=(0.019852, 0.7)),
        "LETd#2": M.LETd(Dwt=1.004e-3, L=1.356, E=10010, T=1.224, theta_range=(0.00625, 0.7)),
        "VG(n)": M.VanGenuchten(n=8.093, l=2.344, Ks=2.079e-6, theta_range=(0.004943, 0.7)),
        "VG(m)": M.VanGenuchten(m=0.8861, l=2.331, Ks=2.105e-6, theta_range=(0.005623, 0.7)),
        "BC": M.BrooksAndCorey(n=0.2837, l=4.795, Ks=3.983e-6, theta_range=(2.378e-5, 0.7)),
        "LETxs": M.LETxs(Lw=1.651, Ew=230.5, Tw=0.9115, Ls=0.517, Es=493.6, Ts=0.3806, Ks=8.900e-3,
                         theta_range=(0.01176, 0.7)),
    }


@pytest.fixture(scope="session")
def paper():
    return paper_models()


@pytest.fixture(scope="session")
def fxo():
    from oracle import fxo as _fxo

    _fxo.build()
    return _fxo


def vg_sweep(n: int, seed: int = 0):
    """BASELINE config C2: van Genuchten sweep."""
    from frontx_b200 import models as M

    rng = np.random.default_rng(seed)
    nn = rng.uniform(1.5, 9.0, n)
    alpha = np.exp(rng.uniform(np.log(0.5), np.log(5.0), n))
    Ks = np.exp(rng.uniform(np.log(1e-7), np.log(1e-4), n))
    return M.VanGenuchten.batch(n=nn, alpha=alpha, Ks=Ks, l=0.5, theta_range=(0.005, 0.7))


def mixed_sweep(n: int, seed: int = 1):
    """BASELINE config C3: Brooks-Corey (even k) / LETd (odd k), interleaved."""
    from frontx_b200 import ModelBatch
    from frontx_b200 import models as M

    rng = np.random.default_rng(seed)
    nb = (n + 1) // 2
    nl = n // 2
    bc = M.BrooksAndCorey.batch(n=rng.uniform(0.2, 0.6, nb), l=rng.uniform(1.0, 5.0, nb),
                                Ks=np.exp(rng.uniform(np.log(1e-6), np.log(1e-5), nb)),
                                theta_range=(2.378e-5, 0.7))
    ld = M.LETd.batch(L=rng.uniform(0.004, 1.5, nl), E=np.exp(rng.uniform(np.log(1e3), np.log(2e4), nl)),
                      T=rng.uniform(1.0, 1.6, nl), Dwt=np.exp(rng.uniform(np.log(1e-4), np.log(2e-3), nl)),
                      theta_range=(0.0198, 0.7))
    ids = np.empty(n, dtype=np.int32)
    par = np.empty((n, 12))
    ids[0::2], par[0::2] = bc.model_id, bc.params
    ids[1::2], par[1::2] = ld.model_id, ld.params
    return ModelBatch(ids, par)


def family_sweep(which: str, n: int, seed: int = 2):
    """Random sweeps of the model families BASELINE configs C2/C3 do not cover, with the boundary values they are
    solved for: (ModelBatch, b, i).  LETxs around the paper set (Gerlero et al. 2022, Table 2); the power law and
    the logarithmic diffusivity are the families of the reference's analytic tests (tests/test_solve.py:40,151)."""
    from frontx_b200 import models as M

    rng = np.random.default_rng(seed)
    lu = lambda lo, hi: np.exp(rng.uniform(np.log(lo), np.log(hi), n))  # noqa: E731
    if which == "letxs":
        mb = M.LETxs.batch(Lw=rng.uniform(1.0, 2.5, n), Ew=lu(20.0, 2000.0), Tw=rng.uniform(0.6, 1.5, n),
                           Ls=rng.uniform(0.3, 1.0, n), Es=lu(50.0, 2000.0), Ts=rng.uniform(0.3, 0.8, n),
                           Ks=lu(1e-3, 3e-2), theta_range=(0.01176, 0.7))
        return mb, B_PAPER, I_PAPER
    if which == "power":
        mb = M.PowerLaw.batch(k=rng.uniform(0.5, 6.0, n), a=lu(0.1, 10.0), epsilon=lu(1e-6, 1e-3))
        return mb, 1.0, 0.05
    if which == "log":
        mb = M.LogDiffusivity.batch(c0=rng.uniform(0.3, 1.0, n), c1=-rng.uniform(0.1, 0.8, n))
        return mb, 1.0, 0.01
    raise ValueError(which)


def interior_boundaries(n: int, seed: int = 11):
    """Per-problem boundary values away from saturation, 40 % of them reversed (i > b: drainage): (b, i)."""
    rng = np.random.default_rng(seed)
    b = rng.uniform(0.3, 0.69, n)
    i = rng.uniform(0.03, 0.25, n)
    swap = rng.random(n) < 0.4
    return np.where(swap, i, b), np.where(swap, b, i)


def all_families(n_each: int, seed: int = 21):
    """Every closed-form family in ONE batch, interleaved, each problem with the boundary values of its family:
    (ModelBatch, b[n], i[n])."""
    from frontx_b200 import ModelBatch
    from frontx_b200 import models as M

    parts = []
    vg = vg_sweep(n_each, seed=seed)
    parts.append((vg, np.full(n_each, B_PAPER), np.full(n_each, I_PAPER)))
    mx = mixed_sweep(2 * n_each, seed=seed + 1)
    parts.append((mx, np.full(2 * n_each, B_PAPER), np.full(2 * n_each, I_PAPER)))
    for k, which in enumerate(("letxs", "power", "log")):
        mb, b, i = family_sweep(which, n_each, seed=seed + 2 + k)
        parts.append((mb, np.full(n_each, b), np.full(n_each, i)))
    rng = np.random.default_rng(seed + 9)
    const = M.Constant.batch(D0=np.exp(rng.uniform(np.log(1e-3), np.log(10.0), n_each)))
    parts.append((const, np.full(n_each, 1.0), np.full(n_each, 0.0)))
    ids = np.concatenate([p[0].model_id for p in parts])
    par = np.concatenate([p[0].params for p in parts])
    b = np.concatenate([p[1] for p in parts])
    i = np.concatenate([p[2] for p in parts])
    order = rng.permutation(ids.size)
    return ModelBatch(ids[order], par[order]), b[order], i[order]
