"""CPU-side checks: the C-ABI library loads and exports every symbol the header
declares; the host-side mirror of the reference interface packs models, parses
arguments and fails the way the reference does; nothing falls back to the CPU.
No compute is launched here (there is no GPU in this container).
"""

from __future__ import annotations

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import frontx_b200 as fx
from frontx_b200 import _lib
from frontx_b200 import models as M
from frontx_b200._boltzmann import AbstractSolution, boltzmannmethod

ROOT = Path(__file__).resolve().parents[1]
HEADER = (ROOT / "include" / "frontx_b200.h").read_text()


def declared_functions() -> list[str]:
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(fx_[a-z0-9_]+)\s*\(", body)))


def test_library_exports_every_declared_symbol():
    assert _lib.LIB_PATH.exists(), "build the extension first: __graft_entry__.build()"
    L = ctypes.CDLL(str(_lib.LIB_PATH))
    names = declared_functions()
    assert len(names) >= 12
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/frontx_b200.h but not exported"
    assert set(_lib.EXPORTS) <= set(names)
    assert L.fx_version() == 100


def test_header_struct_matches_binding():
    o = _lib.default_options()
    assert (o.itol, o.max_bisect, o.max_ode_steps, o.rtol, o.atol) == (1e-3, 100, 4096, 1e-3, 1e-6)
    assert (o.max_expansions, o.radial_k, o.ob) == (64, 0, 0.0)
    with pytest.raises(TypeError):
        _lib.default_options(nonsense=1)


def test_argument_errors_without_a_device():
    L = _lib.lib()
    assert L.fx_solve_batch(-1, None, None, None, None, None, None, None, 0, None, None) == -1
    assert b"negative" in L.fx_last_error()
    assert L.fx_solve_batch(4, None, None, None, None, None, None, None, 0, None, None) == -1
    assert L.fx_inverse_batch(1, 1, None, None, None, None, None, None, None, None) == -1
    assert L.fx_solve_batch(0, None, None, None, None, None, None, None, 0, None, None) == 0
    # flow-rate boundary: needs q and a bracket, and a radial geometry
    assert L.fx_solve_flowrate_batch(3, None, None, None, None, None, None, None, None, None, 0, None, None) == -1
    planar = _lib.default_options()
    dummy = ctypes.c_void_p(8)
    assert L.fx_solve_flowrate_batch(3, dummy, dummy, dummy, dummy, dummy, dummy, planar, dummy, dummy, 0, None,
                                     None) == -1
    assert b"radial" in L.fx_last_error()
    assert L.fx_solve_flowrate_batch(0, None, None, None, None, None, None, None, None, None, 0, None, None) == 0
    # batched D0 fit: mode in {0, 1, 2}, at least one data point, dense output required
    assert L.fx_scaled_fit_batch(2, dummy, dummy, dummy, 8, 5, dummy, dummy, None, 3, dummy, dummy, dummy, None) == -1
    assert L.fx_scaled_fit_batch(2, dummy, dummy, dummy, 8, 0, dummy, dummy, None, 1, dummy, dummy, dummy, None) == -1
    assert L.fx_scaled_fit_batch(2, dummy, dummy, None, 8, 5, dummy, dummy, None, 1, dummy, dummy, dummy, None) == -1
    assert L.fx_scaled_fit_batch(0, None, None, None, 8, 5, None, None, None, 1, None, None, None, None) == 0


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fx.FrontxB200Error, match="no CPU fallback"):
        fx.solve(fx.D.power_law(k=4), b=1, i=0.1)
    with pytest.raises(fx.FrontxB200Error):
        fx.inverse(np.linspace(0, 1, 8), np.linspace(1, 0.1, 8))
    with pytest.raises(fx.FrontxB200Error):
        M.LETd(L=1, E=1, T=1)(0.5)


def test_product_never_imports_the_oracle():
    for path in (ROOT / "frontx_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".h"}:
            for line in path.read_text().splitlines():
                if any(k in line for k in ("import", "#include", "CDLL", "dlopen")):
                    assert "oracle" not in line and "fxo" not in line, f"{path}: {line.strip()}"


def test_arbitrary_callable_is_rejected():
    with pytest.raises(TypeError, match="tagged|frontx_b200 model"):
        fx.solve_batch(lambda theta: theta, b=1, i=0)
    with pytest.raises(TypeError):
        fx.ModelBatch.from_models([lambda theta: theta])


# ---- model packing mirrors src/frontx/models.py -----------------------------------
def test_van_genuchten_n_m_relation():
    a = M.VanGenuchten(n=8.093, l=2.344, Ks=2.079e-6, theta_range=(0.004943, 0.7)).params()
    assert a[0] == 8.093 and a[1] == 1 - 1 / 8.093
    b = M.VanGenuchten(m=0.8861, l=2.331, Ks=2.105e-6, theta_range=(0.005623, 0.7)).params()
    assert b[1] == 0.8861 and b[0] == 1 / (1 - 0.8861)
    with pytest.raises(ValueError, match="Either n or m must be set"):
        M.VanGenuchten().params()


def test_Ks_resolution():
    # models.py:76-86: Ks given | k given -> rho*g*k/mu | neither -> 1 | both -> ValueError
    assert M.BrooksAndCorey(n=0.3, Ks=2.0).params()[2] == 2.0
    assert M.BrooksAndCorey(n=0.3, k=1e-12).params()[2] == 1e3 * 9.81 * 1e-12 / 1e-3
    assert M.BrooksAndCorey(n=0.3).params()[2] == 1.0
    with pytest.raises(ValueError, match="Cannot set both Ks and k"):
        M.LETxs(Lw=1, Ew=1, Tw=1, Ls=1, Es=1, Ts=1, Ks=1.0, k=1e-12).params()


def test_defaults_match_reference():
    assert M.LETd(L=1, E=2, T=3).params()[:6].tolist() == [1, 2, 3, 1, 0, 1]
    assert M.BrooksAndCorey(n=0.5).l == 1 and M.VanGenuchten(n=2).l == 0.5
    assert M.BrooksAndCorey(n=0.5).params()[3:6].tolist() == [1, 0, 1]


def test_north_star_names():
    assert fx.D.constant(2.0).model_id == _lib.CONSTANT
    assert fx.D.power_law(k=4).params()[:3].tolist() == [1.0, 4.0, 0.0]
    assert fx.D.van_genuchten(n=2).model_id == _lib.VAN_GENUCHTEN
    assert fx.D.brooks_and_corey(n=0.3).model_id == _lib.BROOKS_COREY
    assert fx.D.letxs(Lw=1, Ew=1, Tw=1, Ls=1, Es=1, Ts=1).model_id == _lib.LETXS
    assert fx.D.letd(L=1, E=1, T=1).model_id == _lib.LETD


def test_vectorised_batch_equals_scalar_packing():
    n = np.array([2.0, 3.5, 8.093])
    Ks = np.array([1e-6, 2e-6, 2.079e-6])
    mb = M.VanGenuchten.batch(n=n, Ks=Ks, l=0.5, alpha=np.array([1.0, 2.0, 0.5]), theta_range=(0.005, 0.7))
    assert len(mb) == 3 and mb.model_id.tolist() == [3, 3, 3]
    for j in range(3):
        one = M.VanGenuchten(n=float(n[j]), Ks=float(Ks[j]), l=0.5, alpha=[1.0, 2.0, 0.5][j],
                             theta_range=(0.005, 0.7)).params()
        np.testing.assert_array_equal(mb.params[j], one)
    cat = fx.ModelBatch.concat([mb, fx.ModelBatch.from_models([fx.D.constant(1.0)])])
    assert len(cat) == 4 and cat[3].model_id.tolist() == [0]


# ---- Boltzmann-variable argument parsing (src/frontx/_boltzmann.py:55-98) ----------
class _Probe(AbstractSolution):
    oi = 2.0

    def D(self, theta):
        return np.ones_like(np.asarray(theta, dtype=float))

    @boltzmannmethod
    def __call__(self, o):
        return np.asarray(o, dtype=float) * 1.0

    @boltzmannmethod
    def d_do(self, o):
        return np.ones_like(np.asarray(o, dtype=float))


def test_boltzmannmethod_signatures():
    p = _Probe()
    assert p(3.0) == 3.0
    assert p(4.0, 4.0) == 2.0  # o = r/sqrt(t)
    assert p(o=1.5) == 1.5
    assert p(r=6.0, t=9.0) == 2.0
    with pytest.raises(TypeError, match="takes \\(r, t\\) or \\(o,\\)"):
        p(1.0, 2.0, 3.0)
    with pytest.raises(TypeError):
        p(r=1.0)
    # derived quantities (_boltzmann.py:137-161)
    assert p.d_dr(2.0, 4.0) == 0.5
    assert p.d_dt(2.0, 4.0) == -2.0 / (2.0 * 2 * 4.0)
    assert p.flux(2.0, 4.0) == -0.5
    assert p.sorptivity() == -2.0
    assert (p.b, p.i, p.d_dob) == (0.0, 2.0, 1.0)


def test_results_enum():
    assert fx.RESULTS.successful == 0 and fx.RESULTS.max_steps_reached != fx.RESULTS.successful
