"""CPU-only: the kernel's arithmetic (host build, oracle/fx_twin.cpp) against the
independent literal oracle (oracle/fx_oracle.c) and against libm / mpmath.

The GPU tests demand bit-for-bit equality with the twin; these tests are what ties
the twin -- i.e. the product's fx_math.cuh / fx_models.cuh / fx_arith.cuh -- to the
reference algorithm as restated literally, and they measure how much of the 1e-6
criterion survives a change of pow() implementation (the situation of ANY two
implementations of the reference, the reference on two backends included).
"""

from __future__ import annotations

import json
from pathlib import Path

import numpy as np
import pytest
from conftest import B_PAPER, I_PAPER, mixed_sweep, vg_sweep

GOLDEN = json.loads((Path(__file__).parent / "golden" / "diffusivity_golden.json").read_text())


def ulps(got, want):
    return np.max(np.abs(got - want) / np.spacing(np.abs(want)))


def test_elementary_functions_vs_libm(fxo):
    T = fxo.twin()
    rng = np.random.default_rng(0)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 4000)), 1 + rng.uniform(-1e-6, 1e-6, 500),
                        rng.uniform(0.5, 2, 1500), [5e-324, 2.2e-308, 1.0, 1.7e308]])
    assert ulps(np.array([T.fxt_log(float(v)) for v in x]), np.log(x)) <= 1.0
    y = np.concatenate([rng.uniform(-700, 700, 4000), rng.uniform(-1, 1, 1500), rng.uniform(-1e-8, 1e-8, 500)])
    assert ulps(np.array([T.fxt_exp(float(v)) for v in y]), np.exp(y)) <= 1.0
    assert ulps(np.array([T.fxt_expm1(float(v)) for v in y]), np.expm1(y)) <= 4.0
    # special values: pow()-like domain behaviour the solver relies on
    assert T.fxt_log(0.0) == -np.inf and np.isnan(T.fxt_log(-1.0)) and T.fxt_log(np.inf) == np.inf
    assert np.isnan(T.fxt_log(np.nan)) and np.isnan(T.fxt_exp(np.nan))
    assert T.fxt_exp(-np.inf) == 0.0 and T.fxt_exp(np.inf) == np.inf and T.fxt_exp(710.0) == np.inf
    assert T.fxt_exp(-745.0) == np.exp(-745.0) and T.fxt_exp(0.0) == 1.0 and T.fxt_expm1(-np.inf) == -1.0


@pytest.mark.parametrize("name", list(GOLDEN["sets"]))
def test_kernel_diffusivity_vs_golden_and_literal(fxo, paper, name):
    """The product's closed forms (host build) vs mpmath golden values and vs the jets."""
    m = paper[name]
    for pname, row in GOLDEN["sets"][name]["values"].items():
        got = fxo.twin_D(m.model_id, m.params(), [row["theta"]])[0]
        want = np.array([float(row["D"]), float(row["D1"]), float(row["D2"])])
        rtol = 3e-9 if pname.startswith("b=") else 5e-12
        np.testing.assert_allclose(got, want, rtol=rtol, err_msg=f"{name} @ {pname}")
    tr, ts = m.theta_range
    th = tr + (ts - tr) * np.concatenate([np.geomspace(1e-4, 0.5, 60), 1 - np.geomspace(2e-7, 0.5, 60)])
    np.testing.assert_allclose(fxo.twin_D(m.model_id, m.params(), th), fxo.D(m.model_id, m.params(), th),
                               rtol=2e-11, err_msg=name)


def test_paper_sets_twin_vs_literal(fxo, paper):
    ms = list(paper.values())
    ids = np.array([m.model_id for m in ms], dtype=np.int32)
    par = np.stack([m.params() for m in ms])
    tw = fxo.twin_solve_batch(ids, par, B_PAPER, I_PAPER, k_max=512)
    lit = fxo.solve_batch(ids, par, B_PAPER, I_PAPER, k_max=512)
    seq = [0, 1, 2, 3, 4, 6, 7]  # status, shots, expansions, attempts, rejected, Jacobians, knots
    # The accept/reject sequence of a shot is chaotic in the last bits of D (module docstring): the
    # two implementations take the same decisions in all but at most one of the six sets, and the
    # bisection (shots, status), the accepted d_dob and the final shot's knot count agree in all six.
    same = (tw["counters"][:, seq] == lit["counters"][:, seq]).all(axis=1)
    assert same.sum() >= len(ms) - 1
    np.testing.assert_array_equal(tw["counters"][:, [0, 1, 2, 7]], lit["counters"][:, [0, 1, 2, 7]])
    np.testing.assert_allclose(tw["summary"][:, 0], lit["summary"][:, 0], rtol=1e-13)  # d_dob
    np.testing.assert_allclose(tw["summary"][:, 3], lit["summary"][:, 3], rtol=1e-9)   # sorptivity
    np.testing.assert_allclose(tw["summary"][:, 2], lit["summary"][:, 2], rtol=1e-6)   # theta(oi)
    for j in range(len(ms)):
        nk = int(lit["counters"][j, 7])
        o = np.linspace(0, lit["summary"][j, 1], 200)
        th_l, _ = fxo.evaluate(lit["knots"][j, :nk], o)
        k4 = tw["knots"][j, :nk]
        k5 = np.concatenate([k4[:, :3], k4[:, 2:3], k4[:, 3:4]], axis=1)  # (o, th, p, F0=p, F1)
        th_t, _ = fxo.evaluate(k5, o)
        np.testing.assert_allclose(th_t, th_l, rtol=1e-6)


@pytest.mark.parametrize("which", ["vg", "mixed"])
def test_sweep_twin_vs_literal_statistics(fxo, which):
    """How much of the 1e-6 criterion survives a different pow()/fma: reported and bounded."""
    mb = vg_sweep(1200) if which == "vg" else mixed_sweep(1200)
    tw = fxo.twin_solve_batch(mb.model_id, mb.params, B_PAPER, I_PAPER)
    lit = fxo.solve_batch(mb.model_id, mb.params, B_PAPER, I_PAPER)
    seq = [0, 1, 2, 3, 4, 6, 7]
    same = (tw["counters"][:, seq] == lit["counters"][:, seq]).all(axis=1)
    rel_S = np.abs(tw["summary"][:, 3] / lit["summary"][:, 3] - 1)
    rel_i = np.abs(tw["summary"][:, 2] / lit["summary"][:, 2] - 1)
    print(f"\n{which}: same step sequence {same.mean():.4f}; sorptivity within 1e-6 {(rel_S < 1e-6).mean():.4f}; "
          f"theta(oi) within 1e-6 {(rel_i < 1e-6).mean():.4f}, within 1e-3 {(rel_i < 1e-3).mean():.4f}")
    assert (tw["counters"][:, 0] == lit["counters"][:, 0]).mean() >= 0.995  # same status
    assert same.mean() >= 0.95
    assert (rel_S < 1e-6).mean() >= 0.99
    assert (rel_i < 1e-6).mean() >= (0.90 if which == "vg" else 0.97)
    assert (rel_i < 1e-3).mean() >= 0.995
