"""Radial problems with a flow-rate boundary (north-star config C4).

``solve_flowrate`` is NOT in the reference: its semantics are those of
``fronts.solve_flowrate`` built on the reference's own algorithm -- the same Kvaerno5 shots, the
same event, the same bisection, run on the boundary value b with d_dob = -Qb/(D(b) angle ob height).
Parity is therefore UNPINNED against the reference; what is checked:
  * CPU: the twin against a bisection on b driven from Python over the LITERAL oracle's own
    shot function (oracle/fx_oracle.c), and the physics (the imposed flow rate is recovered
    from the solution, b grows with Qb);
  * GPU: the kernel against the twin, bit for bit, and the Python surface.
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import pytest
from conftest import I_PAPER, paper_models

OB, HEIGHT, ANGLE = 1e-6, 1e-3, 2 * np.pi
AREA = ANGLE * OB * HEIGHT


def c4_workload(n: int, seed: int = 2):
    """C4: the paper's van Genuchten set, cylindrical, ob = 1e-6, angle 2 pi, height 1e-3 and
    Qb ~ logU[1e-13, 5e-10] -- the range over which a boundary value below saturation can carry
    the imposed flow (SURVEY.md section 8d's [1e-9, 1e-7] exceeds what this soil takes at ob = 1e-6:
    the residual is negative on the whole bracket and the solve reports FX_NO_BRACKET)."""
    from frontx_b200 import ModelBatch

    vg = paper_models()["VG(n)"]
    rng = np.random.default_rng(seed)
    Qb = np.exp(rng.uniform(np.log(1e-13), np.log(5e-10), n))
    mb = ModelBatch(np.full(n, vg.model_id, dtype=np.int32), np.tile(vg.params(), (n, 1)))
    return mb, Qb


def literal_flowrate(fxo, model, params, q, i, b_lo, b_hi, itol=1e-3, max_bisect=100):
    """Bisection on b (src/frontx/_forward.py:104-112 restated) over the literal oracle's shots."""
    p = fxo.pad_params(params)
    opt = fxo.default_options(radial_k=1, ob=OB)
    cnt = np.zeros(8, dtype=np.int32)
    nk = C.c_int(0)

    def shot(b):
        D_b = fxo.D(model, p, b)[0, 0]
        return fxo.lib().fxo_shoot(model, fxo._dp(p), float(b), float(i), float(itol), float(-q / D_b),
                                   C.byref(opt), 0, None, C.byref(nk), fxo._ip(cnt))

    r_lo, r_hi = shot(b_lo), shot(b_hi)
    shots = 2
    if (r_lo < 0) == (r_hi < 0):
        return 2, b_hi, shots
    lo_neg, lower, upper = r_lo < 0, b_lo, b_hi
    for _ in range(max_bisect):
        mid = lower + 0.5 * (upper - lower)
        res = shot(mid)
        shots += 1
        if (res < 0) == lo_neg:
            lower = mid
        else:
            upper = mid
        if abs(res) < itol:
            return 0, mid, shots
    return 1, mid, shots


def test_twin_flowrate_physics_and_literal_oracle(fxo):
    mb, Qb = c4_workload(24)
    q = Qb / AREA
    hi = mb.saturation()
    tw = fxo.twin_solve_flowrate_batch(mb.model_id, mb.params, q, I_PAPER, I_PAPER, hi, radial_k=1, ob=OB)
    s, c = tw["summary"], tw["counters"]
    assert (c[:, 0] == 0).all()
    # the imposed flow rate is what the solution carries through the inlet: Qb = -D(b) d_dob angle ob height
    np.testing.assert_allclose(-s[:, 4] * s[:, 3] * AREA, Qb, rtol=1e-14)
    # far field met within itol; b strictly inside the bracket and increasing with Qb
    assert np.all(np.abs(s[:, 2] - I_PAPER) < 1e-3)
    order = np.argsort(Qb)
    assert np.all(np.diff(s[order, 0]) > 0) and np.all(s[:, 0] > I_PAPER) and np.all(s[:, 0] < hi)
    # independent driver over the literal oracle's shot: same bisection path, same b
    same = 0
    for j in range(8):
        st, b_lit, shots = literal_flowrate(fxo, int(mb.model_id[j]), mb.params[j], q[j], I_PAPER, I_PAPER, hi[j])
        assert st == 0
        assert abs(b_lit / s[j, 0] - 1) < 1e-6
        same += int(shots == c[j, 1] and b_lit == s[j, 0])
    assert same >= 7  # the same midpoints, bit for bit, in (nearly) every case


def test_twin_flowrate_reports_an_impossible_flow(fxo):
    """A flow rate no boundary value below saturation can carry: both ends of the bracket overshoot."""
    mb, _ = c4_workload(2)
    q = np.array([1e-7, 1e-8]) / AREA
    tw = fxo.twin_solve_flowrate_batch(mb.model_id, mb.params, q, I_PAPER, I_PAPER, mb.saturation(), ob=OB)
    assert (tw["counters"][:, 0] == 2).all() and (tw["counters"][:, 1] == 2).all()


@pytest.mark.gpu
def test_gpu_flowrate_bitwise_vs_twin(fxo):
    import torch

    import frontx_b200 as fx
    from frontx_b200 import _lib

    for n, k_max in ((3000, 0), (64, 512)):
        mb, Qb = c4_workload(n)
        Qb[::17] = 3e-8  # some impossible flows in the batch
        q = Qb / AREA
        hi = mb.saturation()
        tw = fxo.twin_solve_flowrate_batch(mb.model_id, mb.params, q, I_PAPER, I_PAPER, hi, k_max=k_max, ob=OB)
        dev = torch.device("cuda", 0)
        ids, par = mb.to_device(dev)
        t = lambda a: torch.as_tensor(np.broadcast_to(a, (n,)).copy(), dtype=torch.float64, device=dev)  # noqa: E731
        q_t, i_t, lo_t, hi_t = t(q), t(I_PAPER), t(I_PAPER), t(hi)
        summary = torch.empty((n, 8), dtype=torch.float64, device=dev)
        counters = torch.zeros((n, 8), dtype=torch.int32, device=dev)
        knots = torch.zeros((n, k_max, 4), dtype=torch.float64, device=dev) if k_max else None
        opt = _lib.default_options(radial_k=1, ob=OB)
        _lib.check(_lib.lib().fx_solve_flowrate_batch(n, ids.data_ptr(), par.data_ptr(), q_t.data_ptr(),
                                                      i_t.data_ptr(), lo_t.data_ptr(), hi_t.data_ptr(), opt,
                                                      summary.data_ptr(), counters.data_ptr(), k_max,
                                                      knots.data_ptr() if k_max else None,
                                                      torch.cuda.current_stream(dev).cuda_stream))
        torch.cuda.synchronize()
        np.testing.assert_array_equal(counters.cpu().numpy(), tw["counters"])
        ok = tw["counters"][:, 0] == 0
        assert ok.sum() > 0.9 * n and (tw["counters"][::17, 0] == 2).all()
        np.testing.assert_array_equal(summary.cpu().numpy()[ok], tw["summary"][ok])
        np.testing.assert_array_equal(summary.cpu().numpy()[~ok][:, [0, 1, 2, 5, 6, 7]],
                                      tw["summary"][~ok][:, [0, 1, 2, 5, 6, 7]])
        if k_max:  # the rows that are the last shot's (speculative shots write no knots; the accepted one is re-run)
            g, w, nk = knots.cpu().numpy(), tw["knots"], tw["counters"][:, 7]
            valid = np.arange(k_max)[None, :] < np.minimum(nk, k_max)[:, None]
            np.testing.assert_array_equal(g[valid], w[valid])
    assert fx.solve_flowrate is not None


@pytest.mark.gpu
def test_gpu_solve_flowrate_surface():
    import frontx_b200 as fx

    vg = paper_models()["VG(n)"]
    Qb = 2e-11
    sol = fx.solve_flowrate(vg, Qb=Qb, i=I_PAPER, radial="cylindrical", ob=OB, height=HEIGHT)
    assert sol.result == fx.RESULTS.successful
    assert I_PAPER < sol.b < 0.7
    assert abs(sol.i - I_PAPER) < 1e-3
    # the profile starts at (ob, b) with the gradient the flow rate imposes
    assert sol(OB) == pytest.approx(sol.b, rel=1e-15)
    assert -float(sol.D(sol.b)) * sol.d_dob * AREA == pytest.approx(Qb, rel=1e-13)
    assert sol.d_do(OB) == pytest.approx(sol.d_dob, rel=1e-15)
    th = sol(np.linspace(OB, sol.oi, 50))
    assert np.all(np.diff(th) <= 1e-12)  # wetting from the inlet: theta decreases outwards
    batch = fx.solve_flowrate_batch([vg, vg], Qb=[1e-12, 1e-10], i=I_PAPER, ob=OB, height=HEIGHT, k_max=0)
    b = batch.b.cpu().numpy()
    assert b[0] < b[1]
    with pytest.raises(fx.SolveError):
        fx.solve_flowrate(vg, Qb=1e-7, i=I_PAPER, ob=OB, height=HEIGHT)
    with pytest.raises(ValueError):
        fx.solve_flowrate(fx.D.power_law(k=4), Qb=1e-3, i=0.1, ob=OB, height=HEIGHT)
