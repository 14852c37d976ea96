"""Developer check: GPU kernels vs the CPU twin (bitwise) and the literal oracle (statistical)."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import torch  # noqa: E402
from conftest import B_PAPER as b  # noqa: E402
from conftest import I_PAPER as i  # noqa: E402
from conftest import mixed_sweep, paper_models, vg_sweep  # noqa: E402

import frontx_b200 as fx  # noqa: E402
from oracle import fxo  # noqa: E402

CNT = ("result", "n_shots", "n_expansions", "n_attempts", "n_rejected", "n_rhs", "n_jac", "n_knots")
print(torch.cuda.get_device_name(0))
pm = paper_models()
for name, m in pm.items():
    th = np.array([0.5, 0.025, b, 0.3, 0.69])
    g = m.derivatives(th)
    c = fxo.twin_D(m.model_id, m.params(), th)
    print(name, "D bitwise equal to twin:", np.array_equal(g, c))
models = fx.ModelBatch.from_models(list(pm.values()) + [fx.D.power_law(k=4), fx.D.philip_exact(), fx.D.constant(1.0)])
bb = np.array([b] * 6 + [1.0, 1.0, 1.0])
ii = np.array([i] * 6 + [0.1, 0.0, 0.0])
sol = fx.solve_batch(models, b=bb, i=ii, k_max=512, throw=False)
tw = fxo.twin_solve_batch(models.model_id, models.params, bb, ii, k_max=512)
print("paper sets: summary bitwise", np.array_equal(sol.summary.cpu().numpy(), tw["summary"]),
      "counters", np.array_equal(sol.counters.cpu().numpy(), tw["counters"]),
      "knots", np.array_equal(sol.knots.cpu().numpy(), tw["knots"]))
for name, mb in (("vg", vg_sweep(4000)), ("mixed", mixed_sweep(4000))):
    t = time.time()
    s = fx.solve_batch(mb, b=b, i=i, k_max=0, throw=False)
    torch.cuda.synchronize()
    tg = time.time() - t
    t = time.time()
    tw = fxo.twin_solve_batch(mb.model_id, mb.params, b, i)
    tt = time.time() - t
    gs, gc = s.summary.cpu().numpy(), s.counters.cpu().numpy()
    eq_s = (gs == tw["summary"]).all(1)
    eq_c = (gc == tw["counters"]).all(1)
    print(name, "gpu s", round(tg, 3), "twin s", round(tt, 2), "summary bitwise frac", eq_s.mean(), "counters frac", eq_c.mean())
    if not eq_s.all():
        j = np.where(~eq_s)[0][:5]
        print("  first mismatches", j, gs[j], tw["summary"][j], gc[j], tw["counters"][j])
for N in (20000, 100000):
    mb = vg_sweep(N)
    mb.to_device()
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.time()
        fx.solve_batch(mb, b=b, i=i, k_max=0, throw=False)
        torch.cuda.synchronize()
        tg = time.time() - t
        print("N", N, "time", round(tg, 4), "solves/s", round(N / tg))
