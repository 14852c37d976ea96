"""Small end-to-end run of every kernel (developer tool): a quick look that each entry point still
produces sane numbers after a change; small enough for `compute-sanitizer --tool memcheck` where
that tool is available (it is closed on the measurement pool)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from conftest import B_PAPER, I_PAPER, mixed_sweep, paper_models, vg_sweep  # noqa: E402

import frontx_b200 as fx  # noqa: E402

pm = paper_models()
sol = fx.solve_batch(list(pm.values()), b=B_PAPER, i=I_PAPER, k_max=256)
print("paper sets: shots", sol.n_shots.tolist())
s2 = fx.solve_batch(vg_sweep(300), b=B_PAPER, i=I_PAPER, k_max=0, throw=False)
s3 = fx.solve_batch(mixed_sweep(300), b=B_PAPER, i=I_PAPER, k_max=0, throw=False)
print("sweeps ok", int((s2.result == 0).sum()), int((s3.result == 0).sum()))
fl = fx.solve_flowrate_batch([pm["VG(n)"]] * 8, Qb=np.geomspace(1e-13, 1e-10, 8), i=I_PAPER, ob=1e-6, height=1e-3,
                             k_max=0)
print("flowrate b", fl.b.tolist()[:3])
th, dth = sol.evaluate(np.linspace(0, 0.002, 50))
print("eval", float(th[0, 0]), float(dth[0, 0]))
o = np.linspace(0, 20, 5000)
inv = fx.inverse(np.stack([o, o / 1.5]), np.stack([np.exp(-o), np.exp(-o)]))
print("inverse", float(inv.D_at_knots[0, 100]), inv.status.tolist())
one = fx.solve(pm["LETd#1"], b=B_PAPER, i=I_PAPER)
oo = np.linspace(0, one.oi, 80)
sc = fx.ScaledSolution.fitting_data(fx.ScaledSolution(one, D0=2.0), oo, one(oo))
print("scaled D0", sc.D0)
