"""One-off check at the benchmark's full size: every one of the 10^5 C2 problems, GPU (shot queue,
speculative bisection) against the CPU twin, bit for bit; plus the flow-rate boundary at 2*10^4."""
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import bench  # noqa: E402
from conftest import mixed_sweep  # noqa: E402

import frontx_b200 as fx  # noqa: E402
from oracle import fxo  # noqa: E402

for name, mb in (("C2 van Genuchten 10^5", bench.c2_workload(100_000, seed=0)), ("C3 mixed 5*10^4", mixed_sweep(50_000))):
    t = time.time()
    tw = fxo.twin_solve_batch(mb.model_id, mb.params, bench.B_BOUNDARY, bench.I_INITIAL)
    t_twin = time.time() - t
    sol = fx.solve_batch(mb, b=bench.B_BOUNDARY, i=bench.I_INITIAL, k_max=0, throw=False)
    s, c = sol.summary.cpu().numpy(), sol.counters.cpu().numpy()
    eq_s = (s.view(np.int64) == tw["summary"].view(np.int64)).all(axis=1)
    eq_c = (c == tw["counters"]).all(axis=1)
    print(f"{name}: summaries bitwise equal {eq_s.sum()}/{len(eq_s)}, counters equal {eq_c.sum()}/{len(eq_c)} "
          f"(twin {t_twin:.0f} s on {fxo.num_threads()} threads)")
