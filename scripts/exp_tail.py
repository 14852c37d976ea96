"""Developer experiment: where does the C2 kernel time go (bulk rounds vs failing-problem tail)?"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from frontx_b200 import _lib  # noqa: E402

import os
if os.environ.get("FX_LIB"):
    _lib.LIB_PATH = Path(os.environ["FX_LIB"]).resolve()
L = _lib.lib()
dev = torch.device("cuda", 0)
N = 100_000
models = bench.c2_workload(N, seed=0)
opt = _lib.default_options(itol=bench.ITOL)
stream = torch.cuda.current_stream(dev)


def run(ids_np, par_np, reps=3):
    n = len(ids_np)
    ids = torch.from_numpy(np.ascontiguousarray(ids_np)).to(dev)
    par = torch.from_numpy(np.ascontiguousarray(par_np)).to(dev)
    b = torch.full((n,), bench.B_BOUNDARY, dtype=torch.float64, device=dev)
    i = torch.full((n,), bench.I_INITIAL, dtype=torch.float64, device=dev)
    summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
    counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(L.fx_solve_batch(n, ids.data_ptr(), par.data_ptr(), b.data_ptr(), i.data_ptr(), opt,
                                    summary.data_ptr(), counters.data_ptr(), 0, None, stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, counters.cpu().numpy()


t_all, cnt = run(models.model_id, models.params)
ok = cnt[:, 0] == 0
print(f"all {N}: {t_all:.1f} ms; failing {int((~ok).sum())}; rhs mean {cnt[:,5].mean():.0f} max {cnt[:,5].max()}")
t_ok, _ = run(models.model_id[ok], models.params[ok])
print(f"successes only ({int(ok.sum())}): {t_ok:.1f} ms")
big = bench.c2_workload(400_000, seed=5)
for n in (148 * 128 * 2, 148 * 128 * 4, 148 * 128 * 8, 148 * 128 * 16):
    t, c = run(big.model_id[:n], big.params[:n], reps=2)
    print(f"n = {n} ({n / (148 * 4 * 32):.1f} warps/SMSP of work): {t:.1f} ms -> {n / t * 1e3:.0f} solves/s; "
          f"rhs mean {c[:,5].mean():.0f} max {c[:,5].max()}")
