"""Developer experiment: finish-time distribution of the problems of one launch (debug build, -DFX_DEBUG_TIMES into
scripts/_dbg/libfrontx_b200_dbg.so): python scripts/exp_timeline.py N [C2|C3]."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from frontx_b200 import _lib  # noqa: E402

_lib.LIB_PATH = Path(__file__).resolve().parent / "_dbg" / "libfrontx_b200_dbg.so"
L = _lib.lib()
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
CFG = sys.argv[2] if len(sys.argv) > 2 else "C2"
models = bench.c2_workload(N, seed=0) if CFG == "C2" else bench.c3_rows(np.arange(N))
opt = _lib.default_options(itol=bench.ITOL, model_mask=models.mask())
stream = torch.cuda.current_stream(dev)
ids, par = models.to_device(dev)
b = torch.full((N,), bench.B_BOUNDARY, dtype=torch.float64, device=dev)
i = torch.full((N,), bench.I_INITIAL, dtype=torch.float64, device=dev)
summary = torch.empty((N, _lib.NSUM), dtype=torch.float64, device=dev)
counters = torch.zeros((N, _lib.NCNT), dtype=torch.int32, device=dev)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    _lib.check(L.fx_solve_batch(N, ids.data_ptr(), par.data_ptr(), b.data_ptr(), i.data_ptr(), opt,
                                summary.data_ptr(), counters.data_ptr(), 0, None, stream.cuda_stream))
    e1.record(stream)
    torch.cuda.synchronize()
print("kernel+setup ms", e0.elapsed_time(e1))
t = summary[:, 5].cpu().numpy()
c = counters.cpu().numpy()
t = (t - t.min()) * 1e-6  # ms after the first finish
order = np.argsort(t)
print("finish time (ms after first finish) percentiles:")
for q in (1, 10, 25, 50, 75, 90, 99, 99.5, 99.9, 100):
    print(f"  p{q}: {np.percentile(t, q):.2f}")
ok = c[:, 0] == 0
print("last regular problem finishes at", t[ok].max(), "failing: first", t[~ok].min() if (~ok).any() else None, "last", t[~ok].max() if (~ok).any() else None)
last = order[-10:]
print("last 10: shots", c[last, 1], "rhs", c[last, 5], "status", c[last, 0])
for lo, hi in ((0, 10), (10, 20), (20, 30), (30, 40), (40, 60), (60, 80), (80, 120)):
    m = (t >= lo) & (t < hi)
    print(f"  finish in [{lo},{hi}) ms: {int(m.sum())} problems, mean shots {c[m, 1].mean() if m.any() else 0:.1f}")

for m in np.unique(models.model_id):
    sel = models.model_id == m
    tm = t[sel]
    print(f"model {m}: {int(sel.sum())} problems, first finish {tm.min():.1f} ms, p50 {np.percentile(tm, 50):.1f}, p99 {np.percentile(tm, 99):.1f}, "
          f"p99.9 {np.percentile(tm, 99.9):.1f}, last {tm.max():.1f}; failing {int((c[sel, 0] != 0).sum())}")
