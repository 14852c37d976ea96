"""Developer measurement of the other BASELINE configs on one GPU (C3 at 10^6 and 10^7, C4 at 10^5);
C2 is bench.py, C5 is scripts/exp_inverse.py."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import bench  # noqa: E402
from conftest import mixed_sweep, paper_models  # noqa: E402
from frontx_b200 import ModelBatch, _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)


def timed(fn, reps=2):
    best = 1e18
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


big = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
for N in (1_000_000, big):
    mb = mixed_sweep(N)
    ids, par = mb.to_device(dev)
    b = torch.full((N,), bench.B_BOUNDARY, dtype=torch.float64, device=dev)
    i = torch.full((N,), bench.I_INITIAL, dtype=torch.float64, device=dev)
    summary = torch.empty((N, _lib.NSUM), dtype=torch.float64, device=dev)
    counters = torch.zeros((N, _lib.NCNT), dtype=torch.int32, device=dev)
    opt = _lib.default_options(itol=bench.ITOL)
    ms = timed(lambda: _lib.check(L.fx_solve_batch(N, ids.data_ptr(), par.data_ptr(), b.data_ptr(), i.data_ptr(), opt,
                                                   summary.data_ptr(), counters.data_ptr(), 0, None,
                                                   stream.cuda_stream)), reps=1 if N > 2_000_000 else 2)
    c = counters.cpu().numpy()
    print(f"C3 mixed Brooks-Corey/LETd, N = {N}: {ms:.0f} ms -> {N / ms * 1e3:.3e} solves/s; "
          f"successful {np.mean(c[:, 0] == 0):.4f}; shots mean {c[:, 1].mean():.2f}; rhs mean {c[:, 5].mean():.0f}")
    del ids, par, b, i, summary, counters, mb

# C4: 10^5 flow rates, the paper's van Genuchten set, cylindrical
N = 100_000
vg = paper_models()["VG(n)"]
rng = np.random.default_rng(2)
Qb = np.exp(rng.uniform(np.log(1e-13), np.log(5e-10), N))
ob, height = 1e-6, 1e-3
q = torch.as_tensor(Qb / (2 * np.pi * ob * height), device=dev)
mb = ModelBatch(np.full(N, vg.model_id, dtype=np.int32), np.tile(vg.params(), (N, 1)))
ids, par = mb.to_device(dev)
i = torch.full((N,), bench.I_INITIAL, dtype=torch.float64, device=dev)
lo = i.clone()
hi = torch.as_tensor(mb.saturation(), device=dev)
summary = torch.empty((N, _lib.NSUM), dtype=torch.float64, device=dev)
counters = torch.zeros((N, _lib.NCNT), dtype=torch.int32, device=dev)
opt = _lib.default_options(itol=bench.ITOL, radial_k=1, ob=ob)
ms = timed(lambda: _lib.check(L.fx_solve_flowrate_batch(N, ids.data_ptr(), par.data_ptr(), q.data_ptr(), i.data_ptr(),
                                                        lo.data_ptr(), hi.data_ptr(), opt, summary.data_ptr(),
                                                        counters.data_ptr(), 0, None, stream.cuda_stream)))
c = counters.cpu().numpy()
s = summary.cpu().numpy()
ok = c[:, 0] == 0
back = -s[ok, 4] * s[ok, 3] * (2 * np.pi * ob * height)
print(f"C4 flow-rate boundary, N = {N}: {ms:.0f} ms -> {N / ms * 1e3:.3e} solves/s; successful {ok.mean():.4f}; "
      f"shots mean {c[:, 1].mean():.2f}; rhs mean {c[:, 5].mean():.0f}; max |Qb recovered / Qb - 1| "
      f"{np.max(np.abs(back / Qb[ok] - 1)):.1e}")
