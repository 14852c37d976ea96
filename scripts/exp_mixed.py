"""Developer experiment: mixed Brooks-Corey / LETd batch (C3 shape) throughput."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
import bench  # noqa: E402
from conftest import mixed_sweep  # noqa: E402
from frontx_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda", 0)
opt = _lib.default_options(itol=bench.ITOL)
stream = torch.cuda.current_stream(dev)
for N in (100_000, 400_000, 1_000_000):
    mb = mixed_sweep(N)
    ids, par = mb.to_device(dev)
    b = torch.full((N,), bench.B_BOUNDARY, dtype=torch.float64, device=dev)
    i = torch.full((N,), bench.I_INITIAL, dtype=torch.float64, device=dev)
    summary = torch.empty((N, _lib.NSUM), dtype=torch.float64, device=dev)
    counters = torch.zeros((N, _lib.NCNT), dtype=torch.int32, device=dev)
    best = 1e9
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(L.fx_solve_batch(N, ids.data_ptr(), par.data_ptr(), b.data_ptr(), i.data_ptr(), opt,
                                    summary.data_ptr(), counters.data_ptr(), 0, None, stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    c = counters.cpu().numpy()
    print(f"mixed BC/LETd N={N}: {best:.1f} ms -> {N / best * 1e3:.0f} solves/s; ok {np.mean(c[:,0]==0):.4f}; "
          f"shots mean {c[:,1].mean():.1f}; rhs mean {c[:,5].mean():.0f}")
