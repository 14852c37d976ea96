"""Developer experiment: small-batch latency of fx_solve_batch (config C1 = one power-law solve with dense
output, and a 64-candidate generation of `fit`), library given by FX_LIB (default: the in-tree build)."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from frontx_b200 import _lib  # noqa: E402
from frontx_b200 import models as M  # noqa: E402

if os.environ.get("FX_LIB"):
    _lib.LIB_PATH = Path(os.environ["FX_LIB"]).resolve()
L = _lib.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
use_mask = os.environ.get("FX_MASK", "1") == "1"


def timeit(models, b, i, k_max, reps=12):
    n = len(models)
    ids, par = models.to_device(dev)
    bt = torch.full((n,), b, dtype=torch.float64, device=dev)
    it = torch.full((n,), i, dtype=torch.float64, device=dev)
    summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
    counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
    knots = torch.zeros((n, max(k_max, 1), 4), dtype=torch.float64, device=dev)
    mask = int(np.bitwise_or.reduce(1 << models.model_id.astype(np.int64))) if use_mask else 0
    opt = _lib.default_options(itol=1e-3, model_mask=mask)
    dev_ms, wall_ms = [], []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(L.fx_solve_batch(n, ids.data_ptr(), par.data_ptr(), bt.data_ptr(), it.data_ptr(), opt,
                                    summary.data_ptr(), counters.data_ptr(), k_max,
                                    knots.data_ptr() if k_max else None, stream.cuda_stream))
        e1.record(stream)
        torch.cuda.synchronize()
        wall_ms.append((time.perf_counter() - t0) * 1e3)
        dev_ms.append(e0.elapsed_time(e1))
    c = counters.cpu().numpy()
    return np.median(dev_ms[2:]), np.median(wall_ms[2:]), c[:, 1].mean(), int((c[:, 0] != 0).sum()), c[:, 7].max()


tag = os.environ.get("FX_VARIANT", "in-tree")
c1 = M.ModelBatch.from_models([M.PowerLaw(k=4.0)])
for k_max in (512, 0):
    d, w, shots, bad, nk = timeit(c1, 1.0, 0.1, k_max)
    print(f"{tag}: C1 single solve k_max={k_max}: device {d:.3f} ms, wall {w:.3f} ms, shots {shots:.0f}, failed {bad}, knots {nk}")
vg = bench.c2_workload(64, seed=7)
for k_max in (512, 0):
    d, w, shots, bad, nk = timeit(vg, bench.B_BOUNDARY, bench.I_INITIAL, k_max)
    print(f"{tag}: 64 van Genuchten candidates k_max={k_max}: device {d:.3f} ms, wall {w:.3f} ms, mean shots {shots:.1f}, failed {bad}, knots {nk}")
