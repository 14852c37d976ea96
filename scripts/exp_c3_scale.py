"""Developer experiment: config C3 (or C2 with FX_CFG=C2) at several sizes through fx_solve_batch of a given library
(FX_LIB), raw ctypes (works with round 1's library too)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from frontx_b200 import _lib  # noqa: E402

L = C.CDLL(os.environ.get("FX_LIB", str(_lib.LIB_PATH)))
vp = C.c_void_p
L.fx_solve_batch.argtypes = [C.c_int64, vp, vp, vp, vp, C.POINTER(_lib.SolveOptions), vp, vp, C.c_int32, vp, vp]
L.fx_default_solve_options.argtypes = [C.POINTER(_lib.SolveOptions)]
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
tag = os.environ.get("FX_VARIANT", "in-tree")
cfg = os.environ.get("FX_CFG", "C3")
reps = int(os.environ.get("FX_REPS", "2"))
for n in [int(x) for x in sys.argv[1:]] or [1_000_000, 4_000_000, 10_000_000]:
    models = bench.c3_rows(np.arange(n)) if cfg == 'C3' else bench.c2_workload(n, seed=0)
    ids, par = models.to_device(dev)
    if os.environ.get('FX_VGFORM') == '1':
        par[:, 7] = 1.0  # van Genuchten, cancellation-free form (what round 1's library always computed)
    b = torch.full((n,), bench.B_BOUNDARY, dtype=torch.float64, device=dev)
    i = torch.full((n,), bench.I_INITIAL, dtype=torch.float64, device=dev)
    summary = torch.empty((n, 8), dtype=torch.float64, device=dev)
    counters = torch.zeros((n, 8), dtype=torch.int32, device=dev)
    opt = _lib.SolveOptions()
    L.fx_default_solve_options(C.byref(opt))
    opt.model_mask = models.mask() if os.environ.get("FX_MASK", "1") == "1" else 0
    best = 1e18
    for rep in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        rc = L.fx_solve_batch(n, ids.data_ptr(), par.data_ptr(), b.data_ptr(), i.data_ptr(), C.byref(opt),
                              summary.data_ptr(), counters.data_ptr(), 0, None, stream.cuda_stream)
        assert rc == 0
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    c = counters.cpu().numpy()
    print(f"{tag} {cfg} n={n}: {best:9.1f} ms {n / best / 1e3:6.3f} Msolves/s fail {int((c[:, 0] != 0).sum())} "
          f"rhs mean {c[:, 5].mean():.0f}", flush=True)
    del ids, par, b, i, summary, counters, models
    torch.cuda.empty_cache()
