"""Profiling driver: BASELINE config C2 (10^5 van Genuchten problems) through fx_solve_batch, twice (the second
launch of shoot_bisect_kernel is the one to capture: `ncu -k regex:shoot_bisect --launch-skip 1 --launch-count 1`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from frontx_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
models = bench.c2_workload(n, seed=0)
ids, par = models.to_device(dev)
b = torch.full((n,), bench.B_BOUNDARY, dtype=torch.float64, device=dev)
i = torch.full((n,), bench.I_INITIAL, dtype=torch.float64, device=dev)
summary = torch.empty((n, _lib.NSUM), dtype=torch.float64, device=dev)
counters = torch.zeros((n, _lib.NCNT), dtype=torch.int32, device=dev)
opt = _lib.default_options(itol=bench.ITOL, model_mask=models.mask())
for _ in range(2):
    _lib.check(L.fx_solve_batch(n, ids.data_ptr(), par.data_ptr(), b.data_ptr(), i.data_ptr(), opt,
                                summary.data_ptr(), counters.data_ptr(), 0, None, stream.cuda_stream))
    torch.cuda.synchronize()
print("ok", int((counters[:, 0] == 0).sum()))
