"""Developer measurement: the inverse scan kernel (BASELINE config C5: 10^3 profiles x 10^6 points)
against the measured HBM copy bandwidth.  Algorithmic bytes: 16 B/point read + 8 B/point written."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from frontx_b200 import _lib  # noqa: E402

import os
if os.environ.get("FX_LIB"):
    _lib.LIB_PATH = Path(os.environ["FX_LIB"]).resolve()
L = _lib.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
peaks = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
peak = json.loads(peaks.read_text())["hbm_gbs"] if peaks.exists() else 6527.5
for batch, npts in ((1000, 1_000_000), (100, 1_000_000), (1000, 100_000)):
    a = 0.5 + 1.5 * torch.rand((batch, 1), dtype=torch.float64, device=dev)
    o = torch.linspace(0, 20, npts, dtype=torch.float64, device=dev).unsqueeze(0) / a
    theta = torch.exp(-a * o)
    o = o.contiguous()
    D = torch.empty_like(o)
    sorp = torch.empty((batch, 4), dtype=torch.float64, device=dev)
    status = torch.empty((batch,), dtype=torch.int32, device=dev)
    best = {}
    for with_sorp in (True, False):  # with the sorptivity integrals of the forward interpolant, and D alone
        best[with_sorp] = 1e9
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _lib.check(L.fx_inverse_batch(batch, npts, o.data_ptr(), theta.data_ptr(), D.data_ptr(),
                                          sorp.data_ptr() if with_sorp else None, status.data_ptr(), None, None,
                                          stream.cuda_stream))
            e1.record(stream)
            torch.cuda.synchronize()
            best[with_sorp] = min(best[with_sorp], e0.elapsed_time(e1))
    print(f"inverse {batch} x {npts} with sorptivity sums: {best[True]:.2f} ms = "
          f"{batch * npts * 24 / 1e6 / best[True] / peak * 100:.1f} % of the copy bandwidth")
    best = best[False]
    gb = batch * npts * 24 / 1e9
    want = (1 - torch.log(theta)) / (2 * a * a)
    sel = slice(npts // 100, npts // 2)
    err = float(((D[:, sel] - want[:, sel]).abs() / want[:, sel]).max())
    print(f"inverse {batch} x {npts}: {best:.2f} ms, {gb / best * 1e3:.0f} GB/s algorithmic = "
          f"{gb / best * 1e3 / peak * 100:.1f} % of the measured {peak} GB/s copy bandwidth; status ok "
          f"{int((status == 0).sum())}/{batch}; max rel. deviation from the analytic D {err:.2e}")
    del o, theta, D, want
    torch.cuda.empty_cache()
