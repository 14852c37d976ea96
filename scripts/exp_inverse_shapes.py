import sys, torch
sys.path.insert(0, "/root/repo")
from frontx_b200 import _lib
L = _lib.lib(); dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(dev)
for batch, npts in ((500, 2_000_000), (2000, 500_000), (50, 10_000_000), (10000, 100_000), (1000, 999_999), (100_000, 10_000), (4, 100_000_000)):
    a = 0.5 + 1.5 * torch.rand((batch, 1), dtype=torch.float64, device=dev)
    o = (torch.linspace(0, 20, npts, dtype=torch.float64, device=dev).unsqueeze(0) / a).contiguous()
    theta = torch.exp(-a * o)
    D = torch.empty_like(o); sorp = torch.empty((batch, 4), dtype=torch.float64, device=dev)
    status = torch.empty((batch,), dtype=torch.int32, device=dev)
    res = {}
    for ws in (True, False):
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _lib.check(L.fx_inverse_batch(batch, npts, o.data_ptr(), theta.data_ptr(), D.data_ptr(), sorp.data_ptr() if ws else None, status.data_ptr(), None, None, stream.cuda_stream))
            e1.record(stream); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[ws] = best
    gb = batch * npts * 24 / 1e6
    print(f"{batch} x {npts}: D alone {res[False]:.2f} ms = {gb/res[False]/6527.5*100:.1f} %; with sorptivity {res[True]:.2f} ms = {gb/res[True]/6527.5*100:.1f} %; ok {int((status==0).sum())}/{batch}", flush=True)
    del o, theta, D, a
    torch.cuda.empty_cache()
