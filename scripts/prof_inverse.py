import sys, torch
from pathlib import Path
sys.path.insert(0, "/root/repo")
from frontx_b200 import _lib
L = _lib.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
batch, npts = 200, 1_000_000
a = 0.5 + 1.5 * torch.rand((batch, 1), dtype=torch.float64, device=dev)
o = (torch.linspace(0, 20, npts, dtype=torch.float64, device=dev).unsqueeze(0) / a).contiguous()
theta = torch.exp(-a * o)
D = torch.empty_like(o)
sorp = torch.empty((batch, 4), dtype=torch.float64, device=dev)
status = torch.empty((batch,), dtype=torch.int32, device=dev)
for ws in (False, True):
    _lib.check(L.fx_inverse_batch(batch, npts, o.data_ptr(), theta.data_ptr(), D.data_ptr(),
                                  sorp.data_ptr() if ws else None, status.data_ptr(), None, None, stream.cuda_stream))
torch.cuda.synchronize()
print("ok")
