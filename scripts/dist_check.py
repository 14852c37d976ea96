"""Developer check under torchrun (NCCL): the sharded solve and the sharded inverse give, on every rank,
what one GPU gives for the whole batch.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py"""
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tests"))
from conftest import B_PAPER, I_PAPER, mixed_sweep  # noqa: E402

import frontx_b200 as fx  # noqa: E402
from frontx_b200 import _dist  # noqa: E402

rank, world, local = _dist.init_from_env()
models = mixed_sweep(20_000, seed=3)
summary, counters, mine = _dist.solve_batch_distributed(models, b=B_PAPER, i=I_PAPER)
single = fx.solve_batch(models, b=B_PAPER, i=I_PAPER, k_max=0, throw=False)
ok_solve = torch.equal(summary, single.summary) and torch.equal(counters, single.counters)
a = np.linspace(0.5, 2.0, 11)
o = np.linspace(0, 20, 50_000)[None, :] / a[:, None]
theta = np.exp(-a[:, None] * o)
D, sorp, status, sl = _dist.inverse_distributed(o, theta, gather_D=True)
one = fx.inverse(o, theta)
ok_inv = torch.equal(D, one.D_at_knots) and torch.equal(sorp, one.sorptivity()) and torch.equal(status, one.status)
print(f"rank {rank}/{world}: solve shard {mine.size} problems, gathered == single GPU: {ok_solve}; "
      f"inverse profiles {sl.start}:{sl.stop}, gathered == single GPU: {ok_inv}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
assert ok_solve and ok_inv
