"""The N>1 path on CPU: block-cyclic sharding + the all-gather of result records,
world_size 2 over gloo.  The local solver is the oracle here (no GPU in this
container); on the GPU box the same code runs the CUDA path over NCCL."""

from __future__ import annotations

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from frontx_b200 import _dist  # noqa: E402


def test_shard_indices_partition():
    for n, world, block in [(10, 2, 2), (0, 3, 4), (4097, 8, 4096), (100_000, 8, 4096), (7, 4, 16), (33, 2, 1)]:
        parts = [_dist.shard_indices(n, r, world, block) for r in range(world)]
        allidx = np.concatenate(parts) if parts else np.zeros(0)
        assert sorted(allidx.tolist()) == list(range(n))
        assert _dist.shard_sizes(n, world, block).tolist() == [p.size for p in parts]
        for p in parts:
            assert np.all(np.diff(p) > 0)
    # block-cyclic: consecutive blocks go to consecutive ranks
    assert _dist.shard_indices(10, 1, 2, 2).tolist() == [2, 3, 6, 7]
    with pytest.raises(ValueError):
        _dist.shard_indices(10, 2, 2)


def _worker(rank: int, world: int, port: int, n: int, out_dir: str) -> None:
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from conftest import B_PAPER, I_PAPER, mixed_sweep

    from oracle import fxo

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    models = mixed_sweep(n, seed=11)

    def oracle_solver(sub, b, i):
        r = fxo.solve_batch(sub.model_id, sub.params, b, i, nthreads=1)
        return torch.from_numpy(r["summary"]), torch.from_numpy(r["counters"])

    summary, counters, mine = _dist.solve_batch_distributed(models, b=B_PAPER, i=I_PAPER, block=3,
                                                            local_solver=oracle_solver)
    np.save(os.path.join(out_dir, f"summary{rank}.npy"), summary.numpy())
    np.save(os.path.join(out_dir, f"counters{rank}.npy"), counters.numpy())
    np.save(os.path.join(out_dir, f"mine{rank}.npy"), mine)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_matches_single_process(tmp_path, fxo):
    import torch.multiprocessing as mp
    from conftest import B_PAPER, I_PAPER, mixed_sweep

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n, world = 14, 2
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    models = mixed_sweep(n, seed=11)
    ref = fxo.solve_batch(models.model_id, models.params, B_PAPER, I_PAPER)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"summary{r}.npy"), ref["summary"])
        np.testing.assert_array_equal(np.load(tmp_path / f"counters{r}.npy"), ref["counters"])
    mine = [np.load(tmp_path / f"mine{r}.npy") for r in range(world)]
    assert sorted(np.concatenate(mine).tolist()) == list(range(n))
    assert mine[0].tolist() == [0, 1, 2, 6, 7, 8, 12, 13]
