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


def test_profile_slices_partition():
    for batch, world in [(1000, 8), (10, 3), (2, 4), (0, 2), (7, 7)]:
        sls = [_dist.profile_slice(batch, r, world) for r in range(world)]
        covered = [j for sl in sls for j in range(sl.start, sl.stop)]
        assert covered == list(range(batch))
        sizes = [sl.stop - sl.start for sl in sls]
        assert max(sizes) - min(sizes) <= 1
    assert _dist.profile_slice(1000, 3, 8) == slice(375, 500)  # C5: 125 profiles per GPU


def _inverse_worker(rank: int, world: int, port: int, out_dir: str) -> None:
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    from oracle import fxo

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o, theta = _inverse_profiles()

    def oracle_inverse(oo, th):
        D, S, st = [], [], []
        for j in range(oo.shape[0]):
            rc, Dk, s1, _ = fxo.inverse(oo[j], th[j])
            D.append(Dk)
            S.append(s1)
            st.append(rc)
        return (torch.from_numpy(np.array(D).reshape(len(D), oo.shape[1])), torch.tensor(S, dtype=torch.float64),
                torch.tensor(st, dtype=torch.int32))

    D, sorp, status, mine = _dist.inverse_distributed(o, theta, gather_D=True, local_inverse=oracle_inverse)
    np.save(os.path.join(out_dir, f"invD{rank}.npy"), D.numpy())
    np.save(os.path.join(out_dir, f"invS{rank}.npy"), sorp.numpy())
    np.save(os.path.join(out_dir, f"invM{rank}.npy"), np.array([mine.start, mine.stop]))
    dist.barrier()
    dist.destroy_process_group()


def _inverse_profiles():
    a = np.array([0.5, 0.8, 1.0, 1.3, 2.0])
    o = np.linspace(0, 20, 400)[None, :] / a[:, None]
    return o, np.exp(-a[:, None] * o)


def test_two_rank_inverse_matches_single_process(tmp_path, fxo):
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    world = 2
    mp.spawn(_inverse_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    o, theta = _inverse_profiles()
    ref_D = np.array([fxo.inverse(o[j], theta[j])[1] for j in range(o.shape[0])])
    ref_S = np.array([fxo.inverse(o[j], theta[j])[2] for j in range(o.shape[0])])
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"invD{r}.npy"), ref_D)
        np.testing.assert_array_equal(np.load(tmp_path / f"invS{r}.npy"), ref_S)
    assert np.load(tmp_path / "invM0.npy").tolist() == [0, 3] and np.load(tmp_path / "invM1.npy").tolist() == [3, 5]


def test_c3_rows_are_block_seeded():
    """bench.c3_rows: any rank can build any shard of config C3 without the rest of the batch."""
    import bench

    full = bench.c3_rows(np.arange(3 * bench.C3_BLOCK + 77))
    idx = np.concatenate([np.arange(5, 40), np.arange(bench.C3_BLOCK - 3, bench.C3_BLOCK + 9), np.arange(2 * bench.C3_BLOCK, 2 * bench.C3_BLOCK + 4)])
    part = bench.c3_rows(idx)
    np.testing.assert_array_equal(part.params, full.params[idx])
    np.testing.assert_array_equal(part.model_id, full.model_id[idx])
    assert (full.model_id[0::2] == bench.BROOKS_COREY).all() and (full.model_id[1::2] == bench.LETD).all()
    p = full.params
    bc, ld = p[0::2], p[1::2]
    assert bc[:, 0].min() >= 0.2 and bc[:, 0].max() <= 0.6 and bc[:, 2].min() >= 1e-6 and bc[:, 2].max() <= 1e-5
    assert ld[:, 1].min() >= 1e3 and ld[:, 1].max() <= 2e4 and ld[:, 2].min() >= 1.0 and ld[:, 2].max() <= 1.6


def _shards_worker(rank: int, world: int, port: int, n: int, out_dir: str) -> None:
    import torch
    import torch.distributed as dist

    sys.path.insert(0, str(ROOT))
    import bench
    from oracle import fxo

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    built = []

    def build(idx):  # the rank builds ONLY its own rows
        built.append(idx.copy())
        return bench.c3_rows(idx), bench.B_BOUNDARY, bench.I_INITIAL

    def oracle_solver(sub, b, i):
        r = fxo.solve_batch(sub.model_id, sub.params, b, i, nthreads=1)
        return torch.from_numpy(r["summary"]), torch.from_numpy(r["counters"])

    summary, counters, mine = _dist.solve_shards_distributed(build, n, block=4, local_solver=oracle_solver)
    assert len(built) == 1 and np.array_equal(built[0], mine)
    np.save(os.path.join(out_dir, f"ssum{rank}.npy"), summary.numpy())
    np.save(os.path.join(out_dir, f"scnt{rank}.npy"), counters.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_built_per_rank(tmp_path, fxo):
    """C3 through solve_shards_distributed at world_size 2: every rank builds only its block-cyclic shard, one
    all-gather + one indexed copy give every rank the global records in the caller's order."""
    import torch.multiprocessing as mp

    import bench

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n, world = 22, 2
    mp.spawn(_shards_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    models = bench.c3_rows(np.arange(n))
    ref = fxo.solve_batch(models.model_id, models.params, bench.B_BOUNDARY, bench.I_INITIAL)
    for r in range(world):
        np.testing.assert_array_equal(np.load(tmp_path / f"ssum{r}.npy"), ref["summary"])
        np.testing.assert_array_equal(np.load(tmp_path / f"scnt{r}.npy"), ref["counters"])
    order = _dist.gather_order(n, world, 4)
    assert sorted(order[order >= 0].tolist()) == list(range(n)) and (order >= 0).sum() == n
