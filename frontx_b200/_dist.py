"""Multi-GPU: the problem batch is split across the ranks of one box with no
inter-GPU dependency; the only collective is an all-gather of the fixed-size
result records.

One process per GPU (``torchrun``), ``torch.distributed`` for the plumbing
(NCCL over NVLink on GPUs; the same code runs on ``gloo`` for the CPU tests of
the sharding logic).  Problems are dealt block-cyclically so that every rank
sees the same cost distribution whatever the order of the caller's sweep.
"""

from __future__ import annotations

from typing import Any, Callable

import numpy as np

from . import _lib
from .models import ModelBatch

BLOCK = 4096  # problems per block of the block-cyclic deal


def shard_indices(n: int, rank: int, world: int, block: int = BLOCK) -> np.ndarray:
    """Global indices owned by ``rank``: blocks rank, rank+world, rank+2*world, ..."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    nblocks = (n + block - 1) // block
    mine = np.arange(rank, nblocks, world, dtype=np.int64)
    if mine.size == 0:
        return np.zeros(0, dtype=np.int64)
    idx = (mine[:, None] * block + np.arange(block, dtype=np.int64)[None, :]).reshape(-1)
    return idx[idx < n]


def shard_sizes(n: int, world: int, block: int = BLOCK) -> np.ndarray:
    return np.array([shard_indices(n, r, world, block).size for r in range(world)], dtype=np.int64)


def gather_order(n: int, world: int, block: int = BLOCK) -> np.ndarray:
    """Global problem index of every row of the padded all-gather buffer (``world`` shards of the largest
    shard's size, rank-major), -1 for padding rows."""
    sizes = shard_sizes(n, world, block)
    pad = int(sizes.max()) if sizes.size else 0
    order = np.full(world * pad, -1, dtype=np.int64)
    for r in range(world):
        order[r * pad: r * pad + int(sizes[r])] = shard_indices(n, r, world, block)
    return order


def gather_records(local, n: int, world: int, rank: int, block: int = BLOCK, group: Any = None, order=None):
    """All-gather per-problem records ([n_local, w] tensors) into global order [n, w].

    Shards are padded to the largest shard so that ONE ``all_gather_into_tensor`` moves everything
    (64-byte summaries: 640 MB at n = 1e7, ~1 ms over NVSwitch), and ONE indexed copy puts the rows in
    the caller's order (``order`` = ``gather_order(n, world, block)`` as a tensor on the records' device;
    pass it to reuse it across calls).
    """
    import torch
    import torch.distributed as dist

    if world == 1:
        return local  # one rank owns every block, in order
    if order is None:
        order = torch.as_tensor(gather_order(n, world, block), device=local.device)
    pad = order.numel() // world if world else 0
    width = local.shape[1]
    if local.shape[0] == pad:
        send = local.contiguous()
    else:
        send = torch.zeros((pad, width), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    recv = torch.empty((world * pad, width), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    valid = order >= 0
    out = torch.empty((n, width), dtype=local.dtype, device=local.device)
    out.index_copy_(0, order[valid], recv[valid])
    return out


def solve_shards_distributed(build_shard: Callable, n: int, *, itol: float = 1e-3, max_steps: int = 100,
                             gather: bool = True, block: int = BLOCK, group: Any = None,
                             local_solver: Callable | None = None, order=None):
    """Solve a batch of ``n`` problems across all ranks, every rank BUILDING ONLY ITS OWN SHARD:
    ``build_shard(indices) -> (ModelBatch, b, i)`` receives the global indices the rank owns (blocks
    rank, rank + world, ... of ``block`` problems) and returns their models and boundary values -- from a
    generator seeded by block, a memory-mapped file, ... -- so no rank ever holds the whole batch.
    Returns ``(summary, counters, indices)``: the global ``[n, 8]`` records with ``gather=True`` (one
    all-gather each + one indexed copy), else the rank's own.
    """
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_indices(n, rank, world, block)
    sub, b_loc, i_loc = build_shard(mine)
    if len(sub) != mine.size:
        raise ValueError("build_shard must return one model per index it was given")
    if local_solver is None:
        from ._forward import solve_batch

        sol = solve_batch(sub, b=b_loc, i=i_loc, itol=itol, max_steps=max_steps, throw=False, k_max=0)
        summary, counters = sol.summary, sol.counters
    else:
        summary, counters = local_solver(sub, np.broadcast_to(np.asarray(b_loc, dtype=np.float64), (mine.size,)),
                                         np.broadcast_to(np.asarray(i_loc, dtype=np.float64), (mine.size,)))
    if not gather:
        return summary, counters, mine
    return (gather_records(summary, n, world, rank, block, group, order),
            gather_records(counters, n, world, rank, block, group, order), mine)


def solve_batch_distributed(models: ModelBatch, *, b: Any, i: Any, itol: float = 1e-3, max_steps: int = 100,
                            gather: bool = True, block: int = BLOCK, group: Any = None,
                            local_solver: Callable | None = None):
    """Convenience form of ``solve_shards_distributed`` for a batch every rank already holds: every rank
    passes the SAME global batch and solves ``shard_indices(n, rank, world)`` of it.  (A 10^7-problem batch is
    960 MB of parameters per rank this way; build the shards per rank instead when that matters.)
    ``local_solver(models, b, i)`` -> (summary, counters) tensors exists so that the CPU tests can exercise
    the sharding and the collective without a GPU; the default is the CUDA path.
    """
    n = len(models)
    b_all = np.broadcast_to(np.asarray(b, dtype=np.float64), (n,))
    i_all = np.broadcast_to(np.asarray(i, dtype=np.float64), (n,))
    empty = ModelBatch(np.zeros(0, np.int32), np.zeros((0, _lib.NPARAM)))

    def build(idx):
        return (models[idx] if idx.size else empty), b_all[idx], i_all[idx]

    return solve_shards_distributed(build, n, itol=itol, max_steps=max_steps, gather=gather, block=block,
                                    group=group, local_solver=local_solver)


def profile_slice(batch: int, rank: int, world: int) -> slice:
    """Contiguous profiles owned by ``rank`` (C5 shards by profile: 125 of 10^3 per GPU on 8 GPUs)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    per, extra = divmod(batch, world)
    start = rank * per + min(rank, extra)
    return slice(start, start + per + (1 if rank < extra else 0))


def inverse_distributed(o: Any, theta: Any, *, gather_D: bool = False, group: Any = None,  # noqa: N803
                        local_inverse: Callable | None = None):
    """Profile -> D(theta) for ``[batch, npts]`` profiles across all ranks of the default process group.

    Every rank passes the SAME global arrays (numpy); rank r inverts ``profile_slice(batch, r, world)`` on its
    own GPU.  Profiles are independent, so nothing is exchanged on the data path; the per-profile sorptivity
    and status records (and, with ``gather_D=True``, the D arrays: 8 bytes per point) are all-gathered.
    Returns ``(D, sorptivity[batch], status[batch], my_slice)`` with ``D`` the local ``[n_local, npts]``
    tensor unless gathered.  ``local_inverse(o, theta) -> (D, sorptivity, status)`` tensors lets the CPU tests
    run the sharding and the collective without a GPU; the default is the CUDA path.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    o = np.asarray(o, dtype=np.float64)
    theta = np.asarray(theta, dtype=np.float64)
    if o.ndim != 2 or o.shape != theta.shape:
        raise ValueError("o and theta must have the same [batch, npts] shape")
    batch, npts = o.shape
    mine = profile_slice(batch, rank, world)
    if local_inverse is None:
        from ._inverse import inverse

        if mine.stop > mine.start:
            res = inverse(o[mine], theta[mine], throw=False)
            D, sorp, status = res.D_at_knots, res.sorptivity(), res.status
        else:
            dev = torch.device("cuda", torch.cuda.current_device())
            D = torch.empty((0, npts), dtype=torch.float64, device=dev)
            sorp = torch.empty((0,), dtype=torch.float64, device=dev)
            status = torch.empty((0,), dtype=torch.int32, device=dev)
    else:
        D, sorp, status = local_inverse(o[mine], theta[mine])
    pad = -(-batch // world)

    def gather(local, width):
        send = torch.zeros((pad, width), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local.reshape(local.shape[0], width)
        recv = torch.empty((world * pad, width), dtype=local.dtype, device=local.device)
        if world == 1:
            recv.copy_(send)
        else:
            dist.all_gather_into_tensor(recv, send, group=group)
        parts = []
        for r in range(world):
            sl = profile_slice(batch, r, world)
            parts.append(recv[r * pad: r * pad + (sl.stop - sl.start)])
        return torch.cat(parts)

    sorp_all = gather(sorp, 1).reshape(batch)
    status_all = gather(status, 1).reshape(batch)
    D_out = gather(D, npts) if gather_D else D
    return D_out, sorp_all, status_all, mine


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment (RANK, WORLD_SIZE, LOCAL_RANK)."""
    import os

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world)
    return rank, world, local
