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


def gather_records(local, n: int, world: int, rank: int, block: int = BLOCK, group: Any = None):
    """All-gather per-problem records ([n_local, w] tensors) into global order [n, w].

    Shards are padded to the largest shard so that one ``all_gather_into_tensor``
    moves everything (64-byte summaries: 640 MB at n = 1e7, ~1 ms over NVSwitch).
    """
    import torch
    import torch.distributed as dist

    sizes = shard_sizes(n, world, block)
    pad = int(sizes.max()) if sizes.size else 0
    width = local.shape[1]
    send = torch.zeros((pad, width), dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    recv = torch.empty((world * pad, width), dtype=local.dtype, device=local.device)
    if world == 1:
        recv.copy_(send)
    else:
        dist.all_gather_into_tensor(recv, send, group=group)
    out = torch.empty((n, width), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = torch.as_tensor(shard_indices(n, r, world, block), device=local.device)
        out[idx] = recv[r * pad: r * pad + int(sizes[r])]
    return out


def solve_batch_distributed(models: ModelBatch, *, b: Any, i: Any, itol: float = 1e-3, max_steps: int = 100,
                            gather: bool = True, block: int = BLOCK, group: Any = None,
                            local_solver: Callable | None = None):
    """Solve a batch across all ranks of the default process group.

    Every rank passes the SAME global batch; rank r solves ``shard_indices(n, r, world)``
    on its own GPU and, with ``gather=True``, all ranks return the global
    ``(summary[n, 8], counters[n, 8])`` tensors.  ``local_solver(models, b, i)`` ->
    (summary, counters) tensors exists so that the CPU tests can exercise the
    sharding and the collective without a GPU; the default is the CUDA path.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = len(models)
    mine = shard_indices(n, rank, world, block)
    b_all = np.broadcast_to(np.asarray(b, dtype=np.float64), (n,))
    i_all = np.broadcast_to(np.asarray(i, dtype=np.float64), (n,))
    sub = models[mine] if mine.size else ModelBatch(np.zeros(0, np.int32), np.zeros((0, _lib.NPARAM)))
    if local_solver is None:
        from ._forward import solve_batch

        sol = solve_batch(sub, b=b_all[mine], i=i_all[mine], itol=itol, max_steps=max_steps, throw=False, k_max=0)
        summary, counters = sol.summary, sol.counters
    else:
        summary, counters = local_solver(sub, b_all[mine], i_all[mine])
    if not gather:
        return summary, counters, mine
    return (gather_records(summary, n, world, rank, block, group),
            gather_records(counters, n, world, rank, block, group), mine)


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
