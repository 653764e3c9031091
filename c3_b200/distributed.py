"""Multi-GPU plumbing: the batch axis (ORBIT sequences, noise trajectories, optimiser samples)
shards across ranks; every batch element is an independent ordered product, so there is no
data-path collective -- only one all-gather of the final unitaries (or fidelities) so that
each rank sees the whole batch.  One process per GPU, torch.distributed
(NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(B: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block split of B batch rows; the remainder goes to the low ranks."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_unitaries(U_local: torch.Tensor, B_total: int = None) -> torch.Tensor:
    """Gather the per-rank U[b0:b1] into U[B] on every rank (one collective).

    Complex tensors travel as their float64 view.  Ragged shards (B not divisible by the world
    size) are padded to the largest shard and trimmed after the gather."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return U_local
    world = dist.get_world_size()
    n_local = U_local.shape[0]
    if B_total is None:
        B_total = n_local * world
    sizes = [shard_bounds(B_total, world, r) for r in range(world)]
    n_max = max(hi - lo for lo, hi in sizes)
    is_c = U_local.is_complex()
    x = torch.view_as_real(U_local.contiguous()) if is_c else U_local.contiguous()
    if n_local < n_max:
        pad = torch.zeros((n_max - n_local,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        x = torch.cat([x, pad], 0)
    out = torch.empty((world * n_max,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous())
    if any((hi - lo) != n_max for lo, hi in sizes):
        out = torch.cat([out[r * n_max:r * n_max + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)
    return torch.view_as_complex(out) if is_c else out


def propagate_sharded(fn, signals: torch.Tensor, *args, **kwargs) -> torch.Tensor:
    """Run ``fn(signals[b0:b1], ...) -> U[b1-b0, ...]`` on this rank's shard of the batch and
    all-gather the result.  ``signals`` is the FULL batch [B,K,N] (host or device) on every rank."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return fn(signals, *args, **kwargs)
    B = signals.shape[0]
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = shard_bounds(B, world, rank)
    U_local = fn(signals[lo:hi], *args, **kwargs) if hi > lo else None
    if B < world:
        # more ranks than batch rows: the ranks without work skip the kernel (it rejects B = 0) but still take part in the
        # gather, with an empty shard shaped like rank 0's result (rank 0 always owns a row)
        meta = [(tuple(U_local.shape[1:]), U_local.dtype) if rank == 0 else None]
        dist.broadcast_object_list(meta, src=0)
        if U_local is None:
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
            U_local = torch.empty((0,) + meta[0][0], dtype=meta[0][1], device=dev)
    return all_gather_unitaries(U_local, B)
