"""world_size-2 gloo test (CPU) of the multi-GPU plumbing: batch sharding + the single
all-gather of final unitaries, with a stand-in propagator so no GPU is needed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_propagate(signals, scale):
    """Deterministic per-row 'unitary': depends only on that row's signals."""
    B = signals.shape[0]
    s = signals.sum(dim=(1, 2)).to(torch.complex128)
    eye = torch.eye(3, dtype=torch.complex128).expand(B, 3, 3)
    return eye * (scale * s)[:, None, None] + 1j * eye


def _worker(rank, world, port, B, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from c3_b200.distributed import propagate_sharded, shard_bounds, all_gather_unitaries
        g = torch.Generator().manual_seed(0)
        signals = torch.rand((B, 2, 5), generator=g, dtype=torch.float64)   # same full batch on every rank
        U = propagate_sharded(_fake_propagate, signals, 2.0)
        want = _fake_propagate(signals, 2.0)
        ok = torch.allclose(U, want) and U.shape == want.shape
        lo, hi = shard_bounds(B, world, rank)
        U2 = all_gather_unitaries(want[lo:hi].clone(), B)
        ok = ok and torch.equal(U2, want)
        # fused-fidelity variant: the gather carries one float64 per batch element
        fid = propagate_sharded(lambda s_, c: _fake_propagate(s_, c).diagonal(dim1=1, dim2=2).real.sum(-1), signals, 2.0)
        ok = ok and fid.shape == (B,) and torch.allclose(fid, want.diagonal(dim1=1, dim2=2).real.sum(-1))
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def _run(B):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, port, B, results), nprocs=world, join=True)
    assert dict(results) == {0: True, 1: True}


def test_sharded_batch_even():
    _run(8)


def test_sharded_batch_ragged():
    _run(7)


def test_sharded_batch_smaller_than_world():
    """One batch row on two ranks: the rank without work must still join the gather (no hang, no B = 0 kernel call)."""
    _run(1)
