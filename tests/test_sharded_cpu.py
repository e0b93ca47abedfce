"""Host logic of the multi-GPU path on CPU: world_size-2 gloo.
Each rank aligns its contiguous shard (here with the oracle standing in for the
kernels -- this test is about the sharding and the compact all-gather, not the
DP), the compact forms are all-gathered and must equal the unsharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from torch_tts_b200 import sharded, synthetic


def test_shard_bounds_cover_batch():
    for batch in (1, 2, 7, 64, 127, 512):
        for world in (1, 2, 3, 4, 8):
            spans = [sharded.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
                assert a1 == b0
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, S, T, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import mas_oracle

        t_x, t_y = synthetic.ragged_lengths(batch, S, T, seed=21)
        nc = synthetic.neg_cent_like(batch, S, T, seed=21)
        lo, hi = sharded.shard_bounds(batch, rank, world)
        path = mas_oracle.maximum_path_c(nc[lo:hi].numpy(), t_y[lo:hi].numpy(), t_x[lo:hi].numpy())
        idx = torch.from_numpy(np.where(path.sum(2) > 0, path.argmax(2), -1).astype(np.int32))
        dur = torch.from_numpy(path.sum(1).astype(np.int32))
        g_idx, g_dur = sharded.gather_compact(idx, dur, batch)
        assert g_idx.shape == (batch, T) and g_dur.shape == (batch, S)
        if batch % world == 0:
            # fixed-shape buckets: the one-collective path gives the same result
            u_idx, u_dur = sharded.gather_compact(idx, dur, batch, uniform=True)
            assert torch.equal(u_idx, g_idx) and torch.equal(u_dur, g_dur)
        torch.save((g_idx, g_dur), os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 7])      # even and ragged shards
def test_gather_compact_world2_gloo(tmp_path, batch):
    from oracle import mas_oracle

    S, T, world = 19, 80, 2
    mp.spawn(_worker, args=(world, _free_port(), batch, S, T, str(tmp_path)), nprocs=world, join=True)
    t_x, t_y = synthetic.ragged_lengths(batch, S, T, seed=21)
    nc = synthetic.neg_cent_like(batch, S, T, seed=21)
    full = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
    want_idx = np.where(full.sum(2) > 0, full.argmax(2), -1)
    for r in range(world):
        g_idx, g_dur = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert np.array_equal(g_idx.numpy(), want_idx)
        assert np.array_equal(g_dur.numpy(), full.sum(1))
        assert torch.equal(g_dur.sum(1), t_y)


def _worker_ragged_shapes(rank, world, port, out_dir):
    """every rank pads its own batch to its own T and S, and holds a different number of utterances"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import mas_oracle

        n, S, T = [(3, 19, 80), (5, 23, 64)][rank]
        t_x, t_y = synthetic.ragged_lengths(n, S, T, seed=30 + rank)
        nc = synthetic.neg_cent_like(n, S, T, seed=30 + rank)
        path = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
        idx = torch.from_numpy(np.where(path.sum(2) > 0, path.argmax(2), -1).astype(np.int32))
        dur = torch.from_numpy(path.sum(1).astype(np.int32))
        g_idx, g_dur = sharded.gather_compact(idx, dur)
        torch.save((g_idx, g_dur), os.path.join(out_dir, f"rank{rank}.pt"))
        with pytest.raises(RuntimeError):
            sharded.gather_compact(idx, dur, batch=99)
    finally:
        dist.destroy_process_group()


def test_gather_compact_ranks_with_different_shapes(tmp_path):
    """DDP ranks pad to their own T / S (data_utils.py:168-177); the gather must not assume one shape."""
    from oracle import mas_oracle

    world = 2
    mp.spawn(_worker_ragged_shapes, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    want_idx, want_dur = [], []
    Tg, Sg = 80, 23
    for rank, (n, S, T) in enumerate([(3, 19, 80), (5, 23, 64)]):
        t_x, t_y = synthetic.ragged_lengths(n, S, T, seed=30 + rank)
        nc = synthetic.neg_cent_like(n, S, T, seed=30 + rank)
        path = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
        idx = np.full((n, Tg), -1, np.int64)
        idx[:, :T] = np.where(path.sum(2) > 0, path.argmax(2), -1)
        dur = np.zeros((n, Sg), np.int64)
        dur[:, :S] = path.sum(1)
        want_idx.append(idx)
        want_dur.append(dur)
    want_idx, want_dur = np.concatenate(want_idx), np.concatenate(want_dur)
    for r in range(world):
        g_idx, g_dur = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert g_idx.shape == (8, Tg) and g_dur.shape == (8, Sg)
        assert np.array_equal(g_idx.numpy(), want_idx) and np.array_equal(g_dur.numpy(), want_dur)
