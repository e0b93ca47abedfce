"""Multi-GPU side of the path: the batch shards across
ranks with no data-path collective; only when the caller wants the global
result are the COMPACT forms (idx [B,T] int32, durations [B,S] int32)
all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and re-expanded
locally.  The dense [B,T,S] path is never communicated.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(batch: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a batch for `rank`; sizes differ by at most one."""
    base, rem = divmod(batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_compact(idx: torch.Tensor, dur: torch.Tensor, batch=None, group=None, uniform: bool = False):
    """All-gather per-rank compact results into global [batch, T_max] / [batch, S_max].

    Ranks may hold different numbers of utterances AND different padded sizes: under DDP every rank's batch is
    padded by TextAudioCollate to its own longest text / spectrogram (data_utils.py:168-177), so T and S differ
    from rank to rank.  The ranks first agree on (shard size, T, S) maxima with one small all-gather, then every
    rank pads to them (idx with -1 = "no frame", durations with 0) for the one data collective.
    `batch`, when given, is checked against the gathered total.
    uniform=True: the caller guarantees that every rank holds the same number of utterances with the same padded T and
    S (fixed-shape buckets): no size exchange and no host synchronisation, one collective on the packed
    [n, T + S] int32 block."""
    world = dist.get_world_size(group)
    n, T = idx.shape
    S = dur.shape[1]
    if uniform:
        packed = torch.cat([idx, dur], 1)
        out = torch.empty((world * n, T + S), dtype=torch.int32, device=idx.device)
        dist.all_gather_into_tensor(out, packed, group=group)
        if batch is not None and out.shape[0] != batch:
            raise _lib.MasError(f"gather_compact: ranks hold {out.shape[0]} utterances in total, caller expected {batch}")
        return out[:, :T].contiguous(), out[:, T:].contiguous()
    mine = torch.tensor([n, T, S], dtype=torch.int64, device=idx.device)
    sizes = torch.empty((world * 3,), dtype=torch.int64, device=idx.device)
    dist.all_gather_into_tensor(sizes, mine, group=group)
    sizes = sizes.view(world, 3).cpu()
    per, Tg, Sg = (int(v) for v in sizes.max(0).values)
    packed = torch.empty((per, Tg + Sg), dtype=torch.int32, device=idx.device)
    packed[:, :Tg] = -1
    packed[:, Tg:] = 0
    packed[:n, :T] = idx
    packed[:n, Tg:Tg + S] = dur
    out = torch.empty((world * per, Tg + Sg), dtype=torch.int32, device=idx.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    rows = [out[r * per: r * per + int(sizes[r, 0])] for r in range(world)]
    full = torch.cat(rows, 0)
    if batch is not None and full.shape[0] != batch:
        raise _lib.MasError(f"gather_compact: ranks hold {full.shape[0]} utterances in total, caller expected {batch}")
    return full[:, :Tg].contiguous(), full[:, Tg:].contiguous()


def expand_path(idx: torch.Tensor, S: int, dtype=torch.float32) -> torch.Tensor:
    """Compact idx [B,T] (-1 = padding row) -> dense {0,1} path [B,T,S] on the GPU."""
    _lib.require_cuda(idx, "idx")
    B, T = idx.shape
    idx = idx.to(torch.int32).contiguous()
    kdtype = dtype if dtype in _lib.PATH_DTYPES else torch.float32
    with torch.cuda.device(idx.device):
        path = torch.empty((B, T, S), dtype=kdtype, device=idx.device)
        rc = _lib.lib().mas_expand_path(_lib.ptr(idx), _lib.ptr(path), _lib.PATH_DTYPES[kdtype], B, T, S,
                                        _lib.stream_ptr(idx.device))
    _lib.check(rc, "mas_expand_path")
    return path if kdtype == dtype else path.to(dtype)


def align_sharded(z_p, m_p, logs_p, x_mask, y_mask, mas_noise_scale=None, noise=None, *, gather: bool = False,
                  global_batch=None, group=None):
    """Align this rank's shard (inputs are already the rank's shard, as under DDP
    with DistributedBucketSampler, data_utils.py:514).  With gather=True also
    returns the global (idx, durations) all-gathered in compact form."""
    from .align import align

    attn, w, (idx, dur, status) = align(z_p, m_p, logs_p, x_mask, y_mask, mas_noise_scale, noise,
                                        return_compact=True)
    if not gather:
        return attn, w
    g_idx, g_dur = gather_compact(idx, dur, global_batch, group)
    return attn, w, g_idx, g_dur
