"""Drop-in for the reference's `monotonic_align` module
(vits2/monotonic_align/__init__.py:6-19): same name, same arguments, same
return contract, but the tensors never leave the GPU.
"""
from __future__ import annotations

import torch

from . import _lib


def lengths_from_mask(mask: torch.Tensor):
    """t_t_max = mask.sum(1)[:, 0], t_s_max = mask.sum(2)[:, 0] (__init__.py:16-17)
    reading only column 0 / row 0 of each plane.  Returns int32 CUDA tensors."""
    _lib.require_cuda(mask, "mask")
    B, T, S = mask.shape
    if mask.dtype == torch.float32 and mask.is_contiguous():
        t_ys = torch.empty(B, dtype=torch.int32, device=mask.device)
        t_xs = torch.empty(B, dtype=torch.int32, device=mask.device)
        with torch.cuda.device(mask.device):
            rc = _lib.lib().mas_lengths_from_mask_f32(_lib.ptr(mask), _lib.ptr(t_ys), _lib.ptr(t_xs), B, T, S,
                                                      _lib.stream_ptr(mask.device))
        _lib.check(rc, "mas_lengths_from_mask_f32")
        return t_ys, t_xs
    # other dtypes / strided masks: same two slices, summed by torch
    t_ys = mask[:, :, 0].sum(1).to(torch.int32)
    t_xs = mask[:, 0, :].sum(1).to(torch.int32)
    return t_ys, t_xs


def maximum_path_compact(neg_cent: torch.Tensor, t_ys: torch.Tensor, t_xs: torch.Tensor, *,
                         want_path: bool = True, path_dtype=None):
    """MAS from explicit lengths.  Returns (path | None, durations int32 [B,S],
    idx int32 [B,T] (-1 past t_y), status int32 [B]).  want_path=False: the dense plane is neither
    allocated nor written (the kernel then only reads neg_cent and writes the compact outputs)."""
    _lib.require_cuda(neg_cent, "neg_cent")
    if neg_cent.dim() != 3:
        raise _lib.MasError(f"neg_cent must be [b, t_t, t_s], got {tuple(neg_cent.shape)}")
    B, T, S = neg_cent.shape
    device = neg_cent.device
    out_dtype = path_dtype or neg_cent.dtype
    nc = neg_cent.detach()
    if nc.dtype != torch.float32:
        nc = nc.float()                       # reference: .astype(np.float32), __init__.py:13
    nc = nc.contiguous()
    if nc.data_ptr() % 16:
        nc = nc.clone()
    t_ys = t_ys.to(device=device, dtype=torch.int32).contiguous()
    t_xs = t_xs.to(device=device, dtype=torch.int32).contiguous()
    kdtype = out_dtype if out_dtype in _lib.PATH_DTYPES else torch.float32
    L = _lib.lib()
    with torch.cuda.device(device):
        path = torch.empty((B, T, S), dtype=kdtype, device=device) if want_path else None
        dur = torch.empty((B, S), dtype=torch.int32, device=device)
        idx = torch.empty((B, T), dtype=torch.int32, device=device)
        status = torch.empty((B,), dtype=torch.int32, device=device)
        nbytes = L.mas_maximum_path_workspace_bytes(B, T, S)
        if nbytes == 0:
            raise _lib.MasError(f"unsupported shape B={B} T={T} S={S} (S <= 1024, T <= 65535)")
        ws = _lib.workspace(device, nbytes)
        rc = L.mas_maximum_path_f32(_lib.ptr(nc), _lib.ptr(t_ys), _lib.ptr(t_xs), _lib.ptr(path),
                                    _lib.PATH_DTYPES[kdtype], _lib.ptr(dur), _lib.ptr(idx), _lib.ptr(status),
                                    _lib.ptr(ws), ws.numel(), B, T, S, _lib.stream_ptr(device))
    _lib.check(rc, "mas_maximum_path_f32")
    if want_path and kdtype != out_dtype:
        path = path.to(out_dtype)
    return path, dur, idx, status


def maximum_path(neg_cent: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """neg_cent: [b, t_t, t_s]; mask: [b, t_t, t_s] -> path [b, t_t, t_s] on
    neg_cent.device with neg_cent.dtype (reference __init__.py:6-19).  Inputs
    are not modified."""
    _lib.require_cuda(neg_cent, "neg_cent")
    t_ys, t_xs = lengths_from_mask(mask.to(neg_cent.device))
    path, _, _, _ = maximum_path_compact(neg_cent, t_ys, t_xs)
    return path
