"""The alignment's first consumers in SynthesizerTrn.forward, from the compact MAS outputs
(reference vits2/models.py:1256, 1261, 1270-1271).

    m_p, logs_p = expand_prior(m_p, logs_p, idx, durations)     # instead of two matmuls with the one-hot path
    logw_       = logw(durations, x_lengths)                    # instead of attn.sum(2) + log

`idx` [B,T] int32 and `durations` [B,S] int32 are what `align(..., return_compact=True)` /
`maximum_path_compact` return.  `expand_prior` is differentiable with respect to m_p / logs_p (they come out
of the text encoder and require grad, models.py:1215); the backward is a segmented sum in a fixed order.
CUDA only, through the C ABI; no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib


def _check_shapes(m_p, logs_p, idx, durations):
    if m_p.dim() != 3 or idx.dim() != 2 or durations.dim() != 2:
        raise _lib.MasError("expected m_p [B,D,S], idx [B,T], durations [B,S]")
    B, D, S = m_p.shape
    if logs_p is not None and tuple(logs_p.shape) != (B, D, S):
        raise _lib.MasError(f"logs_p {tuple(logs_p.shape)} does not match m_p {tuple(m_p.shape)}")
    if idx.shape[0] != B or tuple(durations.shape) != (B, S):
        raise _lib.MasError(f"idx {tuple(idx.shape)} / durations {tuple(durations.shape)} do not match m_p {tuple(m_p.shape)}")
    return B, D, idx.shape[1], S


class _ExpandPrior(torch.autograd.Function):
    @staticmethod
    def forward(ctx, m_p, logs_p, idx, durations):
        B, D, T, S = _check_shapes(m_p, logs_p, idx, durations)
        device = m_p.device
        m = m_p.detach().float().contiguous()
        l = logs_p.detach().float().contiguous() if logs_p is not None else None
        idx = idx.to(device=device, dtype=torch.int32).contiguous()
        with torch.cuda.device(device):
            m_out = torch.empty((B, D, T), dtype=torch.float32, device=device)
            l_out = torch.empty((B, D, T), dtype=torch.float32, device=device) if l is not None else None
            rc = _lib.lib().mas_expand_prior_f32(_lib.ptr(m), _lib.ptr(l), _lib.ptr(idx), _lib.ptr(m_out), _lib.ptr(l_out),
                                                 B, D, T, S, _lib.stream_ptr(device))
        _lib.check(rc, "mas_expand_prior_f32")
        ctx.save_for_backward(durations.to(device=device, dtype=torch.int32).contiguous())
        ctx.dims = (B, D, T, S)
        ctx.has_logs = l is not None
        ctx.in_dtype = m_p.dtype
        m_out = m_out.to(m_p.dtype)
        if l_out is not None:
            return m_out, l_out.to(logs_p.dtype)
        return m_out, None

    @staticmethod
    def backward(ctx, g_m, g_l):
        (dur,) = ctx.saved_tensors
        B, D, T, S = ctx.dims
        device = dur.device
        two = ctx.has_logs
        if g_m is None:
            g_m = torch.zeros((B, D, T), dtype=torch.float32, device=device)
        if two and g_l is None:
            g_l = torch.zeros((B, D, T), dtype=torch.float32, device=device)
        g_m = g_m.float().contiguous()
        g_l = g_l.float().contiguous() if two else None
        with torch.cuda.device(device):
            g_m_p = torch.empty((B, D, S), dtype=torch.float32, device=device)
            g_l_p = torch.empty((B, D, S), dtype=torch.float32, device=device) if two else None
            rc = _lib.lib().mas_expand_prior_backward_f32(_lib.ptr(g_m), _lib.ptr(g_l), _lib.ptr(dur), _lib.ptr(g_m_p),
                                                          _lib.ptr(g_l_p), B, D, T, S, _lib.stream_ptr(device))
        _lib.check(rc, "mas_expand_prior_backward_f32")
        return g_m_p.to(ctx.in_dtype), (g_l_p.to(ctx.in_dtype) if two else None), None, None


def expand_prior(m_p: torch.Tensor, logs_p, idx: torch.Tensor, durations: torch.Tensor):
    """Replacement for models.py:1270-1271: (m_p, logs_p) [B,D,S] -> [B,D,T] along the alignment.
    Frames past an utterance's mel length (idx == -1) come out as 0, as the all-zero path rows make them.
    logs_p may be None (then the second result is None)."""
    _lib.require_cuda(m_p, "m_p")
    return _ExpandPrior.apply(m_p, logs_p, idx, durations)


def logw(durations: torch.Tensor, x_lengths: torch.Tensor) -> torch.Tensor:
    """Replacement for models.py:1256 + 1261: logw_ = log(attn.sum(2) + 1e-6) * x_mask, [B,1,S] fp32."""
    _lib.require_cuda(durations, "durations")
    B, S = durations.shape
    device = durations.device
    dur = durations.to(torch.int32).contiguous()
    t_xs = x_lengths.to(device=device, dtype=torch.int32).contiguous()
    with torch.cuda.device(device):
        out = torch.empty((B, S), dtype=torch.float32, device=device)
        rc = _lib.lib().mas_logw_f32(_lib.ptr(dur), _lib.ptr(t_xs), _lib.ptr(out), B, S, _lib.stream_ptr(device))
    _lib.check(rc, "mas_logw_f32")
    return out.unsqueeze(1)


def idx_from_durations(duration: torch.Tensor, x_lengths: torch.Tensor, T: int, y_lengths=None) -> torch.Tensor:
    """Inference side: ceil()ed durations [B,S] or [B,1,S] -> compact alignment idx [B,T] int32 (-1 where no text
    column covers the frame), the compact form of commons.generate_path (commons.py:130-145)."""
    _lib.require_cuda(duration, "duration")
    dur = duration.reshape(duration.shape[0], duration.shape[-1]).float().contiguous()
    B, S = dur.shape
    device = dur.device
    t_xs = x_lengths.to(device=device, dtype=torch.int32).contiguous()
    t_ys = y_lengths.to(device=device, dtype=torch.int32).contiguous() if y_lengths is not None else None
    with torch.cuda.device(device):
        idx = torch.empty((B, T), dtype=torch.int32, device=device)
        rc = _lib.lib().mas_idx_from_durations_f32(_lib.ptr(dur), _lib.ptr(t_xs), _lib.ptr(t_ys), _lib.ptr(idx), B, T, S,
                                                   _lib.stream_ptr(device))
    _lib.check(rc, "mas_idx_from_durations_f32")
    return idx


def generate_path(duration: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """Drop-in for commons.generate_path(duration [b,1,t_x], mask [b,1,t_y,t_x]) -> path [b,1,t_y,t_x] in mask.dtype.
    The lengths are read off the mask's first row / column as `maximum_path` does (the mask is an outer product of
    two length masks, models.py:1309)."""
    from .monotonic_align import lengths_from_mask
    from .sharded import expand_path

    _lib.require_cuda(mask, "mask")
    b, _, t_y, t_x = mask.shape
    t_ys, t_xs = lengths_from_mask(mask.reshape(b, t_y, t_x))
    idx = idx_from_durations(duration, t_xs, t_y, t_ys)
    return expand_path(idx, t_x, mask.dtype).unsqueeze(1)
