"""The SynthesizerTrn alignment call as one unit (reference vits2/models.py:1224-1256):
cost -> optional VITS2 noise-scaled MAS -> path and durations.
"""
from __future__ import annotations

import torch

from . import _lib


class CompactAlignment:
    """The alignment in compact form: `idx` [B,T] int32 (text column of every mel frame, -1 past
    the mel length), `durations` [B,S] int32 (= attn.sum(2), models.py:1256) and the per-utterance `status`.
    The dense one-hot `attn` [B,1,T,S] that the reference builds (models.py:1250-1254) -- half of the alignment's
    memory traffic, and only ever consumed again by two matmuls that `expand_prior` replaces and by the TensorBoard
    image hook (train_ms.py:517-519, cli.py:53-59) -- is built on first use of `.attn()` and cached."""

    def __init__(self, idx, durations, status, S: int, dtype):
        self.idx, self.durations, self.status = idx, durations, status
        self.S, self.dtype = S, dtype
        self._attn = None

    def attn(self):
        if self._attn is None:
            from .sharded import expand_path

            self._attn = expand_path(self.idx, self.S, self.dtype).unsqueeze(1)
        return self._attn

    @property
    def w(self):
        return self.durations.to(self.dtype).unsqueeze(1)

    def check(self):
        """Raise if any utterance violated 1 <= t_x <= t_y <= T (one small device -> host copy)."""
        bad = torch.nonzero(self.status).flatten().tolist()
        if bad:
            raise _lib.MasError(f"utterances {bad[:8]}{'...' if len(bad) > 8 else ''} have text/mel lengths the "
                                "reference leaves undefined (need 1 <= t_x <= t_y <= T, t_x <= S): no alignment")
        return self


def _lengths(x_mask, y_mask):
    # what mask.sum(1)[:,0] / mask.sum(2)[:,0] give for attn_mask = x_mask[:,:,None] * y_mask[...,None]
    # (models.py:1249, __init__.py:16-17)
    xm = x_mask.reshape(x_mask.shape[0], -1).float()
    ym = y_mask.reshape(y_mask.shape[0], -1).float()
    t_ys = (ym.sum(1) * xm[:, 0]).to(torch.int32)
    t_xs = (xm.sum(1) * ym[:, 0]).to(torch.int32)
    return t_ys, t_xs


def neg_cent(z_p: torch.Tensor, m_p: torch.Tensor, logs_p: torch.Tensor) -> torch.Tensor:
    """models.py:1226-1239 -> [B, T, S] float32."""
    for n, t in (("z_p", z_p), ("m_p", m_p), ("logs_p", logs_p)):
        _lib.require_cuda(t, n)
    B, D, T = z_p.shape
    S = m_p.shape[2]
    device = z_p.device
    z, m, l = (t.detach().float().contiguous() for t in (z_p, m_p, logs_p))
    L = _lib.lib()
    with torch.cuda.device(device):
        out = torch.empty((B, T, S), dtype=torch.float32, device=device)
        nbytes = L.mas_neg_cent_workspace_bytes(B, D, T, S)
        if nbytes == 0:
            raise _lib.MasError(f"unsupported shape B={B} D={D} T={T} S={S}")
        ws = _lib.workspace(device, nbytes)
        rc = L.mas_neg_cent_f32(_lib.ptr(z), _lib.ptr(m), _lib.ptr(l), _lib.ptr(out), None, _lib.ptr(ws),
                                ws.numel(), B, D, T, S, _lib.stream_ptr(device))
    _lib.check(rc, "mas_neg_cent_f32")
    return out


def align(z_p, m_p, logs_p, x_mask, y_mask, mas_noise_scale=None, noise=None, *,
          x_lengths=None, y_lengths=None, return_compact: bool = False, return_neg_cent: bool = False,
          zero_scale_is_no_noise: bool = False, dense: bool = True, check_status: bool = False):
    """Replacement for models.py:1224-1256.

    z_p [B,D,T], m_p/logs_p [B,D,S], x_mask [B,1,S], y_mask [B,1,T].
    mas_noise_scale: None, or the scalar of cli.py:268-271 (0 still takes the
    noise branch, as in the reference).  noise: optional [B,T,S] draw standing
    in for torch.randn_like(neg_cent) (models.py:1244); drawn here when absent.
    zero_scale_is_no_noise: the schedule of cli.py:268-271 reaches 0 after 5000 steps and stays there;
    `neg_cent + (std * noise) * 0` is `neg_cent` bit for bit whenever every cost is finite, so with this
    switch a scale of exactly 0 takes the single fused kernel (no draw, no second pass over the plane).
    Off by default because the two differ when neg_cent holds a NaN/Inf: the reference's std is then NaN
    and every cost of the batch with it.
    dense=False: the dense attn is never written (no zero fill, no scatter: the call moves about half the
    bytes); the first result is then a `CompactAlignment` (idx / durations / status, `.attn()` expands lazily).
    check_status=True: raise if an utterance has lengths outside 1 <= t_x <= t_y <= T (costs one tiny
    device -> host copy, i.e. a sync); without it such an utterance silently gets an all-zero path, zero
    durations and status 1 -- where the reference would read out of bounds.  Callers who keep the default must
    look at `status` themselves (return_compact=True).
    Note: the reference's default config (use_noise_scaled_mas, cli.py:268-271) always passes a number, also
    after the schedule has decayed to 0; pass zero_scale_is_no_noise=True to let scale 0 take the no-noise kernel.
    Returns (attn [B,1,T,S] in z_p.dtype, w [B,1,S]) and, on request, the
    compact (idx, durations, status) and the aligned neg_cent.
    """
    if zero_scale_is_no_noise and mas_noise_scale is not None and float(mas_noise_scale) == 0.0:
        mas_noise_scale, noise = None, None
    for n, t in (("z_p", z_p), ("m_p", m_p), ("logs_p", logs_p)):
        _lib.require_cuda(t, n)
    B, D, T = z_p.shape
    S = m_p.shape[2]
    device, dtype = z_p.device, z_p.dtype
    with torch.no_grad():
        z, m, l = (t.detach().float().contiguous() for t in (z_p, m_p, logs_p))
        if x_lengths is not None and y_lengths is not None:
            t_xs = x_lengths.to(device=device, dtype=torch.int32)
            t_ys = y_lengths.to(device=device, dtype=torch.int32)
        else:
            t_ys, t_xs = _lengths(x_mask.to(device), y_mask.to(device))
        nz = None
        scale = 0.0
        if mas_noise_scale is not None:
            scale = float(mas_noise_scale)
            nz = noise if noise is not None else torch.randn((B, T, S), dtype=torch.float32, device=device)
            nz = nz.to(device=device, dtype=torch.float32).contiguous()
        kdtype = dtype if dtype in _lib.PATH_DTYPES else torch.float32
        L = _lib.lib()
        with torch.cuda.device(device):
            path = torch.empty((B, T, S), dtype=kdtype, device=device) if dense else None
            dur = torch.empty((B, S), dtype=torch.int32, device=device)
            idx = torch.empty((B, T), dtype=torch.int32, device=device)
            status = torch.empty((B,), dtype=torch.int32, device=device)
            nc_out = torch.empty((B, T, S), dtype=torch.float32, device=device) if return_neg_cent else None
            nbytes = L.mas_fused_align_workspace_bytes(B, D, T, S, int(nz is not None))
            if nbytes == 0:
                raise _lib.MasError(f"unsupported shape B={B} D={D} T={T} S={S}")
            ws = _lib.workspace(device, nbytes)
            rc = L.mas_fused_align_f32(_lib.ptr(z), _lib.ptr(m), _lib.ptr(l), _lib.ptr(t_ys.contiguous()),
                                       _lib.ptr(t_xs.contiguous()), _lib.ptr(nz), scale, _lib.ptr(path),
                                       _lib.PATH_DTYPES[kdtype], _lib.ptr(dur), _lib.ptr(idx), _lib.ptr(status),
                                       _lib.ptr(nc_out), _lib.ptr(ws), ws.numel(), B, D, T, S,
                                       _lib.stream_ptr(device))
        _lib.check(rc, "mas_fused_align_f32")
        compact = CompactAlignment(idx, dur, status, S, dtype)
        if check_status:
            compact.check()
        if dense:
            if kdtype != dtype:
                path = path.to(dtype)
            attn = path.unsqueeze(1)                   # models.py:1252
        else:
            attn = compact
        w = dur.to(dtype).unsqueeze(1)                 # models.py:1256  attn.sum(2)
    out = (attn, w)
    if return_compact:
        out = out + ((idx, dur, status),)
    if return_neg_cent:
        out = out + (nc_out,)
    return out
