"""AlignPlan: a fixed-shape, pre-allocated launch plan for the fused alignment
(cost -> noise -> MAS).  One ctypes call per step, no allocation, optional CUDA
graph replay -- what a training loop with static shapes (DistributedBucketSampler
buckets, data_utils.py:434-550) would hold per bucket.
"""
from __future__ import annotations

import torch

from . import _lib


class AlignPlan:
    def __init__(self, B: int, D: int, T: int, S: int, device, dtype=torch.float32, with_noise: bool = False,
                 want_path: bool = True):
        """want_path=False: compact outputs only (self.idx / self.dur / self.status); the dense path plane is
        neither allocated nor written -- `expand_path(plan.idx, S)` rebuilds it on demand."""
        self.B, self.D, self.T, self.S = B, D, T, S
        self.device = torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.dtype = dtype
        self.with_noise = with_noise
        self.want_path = want_path
        if dtype not in _lib.PATH_DTYPES:
            raise _lib.MasError(f"unsupported path dtype {dtype}")
        L = _lib.lib()
        nbytes = L.mas_fused_align_workspace_bytes(B, D, T, S, int(with_noise))
        if nbytes == 0:
            raise _lib.MasError(f"unsupported shape B={B} D={D} T={T} S={S}")
        with torch.cuda.device(self.device):
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self.path = torch.empty((B, T, S), dtype=dtype, device=self.device) if want_path else None
            self.dur = torch.empty((B, S), dtype=torch.int32, device=self.device)
            self.idx = torch.empty((B, T), dtype=torch.int32, device=self.device)
            self.status = torch.empty((B,), dtype=torch.int32, device=self.device)
        self._L = L
        self._graphs = {}
        self._inputs = {}

    def run(self, z_p, m_p, logs_p, t_ys, t_xs, noise=None, noise_scale: float = 0.0):
        """Stream-ordered on the current stream of the plan's device; results land in self.path/dur/idx/status.
        Inputs must be fp32, contiguous, on the plan's device, of the plan's shapes (under autocast: upcast first,
        as `align()` does) -- checked here, the kernels read raw pointers."""
        B, D, T, S = self.B, self.D, self.T, self.S
        self._check("z_p", z_p, (B, D, T), torch.float32)
        self._check("m_p", m_p, (B, D, S), torch.float32)
        self._check("logs_p", logs_p, (B, D, S), torch.float32)
        self._check("t_ys", t_ys, (B,), torch.int32)
        self._check("t_xs", t_xs, (B,), torch.int32)
        if noise is not None:
            self._check("noise", noise, (B, T, S), torch.float32)
        with torch.cuda.device(self.device):
            rc = self._L.mas_fused_align_f32(
                z_p.data_ptr(), m_p.data_ptr(), logs_p.data_ptr(), t_ys.data_ptr(), t_xs.data_ptr(),
                None if noise is None else noise.data_ptr(), float(noise_scale),
                None if self.path is None else self.path.data_ptr(),
                _lib.PATH_DTYPES[self.dtype], self.dur.data_ptr(), self.idx.data_ptr(), self.status.data_ptr(), None,
                self.ws.data_ptr(), self.ws.numel(), B, D, T, S,
                torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "mas_fused_align_f32")

    def _check(self, name, t, shape, dtype):
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != shape or not t.is_contiguous():
            raise _lib.MasError(f"AlignPlan.run: {name} must be a contiguous {dtype} tensor of shape {shape} on "
                                f"{self.device}, got {t.dtype} {tuple(t.shape)} on {t.device}"
                                f"{'' if t.is_contiguous() else ' (not contiguous)'}")

    def capture(self, key, z_p, m_p, logs_p, t_ys, t_xs, noise=None, noise_scale: float = 0.0):
        """Capture one step on these (static) input buffers into a CUDA graph; replay(key) launches it."""
        self.run(z_p, m_p, logs_p, t_ys, t_xs, noise, noise_scale)   # warm-up outside capture (func attributes)
        torch.cuda.synchronize(self.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.run(z_p, m_p, logs_p, t_ys, t_xs, noise, noise_scale)
        self._graphs[key] = g
        # the graph reads these buffers on every replay: keep them alive (a caller passing temporaries, e.g.
        # `t_y.to(device)`, would otherwise replay on freed memory)
        self._inputs[key] = (z_p, m_p, logs_p, t_ys, t_xs, noise)
        return g

    def replay(self, key):
        self._graphs[key].replay()
