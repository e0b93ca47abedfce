"""CPU oracle for the VITS2 MAS alignment hot path (python face of oracle/mas_oracle.c).

TEST INFRASTRUCTURE ONLY.  The product package (torch_tts_b200/) never imports
this module; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs do, and only as the checker.

Parity status: PINNED -- see the header of oracle/mas_oracle.c and
tests/test_oracle.py (bit-exact against the compiled reference in oracle/_ref
and against tests/golden/*.npz generated from the reference itself).

Citations are relative to /root/reference/.
"""
from __future__ import annotations

import ctypes
import importlib.util
import math
import os
import subprocess
import sysconfig

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(HERE, "libmas_oracle.so")
_SRC_PATH = os.path.join(HERE, "mas_oracle.c")
_lib = None


def build(force: bool = False) -> str:
    """gcc -O2 (no fast-math: the DP must stay one add + one compare per cell)."""
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(_SRC_PATH)):
        return _LIB_PATH
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.run([cc, "-O2", "-shared", "-fPIC", "-fno-fast-math", "-ffp-contract=off",
                    _SRC_PATH, "-o", _LIB_PATH, "-lm"], check=True)
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(_LIB_PATH)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f32p = ctypes.POINTER(ctypes.c_float)
        lib.mas_oracle_batch.argtypes = [i32p, f32p, i32p, i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.mas_oracle_batch.restype = ctypes.c_int
        lib.mas_oracle_lengths.argtypes = [f32p, i32p, i32p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        lib.mas_oracle_lengths.restype = None
        lib.mas_oracle_neg_cent.argtypes = [f32p, f32p, f32p, f32p] + [ctypes.c_int] * 4
        lib.mas_oracle_neg_cent.restype = None
        _lib = lib
    return _lib


def _f32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _i32(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def maximum_path_c(neg_cent: np.ndarray, t_ys: np.ndarray, t_xs: np.ndarray, strict: bool = True) -> np.ndarray:
    """core.pyx:38-42 over a batch.  neg_cent [B,T,S] float32 is NOT modified
    (a scratch copy is, like __init__.py:13).  Returns int32 path [B,T,S]."""
    neg_cent = np.ascontiguousarray(neg_cent, dtype=np.float32)
    B, T, S = neg_cent.shape
    values = neg_cent.copy()
    paths = np.zeros((B, T, S), dtype=np.int32)
    t_ys = np.ascontiguousarray(t_ys, dtype=np.int32)
    t_xs = np.ascontiguousarray(t_xs, dtype=np.int32)
    bad = _load().mas_oracle_batch(_i32(paths), _f32(values), _i32(t_ys), _i32(t_xs), B, T, S)
    if bad and strict:
        raise ValueError(f"utterance {bad - 1}: lengths are undefined behaviour in the reference "
                         f"(need 1 <= t_x <= t_y <= T and t_x <= S)")
    return paths


def lengths_from_mask(mask: np.ndarray):
    """__init__.py:16-17."""
    mask = np.ascontiguousarray(mask, dtype=np.float32)
    B, T, S = mask.shape
    t_ys = np.zeros(B, np.int32)
    t_xs = np.zeros(B, np.int32)
    _load().mas_oracle_lengths(_f32(mask), _i32(t_ys), _i32(t_xs), B, T, S)
    return t_ys, t_xs


def maximum_path(neg_cent: np.ndarray, mask: np.ndarray) -> np.ndarray:
    """__init__.py:6-19 on numpy arrays (int32 path; the dtype cast of :19 is the caller's)."""
    t_ys, t_xs = lengths_from_mask(mask)
    return maximum_path_c(neg_cent, t_ys, t_xs)


def neg_cent_f64(z_p: np.ndarray, m_p: np.ndarray, logs_p: np.ndarray) -> np.ndarray:
    """models.py:1226-1239 accumulated in fp64 (accuracy yardstick)."""
    z_p = np.ascontiguousarray(z_p, np.float32)
    m_p = np.ascontiguousarray(m_p, np.float32)
    logs_p = np.ascontiguousarray(logs_p, np.float32)
    B, D, T = z_p.shape
    S = m_p.shape[2]
    out = np.empty((B, T, S), np.float32)
    _load().mas_oracle_neg_cent(_f32(z_p), _f32(m_p), _f32(logs_p), _f32(out), B, D, T, S)
    return out


def neg_cent_torch(z_p, m_p, logs_p):
    """models.py:1226-1239 restated op for op in torch fp32 (CPU): the value the
    reference itself feeds to maximum_path.  Term order follows :1239."""
    import torch

    inv_var = torch.exp(-2 * logs_p)                                            # :1226
    c1 = torch.sum(-0.5 * math.log(2 * math.pi) - logs_p, [1], keepdim=True)    # :1227-1229
    c2 = torch.matmul(-0.5 * (z_p ** 2).transpose(1, 2), inv_var)               # :1230-1232
    c3 = torch.matmul(z_p.transpose(1, 2), m_p * inv_var)                       # :1233-1235
    c4 = torch.sum(-0.5 * (m_p ** 2) * inv_var, [1], keepdim=True)              # :1236-1238
    return c1 + c2 + c3 + c4                                                    # :1239


def add_mas_noise(neg_cent, noise, mas_noise_scale):
    """models.py:1241-1247 with the randn_like draw supplied by the caller."""
    import torch

    eps = torch.std(neg_cent) * noise * mas_noise_scale
    return neg_cent + eps


def align_torch(z_p, m_p, logs_p, x_mask, y_mask, mas_noise_scale=None, noise=None):
    """models.py:1224-1256 as a unit on CPU tensors -> (attn [B,1,T,S], w [B,1,S], neg_cent)."""
    import torch

    with torch.no_grad():
        nc = neg_cent_torch(z_p, m_p, logs_p)
        if mas_noise_scale is not None:
            nc = add_mas_noise(nc, noise, mas_noise_scale)
        attn_mask = torch.unsqueeze(x_mask, 2) * torch.unsqueeze(y_mask, -1)    # :1249
        path = maximum_path(nc.numpy(), attn_mask.squeeze(1).numpy())
        attn = torch.from_numpy(path).to(dtype=nc.dtype).unsqueeze(1)
    return attn, attn.sum(2), nc                                                # :1256


def expand_prior_torch(attn, m_p, logs_p):
    """models.py:1270-1271 as written: the prior statistics carried along the one-hot path (CPU torch;
    differentiable, so autograd of these two matmuls is the backward oracle as well)."""
    import torch

    m = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)        # :1270
    l = torch.matmul(attn.squeeze(1), logs_p.transpose(1, 2)).transpose(1, 2)     # :1271
    return m, l


def expand_prior_backward_np(g, dur):
    """Backward of expand_prior_torch with respect to m_p (or logs_p) from the durations: what autograd derives from
    models.py:1270-1271 is `attn^T g`, and with a monotonic one-hot path the frames of column s are the contiguous
    range [start_s, start_s + dur_s), start = exclusive prefix sum of the durations.  g [B,D,T] -> fp64 [B,D,S];
    negative durations count as 0, frames past T do not exist (the contract of mas_expand_prior_backward_f32)."""
    g = np.asarray(g, dtype=np.float64)
    dur = np.asarray(dur)
    B, D, T = g.shape
    S = dur.shape[1]
    out = np.zeros((B, D, S), dtype=np.float64)
    for b in range(B):
        start = 0
        for s in range(S):
            n = max(int(dur[b, s]), 0)
            lo, hi = min(start, T), min(start + n, T)
            if hi > lo:
                out[b, :, s] = g[b, :, lo:hi].sum(1)
            start += n
    return out


def logw_torch(attn, x_mask):
    """models.py:1256 + 1261: w = attn.sum(2); logw_ = log(w + 1e-6) * x_mask."""
    import torch

    w = attn.sum(2)
    return torch.log(w + 1e-6) * x_mask


def generate_path_torch(duration, mask):
    """commons.py:130-145 restated (CPU torch): duration [b,1,t_x], mask [b,1,t_y,t_x] -> path [b,1,t_y,t_x]."""
    import torch
    import torch.nn.functional as F

    b, _, t_y, t_x = mask.shape
    cum = torch.cumsum(duration, -1).view(b * t_x)                                        # :138-140
    path = (torch.arange(t_y, dtype=cum.dtype).unsqueeze(0) < cum.unsqueeze(1)).to(mask.dtype)   # sequence_mask :123-127
    path = path.view(b, t_x, t_y)
    path = path - F.pad(path, [0, 0, 1, 0, 0, 0])[:, :-1]                                  # :143
    return path.unsqueeze(1).transpose(2, 3) * mask                                        # :144


# --------------------------------------------------------------------------
# the compiled, unmodified reference kernel (oracle/_ref, built by build_ref.py)
# --------------------------------------------------------------------------
_ref_core = None


def ref_core():
    """Import oracle/_ref/core*.so (the reference's core.pyx compiled as shipped)
    or return None when it has not been built."""
    global _ref_core
    if _ref_core is None:
        so = os.path.join(HERE, "_ref", "core" + sysconfig.get_config_var("EXT_SUFFIX"))
        if not os.path.exists(so):
            return None
        spec = importlib.util.spec_from_file_location("core", so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_core = mod
    return _ref_core


def ref_maximum_path_c(neg_cent: np.ndarray, t_ys: np.ndarray, t_xs: np.ndarray) -> np.ndarray:
    """The reference kernel itself on a scratch copy (what __init__.py:13-18 does)."""
    core = ref_core()
    if core is None:
        raise RuntimeError("oracle/_ref is not built (run python oracle/build_ref.py where /root/reference exists)")
    values = np.ascontiguousarray(neg_cent, dtype=np.float32).copy()
    paths = np.zeros(values.shape, dtype=np.int32)
    core.maximum_path_c(paths, values, np.ascontiguousarray(t_ys, np.int32), np.ascontiguousarray(t_xs, np.int32))
    return paths
