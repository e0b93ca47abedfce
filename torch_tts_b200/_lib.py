"""ctypes binding of libmas_b200.so (the C ABI declared in include/mas_b200.h).

The library is built in-tree by `__graft_entry__.build()` / `python -m
torch_tts_b200.build`.  If it is missing this module raises: the product path
has no CPU or PyTorch fallback by design.
"""
from __future__ import annotations

import ctypes
import os
import threading

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAS_LIB_PATH") or os.path.join(HERE, "libmas_b200.so")   # override: kernel experiments

# every symbol include/mas_b200.h declares (tests check the library exports them all)
ABI_SYMBOLS = (
    "mas_b200_abi_version", "mas_status_string", "mas_last_cuda_error",
    "mas_lengths_from_mask_f32",
    "mas_maximum_path_workspace_bytes", "mas_maximum_path_f32",
    "mas_neg_cent_workspace_bytes", "mas_neg_cent_f32",
    "mas_fused_align_workspace_bytes", "mas_fused_align_f32",
    "mas_expand_path", "mas_expand_prior_f32", "mas_expand_prior_backward_f32", "mas_logw_f32",
    "mas_idx_from_durations_f32",
    "mas_take_launch_count", "mas_debug_read_trace", "mas_reload_config",
)

PATH_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.int32: 3}

_lib = None
_lock = threading.Lock()


class MasError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise MasError(
                        f"{LIB_PATH} is not built. Run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(nvcc, sm_100a). torch_tts_b200 has no CPU/PyTorch fallback.")
                L = ctypes.CDLL(LIB_PATH)
                vp, i32, sz, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
                L.mas_b200_abi_version.restype = i32
                L.mas_status_string.restype = ctypes.c_char_p
                L.mas_status_string.argtypes = [i32]
                L.mas_last_cuda_error.restype = ctypes.c_char_p
                L.mas_take_launch_count.restype = ctypes.c_long
                L.mas_lengths_from_mask_f32.restype = i32
                L.mas_lengths_from_mask_f32.argtypes = [vp, vp, vp, i32, i32, i32, vp]
                L.mas_maximum_path_workspace_bytes.restype = sz
                L.mas_maximum_path_workspace_bytes.argtypes = [i32, i32, i32]
                L.mas_maximum_path_f32.restype = i32
                L.mas_maximum_path_f32.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, sz, i32, i32, i32, vp]
                L.mas_neg_cent_workspace_bytes.restype = sz
                L.mas_neg_cent_workspace_bytes.argtypes = [i32, i32, i32, i32]
                L.mas_neg_cent_f32.restype = i32
                L.mas_neg_cent_f32.argtypes = [vp, vp, vp, vp, vp, vp, sz, i32, i32, i32, i32, vp]
                L.mas_fused_align_workspace_bytes.restype = sz
                L.mas_fused_align_workspace_bytes.argtypes = [i32, i32, i32, i32, i32]
                L.mas_fused_align_f32.restype = i32
                L.mas_fused_align_f32.argtypes = [vp, vp, vp, vp, vp, vp, f32, vp, i32, vp, vp, vp, vp, vp, sz,
                                                  i32, i32, i32, i32, vp]
                L.mas_reload_config.restype = None
                L.mas_debug_read_trace.restype = i32
                L.mas_debug_read_trace.argtypes = [vp, i32]
                L.mas_expand_path.restype = i32
                L.mas_expand_path.argtypes = [vp, vp, i32, i32, i32, i32, vp]
                L.mas_expand_prior_f32.restype = i32
                L.mas_expand_prior_f32.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
                L.mas_expand_prior_backward_f32.restype = i32
                L.mas_expand_prior_backward_f32.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
                L.mas_logw_f32.restype = i32
                L.mas_logw_f32.argtypes = [vp, vp, vp, i32, i32, vp]
                L.mas_idx_from_durations_f32.restype = i32
                L.mas_idx_from_durations_f32.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
                _lib = L
    return _lib


def reload_config():
    """Re-read the MAS_* tuning knobs from the environment (the library reads them once, on first use)."""
    lib().mas_reload_config()


def check(rc: int, what: str):
    if rc != 0:
        L = lib()
        msg = L.mas_status_string(rc).decode()
        if rc == 7:
            msg += ": " + L.mas_last_cuda_error().decode()
        raise MasError(f"{what} failed: {msg} (code {rc})")


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise MasError(f"{name} must be a CUDA tensor: torch_tts_b200 runs on B200 only and has no CPU fallback "
                       f"(got device {t.device})")


# one growing workspace per (device, stream): calls on one stream are ordered, so reuse is safe
_workspaces: dict = {}


def workspace(device, nbytes: int) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf
