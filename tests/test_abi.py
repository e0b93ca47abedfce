"""The C-ABI boundary without a GPU: libmas_b200.so loads, exports every symbol
include/mas_b200.h declares, and rejects bad arguments on the host before any
CUDA work.  No compute calls here (those are the -m gpu tests)."""
import ctypes
import os
import re

import pytest
import torch

from torch_tts_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mas_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mas_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    assert _declared_functions() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    for name in _declared_functions():
        assert hasattr(L, name), name
    assert L.mas_b200_abi_version() >= 1


def test_no_torch_or_python_in_the_abi():
    """plain pointers and sizes only: the shared object must not link libtorch / libpython."""
    import subprocess

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libpython" not in out and "libc10" not in out


def test_status_strings():
    L = _lib.lib()
    seen = set()
    for code in range(0, 8):
        s = L.mas_status_string(code).decode()
        assert s and s != "unknown status"
        seen.add(s)
    assert len(seen) == 8
    assert L.mas_status_string(99).decode() == "unknown status"


def test_host_side_argument_errors():
    L = _lib.lib()
    null = None
    fake = ctypes.c_void_p(0x1000)         # never dereferenced: validation fails first
    odd = ctypes.c_void_p(0x1004)
    # null pointers
    assert L.mas_maximum_path_f32(null, fake, fake, fake, 0, null, null, null, fake, 0, 1, 8, 4, null) == 1
    assert L.mas_neg_cent_f32(null, fake, fake, fake, null, fake, 0, 1, 4, 8, 4, null) == 1
    assert L.mas_fused_align_f32(fake, fake, fake, null, fake, null, 0.0, fake, 0, null, null, null, null, fake, 0,
                                 1, 4, 8, 4, null) == 1
    assert L.mas_expand_path(null, fake, 0, 1, 8, 4, null) == 1
    assert L.mas_lengths_from_mask_f32(null, fake, fake, 1, 8, 4, null) == 1
    # bad / unsupported shapes
    assert L.mas_maximum_path_f32(fake, fake, fake, fake, 0, null, null, null, fake, 0, 0, 8, 4, null) == 2
    assert L.mas_maximum_path_f32(fake, fake, fake, fake, 0, null, null, null, fake, 0, 1, 8, 2000, null) == 3
    assert L.mas_maximum_path_f32(fake, fake, fake, fake, 0, null, null, null, fake, 0, 1, 70000, 4, null) == 3
    assert L.mas_neg_cent_f32(fake, fake, fake, fake, null, fake, 0, 1, 0, 8, 4, null) == 2
    # dtype, alignment
    assert L.mas_maximum_path_f32(fake, fake, fake, fake, 9, null, null, null, fake, 0, 1, 8, 4, null) == 6
    assert L.mas_maximum_path_f32(odd, fake, fake, fake, 0, null, null, null, fake, 0, 1, 8, 4, null) == 4
    assert L.mas_expand_path(fake, fake, 7, 1, 8, 4, null) == 6
    # fused: workspace too small is caught on the host
    assert L.mas_fused_align_f32(fake, fake, fake, fake, fake, null, 0.0, fake, 0, null, null, null, null, fake, 16,
                                 2, 192, 64, 16, null) == 5


def test_workspace_queries():
    L = _lib.lib()
    assert L.mas_maximum_path_workspace_bytes(0, 8, 4) == 0
    assert L.mas_maximum_path_workspace_bytes(1, 8, 4000) == 0           # S > MAS_MAX_TEXT
    B, D, T, S = 64, 192, 1024, 256
    dp = L.mas_maximum_path_workspace_bytes(B, T, S)
    cost = L.mas_neg_cent_workspace_bytes(B, D, T, S)
    fused = L.mas_fused_align_workspace_bytes(B, D, T, S, 0)
    assert dp > 0 and cost > 0
    assert fused >= dp + cost + B * T * S * 4                            # private neg_cent plane lives in it
    # long utterances spill direction bits to the workspace
    assert L.mas_maximum_path_workspace_bytes(32, 4000, 600) > 32 * 4000 * 600 // 8


def test_cpu_tensors_fail_loudly_no_fallback():
    import torch_tts_b200 as tts

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tts.maximum_path(torch.zeros(1, 4, 2), torch.ones(1, 4, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tts.align(torch.zeros(1, 4, 8), torch.zeros(1, 4, 2), torch.zeros(1, 4, 2), torch.ones(1, 1, 2),
                  torch.ones(1, 1, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tts.neg_cent(torch.zeros(1, 4, 8), torch.zeros(1, 4, 2), torch.zeros(1, 4, 2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tts.expand_path(torch.zeros(1, 4, dtype=torch.int32), 2)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torch_tts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\.*oracle|mas_oracle|libmas_oracle|#include.*oracle", src,
                                     flags=re.M), f"{f} uses the oracle"


def test_alignment_lengths_from_masks():
    """what mask.sum(1)[:,0] / mask.sum(2)[:,0] give for attn_mask = x_mask (x) y_mask (models.py:1249)."""
    from torch_tts_b200.align import _lengths
    from torch_tts_b200 import synthetic

    t_x, t_y = synthetic.ragged_lengths(9, 40, 170, 3)
    x_mask, y_mask = synthetic.masks(t_x, t_y, 40, 170)
    mask = (x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)).squeeze(1)
    ty, tx = _lengths(x_mask, y_mask)
    assert torch.equal(ty, mask.sum(1)[:, 0].to(torch.int32)) and torch.equal(tx, mask.sum(2)[:, 0].to(torch.int32))
    assert torch.equal(ty, t_y) and torch.equal(tx, t_x)
