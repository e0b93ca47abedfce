"""MAS_TRACE=1: where the DP kernel's cycles go (standalone maximum_path at config 2)."""
import os, sys, ctypes
os.environ["MAS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib
B, S, T, ragged = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda:0")
t_x, t_y = synthetic.full_lengths(B, S, T)
nc = (torch.randn((B, T, S), device=dev) * 50 - 470)
L = _lib.lib()
for _ in range(3): tts.maximum_path_compact(nc, t_y.to(dev), t_x.to(dev))
torch.cuda.synchronize()
buf = np.zeros(1 << 16, dtype=np.uint64)
assert L.mas_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size) == 0
tr = buf[40960:40960 + B * 16].reshape(B, 16).astype(np.int64)
n_steps = (T + 31) // 32 + 3
for name, off, labels in [("DP warp 0", 0, ["tile wait", "compute", "bits/hop", "barrier"]),
                          ("DP warp 3", 4, ["tile wait", "compute", "bits/hop", "barrier"]),
                          ("producer", 8, ["zero-fill issue", "barrier", "tile issue"])]:
    tot = tr[:, off:off + len(labels)].sum(1).mean()
    print(f"{name}: {tot:.0f} cycles in the step loop ({tot / n_steps:.0f} per step)")
    for j, lab in enumerate(labels):
        print(f"   {lab:16s} {tr[:, off + j].mean():9.0f} ({tr[:, off + j].mean() / tot:5.1%})")

if os.environ.get("MAS_DP_VK", "1") == "1":
    print("warp split (per utterance, cycles):")
    for name, off, labels in [("value warp 0", 0, ["tile wait", "compute", "bits", "barrier"]), ("origin warp 0", 4, ["compute", "barrier"])]:
        print("  " + name + ": " + ", ".join(f"{lab} {tr[:, off + j].mean():.0f}" for j, lab in enumerate(labels)))
