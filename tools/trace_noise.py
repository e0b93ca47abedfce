"""MAS_TRACE=1: where the DP kernel's cycles go on the noise-scaled path (contraction + statistics, then DP with the
noise tile streamed next to the cost tile) at config 2."""
import os, sys, ctypes
os.environ["MAS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib
B, S, T, ragged = synthetic.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "c2"]
dev = torch.device("cuda:0")
t_x, t_y = synthetic.full_lengths(B, S, T)
z, m, l, xm, ym = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=0)
noise = torch.randn((B, T, S), device=dev)
args = [t.to(dev) for t in (z, m, l, xm, ym)]
for _ in range(3): tts.align(*args, 0.01, noise)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): tts.align(*args, 0.01, noise)
b.record(); torch.cuda.synchronize()
print(f"align + noise: {a.elapsed_time(b) / 5 * 1e3:.1f} us per call (eager, includes host launch gaps)")
buf = np.zeros(1 << 16, dtype=np.uint64)
assert _lib.lib().mas_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size) == 0
tr = buf[40960:40960 + B * 16].reshape(B, 16).astype(np.int64)
for name, off, labels in [("DP warp 0", 0, ["tile wait", "compute", "bits/hop", "barrier"]),
                          ("producer", 8, ["zero-fill issue", "barrier", "tile issue"])]:
    tot = tr[:, off:off + len(labels)].sum(1).mean()
    print(f"{name}: {tot:.0f} cycles in the step loop")
    for j, lab in enumerate(labels):
        print(f"   {lab:16s} {tr[:, off + j].mean():9.0f} ({tr[:, off + j].mean() / tot:5.1%})")
