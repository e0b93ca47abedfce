"""MAS_TRACE=1: per-role timeline of the contraction (standalone or inside the fused kernel)."""
import os, sys, ctypes
os.environ["MAS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib
B, S, T, D = 64, 256, 1024, 192
dev = torch.device("cuda:0")
fused = len(sys.argv) > 1 and sys.argv[1] == "fused"
t_x, t_y = synthetic.full_lengths(B, S, T)
z, m, l, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=0)
zd, md, ld = z.to(dev), m.to(dev), l.to(dev)
L = _lib.lib()
if fused:
    plan = tts.AlignPlan(B, D, T, S, dev)
    args = (zd, md, ld, t_y.to(dev), t_x.to(dev))
    run = lambda: plan.run(*args)
else:
    run = lambda: tts.neg_cent(zd, md, ld)
for _ in range(3): run()
torch.cuda.synchronize()
buf = np.zeros(1 << 16, dtype=np.uint64)
run(); torch.cuda.synchronize()
assert L.mas_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size) == 0
tr = buf[16384:16384 + 148 * 64].reshape(148, 4, 8, 2).astype(np.int64)
used = tr[:, 0, 0, 0] > 0
t0 = tr[used][:, :, 0, 0][tr[used][:, :, 0, 0] > 0].min()
names = ["mma", "epilogue", "converter"]
for cta in ([0, 1, 40, 41] if used.sum() > 41 else [0, 1]):
    print(f"--- cta {cta}")
    for r in range(3):
        line = []
        for u in range(8):
            a, b = tr[cta, r, u]
            if a > 0: line.append(f"[{(a - t0)/1e3:5.1f},{(b - t0)/1e3:5.1f}]")
        print(f"{names[r]:10s}", " ".join(line))
for r in range(3):
    for u in range(8):
        a = tr[used, r, u, 0]; b = tr[used, r, u, 1]
        ok = a > 0
        if ok.sum() == 0: continue
        print(f"{names[r]:10s} unit {u}: n={ok.sum():3d} begin mean {(a[ok]-t0).mean()/1e3:6.1f}  end mean {(b[ok]-t0).mean()/1e3:6.1f}  dur mean {(b[ok]-a[ok]).mean()/1e3:5.2f}")

ph = buf[32768:32768 + 148 * 16].reshape(148, 2, 8).astype(np.int64)
names = ["wait z", "LDS z", "wait empty", "convert+STS", "fence+arrive"]
for g in range(2):
    tot = ph[used, g, :5].sum(1).mean()
    print(f"converter group {g}: cycles per unit-run by phase (mean over CTAs), total {tot:.0f}")
    for j in range(5):
        print(f"   {names[j]:14s} {ph[used, g, j].mean():9.0f}  ({ph[used, g, j].mean() / tot:5.1%})")
