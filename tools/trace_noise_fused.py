"""MAS_TRACE=1 (trace build): timeline of one single-launch noise-scaled alignment step at config 2."""
import os, sys, ctypes
os.environ["MAS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib
B, S, T, D = 64, 256, 1024, 192
dev = torch.device("cuda:0")
t_x, t_y = synthetic.full_lengths(B, S, T)
z, m, l, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=0)
noise = torch.randn((B, T, S), device=dev)
plan = tts.AlignPlan(B, D, T, S, dev, with_noise=True)
args = (z.to(dev), m.to(dev), l.to(dev), t_y.to(dev), t_x.to(dev), noise, 0.01)
L = _lib.lib()
for _ in range(3):
    plan.run(*args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); plan.run(*args); b.record(); torch.cuda.synchronize()
print(f"one eager step: {a.elapsed_time(b) * 1e3:.1f} us")
buf = np.zeros(1 << 16, dtype=np.uint64)
assert L.mas_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size) == 0
dp = buf[12288:12288 + B * 32].reshape(B, 32).astype(np.int64)
G = 148
sec = lambda k: buf[49152 + 256 * k:49152 + 256 * k + G].astype(np.int64)
ent, cdone, bar, nz, zf = (sec(k) for k in range(5))
t0 = ent.min()
rel = lambda x: (x - t0) / 1e3
def st(name, x):
    x = x[x > 0]
    if len(x): print(f"{name:28s} min {rel(x).min():7.1f} mean {rel(x).mean():7.1f} max {rel(x).max():7.1f}  (n={len(x)})")
st("CTA entry", ent)
st("contraction done", cdone)
st("barrier passed", bar)
st("noise applied (appliers)", nz[B:])
st("zero fill done (appliers)", zf[B:])
st("DP start", dp[:, 0])
for k in range(8):
    st(f"tile {k} acquired", dp[:, 2 + k])
st("forward end", dp[:, 1])
st("level-1 hops done", dp[:, 26])
st("level-2 walks done", dp[:, 27])
st("zero flag seen", dp[:, 28])
st("outputs end", dp[:, 30])
# contraction role unit marks of CTA 0 and a late CTA
for c in (0, 2, 146):
    u = buf[16384 + c * 64:16384 + c * 64 + 64].astype(np.int64).reshape(4, 16)
    for role, nm in ((0, "MMA"), (1, "epilogue")):
        v = u[role][u[role] > 0]
        print(f"cta {c} {nm} unit begin/end:", np.round(rel(v), 1).tolist())
tr = buf[40960:40960 + B * 16].reshape(B, 16).astype(np.int64)
n_steps = (T + 31) // 32 + 4
for name, off, labels in [("value warp 0", 0, ["tile wait", "compute", "bits/hop", "barrier"]),
                          ("producer", 8, ["zero-fill issue", "barrier", "tile issue"]),
                          ("helper / feeder thread 0", 12, ["wait (loads)", "read", "wait (credit)", "apply + store"]),
                          ("issue_tile", 0, None)]:
    if labels is None:
        for lab, slot in (("flag/pad/fence", 11), ("arrive.expect_tx", 14), ("bulk issue", 15)):
            print(f"   issue_tile {lab:18s} {tr[:, slot].mean():9.0f}")
        continue
    tot = tr[:, off:off + len(labels)].sum(1).mean()
    print(f"{name}: {tot:.0f} cycles in the step loop ({tot / n_steps:.0f} per step)")
    for j, lab in enumerate(labels):
        print(f"   {lab:16s} {tr[:, off + j].mean():9.0f} ({tr[:, off + j].mean() / max(tot, 1):5.1%})")
