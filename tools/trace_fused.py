"""MAS_TRACE=1 python tools/trace_fused.py: timeline of one fused step (tile publications, DP waits)."""
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
plan = tts.AlignPlan(B, D, T, S, dev)
args = (z.to(dev), m.to(dev), l.to(dev), t_y.to(dev), t_x.to(dev))
L = _lib.lib()
for _ in range(3):
    plan.run(*args)
torch.cuda.synchronize()
# clear trace, run once
buf = np.zeros(1 << 16, dtype=np.uint64)
plan.run(*args); torch.cuda.synchronize()
rc = L.mas_debug_read_trace(buf.ctypes.data_as(ctypes.c_void_p), buf.size)
assert rc == 0, rc
gem = buf[:148 * 64].reshape(148, 64)
dp = buf[12288:12288 + B * 32].reshape(B, 32).astype(np.int64)
t0 = dp[:, 0].min()
# counts accumulate over the 4 runs: last run's publications are the last (count/4) entries... use modular layout
cnt = gem[:, 0].astype(np.int64)
print("gemm ctas with publications:", (cnt > 0).sum(), "counts:", np.unique(cnt))
rel = lambda x: (x - t0) / 1e3
NRUN = 4
rows = []
for c in range(148):
    k = int(cnt[c]) // NRUN
    if k == 0: continue
    tt = gem[c, 1 + (NRUN - 1) * k: 1 + NRUN * k].astype(np.int64)
    rows.append(rel(tt))
mx = max(len(r) for r in rows)
for j in range(mx):
    col = np.array([r[j] for r in rows if len(r) > j])
    print(f"gemm publication {j}: n={len(col):3d} min {col.min():7.1f} mean {col.mean():7.1f} max {col.max():7.1f}")
ent = buf[49152:49152 + 148].astype(np.int64); left = buf[49152 + 256:49152 + 256 + 148].astype(np.int64)
zf = buf[49152 + 512:49152 + 512 + 148].astype(np.int64)
print("CTA entry (us rel): min %.1f max %.1f" % (rel(ent).min(), rel(ent).max()))
print("contraction role left: DP CTAs min %.1f mean %.1f max %.1f; others min %.1f mean %.1f max %.1f" % (
    rel(left[:B]).min(), rel(left[:B]).mean(), rel(left[:B]).max(), rel(left[B:]).min(), rel(left[B:]).mean(), rel(left[B:]).max()))
if (zf[B:] > 0).all():
    print("zero-fill role done: min %.1f mean %.1f max %.1f" % (rel(zf[B:]).min(), rel(zf[B:]).mean(), rel(zf[B:]).max()))
print("DP start  (us rel): min %.1f max %.1f" % (rel(dp[:, 0]).min(), rel(dp[:, 0]).max()))
for k in range(8):
    a = rel(dp[:, 2 + k])
    print(f"tile {k} acquired: min {a.min():7.1f} mean {a.mean():7.1f} max {a.max():7.1f}")
print("forward end: min %.1f mean %.1f max %.1f" % (rel(dp[:, 1]).min(), rel(dp[:, 1]).mean(), rel(dp[:, 1]).max()))
for name, k in (("level-1 hops done", 26), ("level-2 walks done", 27), ("zero flag seen", 28)):
    print("%-18s min %.1f mean %.1f max %.1f" % (name + ":", rel(dp[:, k]).min(), rel(dp[:, k]).mean(), rel(dp[:, k]).max()))
print("outputs end: min %.1f mean %.1f max %.1f" % (rel(dp[:, 30]).min(), rel(dp[:, 30]).mean(), rel(dp[:, 30]).max()))
tr = buf[40960:40960 + B * 16].reshape(B, 16).astype(np.int64)
n_steps = (T + 31) // 32 + 3
for name, off, labels in [("DP warp 0", 0, ["tile wait", "compute", "bits/hop", "barrier"]),
                          ("DP warp 3", 4, ["tile wait", "compute", "bits/hop", "barrier"]),
                          ("producer", 8, ["zero-fill issue", "barrier", "tile issue"])]:
    tot = tr[:, off:off + len(labels)].sum(1).mean()
    print(f"{name}: {tot:.0f} cycles in the step loop ({tot / n_steps:.0f} per step)")
    for j, lab in enumerate(labels):
        print(f"   {lab:16s} {tr[:, off + j].mean():9.0f} ({tr[:, off + j].mean() / tot:5.1%})")
for c in (0, 2, 70, 146):
    u = buf[16384 + c * 64:16384 + c * 64 + 64].astype(np.int64).reshape(4, 16)
    for role, nm in ((0, "MMA"), (1, "epilogue")):
        v = u[role][u[role] > 0]
        print(f"cta {c} {nm} unit begin/end:", np.round(rel(v), 1).tolist())
