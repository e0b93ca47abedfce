"""A few steps of one alignment call at a named config, for ncu: python tools/run_once.py c2 [noise] [compact]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
with_noise, compact = "noise" in sys.argv, "compact" in sys.argv
B, S, T, ragged = synthetic.CONFIGS[name]
dev = torch.device("cuda:0")
t_x, t_y = synthetic.config_lengths(name, seed=3)
z, m, l, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, synthetic.D_PRIOR, seed=0)
noise = torch.randn((B, T, S), device=dev) if with_noise else None
plan = tts.AlignPlan(B, synthetic.D_PRIOR, T, S, dev, with_noise=with_noise, want_path=not compact)
args = (z.to(dev), m.to(dev), l.to(dev), t_y.to(dev), t_x.to(dev), noise, 0.01 if with_noise else 0.0)
for _ in range(5):
    plan.run(*args)
torch.cuda.synchronize()
assert (plan.status == 0).all() and torch.equal(plan.dur.sum(1).cpu(), t_y)
print("ok")
