"""Accuracy + timing of the neg_cent kernel against the torch-fp32 reference expression."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_tts_b200 as tts
from oracle import mas_oracle
from torch_tts_b200 import synthetic, _lib
dev = torch.device("cuda:0")
shapes = [(2, 64, 256, False), (3, 80, 300, True), (4, 256, 1024, False), (2, 200, 800, True), (2, 17, 50, False),
          (2, 600, 1000, False), (3, 300, 700, True), (1, 1024, 1030, False)]
if len(sys.argv) > 1 and sys.argv[1] == "big":
    shapes = [(64, 256, 1024, False)]
for B, S, T, ragged in shapes:
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 1) if ragged else synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=B + S)
    want = mas_oracle.neg_cent_torch(z_p, m_p, logs_p)
    zd, md, ld = z_p.to(dev), m_p.to(dev), logs_p.to(dev)
    got = tts.neg_cent(zd, md, ld)
    torch.cuda.synchronize()
    g = got.cpu()
    rel = ((g - want).abs() / want.abs().clamp_min(1.0))
    print(f"B={B} S={S} T={T}: max rel err {rel.max().item():.3e} mean {rel.mean().item():.3e} max abs {(g-want).abs().max().item():.3e} finite={bool(torch.isfinite(g).all())}")
    if rel.max().item() > 1e-3:
        bad = (rel > 1e-3).nonzero()
        print("   first bad", bad[:5].tolist(), "n bad", len(bad), "got", g[tuple(bad[0])].item(), "want", want[tuple(bad[0])].item())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): tts.neg_cent(zd, md, ld)
    a.record()
    for _ in range(10): tts.neg_cent(zd, md, ld)
    b.record(); torch.cuda.synchronize()
    print(f"   {a.elapsed_time(b)/10*1e3:.1f} us/call")
