"""One small pass over every kernel of the hot path, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool memcheck python tools/sanitize_small.py
Shapes are small (the tools slow kernels down 10-100x) but cover: the fused contraction + DP kernel (flags between
CTAs, TMA loads/stores, tcgen05), the single-launch noise kernel (grid barrier, noise appliers), the separate
contraction / DP kernels (with noise inside the DP), unaligned S / T, compact outputs, the consumers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib
from oracle import mas_oracle

dev = torch.device("cuda:0")
which = sys.argv[1:] or ["fused", "noise", "separate", "odd", "mas", "consumers"]


def check(name, B, S, T, scale, **kw):
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 1)
    z, m, l, xm, ym = synthetic.prior_inputs(B, S, T, t_x, t_y, 192, seed=1)
    nz = None if scale is None else torch.randn((B, T, S), generator=torch.Generator().manual_seed(2))
    args = [t.to(dev) for t in (z, m, l, xm, ym)]
    attn, w, (idx, dur, st), nc = tts.align(*args, scale, None if nz is None else nz.to(dev), return_compact=True,
                                            return_neg_cent=True, **kw)
    attn2, w2, (idx2, dur2, st2) = tts.align(*args, scale, None if nz is None else nz.to(dev), return_compact=True, **kw)
    torch.cuda.synchronize()
    want = mas_oracle.maximum_path_c(nc.cpu().numpy(), t_y.numpy(), t_x.numpy())
    ok = np.array_equal(attn.squeeze(1).cpu().numpy().astype(np.int32), want) and torch.equal(idx, idx2)
    print(f"{name}: B={B} S={S} T={T} scale={scale}: {'ok' if ok else 'MISMATCH'}", flush=True)
    assert ok


if "fused" in which:
    check("fused", 6, 64, 256, None)
if "noise" in which:
    check("noise (single launch)", 6, 64, 256, 0.01)
if "separate" in which:
    os.environ["MAS_NO_FUSED"] = "1"
    _lib.reload_config()
    check("separate launches", 4, 64, 256, None)
    check("separate launches + noise in the DP", 4, 64, 256, 0.01)
    os.environ.pop("MAS_NO_FUSED")
    _lib.reload_config()
if "odd" in which:
    check("odd shapes", 3, 37, 131, None)
    check("odd shapes + noise", 3, 37, 131, 0.01)
if "mas" in which:
    B, S, T = 3, 300, 700
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 3)
    nc = synthetic.neg_cent_like(B, S, T, seed=3)
    p, dur, idx, st = tts.maximum_path_compact(nc.to(dev), t_y.to(dev), t_x.to(dev))
    n, dur2, idx2, _ = tts.maximum_path_compact(nc.to(dev), t_y.to(dev), t_x.to(dev), want_path=False)
    want = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
    assert np.array_equal(p.cpu().numpy().astype(np.int32), want) and torch.equal(idx, idx2)
    print("maximum_path (W=2, C=5; compact): ok", flush=True)
if "consumers" in which:
    B, S, T = 2, 50, 200
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 4)
    nc = synthetic.neg_cent_like(B, S, T, seed=4)
    p, dur, idx, st = tts.maximum_path_compact(nc.to(dev), t_y.to(dev), t_x.to(dev))
    m = torch.randn((B, 8, S), device=dev, requires_grad=True)
    l = torch.randn((B, 8, S), device=dev, requires_grad=True)
    a, b = tts.expand_prior(m, l, idx, dur)
    (a.sum() + b.sum()).backward()
    tts.logw(dur, t_x.to(dev))
    tts.generate_path(dur.float().unsqueeze(1), torch.ones((B, 1, T, S), device=dev))
    torch.cuda.synchronize()
    print("consumers: ok", flush=True)
print("sanitize_small done")
