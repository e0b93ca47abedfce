"""Debug: c3 ragged no-noise dense path vs compact idx (where do they differ?)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib

dev = torch.device("cuda:0")
B, S, T = 128, 256, 1024
seed = int(os.environ.get("SEED", 31))
t_x, t_y = synthetic.config_lengths("c3", seed=3)
z, m, l, xm, ym = synthetic.prior_inputs(B, S, T, t_x, t_y, 192, seed=seed)
args = [t.to(dev) for t in (z, m, l, xm, ym)]


def run(tag, **kw):
    for rep in range(3):
        # poison so a missing zero fill shows
        junk = torch.full((B, T, S), 7.0, device=dev)
        del junk
        attn, w, (idx, dur, st) = tts.align(*args, None, None, return_compact=True, **kw)
        torch.cuda.synchronize()
        want = tts.expand_path(idx, S)
        diff = (attn.squeeze(1) != want)
        n = int(diff.sum())
        print(f"{tag} rep {rep}: differing cells {n}", flush=True)
        if n:
            per_b = diff.flatten(1).sum(1).cpu().numpy()
            bad = np.nonzero(per_b)[0]
            print("  utterances", bad[:20], "counts", per_b[bad][:20])
            b = int(bad[0])
            d = diff[b].nonzero().cpu().numpy()
            print("  first utt", b, "t_y", int(t_y[b]), "t_x", int(t_x[b]), "cells (y,x):", d[:12].tolist(), "...", d[-4:].tolist())
            vals = attn.squeeze(1)[b][diff[b]].cpu().numpy()
            print("  values there:", np.unique(vals)[:8], "want", np.unique(want[b][diff[b]].cpu().numpy()))


run("default")
for env in ({"MAS_FUSED_DP_CTAS": "64"}, {"MAS_FUSED_DP_CTAS": "128"}, {"MAS_NO_FUSED": "1"}):
    os.environ.update(env)
    _lib.reload_config()
    run(str(env))
    for k in env:
        del os.environ[k]
