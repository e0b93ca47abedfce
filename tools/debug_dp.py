import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_tts_b200 as tts
from oracle import mas_oracle
from torch_tts_b200 import synthetic
dev = torch.device("cuda:0")
def case(B, S, T, ragged, seed, ties=False):
    nc = synthetic.neg_cent_like(B, S, T, seed=seed, ties=ties)
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed) if ragged else synthetic.full_lengths(B, S, T)
    want = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
    for rep in range(2):
        path, dur, idx, status = tts.maximum_path_compact(nc.to(dev), t_y.to(dev), t_x.to(dev))
        got = path.cpu().numpy().astype(np.int32)
        widx = np.where(want.sum(2) > 0, want.argmax(2), -1)
        gidx = idx.cpu().numpy()
        bad_path = (got != want).sum(); bad_idx = (gidx != widx).sum()
        print(f"B={B} S={S} T={T} ragged={ragged} rep={rep}: path cells differ={bad_path} idx rows differ={bad_idx} rowsums_ok={np.array_equal(got.sum(2), want.sum(2))}")
        if bad_idx:
            b, y = np.argwhere(gidx != widx)[0]
            ys = np.argwhere(gidx[b] != widx[b])[:, 0]
            print("   first utt", b, "t_x", int(t_x[b]), "t_y", int(t_y[b]), "rows", ys[:10], "...", ys[-5:], "n", len(ys))
            print("   got ", gidx[b, ys[:10]], " want", widx[b, ys[:10]])
        elif bad_path:
            b, y, x = np.argwhere(got != want)[0]
            print("   idx ok but path differs: first", b, y, x, "got", got[b, y, x], "row sum", got[b, y].sum())
case(2, 64, 64, False, 2064)
case(1, 1024, 1030, False, 2024)
case(16, 200, 800, True, 0, True)
case(9, 190, 999, True, 2)
case(2, 256, 1024, False, 2256)
