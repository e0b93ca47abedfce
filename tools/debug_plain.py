"""Trace build, MAS_TC_NO_TMA=2 (plain-store epilogue forced): which cells of neg_cent are wrong, and what are they?
MAS_LIB_PATH=torch_tts_b200/libmas_b200_trace.so MAS_TC_NO_TMA=2 python tools/debug_plain.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib
dev = torch.device("cuda:0")
B, S, T, D = int(os.environ.get("DB", 80)), 256, int(os.environ.get("DT", 384)), 192
t_x, t_y = synthetic.full_lengths(B, S, T)
z, m, l, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=1)
zd, md, ld = z.to(dev), m.to(dev), l.to(dev)
os.environ["MAS_TC_NO_TMA"] = "0"; _lib.reload_config()
ref = tts.neg_cent(zd, md, ld).cpu()
ref2 = tts.neg_cent(zd, md, ld).cpu()
print("tma path deterministic:", torch.equal(ref, ref2))
os.environ["MAS_TC_NO_TMA"] = "2"; _lib.reload_config()
# per-K-block contributions on the host (fp64), and the bias partials
r = torch.exp(-2 * l.double())                                   # [B,D,S]
zz = z.double()
def contrib(b, kb):                                             # [T,S] of K block kb (16 channels)
    sl = slice(16 * kb, 16 * kb + 16)
    return (-0.5 * zz[b, sl] ** 2).T @ r[b, sl] + zz[b, sl].T @ (m.double()[b, sl] * r[b, sl])
def bias(b):
    return ((-0.5 * np.log(2 * np.pi) - l.double()[b]).sum(0) + (-0.5 * m.double()[b] ** 2 * r[b]).sum(0))   # [S]
for run in range(3):
    got = tts.neg_cent(zd, md, ld).cpu()
    bad = (got != ref)
    print(f"run {run}: wrong cells {int(bad.sum())} of {bad.numel()}")
    if not bad.any():
        continue
    bb, tt, ss = np.nonzero(bad.numpy())
    groups = {}
    for b_, t_ in zip(bb, tt):
        groups.setdefault((int(b_), int(t_) // 32), 0)
        groups[(int(b_), int(t_) // 32)] += 1
    print("  (utterance, 32-row group): cells", sorted(groups.items())[:24], "...", len(groups), "groups")
    for (b_, g_), n in sorted(groups.items())[:6]:
        rows = slice(32 * g_, min(32 * g_ + 32, T))
        err = (got[b_, rows].double() - ref[b_, rows].double())          # [rows,S]
        print(f"  b={b_} rows {32 * g_}..: cells {n}, |err| mean {err.abs().mean():.4f}, ref |.| mean {ref[b_, rows].abs().mean():.2f}")
        sv = torch.linalg.svdvals(err)
        print("    singular values of the error block:", [round(float(x), 3) for x in sv[:20]])
        print("    row-to-row variation:", float(err.std(0).mean()), " column-to-column variation:", float(err.std(1).mean()))
        # is it a stale A stage?  err = (A_stale - A) . B_kb for one K block: project err onto the row space of B_kb
        Bm = torch.cat([r[b_], m.double()[b_] * r[b_]], 0)                 # [2D, S]: rows d (z^2 part), D + d (z part)
        for kb in range(12):
            idx = list(range(16 * kb, 16 * kb + 16)) + list(range(D + 16 * kb, D + 16 * kb + 16))
            Bk = Bm[idx]                                                  # [32, S]
            coef = err @ torch.linalg.pinv(Bk)                            # [rows, 32]
            res = (err - coef @ Bk).abs().mean()
            if res < 0.2 * err.abs().mean():
                print(f"    K block {kb}: error lies in the row space of its B rows (residual {float(res):.4f})")
                # coef = A_stale - A: compare A_stale with every (b2, tile, same rows) candidate
                A_true = torch.cat([-0.5 * zz[b_, 16 * kb:16 * kb + 16, rows] ** 2, zz[b_, 16 * kb:16 * kb + 16, rows]], 0).T
                A_stale = coef + A_true
                best = []
                for b2 in range(B):
                    for mt2 in range((T + 127) // 128):
                        for kb2 in (kb,):
                            r0 = mt2 * 128 + (32 * g_) % 128
                            if r0 + 32 > T:
                                continue
                            cand = torch.cat([-0.5 * zz[b2, 16 * kb2:16 * kb2 + 16, r0:r0 + 32] ** 2, zz[b2, 16 * kb2:16 * kb2 + 16, r0:r0 + 32]], 0).T
                            best.append((float((cand - A_stale).abs().mean()), b2, mt2))
                best.sort()
                print("      stale A candidates (residual, utterance, tile):", best[:3], " zero A residual:", float(A_stale.abs().mean()))
                # which channels of the K block are wrong, and where do their z values come from?
                dz = coef[:, 16:]                                         # z part: [rows, 16]
                wrong = [d for d in range(16) if dz[:, d].abs().max() > 1e-3]
                print("      wrong channels of the K block:", wrong)
                n32 = T // 32
                cand = zz[:, :, :n32 * 32].reshape(B, D, n32, 32)          # [B, D, T/32, 32]
                for d in wrong[:3]:
                    zs_ = (dz[:, d] + zz[b_, 16 * kb + d, rows])           # the stale z values [32]
                    dist = (cand - zs_[None, None, None, :]).abs().mean(-1)  # [B, D, T/32]
                    k = int(dist.argmin())
                    b2, ch2, g2 = k // (D * n32), (k // n32) % D, k % n32
                    print(f"        channel {d} (global {16 * kb + d}) holds z[b={b2}, ch={ch2} (kb {ch2 // 16}, d {ch2 % 16}), rows {32 * g2}..] "
                          f"(distance {float(dist.min()):.5f}; |z_stale| mean {float(zs_.abs().mean()):.3f})")
        # columns affected
        cols = np.nonzero(bad[b_, rows].any(0).numpy())[0]
        print("    columns affected:", cols.min(), "..", cols.max(), "count", len(cols), " rows affected:", np.nonzero(bad[b_, rows].any(1).numpy())[0][[0, -1]])
