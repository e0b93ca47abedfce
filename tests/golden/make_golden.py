#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference itself.

Runs only where /root/reference exists (the authoring container); the GPU box
and the test-suite use the committed .npz files.  Nothing from the reference is
copied into the repo: its monotonic_align package is built as shipped
(vits2/monotonic_align/setup.py, `python setup.py build_ext --inplace`) in a
temporary directory, its models.py is imported from where it lies, and only
input/output ARRAYS are saved.

Fixtures written
----------------
mas_small.npz     monotonic_align.maximum_path(neg_cent, mask) (the reference's
                  Python wrapper + Cython kernel, __init__.py:6-19) on small
                  explicit cost planes: ragged, ties, edge lengths.  Inputs and
                  outputs stored in full.
mas_seeded.npz    the same call on BASELINE config 1 (B=16, S=200, T=800, full
                  lengths) and a ragged B=12 batch, inputs regenerated from a
                  torch.Generator seed (torch_tts_b200.synthetic), outputs
                  stored compactly (column index per mel row, int16).
synth_align.npz   the real SynthesizerTrn.forward (models.py:1197-1290) on CPU
                  with random init: z_p, m_p, logs_p, x_mask, y_mask captured at
                  models.py:1220-1222, the neg_cent and mask it hands to
                  maximum_path (:1250), the randn_like draw (:1244) and the
                  attn / w it returns, for mas_noise_scale in {None, 0.01, 0}.

consumers.npz     the alignment's consumers: commons.generate_path (commons.py:130-145) on seeded ceil()ed
                  durations, and the statements of models.py:1256, 1261, 1270-1271 (w, logw_, the two
                  one-hot matmuls) evaluated on that path.  Inputs and outputs stored in full.

Usage:  python tests/golden/make_golden.py [--only consumers]
"""
from __future__ import annotations

import importlib
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/vits2"
sys.path.insert(0, ROOT)

from torch_tts_b200 import synthetic  # noqa: E402  (seeded input generators only; no kernels)


def build_reference_monotonic_align(tmp: str):
    """vits2/README.md:21-26: cd monotonic_align; mkdir monotonic_align; python setup.py build_ext --inplace"""
    pkg = os.path.join(tmp, "monotonic_align")
    shutil.copytree(os.path.join(REF, "monotonic_align"), pkg)
    os.makedirs(os.path.join(pkg, "monotonic_align"), exist_ok=True)
    env = dict(os.environ, CC="/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc")
    subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=pkg, check=True,
                   capture_output=True, env=env)
    sys.path.insert(0, tmp)
    return importlib.import_module("monotonic_align")


def dense_mask(t_x, t_y, S, T):
    x_mask, y_mask = synthetic.masks(t_x, t_y, S, T)
    return (x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)).squeeze(1)      # models.py:1249,1251


def compact(path: np.ndarray) -> np.ndarray:
    """[B,T,S] {0,1} -> int16 [B,T] column per row, -1 for empty rows."""
    return np.where(path.sum(2) > 0, path.argmax(2), -1).astype(np.int16)


def gen_mas_small(ma):
    out = {}
    cases = [
        # name, B, S, T, ragged seed or explicit lengths, ties
        ("ragged_a", 5, 23, 97, 1, False),
        ("ragged_ties", 4, 31, 120, 2, True),
        ("full", 3, 16, 40, None, False),
        ("full_ties", 2, 64, 64, None, True),
    ]
    for name, B, S, T, seed, ties in cases:
        nc = synthetic.neg_cent_like(B, S, T, seed=len(name) * 7 + B, ties=ties)
        t_x, t_y = synthetic.ragged_lengths(B, S, T, seed) if seed is not None else synthetic.full_lengths(B, S, T)
        mask = dense_mask(t_x, t_y, S, T)
        path = ma.maximum_path(nc, mask)
        assert path.dtype == nc.dtype
        out[f"{name}/neg_cent"] = nc.numpy()
        out[f"{name}/t_x"] = t_x.numpy()
        out[f"{name}/t_y"] = t_y.numpy()
        out[f"{name}/path"] = path.numpy().astype(np.int8)
    # edge lengths: t_x == t_y, t_x == 1, t_y == 1, one-off-diagonal
    S, T = 12, 30
    t_x = torch.tensor([12, 1, 1, 7, 12, 2, 11], dtype=torch.int32)
    t_y = torch.tensor([12, 30, 1, 7, 30, 2, 12], dtype=torch.int32)
    nc = synthetic.neg_cent_like(len(t_x), S, T, seed=5)
    path = ma.maximum_path(nc, dense_mask(t_x, t_y, S, T))
    out["edges/neg_cent"] = nc.numpy()
    out["edges/t_x"] = t_x.numpy()
    out["edges/t_y"] = t_y.numpy()
    out["edges/path"] = path.numpy().astype(np.int8)
    # non-finite cells: the compiled comparisons decide (core.pyx:28,32)
    S, T = 10, 40
    t_x, t_y = synthetic.full_lengths(2, S, T)
    nc = synthetic.neg_cent_like(2, S, T, seed=11)
    nc[0, 13, 4] = float("nan")
    nc[0, 14, 5] = float("-inf")
    nc[1, 5, 3] = float("inf")
    path = ma.maximum_path(nc, dense_mask(t_x, t_y, S, T))
    out["nonfinite/neg_cent"] = nc.numpy()
    out["nonfinite/t_x"] = t_x.numpy()
    out["nonfinite/t_y"] = t_y.numpy()
    out["nonfinite/path"] = path.numpy().astype(np.int8)
    np.savez_compressed(os.path.join(HERE, "mas_small.npz"), **out)
    return len(out)


def gen_mas_seeded(ma):
    out = {}
    # BASELINE config 1 and a ragged batch; inputs are regenerated from the seed by the tests
    for name, B, S, T, ragged, seed, ties in [("c1", 16, 200, 800, False, 0, False),
                                               ("c1_ties", 16, 200, 800, False, 1, True),
                                               ("ragged12", 12, 256, 1024, True, 3, False)]:
        nc = synthetic.neg_cent_like(B, S, T, seed=seed, ties=ties)
        t_x, t_y = synthetic.ragged_lengths(B, S, T, seed) if ragged else synthetic.full_lengths(B, S, T)
        path = ma.maximum_path(nc, dense_mask(t_x, t_y, S, T)).numpy()
        out[f"{name}/shape"] = np.array([B, S, T, int(ragged), seed, int(ties)], np.int32)
        out[f"{name}/idx"] = compact(path)
        out[f"{name}/nc_checksum"] = np.array([float(nc.double().sum()), float(nc.double().abs().max())])
    np.savez_compressed(os.path.join(HERE, "mas_seeded.npz"), **out)
    return len(out)


def gen_synth_align(ma):
    """The real model, spied at the alignment call."""
    sys.path.insert(0, REF)
    import commons  # noqa: F401  (reference module)
    import models   # reference vits2/models.py

    torch.manual_seed(1234)
    net = models.SynthesizerTrn(
        n_vocab=60, spec_channels=80, segment_size=8192 // 256, inter_channels=192, hidden_channels=192,
        filter_channels=768, n_heads=2, n_layers=6, kernel_size=3, p_dropout=0.1, resblock="1",
        resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
        upsample_rates=[8, 8, 2, 2], upsample_initial_channel=512, upsample_kernel_sizes=[16, 16, 4, 4],
        n_speakers=0, gin_channels=0, use_sdp=True, use_transformer_flows=True, transformer_flow_type="pre_conv",
    )
    net.eval()
    B, S, T = 3, 24, 96
    g = torch.Generator().manual_seed(7)
    x_lengths = torch.tensor([24, 17, 9])
    y_lengths = torch.tensor([96, 70, 41])
    x = torch.randint(1, 60, (B, S), generator=g)
    y = torch.randn((B, 80, T), generator=g)
    for b in range(B):
        x[b, x_lengths[b]:] = 0
        y[b, :, y_lengths[b]:] = 0

    captured = {}
    real_mp = models.monotonic_align.maximum_path
    real_randn_like = torch.randn_like
    real_flow = net.flow.forward
    real_enc_p = net.enc_p.forward

    def spy_mp(neg_cent, mask):
        captured["neg_cent"] = neg_cent.detach().clone()
        captured["mask"] = mask.detach().clone()
        return real_mp(neg_cent, mask)

    def spy_randn_like(t, *a, **k):
        r = real_randn_like(t, *a, **k)
        if t.dim() == 3 and t.shape[1:] == (T, S):
            captured["noise"] = r.detach().clone()
        return r

    def spy_flow(*a, **k):
        r = real_flow(*a, **k)
        if not k.get("reverse", False):
            captured["z_p"] = r.detach().clone()
        return r

    def spy_enc_p(*a, **k):
        r = real_enc_p(*a, **k)
        captured["m_p"], captured["logs_p"], captured["x_mask"] = (t.detach().clone() for t in r[1:4])
        return r

    models.monotonic_align.maximum_path = spy_mp
    torch.randn_like = spy_randn_like
    net.flow.forward = spy_flow
    net.enc_p.forward = spy_enc_p
    out = {}
    try:
        for tag, scale in [("none", None), ("s001", 0.01), ("s0", 0)]:
            captured.clear()
            torch.manual_seed(99)
            with torch.no_grad():
                res = net(x, x_lengths, y, y_lengths, mas_noise_scale=scale)
            attn, y_mask = res[2], res[5]
            assert torch.equal(attn.sum((1, 2, 3)).long(), y_lengths)
            ins = {"z_p": captured["z_p"], "m_p": captured["m_p"], "logs_p": captured["logs_p"],
                   "x_mask": captured["x_mask"], "y_mask": y_mask}
            for k, v in ins.items():            # identical in the three runs (same seed, eval mode): stored once
                if f"inputs/{k}" in out:
                    assert np.array_equal(out[f"inputs/{k}"], v.numpy()), k
                else:
                    out[f"inputs/{k}"] = v.numpy()
            out[f"{tag}/neg_cent"] = captured["neg_cent"].numpy()       # after the noise add, as handed to MAS
            out[f"{tag}/mask"] = captured["mask"].numpy().astype(np.int8)
            if scale is not None:
                out[f"{tag}/noise"] = captured["noise"].numpy()
            out[f"{tag}/scale"] = np.array([np.nan if scale is None else float(scale)])
            out[f"{tag}/attn"] = attn.squeeze(1).numpy().astype(np.int8)
            out[f"{tag}/w"] = attn.sum(2).squeeze(1).numpy()
    finally:
        models.monotonic_align.maximum_path = real_mp
        torch.randn_like = real_randn_like
    np.savez_compressed(os.path.join(HERE, "synth_align.npz"), **out)
    return len(out)


def gen_consumers():
    sys.path.insert(0, REF)
    commons = importlib.import_module("commons")
    g = torch.Generator().manual_seed(2024)
    B, S, T, D = 3, 23, 96, 5
    t_x = torch.tensor([23, 11, 17])
    x_mask = (torch.arange(S)[None, :] < t_x[:, None]).float().unsqueeze(1)
    w_ceil = torch.ceil(torch.rand((B, 1, S), generator=g) * 7.0) * x_mask
    w_ceil[0, 0, 2] = 0.0
    y_len = torch.clamp_min(torch.sum(w_ceil, [1, 2]), 1).long().clamp_max(T)          # models.py:1304
    y_mask = commons.sequence_mask(y_len, T).unsqueeze(1).to(x_mask.dtype)             # :1305
    attn_mask = torch.unsqueeze(x_mask, 2) * torch.unsqueeze(y_mask, -1)               # :1309
    attn = commons.generate_path(w_ceil, attn_mask)                                    # :1310
    m_p = torch.randn((B, D, S), generator=g) * x_mask
    logs_p = torch.randn((B, D, S), generator=g) * x_mask
    m_e = torch.matmul(attn.squeeze(1), m_p.transpose(1, 2)).transpose(1, 2)           # :1270 / :1312
    l_e = torch.matmul(attn.squeeze(1), logs_p.transpose(1, 2)).transpose(1, 2)        # :1271 / :1315
    w = attn.sum(2)                                                                    # :1256
    logw_ = torch.log(w + 1e-6) * x_mask                                               # :1261
    out = {"w_ceil": w_ceil.numpy(), "x_mask": x_mask.numpy(), "attn_mask": attn_mask.numpy(), "t_x": t_x.numpy(),
           "y_len": y_len.numpy(), "attn": attn.numpy(), "m_p": m_p.numpy(), "logs_p": logs_p.numpy(),
           "m_expanded": m_e.numpy(), "logs_expanded": l_e.numpy(), "w": w.numpy(), "logw_": logw_.numpy()}
    np.savez_compressed(os.path.join(HERE, "consumers.npz"), **out)
    return len(out)


def main():
    if not os.path.isdir(REF):
        sys.exit("/root/reference is absent: golden vectors can only be regenerated in the authoring container")
    torch.set_num_threads(4)
    if "--only" in sys.argv and sys.argv[sys.argv.index("--only") + 1] == "consumers":
        print("consumers.npz arrays:", gen_consumers())
        return
    with tempfile.TemporaryDirectory(prefix="mas_golden_") as tmp:
        ma = build_reference_monotonic_align(tmp)
        # models.py does `import monotonic_align` (models.py:12): make it resolve to the build above
        sys.modules["monotonic_align"] = ma
        n1 = gen_mas_small(ma)
        n2 = gen_mas_seeded(ma)
        n3 = gen_synth_align(ma)
    gen_consumers()
    for f in ("mas_small.npz", "mas_seeded.npz", "synth_align.npz", "consumers.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
    print("arrays:", n1, n2, n3)


if __name__ == "__main__":
    main()
