"""Pins the CPU oracle (oracle/) to the reference itself.

1. against the golden vectors in tests/golden/*.npz, which tests/golden/make_golden.py
   produced by running the unmodified reference (its Cython maximum_path through its own
   Python wrapper, and the real SynthesizerTrn.forward);
2. against the reference kernel compiled as shipped into oracle/_ref (when present);
3. against tests/dp_model.py, the numpy model of the formulation the CUDA kernel uses
   (register row + 1 decision bit per cell + checkpointed two-level backtrack).

No GPU needed.  The reference ships no tests of its own for this path.
"""
import os

import numpy as np
import pytest
import torch

from oracle import mas_oracle
from torch_tts_b200 import synthetic

import dp_model

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _npz(name):
    return np.load(os.path.join(GOLD, name))


def _groups(z):
    return sorted({k.split("/")[0] for k in z.files})


# --------------------------------------------------------------------------
# 1. golden vectors
# --------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["ragged_a", "ragged_ties", "full", "full_ties", "edges", "nonfinite"])
def test_mas_small_golden(case):
    z = _npz("mas_small.npz")
    nc, t_x, t_y, want = (z[f"{case}/{k}"] for k in ("neg_cent", "t_x", "t_y", "path"))
    keep = nc.copy()
    got = mas_oracle.maximum_path_c(nc, t_y, t_x)
    assert got.dtype == np.int32
    assert np.array_equal(got, want.astype(np.int32))
    assert np.array_equal(nc.view(np.int32), keep.view(np.int32))       # oracle works on a scratch copy
    # and through the mask front end (__init__.py:16-17)
    x_mask, y_mask = synthetic.masks(torch.from_numpy(t_x), torch.from_numpy(t_y), nc.shape[2], nc.shape[1])
    mask = (x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)).squeeze(1).numpy()
    ty2, tx2 = mas_oracle.lengths_from_mask(mask)
    assert np.array_equal(ty2, t_y) and np.array_equal(tx2, t_x)
    assert np.array_equal(mas_oracle.maximum_path(nc, mask), want.astype(np.int32))


@pytest.mark.parametrize("case", ["c1", "c1_ties", "ragged12"])
def test_mas_seeded_golden(case):
    z = _npz("mas_seeded.npz")
    B, S, T, ragged, seed, ties = (int(v) for v in z[f"{case}/shape"])
    nc = synthetic.neg_cent_like(B, S, T, seed=seed, ties=bool(ties))
    # the regenerated input is the one the golden was made from
    chk = z[f"{case}/nc_checksum"]
    assert float(nc.double().sum()) == chk[0] and float(nc.double().abs().max()) == chk[1]
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed) if ragged else synthetic.full_lengths(B, S, T)
    got = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
    idx = np.where(got.sum(2) > 0, got.argmax(2), -1)
    assert np.array_equal(idx, z[f"{case}/idx"].astype(np.int64))
    assert np.array_equal(got.sum((1, 2)), t_y.numpy())


@pytest.mark.parametrize("tag", ["none", "s001", "s0"])
def test_synthesizer_golden(tag):
    """models.py:1224-1256 restated in oracle/mas_oracle.py against the real model's tensors."""
    z = _npz("synth_align.npz")
    z_p, m_p, logs_p, x_mask, y_mask = (torch.from_numpy(z[f"inputs/{k}"])
                                        for k in ("z_p", "m_p", "logs_p", "x_mask", "y_mask"))
    scale = z[f"{tag}/scale"][0]
    scale = None if np.isnan(scale) else float(scale)
    noise = torch.from_numpy(z[f"{tag}/noise"]) if scale is not None else None
    want_nc = torch.from_numpy(z[f"{tag}/neg_cent"])
    attn, w, nc = mas_oracle.align_torch(z_p, m_p, logs_p, x_mask, y_mask, scale, noise)
    # same torch expression; only the sgemm blocking (thread count) may differ
    rel = ((nc - want_nc).abs() / want_nc.abs().clamp_min(1.0)).max().item()
    assert rel < 2e-6, rel
    # the mask the model hands to maximum_path is x_mask (x) y_mask
    mask = (x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)).squeeze(1)
    assert np.array_equal(mask.numpy().astype(np.int8), z[f"{tag}/mask"])
    # MAS on the model's own neg_cent is bit-exact
    path = mas_oracle.maximum_path(want_nc.numpy(), mask.numpy())
    assert np.array_equal(path.astype(np.int8), z[f"{tag}/attn"])
    assert np.array_equal(path.sum(1).astype(np.float32), z[f"{tag}/w"])
    # the restated cost leads to the same alignment here as well
    assert np.array_equal(attn.squeeze(1).numpy().astype(np.int8), z[f"{tag}/attn"])
    # fp64 yardstick agrees with the model's fp32 neg_cent (no-noise case)
    if scale is None:
        f64 = torch.from_numpy(mas_oracle.neg_cent_f64(z_p.numpy(), m_p.numpy(), logs_p.numpy()))
        assert ((f64 - want_nc).abs() / want_nc.abs().clamp_min(1.0)).max().item() < 1e-5


def test_noise_scale_zero_still_takes_noise_branch():
    """cli.py:268-271 decays to int 0, not None: the golden 's0' cost equals the 'none' cost bit for bit
    only because std * noise * 0 == 0, and the oracle reproduces that."""
    z = _npz("synth_align.npz")
    assert np.array_equal(z["s0/neg_cent"], z["none/neg_cent"])
    assert not np.array_equal(z["s001/neg_cent"], z["none/neg_cent"])


# --------------------------------------------------------------------------
# 2. the compiled reference kernel (oracle/_ref)
# --------------------------------------------------------------------------
needs_ref = pytest.mark.skipif(mas_oracle.ref_core() is None, reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("B,S,T,seed,ties", [(6, 40, 170, 0, False), (6, 40, 170, 1, True), (3, 256, 1024, 2, False),
                                             (2, 300, 700, 3, True), (9, 7, 33, 4, False)])
def test_port_equals_compiled_reference(B, S, T, seed, ties):
    nc = synthetic.neg_cent_like(B, S, T, seed=seed, ties=ties).numpy()
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed)
    a = mas_oracle.maximum_path_c(nc, t_y.numpy(), t_x.numpy())
    b = mas_oracle.ref_maximum_path_c(nc, t_y.numpy(), t_x.numpy())
    assert np.array_equal(a, b)


@needs_ref
def test_port_equals_compiled_reference_nonfinite_and_extreme():
    rng = np.random.default_rng(0)
    B, S, T = 4, 21, 90
    nc = (rng.standard_normal((B, T, S)) * 50 - 470).astype(np.float32)
    nc[0, rng.integers(0, T, 20), rng.integers(0, S, 20)] = np.nan
    nc[1, rng.integers(0, T, 20), rng.integers(0, S, 20)] = -np.inf
    nc[2] = -3e7                                   # runs below the -1e9 sentinel
    nc[3] = np.round(nc[3] / 64) * 64              # heavy ties
    t_x = np.array([21, 20, 21, 11], np.int32)
    t_y = np.array([90, 77, 90, 45], np.int32)
    a = mas_oracle.maximum_path_c(nc, t_y, t_x)
    b = mas_oracle.ref_maximum_path_c(nc, t_y, t_x)
    assert np.array_equal(a, b)


def test_bad_lengths_flagged():
    nc = synthetic.neg_cent_like(3, 8, 10, seed=0).numpy()
    with pytest.raises(ValueError):
        mas_oracle.maximum_path_c(nc, np.array([10, 5, 10], np.int32), np.array([8, 6, 8], np.int32))
    p = mas_oracle.maximum_path_c(nc, np.array([10, 5, 10], np.int32), np.array([8, 6, 0], np.int32), strict=False)
    assert p[1].sum() == 0 and p[2].sum() == 0 and p[0].sum() == 10


# --------------------------------------------------------------------------
# 3. the kernel's formulation (numpy model) == oracle
# --------------------------------------------------------------------------
@pytest.mark.parametrize("S,T,t_x,t_y,s_pad,ties", [
    (17, 50, 17, 50, 128, False), (17, 50, 9, 31, 128, True), (40, 90, 40, 40, 128, False),
    (40, 90, 1, 90, 128, False), (130, 140, 129, 133, 256, True), (33, 100, 33, 64, 128, True),
    (33, 100, 20, 65, 128, False), (5, 33, 5, 32, 128, True),
])
def test_dp_model_matches_oracle(S, T, t_x, t_y, s_pad, ties):
    nc = synthetic.neg_cent_like(1, S, T, seed=S + T + t_x, ties=ties).numpy()
    want = mas_oracle.maximum_path_c(nc, np.array([t_y], np.int32), np.array([t_x], np.int32))[0]
    got = dp_model.dp_model(nc[0], t_y, t_x, s_pad)
    assert np.array_equal(got, want)


# --------------------------------------------------------------------------
# properties of the oracle output (the same ones the full-size GPU tests use)
# --------------------------------------------------------------------------
def test_path_invariants():
    B, S, T = 8, 50, 210
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 5)
    nc = synthetic.neg_cent_like(B, S, T, seed=5).numpy()
    p = mas_oracle.maximum_path_c(nc, t_y.numpy(), t_x.numpy())
    for b in range(B):
        ty, tx = int(t_y[b]), int(t_x[b])
        assert p[b, ty:].sum() == 0 and p[b, :, tx:].sum() == 0
        idx = p[b, :ty].argmax(1)
        assert (p[b, :ty].sum(1) == 1).all()
        assert idx[0] == 0 and idx[-1] == tx - 1
        assert set(np.diff(idx)) <= {0, 1}
        # optimality: the path's score is the DP optimum
        v = np.full(tx, -np.inf)
        v[0] = nc[b, 0, 0]
        for y in range(1, ty):
            prev = np.concatenate(([-np.inf], v[:-1]))
            v = nc[b, y, :tx].astype(np.float64) + np.maximum(v, prev)
        score = nc[b, np.arange(ty), idx].astype(np.float64).sum()
        assert abs(score - v[tx - 1]) < 1e-6 * abs(score)


def test_consumers_golden():
    """oracle restatements of commons.generate_path and models.py:1256/1261/1270-1271 against the reference's own
    outputs (tests/golden/consumers.npz, made by make_golden.py from /root/reference)."""
    z = _npz("consumers.npz")
    w_ceil, attn_mask, x_mask = (torch.from_numpy(z[k]) for k in ("w_ceil", "attn_mask", "x_mask"))
    attn = mas_oracle.generate_path_torch(w_ceil, attn_mask)
    assert np.array_equal(attn.numpy(), z["attn"])
    m_e, l_e = mas_oracle.expand_prior_torch(attn, torch.from_numpy(z["m_p"]), torch.from_numpy(z["logs_p"]))
    assert np.array_equal(m_e.numpy(), z["m_expanded"]) and np.array_equal(l_e.numpy(), z["logs_expanded"])
    assert np.array_equal(mas_oracle.logw_torch(attn, x_mask).numpy(), z["logw_"])
    # the compact form the kernels use says the same thing: a gather over idx is the one-hot matmul
    path = attn.squeeze(1).numpy()
    idx = np.where(path.sum(2) > 0, path.argmax(2), -1)
    m_p = z["m_p"]
    gathered = np.where(idx[:, None, :] >= 0, np.take_along_axis(m_p, np.maximum(idx, 0)[:, None, :].repeat(m_p.shape[1], 1), 2), 0)
    assert np.array_equal(gathered.astype(np.float32), z["m_expanded"])


def test_expand_prior_backward_restated_from_durations():
    """The durations form of the prior-expansion backward (oracle.expand_prior_backward_np, the checker of the CUDA
    segmented sum) against autograd of the reference's own two matmuls (models.py:1270-1271) on the golden path."""
    z = _npz("consumers.npz")
    attn = torch.from_numpy(z["attn"])
    m_p = torch.from_numpy(z["m_p"]).double().requires_grad_(True)
    logs_p = torch.from_numpy(z["logs_p"]).double().requires_grad_(True)
    m_e, l_e = mas_oracle.expand_prior_torch(attn.double(), m_p, logs_p)
    g = torch.Generator().manual_seed(4)
    gm, gl = torch.randn(m_e.shape, generator=g, dtype=torch.float64), torch.randn(l_e.shape, generator=g, dtype=torch.float64)
    (m_e * gm).sum().add((l_e * gl).sum()).backward()
    dur = attn.squeeze(1).sum(1).round().int().numpy()                 # w = attn.sum(2) of models.py:1256, [B,S]
    assert np.allclose(mas_oracle.expand_prior_backward_np(gm.numpy(), dur), m_p.grad.numpy(), rtol=0, atol=1e-12)
    assert np.allclose(mas_oracle.expand_prior_backward_np(gl.numpy(), dur), logs_p.grad.numpy(), rtol=0, atol=1e-12)

