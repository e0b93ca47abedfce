"""GPU parity of the fused SynthesizerTrn alignment call (models.py:1224-1256):
neg_cent within 1e-4 relative of the reference expression, path agreement
>= 99.99 % of cells and an identical duration-sum invariant (BASELINE.json)."""
import numpy as np
import pytest
import torch

import torch_tts_b200 as tts
from oracle import mas_oracle
from torch_tts_b200 import synthetic

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4          # BASELINE.json north_star: neg_cent within 1e-4 relative
MIN_AGREE = 0.9999      # path cell agreement


def _rel_err(got, want):
    return ((got - want).abs() / want.abs().clamp_min(1.0)).max().item()


@pytest.mark.parametrize("B,S,T,ragged", [(4, 50, 200, True), (3, 130, 515, True), (8, 256, 1024, False),
                                          (2, 97, 333, False)])
def test_neg_cent_matches_reference_expression(cuda_device, B, S, T, ragged):
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 1) if ragged else synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=B + S)
    want = mas_oracle.neg_cent_torch(z_p, m_p, logs_p)
    got = tts.neg_cent(z_p.to(cuda_device), m_p.to(cuda_device), logs_p.to(cuda_device)).cpu()
    assert _rel_err(got, want) < REL_TOL
    # and against the fp64-accumulated yardstick (small cases only: it is a scalar loop)
    if B * S * T <= 4 * 130 * 515:
        f64 = torch.from_numpy(mas_oracle.neg_cent_f64(z_p.numpy(), m_p.numpy(), logs_p.numpy()))
        assert _rel_err(got, f64) < REL_TOL


@pytest.mark.parametrize("B,S,T,ragged", [(6, 80, 320, True), (4, 256, 1024, False), (5, 200, 800, True)])
@pytest.mark.parametrize("scale", [None, 0.01, 0])
def test_fused_align(cuda_device, B, S, T, ragged, scale):
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 2) if ragged else synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=7)
    noise = None
    if scale is not None:
        noise = torch.randn((B, T, S), generator=torch.Generator().manual_seed(5))
    attn_ref, w_ref, nc_ref = mas_oracle.align_torch(z_p, m_p, logs_p, x_mask, y_mask, scale, noise)
    d = cuda_device
    attn, w, (idx, dur, status), nc = tts.align(
        z_p.to(d), m_p.to(d), logs_p.to(d), x_mask.to(d), y_mask.to(d), scale,
        None if noise is None else noise.to(d), return_compact=True, return_neg_cent=True)
    assert attn.shape == (B, 1, T, S) and w.shape == (B, 1, S) and attn.dtype == z_p.dtype
    assert (status == 0).all()
    assert _rel_err(nc.cpu(), nc_ref) < REL_TOL
    agree = (attn.cpu() == attn_ref).float().mean().item()
    assert agree >= MIN_AGREE, agree
    assert torch.equal(w.sum((1, 2)).cpu(), w_ref.sum((1, 2)))          # duration-sum invariant
    assert torch.equal(w.cpu(), attn.sum(2).cpu())
    # the path is the exact MAS optimum of the cost the GPU itself produced
    want = mas_oracle.maximum_path_c(nc.cpu().numpy(), t_y.numpy(), t_x.numpy())
    assert np.array_equal(attn.squeeze(1).cpu().numpy().astype(np.int32), want)


def test_align_draws_noise_like_reference(cuda_device):
    """mas_noise_scale given, noise not: the wrapper draws randn on the device like models.py:1244."""
    B, S, T = 3, 40, 160
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 3)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=3)
    d = cuda_device
    attn, w = tts.align(z_p.to(d), m_p.to(d), logs_p.to(d), x_mask.to(d), y_mask.to(d), 0.01)
    assert torch.equal(w.sum((1, 2)).cpu().to(torch.int32), t_y)


def test_expand_path_roundtrip(cuda_device):
    B, S, T = 4, 33, 120
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 4)
    nc = synthetic.neg_cent_like(B, S, T, seed=4).to(cuda_device)
    path, dur, idx, _ = tts.maximum_path_compact(nc, t_y.to(cuda_device), t_x.to(cuda_device))
    again = tts.expand_path(idx, S, torch.float32)
    assert torch.equal(again, path)
    assert torch.equal(tts.expand_path(idx, S, torch.bfloat16).float(), path)
