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


@pytest.mark.parametrize("B,S,T", [(70, 64, 256), (9, 192, 384), (5, 100, 260), (3, 32, 132), (4, 256, 1100), (66, 32, 128),
                                   (1, 256, 1024), (2, 128, 4096), (150, 16, 64)])
def test_fused_shapes_and_many_utterances(cuda_device, B, S, T):
    """the one-kernel path beyond config 2: more utterances than DP CTAs (a CTA aligns several in turn),
    C = 1..4 columns per thread, a ragged last mel tile, odd numbers of mel tiles for the CTA pairs."""
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 5)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=11)
    d = cuda_device
    attn, w, (idx, dur, status), nc = tts.align(z_p.to(d), m_p.to(d), logs_p.to(d), x_mask.to(d), y_mask.to(d),
                                                return_compact=True, return_neg_cent=True)
    assert (status == 0).all()
    nc_ref = mas_oracle.neg_cent_torch(z_p, m_p, logs_p)
    assert _rel_err(nc.cpu(), nc_ref) < REL_TOL
    want = mas_oracle.maximum_path_c(nc.cpu().numpy(), t_y.numpy(), t_x.numpy())
    assert np.array_equal(attn.squeeze(1).cpu().numpy().astype(np.int32), want)
    assert torch.equal(dur.sum(1).cpu(), t_y)


def test_fused_without_neg_cent_out_skips_dead_tiles(cuda_device):
    """private cost plane: mel tiles past t_y are never computed; the result is unchanged."""
    B, S, T = 12, 96, 640
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 9)
    t_y = torch.clamp(t_y, max=200)               # most tiles are dead
    t_x = torch.minimum(t_x, t_y)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=2)
    d = cuda_device
    a1, w1 = tts.align(z_p.to(d), m_p.to(d), logs_p.to(d), x_mask.to(d), y_mask.to(d))
    a2, w2, nc = tts.align(z_p.to(d), m_p.to(d), logs_p.to(d), x_mask.to(d), y_mask.to(d), return_neg_cent=True)
    assert torch.equal(a1, a2) and torch.equal(w1, w2)
    want = mas_oracle.maximum_path_c(nc.cpu().numpy(), t_y.numpy(), t_x.numpy())
    assert np.array_equal(a1.squeeze(1).cpu().numpy().astype(np.int32), want)


def test_align_plan_graph_replay_is_repeatable(cuda_device):
    """AlignPlan: one C call per step, CUDA-graph capturable; the tile flags clean up after themselves,
    so replays (and a change of inputs between replays) give the right answer every time."""
    B, S, T, D = 8, 128, 512, 192
    d = cuda_device
    plan = tts.AlignPlan(B, D, T, S, d)
    outs = []
    bufs = None
    for seed in (1, 2, 1):
        t_x, t_y = synthetic.ragged_lengths(B, S, T, seed)
        z_p, m_p, logs_p, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=seed)
        new = (z_p.to(d), m_p.to(d), logs_p.to(d), t_y.to(d), t_x.to(d))
        if bufs is None:
            bufs = new
            plan.capture(0, *bufs)
        else:
            for dst, src in zip(bufs, new):
                dst.copy_(src)
        for _ in range(3):
            plan.replay(0)
        torch.cuda.synchronize()
        outs.append((plan.idx.clone(), plan.dur.clone()))
        nc_ref = mas_oracle.neg_cent_torch(z_p, m_p, logs_p)
        ref = mas_oracle.maximum_path_c(nc_ref.numpy(), t_y.numpy(), t_x.numpy())
        agree = (plan.path.cpu().numpy().astype(np.int32) == ref).mean()
        assert agree >= MIN_AGREE
        assert torch.equal(plan.dur.sum(1).cpu(), t_y)
    assert torch.equal(outs[0][0], outs[2][0]) and torch.equal(outs[0][1], outs[2][1])


def test_sharded_align_single_rank_nccl(cuda_device):
    """align_sharded + compact all-gather over NCCL with a world of one rank (the multi-rank host logic is
    covered on CPU by tests/test_sharded_cpu.py; tools/run_sharded_check.py checks 2+ GPUs under torchrun)."""
    import os
    import torch.distributed as dist

    if dist.is_initialized():
        pytest.skip("a process group already exists")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29731")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda_device)
    try:
        B, S, T = 6, 64, 256
        t_x, t_y = synthetic.ragged_lengths(B, S, T, 3)
        z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=3)
        d = cuda_device
        attn, w, g_idx, g_dur = tts.align_sharded(z_p.to(d), m_p.to(d), logs_p.to(d), x_mask.to(d), y_mask.to(d),
                                                  gather=True)
        assert torch.equal(tts.expand_path(g_idx, S), attn.squeeze(1))
        assert torch.equal(g_dur.sum(1).cpu(), t_y)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,S,T,ragged", [(6, 80, 320, True), (4, 256, 1024, False), (9, 200, 800, True), (3, 600, 1200, False)])
@pytest.mark.parametrize("scale", [0.01, 0])
def test_noise_applied_inside_the_dp(cuda_device, B, S, T, ragged, scale):
    """without a neg_cent_out request the noised cost is never materialised: the DP adds (std * noise) * scale
    while the cost streams in.  Same path as the route through the explicit noise pass, bit for bit."""
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 2) if ragged else synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=13)
    noise = torch.randn((B, T, S), generator=torch.Generator().manual_seed(6))
    d = cuda_device
    args = (z_p.to(d), m_p.to(d), logs_p.to(d), x_mask.to(d), y_mask.to(d), scale, noise.to(d))
    attn_a, w_a, (idx_a, dur_a, st_a) = tts.align(*args, return_compact=True)                  # noise inside the DP
    attn_b, w_b, (idx_b, dur_b, st_b), nc = tts.align(*args, return_compact=True, return_neg_cent=True)   # explicit pass
    assert (st_a == 0).all() and (st_b == 0).all()
    assert torch.equal(idx_a, idx_b) and torch.equal(dur_a, dur_b) and torch.equal(attn_a, attn_b)
    # and it is the reference's alignment of the reference's noised cost
    attn_ref, w_ref, nc_ref = mas_oracle.align_torch(z_p, m_p, logs_p, x_mask, y_mask, scale, noise)
    assert _rel_err(nc.cpu(), nc_ref) < REL_TOL
    assert (attn_a.cpu() == attn_ref).float().mean().item() >= MIN_AGREE
    assert torch.equal(w_a.sum((1, 2)).cpu(), w_ref.sum((1, 2)))


def test_zero_scale_shortcut_is_the_noise_branch_on_finite_costs(cuda_device):
    """mas_noise_scale == 0 (where cli.py:268-271 ends up): the noise branch and the fused no-noise kernel agree
    bit for bit on finite costs, which is what `zero_scale_is_no_noise=True` relies on."""
    B, S, T = 5, 120, 500
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 4)
    args = [t.to(cuda_device) for t in synthetic.prior_inputs(B, S, T, t_x, t_y, seed=13)]
    noise = torch.randn((B, T, S), device=cuda_device)
    a0, w0, nc0 = tts.align(*args, 0, noise, return_neg_cent=True)
    a1, w1, nc1 = tts.align(*args, 0, noise, return_neg_cent=True, zero_scale_is_no_noise=True)
    assert torch.equal(a0, a1) and torch.equal(w0, w1) and torch.equal(nc0, nc1)
