"""GPU parity of the alignment's first consumers: the prior expansion of
models.py:1270-1271 as a gather over the compact idx, its backward as a segmented sum, and logw_
(models.py:1256, 1261) from the int32 durations.  Checked against the reference expressions run in CPU torch."""
import numpy as np
import pytest
import torch

import torch_tts_b200 as tts
from oracle import mas_oracle
from torch_tts_b200 import synthetic

pytestmark = pytest.mark.gpu


def _aligned(B, S, T, ragged, dev, seed=3, D=24):
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed) if ragged else synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=seed)
    attn, w, (idx, dur, status) = tts.align(z_p.to(dev), m_p.to(dev), logs_p.to(dev), x_mask.to(dev), y_mask.to(dev),
                                            return_compact=True)
    assert int(status.abs().sum()) == 0
    return t_x, t_y, m_p, logs_p, x_mask, attn, idx, dur


@pytest.mark.parametrize("B,S,T,ragged", [(3, 50, 200, True), (2, 256, 1024, False), (4, 97, 333, True), (1, 7, 9, False)])
def test_expand_prior_is_the_one_hot_matmul(cuda_device, B, S, T, ragged):
    t_x, t_y, m_p, logs_p, x_mask, attn, idx, dur = _aligned(B, S, T, ragged, cuda_device)
    want_m, want_l = mas_oracle.expand_prior_torch(attn.cpu(), m_p, logs_p)
    got_m, got_l = tts.expand_prior(m_p.to(cuda_device), logs_p.to(cuda_device), idx, dur)
    # a one-hot row times finite numbers is a selection: exact, not a tolerance
    assert torch.equal(got_m.cpu(), want_m) and torch.equal(got_l.cpu(), want_l)
    # rows past t_y are zero
    for b in range(B):
        assert float(got_m[b, :, int(t_y[b]):].abs().sum()) == 0.0
    only_m, none = tts.expand_prior(m_p.to(cuda_device), None, idx, dur)
    assert none is None and torch.equal(only_m.cpu(), want_m)


@pytest.mark.parametrize("B,S,T,ragged", [(3, 50, 200, True), (2, 256, 1024, False), (2, 1000, 1003, False)])
def test_expand_prior_backward_matches_autograd_of_the_matmuls(cuda_device, B, S, T, ragged):
    t_x, t_y, m_p, logs_p, x_mask, attn, idx, dur = _aligned(B, S, T, ragged, cuda_device, seed=5, D=8)
    g = torch.Generator().manual_seed(11)
    gm, gl = torch.randn((B, 8, T), generator=g), torch.randn((B, 8, T), generator=g)
    m_ref, l_ref = m_p.clone().requires_grad_(True), logs_p.clone().requires_grad_(True)
    want_m, want_l = mas_oracle.expand_prior_torch(attn.cpu(), m_ref, l_ref)
    (want_m * gm).sum().add((want_l * gl).sum()).backward()
    m_dev = m_p.to(cuda_device).requires_grad_(True)
    l_dev = logs_p.to(cuda_device).requires_grad_(True)
    got_m, got_l = tts.expand_prior(m_dev, l_dev, idx, dur)
    (got_m * gm.to(cuda_device)).sum().add((got_l * gl.to(cuda_device)).sum()).backward()
    # sums of ~T/S terms in a different order than the GEMM: 1e-5 relative to the gradient scale
    for got, want in ((m_dev.grad.cpu(), m_ref.grad), (l_dev.grad.cpu(), l_ref.grad)):
        assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())
    # columns past t_x receive nothing
    for b in range(B):
        assert float(m_dev.grad[b, :, int(t_x[b]):].abs().sum()) == 0.0


def _segsum_reference(g, dur):
    """fp64 segmented sum on the host (the oracle's restatement, pinned against autograd in tests/test_oracle.py)"""
    return torch.from_numpy(mas_oracle.expand_prior_backward_np(g.numpy(), dur.numpy()))


@pytest.mark.parametrize("B,D,S,T", [(3, 192, 256, 1024), (2, 80, 97, 332), (2, 33, 600, 2000), (1, 192, 1000, 1100),
                                     (4, 130, 50, 64), (2, 192, 31, 1000), (2, 64, 77, 301), (70, 192, 40, 512),
                                     (2, 192, 3, 4000), (3, 96, 200, 260)])
@pytest.mark.parametrize("two", [True, False])
def test_prior_backward_kernels_agree(cuda_device, mas_env, B, D, S, T, two):
    """The channels-on-lanes kernel (tensor-map tiles, T % 4 == 0) and the column-per-thread kernel (any T) add the
    same frames in the same ascending order: bit-identical to each other as long as a CTA walks an utterance in one
    run (MAS_SEG_PARTS=1; with 2 or 4 runs a column that straddles a cut is the sum of its pieces), and equal to the
    fp64 sum within fp32 rounding either way.  Durations include empty columns, one very long segment, and (last utterance) columns past t_y."""
    g = torch.Generator().manual_seed(B * D + S)
    dur = torch.zeros((B, S), dtype=torch.int32)
    for b in range(B):
        t_y = T if b == 0 else int(torch.randint(S // 2 + 1, T + 1, (1,), generator=g))
        cuts = torch.sort(torch.randint(0, t_y + 1, (S - 1,), generator=g)).values
        edges = torch.cat([torch.zeros(1, dtype=torch.long), cuts, torch.tensor([t_y])])
        dur[b] = (edges[1:] - edges[:-1]).int()
    if B > 1:   # one column owns nearly everything
        dur[1] = 0
        dur[1, S // 3] = T - 3
        dur[1, S // 3 + 1] = 2
    gm, gl = torch.randn((B, D, T), generator=g), torch.randn((B, D, T), generator=g)
    L, p = tts._lib.lib(), tts._lib.ptr
    gm_d, gl_d, dur_d = gm.to(cuda_device), (gl.to(cuda_device) if two else None), dur.to(cuda_device)

    def run():
        out_m = torch.full((B, D, S), float("nan"), device=cuda_device)
        out_l = torch.full((B, D, S), float("nan"), device=cuda_device) if two else None
        rc = L.mas_expand_prior_backward_f32(p(gm_d), p(gl_d), p(dur_d), p(out_m), p(out_l), B, D, T, S,
                                             torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        return out_m, out_l

    split_m, split_l = run()          # small batches: the frames of an utterance are cut into 2 or 4 runs per CTA
    again_m, again_l = run()
    assert torch.equal(split_m, again_m)            # a fixed order of additions, whatever the timing
    mas_env(MAS_SEG_PARTS=1)
    new_m, new_l = run()
    mas_env(MAS_SEGSUM=0)
    old_m, old_l = run()
    assert torch.equal(new_m, old_m)                # one run per CTA: the very same additions
    want_m = _segsum_reference(gm, dur)
    tol = 1e-5 * max(1.0, want_m.abs().max().item())
    assert (new_m.cpu().double() - want_m).abs().max().item() <= tol
    assert (split_m.cpu().double() - want_m).abs().max().item() <= tol
    if two:
        assert torch.equal(new_l, old_l) and torch.equal(split_l, again_l)
        want_l = _segsum_reference(gl, dur)
        tol = 1e-5 * max(1.0, want_l.abs().max().item())
        assert (new_l.cpu().double() - want_l).abs().max().item() <= tol
        assert (split_l.cpu().double() - want_l).abs().max().item() <= tol


@pytest.mark.parametrize("parts", [0, 1, 2, 4])
def test_prior_backward_degenerate_durations(cuda_device, mas_env, parts):
    """All-empty utterances, a single column, durations that add up to more than T (clipped at T like the
    column-per-thread kernel clips them), negative entries (treated as 0), one channel."""
    mas_env(MAS_SEG_PARTS=parts)
    L, p = tts._lib.lib(), tts._lib.ptr
    cases = [
        (torch.zeros((2, 5), dtype=torch.int32), 3, 8),
        (torch.tensor([[8], [3]], dtype=torch.int32), 1, 8),
        (torch.tensor([[100, 100, 100, 0, 7], [0, 0, 300, 1, 0]], dtype=torch.int32), 33, 256),
        (torch.tensor([[-4, 2, 0, 0, 1022], [1, 1, 1, 1, 1]], dtype=torch.int32), 40, 1024),
        (torch.full((3, 1024), 1, dtype=torch.int32), 96, 1024),
    ]
    for dur, D, T in cases:
        B, S = dur.shape
        g = torch.randn((B, D, T), generator=torch.Generator().manual_seed(S + T))
        out = torch.full((B, D, S), float("nan"), device=cuda_device)
        g_d, dur_d = g.to(cuda_device), dur.to(cuda_device)      # kept alive across the raw-pointer call
        rc = L.mas_expand_prior_backward_f32(p(g_d), None, p(dur_d), p(out), None, B, D, T, S,
                                             torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        torch.cuda.synchronize()
        want = _segsum_reference(g, dur)
        assert (out.cpu().double() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item()), (dur.shape, D, T)


@pytest.mark.parametrize("B,S,T,ragged", [(5, 80, 320, True), (2, 256, 1024, False)])
def test_logw(cuda_device, B, S, T, ragged):
    t_x, t_y, m_p, logs_p, x_mask, attn, idx, dur = _aligned(B, S, T, ragged, cuda_device, seed=9, D=8)
    want = mas_oracle.logw_torch(attn.cpu().float(), x_mask)
    got = tts.logw(dur, t_x.to(cuda_device)).cpu()
    assert got.shape == want.shape == (B, 1, S)
    assert (got - want).abs().max().item() <= 1e-6 * 14.0          # |log| <= 13.8; logf vs torch.log: last-ulp
    assert torch.equal(got == 0, want == 0)                          # masked columns are (signed) zeros on both sides


def test_expand_prior_rejects_cpu_tensors():
    with pytest.raises(tts._lib.MasError):
        tts.expand_prior(torch.zeros(1, 2, 3), None, torch.zeros(1, 4, dtype=torch.int32), torch.zeros(1, 3, dtype=torch.int32))


@pytest.mark.parametrize("B,S,T", [(3, 40, 300), (2, 256, 1500), (1, 5, 8)])
def test_generate_path_matches_commons(cuda_device, B, S, T):
    """Inference: durations -> path (commons.generate_path) and the expansion that follows (models.py:1310-1317)."""
    g = torch.Generator().manual_seed(S)
    t_x = torch.randint(max(1, S // 2), S + 1, (B,), generator=g)
    x_mask = (torch.arange(S)[None, :] < t_x[:, None]).float().unsqueeze(1)                 # [B,1,S]
    w_ceil = torch.ceil(torch.rand((B, 1, S), generator=g) * (2.0 * T / S)) * x_mask         # some zeros, sum may exceed T
    w_ceil[0, 0, 1] = 0.0                                                                     # a token without frames
    y_len = torch.clamp_min(w_ceil.sum([1, 2]), 1).long().clamp_max(T)                        # models.py:1304 (capped: fixed T here)
    y_mask = (torch.arange(T)[None, :] < y_len[:, None]).float().unsqueeze(1)
    attn_mask = x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)                                    # [B,1,T,S]
    want = mas_oracle.generate_path_torch(w_ceil, attn_mask)
    got = tts.generate_path(w_ceil.to(cuda_device), attn_mask.to(cuda_device))
    assert got.shape == want.shape and torch.equal(got.cpu(), want)
    # compact form + expansion == the reference's matmul with that path
    idx = tts.idx_from_durations(w_ceil.to(cuda_device), t_x.to(cuda_device), T, y_len.to(cuda_device))
    m_p = torch.randn((B, 6, S), generator=g) * x_mask
    want_m, _ = mas_oracle.expand_prior_torch(want, m_p, m_p)
    dur_i = want.squeeze(1).sum(1).int()
    got_m, _ = tts.expand_prior(m_p.to(cuda_device), None, idx, dur_i.to(cuda_device))
    assert torch.equal(got_m.cpu(), want_m)


def test_consumers_golden_through_the_c_abi(cuda_device):
    """The reference's own outputs (tests/golden/consumers.npz): generate_path, the expansion and logw_."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "consumers.npz"))
    dev = cuda_device
    w_ceil, attn_mask = torch.from_numpy(z["w_ceil"]).to(dev), torch.from_numpy(z["attn_mask"]).to(dev)
    attn = tts.generate_path(w_ceil, attn_mask)
    assert np.array_equal(attn.cpu().numpy(), z["attn"])
    T = attn_mask.shape[2]
    idx = tts.idx_from_durations(w_ceil, torch.from_numpy(z["t_x"]).to(dev), T, torch.from_numpy(z["y_len"]).to(dev))
    dur = torch.from_numpy(z["w"]).reshape(z["w"].shape[0], -1).int().to(dev)
    m_e, l_e = tts.expand_prior(torch.from_numpy(z["m_p"]).to(dev), torch.from_numpy(z["logs_p"]).to(dev), idx, dur)
    assert np.array_equal(m_e.cpu().numpy(), z["m_expanded"]) and np.array_equal(l_e.cpu().numpy(), z["logs_expanded"])
    lw = tts.logw(dur, torch.from_numpy(z["t_x"]).to(dev)).cpu().numpy()
    assert np.abs(lw - z["logw_"]).max() <= 1.4e-5 and np.array_equal(lw == 0, z["logw_"] == 0)
