"""GPU tests of the round-2 rows: compact outputs through the caller, arbitrary S / T on the
one-kernel path, the single-launch noise-scaled alignment (models.py:1241-1247, the branch cli.py:268-271 takes
every training step), the four-value-warp DP variant, and the host-side guards ADVICE.md asked for."""
import numpy as np
import pytest
import torch

import torch_tts_b200 as tts
from oracle import mas_oracle
from torch_tts_b200 import synthetic, _lib

pytestmark = pytest.mark.gpu
D = synthetic.D_PRIOR
MIN_AGREE = 0.9999


def _inputs(B, S, T, seed, ragged=True, dev=None):
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed) if ragged else synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=seed)
    host = (z_p, m_p, logs_p, x_mask, y_mask)
    return t_x, t_y, host, tuple(t.to(dev) for t in host)


def _launches(fn):
    L = _lib.lib()
    L.mas_take_launch_count()
    out = fn()
    torch.cuda.synchronize()
    return out, int(L.mas_take_launch_count())


# --------------------------------------------------------------------------
# 8f-3: compact path through the caller
# --------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,T", [(5, 77, 301), (3, 256, 1024), (2, 600, 2100)])
def test_maximum_path_compact_without_the_dense_plane(cuda_device, B, S, T):
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed=B)
    nc = synthetic.neg_cent_like(B, S, T, seed=S).to(cuda_device)
    path, dur, idx, status = tts.maximum_path_compact(nc, t_y.to(cuda_device), t_x.to(cuda_device))
    none, dur2, idx2, status2 = tts.maximum_path_compact(nc, t_y.to(cuda_device), t_x.to(cuda_device), want_path=False)
    assert none is None
    assert torch.equal(dur, dur2) and torch.equal(idx, idx2) and torch.equal(status, status2)
    assert torch.equal(tts.expand_path(idx2, S), path)
    want = mas_oracle.maximum_path_c(nc.cpu().numpy(), t_y.numpy(), t_x.numpy())
    assert np.array_equal(tts.expand_path(idx2, S).cpu().numpy().astype(np.int32), want)


@pytest.mark.parametrize("scale", [None, 0.01])
@pytest.mark.parametrize("B,S,T", [(6, 96, 384), (64, 256, 1024), (9, 187, 743)])
def test_align_compact_outputs_only(cuda_device, B, S, T, scale):
    """align(dense=False): no dense attn is allocated or written; idx / durations are those of the dense call,
    and .attn() rebuilds exactly the dense attn on demand (TensorBoard hook, train_ms.py:517-519)."""
    t_x, t_y, host, dev = _inputs(B, S, T, seed=T, dev=cuda_device)
    noise = None if scale is None else torch.randn((B, T, S), generator=torch.Generator().manual_seed(1)).to(cuda_device)
    attn, w, (idx, dur, status) = tts.align(*dev, scale, noise, return_compact=True)
    compact, w2 = tts.align(*dev, scale, noise, dense=False)
    assert isinstance(compact, tts.CompactAlignment)
    assert torch.equal(compact.idx, idx) and torch.equal(compact.durations, dur) and torch.equal(compact.status, status)
    assert torch.equal(w, w2) and torch.equal(compact.w, w)
    assert torch.equal(compact.attn(), attn) and compact.attn() is compact.attn()
    compact.check()


def test_align_plan_without_path(cuda_device):
    B, S, T = 16, 128, 512
    t_x, t_y, host, dev = _inputs(B, S, T, seed=7, dev=cuda_device)
    z, m, l = dev[:3]
    full = tts.AlignPlan(B, D, T, S, cuda_device)
    lean = tts.AlignPlan(B, D, T, S, cuda_device, want_path=False)
    assert lean.path is None
    ty, tx = t_y.to(cuda_device), t_x.to(cuda_device)
    full.run(z, m, l, ty, tx)
    lean.capture(0, z, m, l, ty, tx)
    lean.replay(0)
    torch.cuda.synchronize()
    assert torch.equal(full.idx, lean.idx) and torch.equal(full.dur, lean.dur) and torch.equal(full.status, lean.status)
    assert torch.equal(tts.expand_path(lean.idx, S), full.path)


def test_c_abi_rejects_a_call_with_no_output(cuda_device):
    L = _lib.lib()
    nc = torch.zeros((1, 8, 4), device=cuda_device)
    t = torch.tensor([4], dtype=torch.int32, device=cuda_device)
    ws = torch.empty(max(L.mas_maximum_path_workspace_bytes(1, 8, 4), 256), dtype=torch.uint8, device=cuda_device)
    rc = L.mas_maximum_path_f32(nc.data_ptr(), t.data_ptr(), t.data_ptr(), None, 0, None, None, None, ws.data_ptr(),
                                ws.numel(), 1, 8, 4, None)
    assert rc == 1           # MAS_ERR_NULL_POINTER: nothing to write


# --------------------------------------------------------------------------
# shape generality: real collated batches have arbitrary S and T (data_utils.py:151-214)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,T", [(8, 187, 743), (5, 101, 402), (3, 250, 999), (7, 33, 130), (4, 255, 1021), (2, 1, 5)])
@pytest.mark.parametrize("scale", [None, 0.01])
def test_arbitrary_shapes_take_the_one_kernel_path(cuda_device, B, S, T, scale):
    """S % 4 != 0 and / or T % 4 != 0: still prior preparation + ONE fused kernel (the private cost plane has its own
    16-byte row stride), same alignment as with the explicit cost plane, exact MAS optimum of that cost."""
    t_x, t_y, host, dev = _inputs(B, S, T, seed=S + T, dev=cuda_device)
    noise = None if scale is None else torch.randn((B, T, S), generator=torch.Generator().manual_seed(2)).to(cuda_device)
    (attn, w, (idx, dur, status)), n = _launches(lambda: tts.align(*dev, scale, noise, return_compact=True))
    assert n == 2, n
    attn2, w2, (idx2, dur2, status2), nc = tts.align(*dev, scale, noise, return_compact=True, return_neg_cent=True)
    assert (status == 0).all()
    assert torch.equal(idx, idx2) and torch.equal(dur, dur2) and torch.equal(attn, attn2)
    want = mas_oracle.maximum_path_c(nc.cpu().numpy(), t_y.numpy(), t_x.numpy())
    assert np.array_equal(attn.squeeze(1).cpu().numpy().astype(np.int32), want)
    z_p, m_p, logs_p, x_mask, y_mask = host
    attn_ref, w_ref, nc_ref = mas_oracle.align_torch(z_p, m_p, logs_p, x_mask, y_mask, scale,
                                                     None if noise is None else noise.cpu())
    assert ((nc.cpu() - nc_ref).abs() / nc_ref.abs().clamp_min(1.0)).max().item() < 1e-4
    assert (attn.cpu() == attn_ref).float().mean().item() >= MIN_AGREE
    assert torch.equal(w.sum((1, 2)).cpu(), w_ref.sum((1, 2)))


@pytest.mark.parametrize("B,S,T", [(4, 600, 4000), (3, 300, 1000), (2, 513, 701), (2, 1000, 1100), (5, 431, 1500), (80, 257, 300)])
def test_wide_text_takes_the_one_kernel_path(cuda_device, B, S, T):
    """256 < S <= 1024 (several column blocks per mel tile, single-role DP teams of two or four warps, spilled decision
    words for long utterances -- BASELINE config 4 is (32, 600, 4000)): prior preparation + ONE fused kernel, same
    alignment as with the explicit cost plane, exact MAS optimum of that cost, parity with the reference expression."""
    t_x, t_y, host, dev = _inputs(B, S, T, seed=S + T, dev=cuda_device)
    (attn, w, (idx, dur, status)), n = _launches(lambda: tts.align(*dev, return_compact=True))
    assert n == 2, n
    attn2, w2, (idx2, dur2, status2), nc = tts.align(*dev, return_compact=True, return_neg_cent=True)
    assert (status == 0).all() and torch.equal(dur.sum(1).cpu(), t_y)
    assert torch.equal(idx, idx2) and torch.equal(dur, dur2) and torch.equal(attn, attn2)
    assert torch.equal(tts.expand_path(idx, S), attn.squeeze(1))
    sel = list(range(min(B, 4)))
    want = mas_oracle.maximum_path_c(nc[sel].cpu().numpy(), t_y[sel].numpy(), t_x[sel].numpy())
    assert np.array_equal(attn[sel].squeeze(1).cpu().numpy().astype(np.int32), want)
    z_p, m_p, logs_p, x_mask, y_mask = (t[sel] for t in host)
    attn_ref, w_ref, nc_ref = mas_oracle.align_torch(z_p, m_p, logs_p, x_mask, y_mask, None, None)
    assert ((nc[sel].cpu() - nc_ref).abs() / nc_ref.abs().clamp_min(1.0)).max().item() < 1e-4
    assert (attn[sel].cpu() == attn_ref).float().mean().item() >= MIN_AGREE


@pytest.mark.parametrize("B,S,T", [(80, 257, 300), (160, 131, 300), (100, 33, 700), (3, 1001, 1515), (160, 131, 301),
                                   (90, 200, 518), (150, 256, 262)])
@pytest.mark.parametrize("with_noise", [False, True])
def test_explicit_cost_plane_with_unaligned_rows(cuda_device, B, S, T, with_noise):
    """S % 4 != 0 and enough units that every contraction CTA takes several rounds: tts.neg_cent and
    align(return_neg_cent=True) store 16-byte rows into the padded workspace plane and pack them afterwards.
    (T % 4 != 0 in the last cases: z_p has no tensor map then, the raw-z producer uses plain loads.)  The
    plane is the same run after run, matches the reference expression, and the path is its exact MAS optimum."""
    t_x, t_y, host, dev = _inputs(B, S, T, seed=S * T, dev=cuda_device)
    z_p, m_p, logs_p, x_mask, y_mask = dev
    nc0 = tts.neg_cent(z_p, m_p, logs_p)
    for _ in range(3):
        assert torch.equal(tts.neg_cent(z_p, m_p, logs_p), nc0)
    want = mas_oracle.neg_cent_torch(*(t[:6] for t in host[:3]))
    assert ((nc0[:6].cpu() - want).abs() / want.abs().clamp_min(1.0)).max().item() < 1e-4
    noise = torch.randn((B, T, S), generator=torch.Generator().manual_seed(5)).to(cuda_device) if with_noise else None
    scale = 0.01 if with_noise else None
    attn, w, (idx, dur, status), nc = tts.align(*dev, scale, noise, return_compact=True, return_neg_cent=True)
    if with_noise:
        std = nc0.double().std().float()
        assert ((nc - (nc0 + std * noise * 0.01)).abs() / nc0.abs().clamp_min(1.0)).max().item() < 1e-5
    else:
        assert torch.equal(nc, nc0)
    assert (status == 0).all() and torch.equal(dur.sum(1).cpu(), t_y)
    sel = list(range(0, B, max(1, B // 5)))
    want_path = mas_oracle.maximum_path_c(nc[sel].cpu().numpy(), t_y[sel].numpy(), t_x[sel].numpy())
    assert np.array_equal(attn[sel].squeeze(1).cpu().numpy().astype(np.int32), want_path)
    # and the private-plane route (one kernel, or contraction + pass + DP when the noise kernel does not cover S)
    attn2, w2, (idx2, dur2, status2) = tts.align(*dev, scale, noise, return_compact=True)
    assert torch.equal(idx, idx2) and torch.equal(dur, dur2)


@pytest.mark.parametrize("B,S,T", [(80, 256, 384), (160, 131, 300), (76, 257, 301), (64, 256, 1024)])
@pytest.mark.parametrize("no_tma", [1, 2, 3])
def test_contraction_without_tensor_maps(cuda_device, mas_env, B, S, T, no_tma):
    """The fallbacks for planes a tensor map cannot describe -- plain z loads (bit 1), per-cell output stores (bit 2) --
    give the same bits as the tensor-map paths, run after run, with several rounds of units per CTA.  (The per-cell
    stores keep the SM's load / store queue full; that is how round 2 found the converter warps handing a z stage back
    before their shared-memory loads had read it: whole 32-row groups of neg_cent with a few channels of the wrong
    K block.  The stage is now released after the values have been used.)"""
    t_x, t_y, host, dev = _inputs(B, S, T, seed=B + S, dev=cuda_device)
    z_p, m_p, logs_p, x_mask, y_mask = dev
    want = tts.neg_cent(z_p, m_p, logs_p)
    a0, w0, (idx0, dur0, st0) = tts.align(*dev, return_compact=True)
    mas_env(MAS_TC_NO_TMA=no_tma)
    for _ in range(3):
        assert torch.equal(tts.neg_cent(z_p, m_p, logs_p), want)
    a1, w1, (idx1, dur1, st1) = tts.align(*dev, return_compact=True)
    assert torch.equal(idx0, idx1) and torch.equal(dur0, dur1) and torch.equal(a0, a1)


# --------------------------------------------------------------------------
# noise-scaled alignment in one launch
# --------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,T,ragged", [(64, 256, 1024, False), (74, 64, 256, True), (1, 256, 1024, False),
                                          (20, 200, 800, True), (80, 64, 256, True), (200, 96, 300, True),
                                          (9, 187, 743, True), (5, 33, 130, True)])
def test_noise_branch_is_two_launches_and_matches_the_separate_kernels(cuda_device, mas_env, B, S, T, ragged):
    """Any batch size: prior preparation + ONE kernel (contraction + statistics | grid barrier | DP whose helper
    warps add the noise to the cost tiles in shared memory).  Bit-identical to the three-launch route
    (MAS_NOISE_FUSED=0), which adds the noise inside the DP loop (or, for rows that are not 16-byte, in a separate
    pass over the plane); B > 148 puts several utterances on one DP CTA; S, T need not be multiples of 4."""
    t_x, t_y, host, dev = _inputs(B, S, T, seed=11, ragged=ragged, dev=cuda_device)
    noise = torch.randn((B, T, S), generator=torch.Generator().manual_seed(3)).to(cuda_device)
    (a, w, (idx, dur, status)), n = _launches(lambda: tts.align(*dev, 0.01, noise, return_compact=True))
    assert n == 2, n
    assert (status == 0).all() and torch.equal(dur.sum(1).cpu(), t_y)
    mas_env(MAS_NOISE_FUSED=0)
    (a3, w3, (idx3, dur3, status3)), n3 = _launches(lambda: tts.align(*dev, 0.01, noise, return_compact=True))
    assert n3 >= 3, n3
    assert torch.equal(idx, idx3) and torch.equal(dur, dur3) and torch.equal(a, a3)


def test_noise_graph_replay_is_repeatable(cuda_device):
    """the grid-barrier counter and the tile flags are cleared by the prior kernel of every step"""
    B, S, T = 12, 128, 512
    t_x, t_y, host, dev = _inputs(B, S, T, seed=13, dev=cuda_device)
    noise = torch.randn((B, T, S), device=cuda_device)
    plan = tts.AlignPlan(B, D, T, S, cuda_device, with_noise=True)
    z, m, l = dev[:3]
    plan.capture(0, z, m, l, t_y.to(cuda_device), t_x.to(cuda_device), noise, 0.01)
    plan.replay(0)
    torch.cuda.synchronize()
    first = (plan.idx.clone(), plan.dur.clone())
    for _ in range(4):
        plan.replay(0)
    torch.cuda.synchronize()
    assert torch.equal(plan.idx, first[0]) and torch.equal(plan.dur, first[1])
    attn, w, (idx, dur, status) = tts.align(*dev, 0.01, noise, return_compact=True)
    assert torch.equal(idx, plan.idx) and torch.equal(attn.squeeze(1), plan.path)


# --------------------------------------------------------------------------
# DP variants
# --------------------------------------------------------------------------
@pytest.mark.parametrize("B,S,T,ties", [(4, 256, 1024, False), (3, 200, 700, True), (2, 129, 40 * 32 + 1, True)])
def test_four_value_warps(cuda_device, mas_env, B, S, T, ties):
    """MAS_DP_WARPS=4: four value warps with two columns per thread (+ four origin warps) for 128 < S <= 256."""
    mas_env(MAS_DP_WARPS=4)
    nc = synthetic.neg_cent_like(B, S, T, seed=S + T, ties=ties)
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed=B)
    nc[0, T // 3, S // 2] = float("nan")          # the exact second pass as well
    want = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
    path, dur, idx, status = tts.maximum_path_compact(nc.to(cuda_device), t_y.to(cuda_device), t_x.to(cuda_device))
    assert np.array_equal(path.cpu().numpy().astype(np.int32), want)
    # and inside the fused kernel
    t_x, t_y, host, dev = _inputs(6, S, T, seed=3, dev=cuda_device)
    (attn, w, (idx, dur, status), nc_gpu), n = _launches(
        lambda: tts.align(*dev, return_compact=True, return_neg_cent=True))
    attn_b, w_b = tts.align(*dev)
    assert torch.equal(attn, attn_b)
    want = mas_oracle.maximum_path_c(nc_gpu.cpu().numpy(), t_y.numpy(), t_x.numpy())
    assert np.array_equal(attn.squeeze(1).cpu().numpy().astype(np.int32), want)


def test_no_fused_knob_gives_the_same_alignment(cuda_device, mas_env):
    B, S, T = 10, 192, 640
    t_x, t_y, host, dev = _inputs(B, S, T, seed=21, dev=cuda_device)
    (a, w, (idx, dur, st)), n = _launches(lambda: tts.align(*dev, return_compact=True))
    mas_env(MAS_NO_FUSED=1)
    (a2, w2, (idx2, dur2, st2)), n2 = _launches(lambda: tts.align(*dev, return_compact=True))
    assert n == 2 and n2 == 3
    assert torch.equal(idx, idx2) and torch.equal(dur, dur2) and torch.equal(a, a2)


# --------------------------------------------------------------------------
# host-side guards
# --------------------------------------------------------------------------
def test_check_status_raises_on_undefined_lengths(cuda_device):
    B, S, T = 3, 20, 30
    t_x = torch.tensor([10, 25, 5], dtype=torch.int32)
    t_y = torch.tensor([30, 20, 30], dtype=torch.int32)          # utterance 1: more text than frames
    z_p, m_p, logs_p, x_mask, y_mask = (t.to(cuda_device) for t in synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=1))
    kw = dict(x_lengths=t_x, y_lengths=t_y)
    attn, w, (idx, dur, status) = tts.align(z_p, m_p, logs_p, x_mask, y_mask, return_compact=True, **kw)
    assert status.cpu().tolist() == [0, 1, 0] and float(attn[1].abs().sum()) == 0.0
    with pytest.raises(_lib.MasError, match="undefined"):
        tts.align(z_p, m_p, logs_p, x_mask, y_mask, check_status=True, **kw)


def test_align_plan_validates_its_inputs(cuda_device):
    B, S, T = 2, 32, 64
    plan = tts.AlignPlan(B, D, T, S, cuda_device)
    t_x, t_y = synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, _, _ = (t.to(cuda_device) for t in synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=1))
    ty, tx = t_y.to(cuda_device), t_x.to(cuda_device)
    plan.run(z_p, m_p, logs_p, ty, tx)
    with pytest.raises(_lib.MasError, match="z_p"):
        plan.run(z_p.half(), m_p, logs_p, ty, tx)                 # autocast output handed in without the upcast
    with pytest.raises(_lib.MasError, match="m_p"):
        plan.run(z_p, m_p.transpose(1, 2), logs_p, ty, tx)
    with pytest.raises(_lib.MasError, match="t_ys"):
        plan.run(z_p, m_p, logs_p, t_y, tx)                       # CPU tensor
    with pytest.raises(_lib.MasError, match="logs_p"):
        plan.run(z_p, m_p, logs_p[:, :, :-1].contiguous(), ty, tx)


def test_generate_path_fractional_durations(cuda_device):
    """commons.generate_path compares frame indices with the FLOAT running sum (commons.py:138-143): durations that
    are not integers must not be truncated before the sum."""
    B, S, T = 3, 12, 40
    g = torch.Generator().manual_seed(4)
    dur = torch.randint(0, 16, (B, 1, S), generator=g).float() * 0.25          # multiples of 1/4: sums exact in any order
    t_x = torch.tensor([12, 9, 5])
    x_mask = (torch.arange(S)[None, :] < t_x[:, None]).float().unsqueeze(1)
    dur = dur * x_mask
    y_len = torch.clamp_min(torch.sum(dur, [1, 2]), 1).long()
    y_mask = (torch.arange(T)[None, :] < y_len[:, None]).float().unsqueeze(1)
    mask = x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)
    want = mas_oracle.generate_path_torch(dur, mask)
    got = tts.generate_path(dur.to(cuda_device), mask.to(cuda_device))
    assert torch.equal(got.cpu(), want)
    # huge / non-finite entries saturate instead of overflowing an int cast
    dur[0, 0, 2] = float("inf")
    dur[1, 0, 1] = 3.0e38
    got = tts.generate_path(dur.to(cuda_device), mask.to(cuda_device))
    assert torch.isfinite(got).all() and float(got.sum(-1).max()) <= 1.0


# --------------------------------------------------------------------------
# protocol edges of the single-launch noise kernel (feeder pairs for B <= 74, helper warps above)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("B", [6, 80])
def test_noise_kernel_with_an_invalid_utterance(cuda_device, mas_env, B):
    """An utterance the reference leaves undefined (t_x > t_y) is skipped by the DP CTA and by its feeder alike: zero
    path, zero durations, idx -1, status 1 -- and the others are aligned exactly as by the separate launches."""
    S, T = 64, 256
    t_x, t_y, host, dev = _inputs(B, S, T, seed=17, dev=cuda_device)
    bad = 2
    x_mask, y_mask = host[3].clone(), host[4].clone()
    y_mask[bad, 0, 5:] = 0.0          # 5 mel frames, far fewer than the text tokens
    x_mask[bad, 0, :40] = 1.0
    dev = dev[:3] + (x_mask.to(cuda_device), y_mask.to(cuda_device))
    noise = torch.randn((B, T, S), generator=torch.Generator().manual_seed(5)).to(cuda_device)
    (a, w, (idx, dur, status)), n = _launches(lambda: tts.align(*dev, 0.01, noise, return_compact=True))
    assert n == 2
    assert int(status[bad]) == 1 and int(status.sum()) == 1
    assert not a[bad].any() and not dur[bad].any() and bool((idx[bad] == -1).all())
    mas_env(MAS_NOISE_FUSED=0)
    a3, w3, (idx3, dur3, status3) = tts.align(*dev, 0.01, noise, return_compact=True)
    assert torch.equal(idx, idx3) and torch.equal(dur, dur3) and torch.equal(a, a3) and torch.equal(status, status3)


@pytest.mark.parametrize("B", [5, 80])
def test_noise_kernel_with_a_non_finite_input(cuda_device, mas_env, B):
    """A NaN in z_p makes the std -- and with it every cost of the batch -- NaN (models.py:1243): every utterance takes
    the exact second pass, i.e. the DP asks its feeder (or its helper warps) for all the tiles once more.  Same paths
    as the separate launches, which implement the reference's compare-select on NaN the same way."""
    S, T = 96, 320
    t_x, t_y, host, dev = _inputs(B, S, T, seed=19, dev=cuda_device)
    z = dev[0].clone()
    z[1, 3, 7] = float("nan")
    dev = (z,) + dev[1:]
    noise = torch.randn((B, T, S), generator=torch.Generator().manual_seed(6)).to(cuda_device)
    (a, w, (idx, dur, status)), n = _launches(lambda: tts.align(*dev, 0.01, noise, return_compact=True))
    assert n == 2 and (status == 0).all() and torch.equal(dur.sum(1).cpu(), t_y)
    mas_env(MAS_NOISE_FUSED=0)
    a3, w3, (idx3, dur3, status3) = tts.align(*dev, 0.01, noise, return_compact=True)
    assert torch.equal(idx, idx3) and torch.equal(dur, dur3) and torch.equal(a, a3)
