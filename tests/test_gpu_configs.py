"""GPU parity at the shapes BASELINE.json names (configs 1-5) and on the reference's own tensors.

Every case goes through the public calls (tts.align / AlignPlan / tts.maximum_path -> C ABI) and is compared
with the CPU oracle, which tests/test_oracle.py pins to the reference.  What "parity" means here (BASELINE.json):
  * MAS on identical neg_cent: bit-exact;
  * the alignment call: neg_cent within 1e-4 relative, path agreement >= 99.99 % of cells, identical duration
    sums -- and the path is the EXACT MAS optimum of the cost the GPU produced.
The measured agreement is printed (run with -s to see it) and asserted.
"""
import os

import numpy as np
import pytest
import torch

import torch_tts_b200 as tts
from oracle import mas_oracle
from torch_tts_b200 import synthetic, _lib

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
MIN_AGREE = 0.9999
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
D = synthetic.D_PRIOR


def _rel_err(got, want):
    return ((got - want).abs() / want.abs().clamp_min(1.0)).max().item()


def _idx_of(path_np):
    return np.where(path_np.sum(2) > 0, path_np.argmax(2), -1)


def _align_and_check(dev, B, S, T, t_x, t_y, scale, seed, label, sample=None):
    """tts.align without a neg_cent request (the product schedule: private cost plane, dead tiles skipped, noise
    applied inside the DP) against the oracle; then the same call WITH the cost returned, which must give the same
    alignment bit for bit and be the exact MAS optimum of that cost."""
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=seed)
    noise = None
    if scale is not None:
        noise = torch.randn((B, T, S), generator=torch.Generator().manual_seed(seed + 1))
    dz, dm, dl, dxm, dym = (t.to(dev) for t in (z_p, m_p, logs_p, x_mask, y_mask))
    dn = None if noise is None else noise.to(dev)
    L = _lib.lib()
    L.mas_take_launch_count()
    attn, w, (idx, dur, status) = tts.align(dz, dm, dl, dxm, dym, scale, dn, return_compact=True)
    launches = L.mas_take_launch_count()
    attn2, w2, (idx2, dur2, status2), nc = tts.align(dz, dm, dl, dxm, dym, scale, dn, return_compact=True,
                                                     return_neg_cent=True)
    torch.cuda.synchronize()
    assert (status == 0).all() and (status2 == 0).all()
    assert torch.equal(idx, idx2) and torch.equal(dur, dur2), f"{label}: private-plane schedule differs from the explicit one"
    assert torch.equal(attn, attn2)
    assert torch.equal(dur.sum(1).cpu(), t_y)                                  # duration-sum invariant
    assert torch.equal(w.squeeze(1).to(torch.int32), dur)
    assert torch.equal(tts.expand_path(idx, S), attn.squeeze(1))               # dense and compact forms agree
    sel = list(range(B)) if sample is None else sample
    sel_t = torch.tensor(sel)
    # the exact MAS optimum of the cost the GPU produced (rows past t_y of a private plane are never read)
    want_gpu = mas_oracle.maximum_path_c(nc[sel_t].cpu().numpy(), t_y[sel_t].numpy(), t_x[sel_t].numpy())
    assert np.array_equal(_idx_of(want_gpu), idx[sel_t].cpu().numpy()), f"{label}: not the MAS optimum of the GPU cost"
    # against the reference expression (oracle on the CPU, same inputs, same noise draw)
    attn_ref, w_ref, nc_ref = mas_oracle.align_torch(z_p[sel_t], m_p[sel_t], logs_p[sel_t], x_mask[sel_t], y_mask[sel_t],
                                                     None, None)
    if scale is not None:
        # torch.std covers ALL cells of the batch (models.py:1243): take it from the whole batch
        nc_all = mas_oracle.neg_cent_torch(z_p, m_p, logs_p)
        nc_ref = nc_all[sel_t] + torch.std(nc_all) * noise[sel_t] * scale
        mask = (x_mask[sel_t].unsqueeze(2) * y_mask[sel_t].unsqueeze(-1)).squeeze(1)
        attn_ref = torch.from_numpy(mas_oracle.maximum_path(nc_ref.numpy(), mask.numpy())).float().unsqueeze(1)
        w_ref = attn_ref.sum(2)
    rel = _rel_err(nc[sel_t].cpu(), nc_ref)
    agree = (attn[sel_t].cpu() == attn_ref).float().mean().item()
    rows = (idx[sel_t].cpu() == torch.from_numpy(_idx_of(attn_ref.squeeze(1).numpy().astype(np.int32)))).float().mean().item()
    print(f"\n[{label}] launches={launches} neg_cent rel err {rel:.2e}, path cell agreement {agree:.8f}, "
          f"mel rows on the reference's column {rows:.6f}")
    assert rel < REL_TOL
    assert agree >= MIN_AGREE, agree
    assert torch.equal(w[sel_t].sum((1, 2)).cpu(), w_ref.sum((1, 2)))
    return launches


# --------------------------------------------------------------------------
# BASELINE.json configs
# --------------------------------------------------------------------------
def test_config1_maximum_path(cuda_device):
    """configs[0]: maximum_path, B=16, T_text=200, T_mel=800 -- bit-exact, with and without forced ties."""
    B, S, T, _ = synthetic.CONFIGS["c1"]
    t_x, t_y = synthetic.full_lengths(B, S, T)
    for ties in (False, True):
        nc = synthetic.neg_cent_like(B, S, T, seed=0, ties=ties)
        x_mask, y_mask = synthetic.masks(t_x, t_y, S, T)
        mask = (x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)).squeeze(1)
        want = mas_oracle.maximum_path(nc.numpy(), mask.numpy())
        got = tts.maximum_path(nc.to(cuda_device), mask.to(cuda_device))
        assert np.array_equal(got.cpu().numpy().astype(np.int32), want)


def test_config2_fused_exact_shape(cuda_device):
    """configs[1], the benchmarked shape exactly: B=64, S=256, T=1024, D=192, full lengths, one fused kernel."""
    B, S, T, _ = synthetic.CONFIGS["c2"]
    t_x, t_y = synthetic.full_lengths(B, S, T)
    launches = _align_and_check(cuda_device, B, S, T, t_x, t_y, None, 0, "c2 fused")
    assert launches == 2          # prior preparation + the fused contraction/DP kernel


def test_config2_align_plan_graph_is_the_benchmarked_call(cuda_device):
    """bench.py's call: AlignPlan at config 2 replayed from a CUDA graph; idx/dur against the oracle."""
    B, S, T, _ = synthetic.CONFIGS["c2"]
    t_x, t_y = synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=100)
    d = cuda_device
    plan = tts.AlignPlan(B, D, T, S, d)
    plan.capture(0, z_p.to(d), m_p.to(d), logs_p.to(d), t_y.to(d), t_x.to(d))
    for _ in range(3):
        plan.replay(0)
    torch.cuda.synchronize()
    nc_ref = mas_oracle.neg_cent_torch(z_p, m_p, logs_p)
    ref = mas_oracle.maximum_path_c(nc_ref.numpy(), t_y.numpy(), t_x.numpy())
    agree = (plan.path.cpu().numpy().astype(np.int32) == ref).mean()
    print(f"\n[c2 AlignPlan graph] path cell agreement {agree:.8f}")
    assert agree >= MIN_AGREE
    assert torch.equal(plan.dur.sum(1).cpu(), t_y) and (plan.status == 0).all()
    assert torch.equal(tts.expand_path(plan.idx, S), plan.path)


@pytest.mark.parametrize("scale", [0.01, 0.005, 0])
def test_config3_noise_ragged(cuda_device, scale):
    """configs[2]: VITS2 noise-scaled MAS, ragged masks, B=128 -- the branch cli.py:268-271 takes every step."""
    B, S, T, _ = synthetic.CONFIGS["c3"]
    t_x, t_y = synthetic.config_lengths("c3", seed=3)
    _align_and_check(cuda_device, B, S, T, t_x, t_y, scale, 30, f"c3 noise scale={scale}")


def test_config3_no_noise_ragged(cuda_device):
    B, S, T, _ = synthetic.CONFIGS["c3"]
    t_x, t_y = synthetic.config_lengths("c3", seed=3)
    _align_and_check(cuda_device, B, S, T, t_x, t_y, None, 31, "c3 no noise")


@pytest.mark.parametrize("scale", [None, 0.01])
def test_config4_long_utterances(cuda_device, scale):
    """configs[3]: S=600, T=4000 (direction bits exceed shared memory), through align(); B=4 keeps the oracle fast."""
    B, S, T = 4, 600, 4000
    t_x = torch.tensor([600, 600, 431, 150], dtype=torch.int32)
    t_y = torch.tensor([4000, 3999, 3127, 700], dtype=torch.int32)
    _align_and_check(cuda_device, B, S, T, t_x, t_y, scale, 40, f"c4 scale={scale}")


@pytest.mark.parametrize("scale", [None, 0.01])
def test_config5_b512_fused(cuda_device, scale):
    """configs[4] per-GPU worst case (the whole B=512 batch on one GPU): size-independent properties on every
    utterance, oracle parity on a sample."""
    B, S, T, _ = synthetic.CONFIGS["c5"]
    t_x, t_y = synthetic.config_lengths("c5", seed=5)
    sample = [0, 1, 63, 64, 147, 148, 300, 511]
    _align_and_check(cuda_device, B, S, T, t_x, t_y, scale, 50, f"c5 B=512 scale={scale}", sample=sample)
    # properties at full size
    d = cuda_device
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=50)
    noise = None if scale is None else torch.randn((B, T, S), generator=torch.Generator().manual_seed(51)).to(d)
    attn, w, (idx, dur, status) = tts.align(z_p.to(d), m_p.to(d), logs_p.to(d), x_mask.to(d), y_mask.to(d), scale, noise,
                                            return_compact=True)
    path = attn.squeeze(1)
    rows = torch.arange(T, device=d)[None, :] < t_y.to(d)[:, None]
    assert torch.equal(path.sum((1, 2)).to(torch.int32).cpu(), t_y)           # one 1 per valid row, none elsewhere
    assert torch.equal(path.sum(2) > 0, rows)
    step = idx[:, 1:] - idx[:, :-1]
    assert ((step[rows[:, 1:]] == 0) | (step[rows[:, 1:]] == 1)).all()        # monotone
    assert (idx[:, 0] == 0).all()
    last = idx.gather(1, (t_y.to(d).long() - 1)[:, None])[:, 0]
    assert torch.equal(last.cpu(), t_x - 1)


# --------------------------------------------------------------------------
# the reference's own tensors (tests/golden, generated by running the reference)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["none", "s001", "s0"])
def test_real_synthesizer_tensors_through_align(cuda_device, tag):
    """z_p / m_p / logs_p, masks and the randn_like draw captured inside the real SynthesizerTrn.forward
    (models.py:1214-1290) -> tts.align on the GPU -> the attn and w the reference model itself produced."""
    z = np.load(os.path.join(GOLD, "synth_align.npz"))
    d = cuda_device
    z_p, m_p, logs_p, x_mask, y_mask = (torch.from_numpy(z[f"inputs/{k}"]).to(d)
                                        for k in ("z_p", "m_p", "logs_p", "x_mask", "y_mask"))
    scale = z[f"{tag}/scale"][0]
    scale = None if np.isnan(scale) else float(scale)
    noise = torch.from_numpy(z[f"{tag}/noise"]).to(d) if scale is not None else None
    attn, w, (idx, dur, status), nc = tts.align(z_p, m_p, logs_p, x_mask, y_mask, scale, noise, return_compact=True,
                                                return_neg_cent=True)
    attn_b, w_b = tts.align(z_p, m_p, logs_p, x_mask, y_mask, scale, noise)          # product schedule
    assert torch.equal(attn, attn_b) and torch.equal(w, w_b)
    want_nc = torch.from_numpy(z[f"{tag}/neg_cent"])
    want_attn = torch.from_numpy(z[f"{tag}/attn"].astype(np.float32))
    rel = _rel_err(nc.cpu(), want_nc)
    agree = (attn.squeeze(1).cpu() == want_attn).float().mean().item()
    print(f"\n[real SynthesizerTrn tensors, {tag}] neg_cent rel err {rel:.2e}, path cell agreement {agree:.8f}")
    assert attn.shape == (3, 1, 96, 24) and w.shape == (3, 1, 24)
    assert rel < REL_TOL
    assert agree >= MIN_AGREE
    assert torch.equal(w.squeeze(1).cpu().sum(1), torch.from_numpy(z[f"{tag}/w"]).sum(1))
    # MAS on the model's own neg_cent: bit-exact against the model's own attn
    mask = torch.from_numpy(z[f"{tag}/mask"].astype(np.float32)).to(d)
    path = tts.maximum_path(want_nc.to(d), mask)
    assert torch.equal(path.cpu(), want_attn)


@pytest.mark.parametrize("case", ["ragged_a", "ragged_ties", "full", "full_ties", "edges", "nonfinite"])
def test_reference_generated_mas_vectors(cuda_device, case):
    """tests/golden/mas_small.npz (outputs of the reference's Cython kernel) straight into the CUDA kernel."""
    z = np.load(os.path.join(GOLD, "mas_small.npz"))
    nc, t_x, t_y, want = (z[f"{case}/{k}"] for k in ("neg_cent", "t_x", "t_y", "path"))
    d = cuda_device
    path, dur, idx, status = tts.maximum_path_compact(torch.from_numpy(nc).to(d), torch.from_numpy(t_y).to(d),
                                                      torch.from_numpy(t_x).to(d))
    assert (status == 0).all()
    assert np.array_equal(path.cpu().numpy().astype(np.int8), want)
    assert np.array_equal(dur.cpu().numpy(), want.sum(1))


@pytest.mark.parametrize("case", ["c1", "c1_ties", "ragged12"])
def test_reference_generated_seeded_vectors(cuda_device, case):
    """tests/golden/mas_seeded.npz: config-1-sized inputs regenerated from the seed, the reference's idx as golden."""
    z = np.load(os.path.join(GOLD, "mas_seeded.npz"))
    B, S, T, ragged, seed, ties = (int(v) for v in z[f"{case}/shape"])
    nc = synthetic.neg_cent_like(B, S, T, seed=seed, ties=bool(ties))
    chk = z[f"{case}/nc_checksum"]
    assert float(nc.double().sum()) == chk[0] and float(nc.double().abs().max()) == chk[1]
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed) if ragged else synthetic.full_lengths(B, S, T)
    d = cuda_device
    path, dur, idx, status = tts.maximum_path_compact(nc.to(d), t_y.to(d), t_x.to(d))
    assert np.array_equal(idx.cpu().numpy(), z[f"{case}/idx"].astype(np.int64))


# --------------------------------------------------------------------------
# autocast dtypes (train_ms.py:351): half-precision z_p / m_p / logs_p
# --------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("scale", [None, 0.01])
def test_align_with_autocast_dtypes(cuda_device, dtype, scale):
    """Under fp16 autocast the encoders hand half tensors to the alignment block; the reference's cost then runs
    in that dtype while MAS upcasts (__init__.py:13).  Here the inputs are upcast once and everything runs in
    fp32: the result must be the fp32 alignment of the upcast inputs, returned in the caller's dtype."""
    B, S, T = 6, 96, 384
    t_x, t_y = synthetic.ragged_lengths(B, S, T, 6)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=6)
    zh, mh, lh = z_p.to(dtype), m_p.to(dtype), logs_p.to(dtype)
    noise = None if scale is None else torch.randn((B, T, S), generator=torch.Generator().manual_seed(8))
    d = cuda_device
    attn, w, (idx, dur, status), nc = tts.align(zh.to(d), mh.to(d), lh.to(d), x_mask.to(d).to(dtype), y_mask.to(d).to(dtype),
                                                scale, None if noise is None else noise.to(d), return_compact=True,
                                                return_neg_cent=True)
    assert attn.dtype == dtype and w.dtype == dtype and attn.shape == (B, 1, T, S)
    assert (status == 0).all()
    attn_ref, w_ref, nc_ref = mas_oracle.align_torch(zh.float(), mh.float(), lh.float(), x_mask, y_mask, scale, noise)
    assert _rel_err(nc.cpu(), nc_ref) < REL_TOL
    agree = (attn.float().cpu() == attn_ref).float().mean().item()
    assert agree >= MIN_AGREE, agree
    assert torch.equal(w.float().sum((1, 2)).cpu(), w_ref.sum((1, 2)))
    want = mas_oracle.maximum_path_c(nc.cpu().numpy(), t_y.numpy(), t_x.numpy())
    assert np.array_equal(attn.squeeze(1).float().cpu().numpy().astype(np.int32), want)
