"""GPU parity: torch_tts_b200.maximum_path (CUDA, through the C ABI) against the
CPU oracle on identical neg_cent -- bit-exact, as BASELINE.json's north_star
requires.  Mirrors what tests of vits2/monotonic_align would look like."""
import numpy as np
import pytest
import torch

import torch_tts_b200 as tts
from oracle import mas_oracle
from torch_tts_b200 import synthetic

pytestmark = pytest.mark.gpu


def _mask(t_x, t_y, S, T):
    x_mask, y_mask = synthetic.masks(t_x, t_y, S, T)
    return (x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)).squeeze(1)


def _check(nc, t_x, t_y, dev, dtype=torch.float32):
    B, T, S = nc.shape
    want = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
    nc_dev = nc.to(dev)
    keep = nc_dev.clone()
    path, dur, idx, status = tts.maximum_path_compact(nc_dev.to(dtype), t_y.to(dev), t_x.to(dev))
    torch.cuda.synchronize()
    assert path.dtype == dtype and path.device == nc_dev.device
    got = path.float().cpu().numpy().astype(np.int32)
    assert np.array_equal(got, want), f"path differs in {(got != want).sum()} cells"
    assert torch.equal(nc_dev.view(torch.int32), keep.view(torch.int32)), "input was modified"
    assert (status == 0).all()
    assert np.array_equal(dur.cpu().numpy(), want.sum(1))
    assert torch.equal(dur.sum(1).cpu(), t_y)          # duration-sum invariant
    widx = np.where(want.sum(2) > 0, want.argmax(2), -1)
    assert np.array_equal(idx.cpu().numpy(), widx)


@pytest.mark.parametrize("B,S,T", [(3, 17, 50), (2, 128, 130), (4, 129, 400), (2, 256, 1024), (2, 300, 700),
                                   (1, 515, 1100), (1, 1024, 1030), (5, 7, 33), (2, 64, 64)])
@pytest.mark.parametrize("ties", [False, True])
def test_full_lengths(cuda_device, B, S, T, ties):
    nc = synthetic.neg_cent_like(B, S, T, seed=B * 1000 + S, ties=ties)
    t_x, t_y = synthetic.full_lengths(B, S, T)
    _check(nc, t_x, t_y, cuda_device)


@pytest.mark.parametrize("B,S,T,seed", [(16, 200, 800, 0), (32, 256, 1024, 1), (9, 190, 999, 2), (7, 333, 1500, 3)])
@pytest.mark.parametrize("ties", [False, True])
def test_ragged(cuda_device, B, S, T, seed, ties):
    nc = synthetic.neg_cent_like(B, S, T, seed=seed, ties=ties)
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed=seed)
    _check(nc, t_x, t_y, cuda_device)


def test_edge_lengths(cuda_device):
    """t_x == t_y (identity diagonal), t_x == 1 (column 0), t_y == 1, rows/cols past the lengths stay zero."""
    S, T = 40, 90
    t_x = torch.tensor([40, 1, 1, 13, 40, 2, 33], dtype=torch.int32)
    t_y = torch.tensor([40, 90, 1, 13, 90, 2, 34], dtype=torch.int32)
    nc = synthetic.neg_cent_like(len(t_x), S, T, seed=5)
    _check(nc, t_x, t_y, cuda_device)


def test_degenerate_values(cuda_device):
    """all-zero cost, huge negative cost (below the -1e9 sentinel), constant rows."""
    S, T = 50, 200
    t_x, t_y = synthetic.full_lengths(4, S, T)
    nc = torch.zeros(4, T, S)
    nc[1] = -3e7          # accumulates past -1e9 after ~33 rows
    nc[2] = synthetic.neg_cent_like(1, S, T, seed=9)[0].round()
    nc[3, :, ::2] = -1.0
    _check(nc, t_x, t_y, cuda_device)


def test_nan_inf_inputs(cuda_device):
    """NaN/Inf cells take the same branches as the reference's compiled comparisons."""
    S, T = 30, 120
    t_x, t_y = synthetic.full_lengths(2, S, T)
    nc = synthetic.neg_cent_like(2, S, T, seed=11)
    nc[0, 40, 10] = float("nan")
    nc[0, 41, 11] = float("-inf")
    nc[1, 5, 3] = float("inf")
    _check(nc, t_x, t_y, cuda_device)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float64])
def test_dtypes(cuda_device, dtype):
    """path comes back in neg_cent.dtype (__init__.py:19); the DP itself runs on the fp32 cast (:13)."""
    B, S, T = 3, 60, 250
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed=4)
    nc = synthetic.neg_cent_like(B, S, T, seed=4).to(dtype)
    want = mas_oracle.maximum_path_c(nc.float().numpy(), t_y.numpy(), t_x.numpy())
    path = tts.maximum_path(nc.to(cuda_device), _mask(t_x, t_y, S, T).to(cuda_device))
    assert path.dtype == dtype
    assert np.array_equal(path.float().cpu().numpy().astype(np.int32), want)


def test_mask_api_matches_reference_wrapper(cuda_device):
    """maximum_path(neg_cent, mask) with the dense mask of models.py:1249."""
    B, S, T = 6, 77, 301
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed=8)
    nc = synthetic.neg_cent_like(B, S, T, seed=8)
    mask = _mask(t_x, t_y, S, T)
    want = mas_oracle.maximum_path(nc.numpy(), mask.numpy())
    got = tts.maximum_path(nc.to(cuda_device), mask.to(cuda_device))
    assert np.array_equal(got.cpu().numpy().astype(np.int32), want)
    ty2, tx2 = tts.lengths_from_mask(mask.to(cuda_device))
    assert torch.equal(ty2.cpu(), t_y) and torch.equal(tx2.cpu(), t_x)
    # bool mask takes the torch slice path
    got2 = tts.maximum_path(nc.to(cuda_device), mask.bool().to(cuda_device))
    assert torch.equal(got, got2)


def test_bad_lengths_are_rejected_not_emulated(cuda_device):
    """t_x > t_y and t_x == 0 are UB in the reference (core.pyx:30-33); here: zero path + status."""
    S, T = 20, 30
    t_x = torch.tensor([10, 0, 20, 5], dtype=torch.int32)
    t_y = torch.tensor([5, 10, 31, 30], dtype=torch.int32)
    nc = synthetic.neg_cent_like(4, S, T, seed=2)
    path, dur, idx, status = tts.maximum_path_compact(nc.to(cuda_device), t_y.to(cuda_device), t_x.to(cuda_device))
    assert status.cpu().tolist() == [1, 1, 1, 0]
    assert path[:3].abs().sum().item() == 0 and dur[:3].sum().item() == 0 and (idx[:3] == -1).all()
    want = mas_oracle.maximum_path_c(nc[3:].numpy(), t_y[3:].numpy(), t_x[3:].numpy())
    assert np.array_equal(path[3:].cpu().numpy().astype(np.int32), want)


def test_cpu_tensor_fails_loudly():
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        tts.maximum_path(torch.zeros(1, 4, 2), torch.ones(1, 4, 2))


def test_long_utterance_spill_path(cuda_device):
    """BASELINE config 4 shape (direction bits exceed shared memory -> workspace spill), small batch."""
    B, S, T = 2, 600, 4000
    nc = synthetic.neg_cent_like(B, S, T, seed=6)
    t_x = torch.tensor([600, 431], dtype=torch.int32)
    t_y = torch.tensor([4000, 3127], dtype=torch.int32)
    _check(nc, t_x, t_y, cuda_device)


def test_full_size_invariants(cuda_device):
    """BASELINE config 5 size (B=512 ragged): properties that need no oracle run at this size,
    plus bit-exactness on a sample of utterances."""
    B, S, T = 512, 256, 1024
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed=12)
    g = torch.Generator(device=cuda_device).manual_seed(3)
    nc = torch.randn((B, T, S), generator=g, device=cuda_device) * 50 - 470
    path, dur, idx, status = tts.maximum_path_compact(nc, t_y.to(cuda_device), t_x.to(cuda_device))
    assert (status == 0).all()
    assert torch.equal(path.sum((1, 2)).to(torch.int32).cpu(), t_y)       # exactly one 1 per valid row
    assert torch.equal(dur.sum(1).cpu(), t_y)
    rows = torch.arange(T, device=cuda_device)[None, :] < t_y.to(cuda_device)[:, None]
    assert torch.equal(path.sum(2) > 0, rows)
    step = idx[:, 1:] - idx[:, :-1]
    valid = rows[:, 1:]
    assert ((step[valid] == 0) | (step[valid] == 1)).all()               # monotone, steps of 0/1
    assert (idx[:, 0] == 0).all()
    last = idx.gather(1, (t_y.to(cuda_device).long() - 1)[:, None])[:, 0]
    assert torch.equal(last.cpu(), t_x - 1)
    sample = [0, 1, 100, 255, 511]
    want = mas_oracle.maximum_path_c(nc[sample].cpu().numpy(), t_y[sample].numpy(), t_x[sample].numpy())
    assert np.array_equal(path[sample].cpu().numpy().astype(np.int32), want)


@pytest.mark.parametrize("vk", ["1", "0"])
@pytest.mark.parametrize("B,S,T,ties", [(4, 256, 1024, False), (3, 100, 333, True), (5, 64, 64, False), (2, 200, 40 * 32 + 1, True)])
def test_value_origin_warp_split(cuda_device, mas_env, B, S, T, ties, vk):
    """The forward DP on value warps + origin warps (MAS_DP_VK, the default where the shape allows it) and on
    single-role warps (MAS_DP_VK=0, what noise-scaled MAS and long texts use) are both bit-exact."""
    mas_env(MAS_DP_VK=vk)
    nc = synthetic.neg_cent_like(B, S, T, seed=S + T, ties=ties)
    t_x, t_y = synthetic.ragged_lengths(B, S, T, seed=B)
    t_y = torch.maximum(t_y, t_x)
    nc[0, T // 3, S // 2] = float("nan")          # the exact second pass as well
    _check(nc, t_x, t_y, cuda_device)
