"""Seeded synthetic LJSpeech-shaped inputs for tests and bench.
Generated on the CPU with a torch.Generator so CPU oracle and GPU path see
identical bits; callers move them to the device.
"""
from __future__ import annotations

import torch

# BASELINE.json configs: name -> (B, S=T_text, T=T_mel, ragged)
CONFIGS = {
    "c1": (16, 200, 800, False),
    "c2": (64, 256, 1024, False),
    "c3": (128, 256, 1024, True),
    "c4": (32, 600, 4000, False),
    "c5": (512, 256, 1024, True),
}
D_PRIOR = 192  # inter_channels, cli.py:159


def ragged_lengths(B: int, S: int, T: int, seed: int = 0):
    """t_x ~ U[S/4, S], t_y = clamp(4 t_x +- 20, t_x, T), sorted by t_y descending
    like TextAudioCollate (data_utils.py:164-166)."""
    g = torch.Generator().manual_seed(seed)
    t_x = torch.randint(max(1, S // 4), S + 1, (B,), generator=g)
    t_y = 4 * t_x + torch.randint(-20, 21, (B,), generator=g)
    t_y = torch.minimum(torch.maximum(t_y, t_x), torch.tensor(T))
    order = torch.argsort(t_y, descending=True, stable=True)
    return t_x[order].to(torch.int32), t_y[order].to(torch.int32)


def full_lengths(B: int, S: int, T: int):
    return torch.full((B,), S, dtype=torch.int32), torch.full((B,), T, dtype=torch.int32)


def masks(t_x, t_y, S: int, T: int):
    """x_mask [B,1,S], y_mask [B,1,T] as float (commons.sequence_mask, models.py:372-374)."""
    x_mask = (torch.arange(S)[None, :] < t_x[:, None]).float().unsqueeze(1)
    y_mask = (torch.arange(T)[None, :] < t_y[:, None]).float().unsqueeze(1)
    return x_mask, y_mask


def neg_cent_like(B: int, S: int, T: int, seed: int = 0, ties: bool = False):
    """Cost planes at the scale the real model produces at init (mean -470, std 50);
    ties=True rounds to multiples of 16 to force equal-value decisions."""
    g = torch.Generator().manual_seed(seed)
    nc = torch.randn((B, T, S), generator=g) * 50.0 - 470.0
    if ties:
        nc = torch.round(nc / 16.0) * 16.0
    return nc


def prior_inputs(B: int, S: int, T: int, t_x, t_y, D: int = D_PRIOR, seed: int = 0):
    """z_p ~ N(0,1) y_mask, m_p ~ N(0,0.5^2) x_mask, logs_p ~ N(-0.5,0.3^2) x_mask
    (padded positions exactly 0, as TextEncoder / the flow leave them)."""
    g = torch.Generator().manual_seed(seed)
    x_mask, y_mask = masks(t_x, t_y, S, T)
    z_p = torch.randn((B, D, T), generator=g) * y_mask
    m_p = torch.randn((B, D, S), generator=g) * 0.5 * x_mask
    logs_p = (torch.randn((B, D, S), generator=g) * 0.3 - 0.5) * x_mask
    return z_p, m_p, logs_p, x_mask, y_mask


def config_lengths(name: str, seed: int = 0):
    B, S, T, ragged = CONFIGS[name]
    return ragged_lengths(B, S, T, seed) if ragged else full_lengths(B, S, T)
