"""Numpy model of the CUDA DP kernel's algorithm (torch_tts_b200/csrc/mas_dp.cu):
register-row forward DP over ALL columns (no lower band edge), 1 decision bit per
cell, checkpointed origins every 32 rows, two-level backtrack.  Used on CPU to
check the formulation against the oracle; the kernel itself is checked on the GPU.
"""
import numpy as np

NEG = np.float32(-1e9)
CHECK = 32


def dp_model(cost: np.ndarray, t_y: int, t_x: int, s_pad: int):
    T, S = cost.shape
    pad = np.zeros((T, s_pad), np.float32)
    pad[:, :S] = cost
    # columns >= S see garbage in the kernel; model that with large noise
    pad[:, S:] = np.float32(123.0)
    x = np.arange(s_pad)
    v = np.full(s_pad, NEG, np.float32)
    org = x.copy()
    bits = np.zeros((t_y, s_pad), bool)
    hop = np.zeros((T // CHECK + 2, s_pad), np.int64)
    for y in range(t_y):
        v_prev = np.empty_like(v)
        v_prev[1:] = v[:-1]
        v_prev[0] = np.float32(0.0) if y == 0 else NEG
        o_prev = np.empty_like(org)
        o_prev[1:] = org[:-1]
        o_prev[0] = 0
        m = np.where(v > v_prev, v, v_prev)
        diag = (v < v_prev) | (x == y)
        diag[0] = False
        nv = (pad[y] + m).astype(np.float32)
        no = np.where(diag, o_prev, org)
        band = x <= y
        v = np.where(band, nv, v)
        org = np.where(band, no, org)
        bits[y] = diag
        if y % CHECK == 0 and y > 0:
            hop[y // CHECK] = x - org
            org = x.copy()
    hop[0] = x - org
    y_last = t_y - 1
    J = y_last // CHECK
    entry = np.zeros(J + 2, np.int64)
    c = t_x - 1
    c -= hop[0][c]
    entry[J] = c
    for j in range(J, 0, -1):
        c -= hop[j][c]
        entry[j - 1] = c
    idx = np.zeros(t_y, np.int64)
    for j in range(J + 1):
        y_lo = CHECK * j + 1
        y_top = CHECK * j + CHECK
        if y_top <= y_last:
            cur = entry[j + 1]
        else:
            y_top = y_last
            cur = t_x - 1
        for y in range(y_top, y_lo - 1, -1):
            idx[y] = cur
            cur -= int(bits[y, cur])
        if j == 0:
            assert cur == 0
        else:
            assert cur == entry[j], (j, cur, entry[j])
    path = np.zeros((T, S), np.int32)
    path[np.arange(t_y), idx] = 1
    return path
