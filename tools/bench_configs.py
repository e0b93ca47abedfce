"""Device time of the hot path on every BASELINE.json config (CUDA-graph replay, CUDA events, one GPU):
standalone maximum_path, the alignment call without noise (fused kernel where the shape allows it) and
with VITS2 noise.  Inputs rotate over 2 sets; sizes are the configs' own."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib
dev = torch.device("cuda:0")
L = _lib.lib()
D = 192

def graph_time(fn, n_sets):
    for i in range(n_sets): fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(2 * n_sets): fn(i)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (3 * 2 * n_sets) * 1e3

names = sys.argv[1:] or ["c1", "c2", "c3", "c4"]
for name in names:
    B, S, T, ragged = synthetic.CONFIGS[name]
    t_x, t_y = synthetic.config_lengths(name)
    ty, tx = t_y.to(dev), t_x.to(dev)
    n_sets = 2
    ins = []
    for i in range(n_sets):
        z, m, l, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=i)
        ins.append((z.to(dev), m.to(dev), l.to(dev)))
    ncs = [torch.randn((B, T, S), device=dev) * 50 - 470 for _ in range(n_sets)]
    noise = torch.randn((B, T, S), device=dev)
    path = torch.empty((B, T, S), device=dev); dur = torch.empty((B, S), dtype=torch.int32, device=dev)
    idx = torch.empty((B, T), dtype=torch.int32, device=dev); status = torch.empty(B, dtype=torch.int32, device=dev)
    ws = torch.empty(max(L.mas_fused_align_workspace_bytes(B, D, T, S, 1), L.mas_maximum_path_workspace_bytes(B, T, S), 256),
                     dtype=torch.uint8, device=dev)
    def st(): return torch.cuda.current_stream().cuda_stream
    def mp(i):
        rc = L.mas_maximum_path_f32(ncs[i % n_sets].data_ptr(), ty.data_ptr(), tx.data_ptr(), path.data_ptr(), 0, dur.data_ptr(),
                                    idx.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), B, T, S, st())
        assert rc == 0, rc
    def al(i, nz=None, scale=0.0):
        z, m, l = ins[i % n_sets]
        rc = L.mas_fused_align_f32(z.data_ptr(), m.data_ptr(), l.data_ptr(), ty.data_ptr(), tx.data_ptr(),
                                   None if nz is None else nz.data_ptr(), scale, path.data_ptr(), 0, dur.data_ptr(),
                                   idx.data_ptr(), status.data_ptr(), None, ws.data_ptr(), ws.numel(), B, D, T, S, st())
        assert rc == 0, rc
    t_mp = graph_time(mp, n_sets)
    t_al = graph_time(lambda i: al(i), n_sets)
    t_nz = graph_time(lambda i: al(i, noise, 0.01), n_sets)
    ok = bool((status == 0).all())
    mas_b = 2 * 4 * T * S; fus_b = 4 * D * T + 8 * D * S + 4 * T * S
    print(f"{name}: B={B} S={S} T={T} ragged={ragged} | maximum_path {t_mp:7.1f} us ({B / t_mp:6.3f} M align/s, "
          f"{mas_b * B / t_mp / 1e3:6.0f} GB/s) | align {t_al:7.1f} us ({B / t_al:6.3f} M/s, {fus_b * B / t_al / 1e3:6.0f} GB/s) | "
          f"align+noise {t_nz:7.1f} us ({B / t_nz:6.3f} M/s) | status ok={ok}", flush=True)
    del ins, ncs, noise, path, ws
    torch.cuda.empty_cache()
