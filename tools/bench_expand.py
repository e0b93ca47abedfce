"""Device time of the prior expansion (gather), its backward (segmented sum) and logw_ at a BASELINE shape,
CUDA-graph replay, against their algorithmic HBM bytes; the reference's one-hot matmuls on the same GPU beside it."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
B, S, T, ragged = synthetic.CONFIGS[name]
D = synthetic.D_PRIOR
dev = torch.device("cuda:0")
t_x, t_y = synthetic.config_lengths(name)
z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=0)
m_d, l_d = m_p.to(dev), logs_p.to(dev)
attn, w, (idx, dur, status) = tts.align(z_p.to(dev), m_d, l_d, x_mask.to(dev), y_mask.to(dev), return_compact=True)
g_m, g_l = torch.randn((B, D, T), device=dev), torch.randn((B, D, T), device=dev)
L = tts._lib.lib()
p = tts._lib.ptr
m_out, l_out = torch.empty((B, D, T), device=dev), torch.empty((B, D, T), device=dev)
gm_p, gl_p = torch.empty((B, D, S), device=dev), torch.empty((B, D, S), device=dev)
lw = torch.empty((B, S), device=dev)
txd = t_x.to(dev).int()


def timed(fn, reps=20):
    for _ in range(3):
        rc = fn()
        assert not isinstance(rc, int) or rc == 0, f"C ABI call failed with status {rc}"
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(5): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * 5) * 1e3


st = lambda: torch.cuda.current_stream().cuda_stream
fwd = timed(lambda: L.mas_expand_prior_f32(p(m_d), p(l_d), p(idx), p(m_out), p(l_out), B, D, T, S, st()))
bwd = timed(lambda: L.mas_expand_prior_backward_f32(p(g_m), p(g_l), p(dur), p(gm_p), p(gl_p), B, D, T, S, st()))
lg = timed(lambda: L.mas_logw_f32(p(dur), p(txd), p(lw), B, S, st()))
a2 = attn.squeeze(1)
ref = timed(lambda: (torch.matmul(a2, m_d.transpose(1, 2)), torch.matmul(a2, l_d.transpose(1, 2))), reps=5)
bytes_fwd = 2 * 4 * B * D * (S + T) + 4 * B * T
bytes_bwd = 2 * 4 * B * D * (S + T) + 4 * B * S
print(json.dumps({"config": name, "B": B, "S": S, "T": T, "D": D,
                  "expand_prior_us": round(fwd, 1), "expand_prior_GBps": round(bytes_fwd / fwd / 1e3, 0),
                  "expand_prior_backward_us": round(bwd, 1), "expand_prior_backward_GBps": round(bytes_bwd / bwd / 1e3, 0),
                  "logw_us": round(lg, 2), "reference_two_matmuls_on_gpu_us": round(ref, 1)}))
