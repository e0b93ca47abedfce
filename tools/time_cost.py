"""Time mas_neg_cent_f32 at BASELINE config 2 under the MAS_TC_DEBUG experiment masks
(1: no A stores, 2: no epilogue stores, 4: no MMA)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch_tts_b200 import synthetic, _lib
B, S, T, D = 64, 256, 1024, 192
dev = torch.device("cuda:0")
L = _lib.lib()
sets = []
for i in range(3):
    t_x, t_y = synthetic.full_lengths(B, S, T)
    z, m, l, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=i)
    sets.append((z.to(dev), m.to(dev), l.to(dev), torch.empty((B, T, S), device=dev)))
ws = torch.empty(L.mas_neg_cent_workspace_bytes(B, D, T, S), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
def run(i):
    z, m, l, o = sets[i % 3]
    rc = L.mas_neg_cent_f32(z.data_ptr(), m.data_ptr(), l.data_ptr(), o.data_ptr(), None, ws.data_ptr(), ws.numel(), B, D, T, S, st)
    assert rc == 0, rc
masks = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 3, 5, 6, 7]
for mask in masks:
    os.environ["MAS_TC_DEBUG"] = str(mask)
    for i in range(3): run(i)
    torch.cuda.synchronize()
    # one CUDA graph of 6 calls (2 per buffer set): GPU time only, no host launch overhead
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        st = torch.cuda.current_stream().cuda_stream
        for i in range(6): run(i)
    st = torch.cuda.current_stream().cuda_stream
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    print(f"MAS_TC_DEBUG={mask}: {a.elapsed_time(b)/30*1e3:.1f} us/call (prior images + contraction, graph replay)")
os.environ["MAS_TC_DEBUG"] = "0"
