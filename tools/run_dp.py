"""Run the standalone MAS kernel a few times on a BASELINE shape (for ncu / timing)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B, S, T, ragged = synthetic.CONFIGS[name]
dev = torch.device("cuda:0")
t_x, t_y = synthetic.config_lengths(name)
g = torch.Generator(device=dev).manual_seed(0)
ncs = [torch.randn((B, T, S), generator=g, device=dev) * 50 - 470 for _ in range(3)]
L = _lib.lib()
ws = torch.empty(max(L.mas_maximum_path_workspace_bytes(B, T, S), 256), dtype=torch.uint8, device=dev)
path = torch.empty((B, T, S), device=dev); dur = torch.empty((B, S), dtype=torch.int32, device=dev)
idx = torch.empty((B, T), dtype=torch.int32, device=dev); status = torch.empty(B, dtype=torch.int32, device=dev)
ty, tx = t_y.to(dev), t_x.to(dev)
st = torch.cuda.current_stream().cuda_stream
def run(i):
    rc = L.mas_maximum_path_f32(ncs[i % 3].data_ptr(), ty.data_ptr(), tx.data_ptr(), path.data_ptr(), 0, dur.data_ptr(),
                                idx.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), B, T, S, st)
    assert rc == 0, rc
for i in range(3): run(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    st = torch.cuda.current_stream().cuda_stream
    for i in range(6): run(i)
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(max(1, reps // 6)): g.replay()
b.record(); torch.cuda.synchronize()
n = max(1, reps // 6) * 6
print(f"{name}: B={B} S={S} T={T} maximum_path {a.elapsed_time(b)/n*1e3:.1f} us/call (graph replay), "
      f"{B/(a.elapsed_time(b)/n*1e-3):.0f} align/s, status ok={bool((status==0).all())}")
