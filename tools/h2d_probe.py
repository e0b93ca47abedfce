"""H2D copy bandwidth from pinned host memory with 1, 2, 4 concurrent streams (what bounds bench.py's e2e number)."""
import torch, time
dev = torch.device("cuda:0")
N = 75_497_984
h = torch.empty(N, dtype=torch.uint8).pin_memory()
d = torch.empty(N, dtype=torch.uint8, device=dev)
for ns in (1, 2, 3, 4, 8):
    streams = [torch.cuda.Stream(dev) for _ in range(ns)]
    cuts = [N * i // ns for i in range(ns + 1)]
    def go():
        for s, a, b in zip(streams, cuts[:-1], cuts[1:]):
            with torch.cuda.stream(s):
                d[a:b].copy_(h[a:b], non_blocking=True)
    for _ in range(3): go()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print(f"{ns} streams: {N / dt / 1e9:.1f} GB/s ({dt * 1e3:.3f} ms per 75.5 MB)")
