#!/usr/bin/env python
"""bench.py -- MAS alignments/sec for the fused neg_cent + maximum_path hot path.

  python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A step is one pass of the hot path over one batch of synthetic input:
BASELINE.json configs[1] -- fused neg_cent + MAS, B=64, T_text=256, T_mel=1024,
192-channel prior -- per GPU (weak scaling: every rank aligns its own B=64 shard,
no data-path collective).  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B, S, T, D = 64, 256, 1024, 192
METRIC = "MAS alignments/sec (B=64,T_text=256,T_mel=1024)"
UNIT = "alignments/s"
# algorithmic bytes / flops per alignment
FUSED_BYTES = 4 * D * T + 2 * 4 * D * S + 4 * T * S        # read z_p, m_p, logs_p; write dense fp32 path
MAS_BYTES = 2 * 4 * T * S                                   # read neg_cent, write path
COST_BYTES = 4 * D * T + 2 * 4 * D * S + 4 * T * S          # read z_p, m_p, logs_p; write neg_cent
COST_FLOPS = 4 * T * S * D                                  # two K=D contractions
NOISE_BYTES = FUSED_BYTES + 4 * T * S                       # + the randn_like draw (models.py:1244)
COMPACT_BYTES = FUSED_BYTES - 4 * T * S + 4 * (T + S)       # no dense path: idx [T] + durations [S] instead
CPU_BASELINE_REPS = 120                                     # ~10 s of host work at ~90 ms per batch
REGIONS = 41                                                # timed regions of exactly K steps each; the median is reported
WORKLOAD = "fused neg_cent+MAS, B=64 per GPU, T_text=256, T_mel=1024, D=192 (BASELINE configs[1])"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="launch through the C ABI every step instead of graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------
# the reference's CPU path (oracle/_ref when it was compiled, else the C port)
# --------------------------------------------------------------------------
def cpu_reference_step(inputs):
    """models.py:1224-1256 on the host, as the reference runs it without a GPU:
    torch CPU ops for neg_cent (all threads) + the Cython kernel (serial as shipped)."""
    import torch
    from oracle import mas_oracle

    z_p, m_p, logs_p, x_mask, y_mask = inputs
    with torch.no_grad():
        nc = mas_oracle.neg_cent_torch(z_p, m_p, logs_p)
        mask = (x_mask.unsqueeze(2) * y_mask.unsqueeze(-1)).squeeze(1)          # models.py:1249
        values = nc.numpy().astype("float32")                                    # __init__.py:13
        t_ys = mask.sum(1)[:, 0].numpy().astype("int32")                         # __init__.py:16
        t_xs = mask.sum(2)[:, 0].numpy().astype("int32")                         # __init__.py:17
        if mas_oracle.ref_core() is not None:
            path = mas_oracle.ref_maximum_path_c(values, t_ys, t_xs)
        else:
            path = mas_oracle.maximum_path_c(values, t_ys, t_xs)
        attn = torch.from_numpy(path).to(dtype=nc.dtype)
        w = attn.sum(1)                                                          # models.py:1256
    return w


def cpu_kind():
    from oracle import mas_oracle

    return "reference" if mas_oracle.ref_core() is not None else "port"


def time_cpu(inputs, reps: int):
    import torch

    cpu_reference_step(inputs)  # warm
    best = []
    for _ in range(reps):
        t0 = time.perf_counter()
        cpu_reference_step(inputs)
        best.append(time.perf_counter() - t0)
    mean = sum(best) / len(best)
    return B / mean, mean, torch.get_num_threads()


def make_inputs(seed: int):
    from torch_tts_b200 import synthetic

    t_x, t_y = synthetic.full_lengths(B, S, T)
    z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=seed)
    return (z_p, m_p, logs_p, x_mask, y_mask), t_x, t_y


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    # torchrun exports OMP_NUM_THREADS=1 to its workers: give the reference every host core it can use
    torch.set_num_threads(os.cpu_count() or 1)
    inputs, _, _ = make_inputs(0)
    for _ in range(max(args.warmup, 1)):
        cpu_reference_step(inputs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(inputs)
    dt = time.perf_counter() - t0
    value = B * args.steps / dt
    threads = torch.get_num_threads()
    sample = f"{args.steps} passes over one full B={B} batch (torch CPU neg_cent on {threads} threads + serial Cython MAS)"
    emit(({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "host": f"reference CPU path on the host cores, rank 0 only, one B={B} batch per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": cpu_kind(), "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args):
    import torch
    import torch.distributed as dist

    import torch_tts_b200 as tts
    from torch_tts_b200 import _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    # ---- inputs: NSETS rotating buffer sets so consecutive steps never reuse L2-resident data
    NSETS = 3
    host_sets, dev_sets = [], []
    for i in range(NSETS):
        (z_p, m_p, logs_p, x_mask, y_mask), t_x, t_y = make_inputs(rank * 100 + i)
        host_sets.append(tuple(t.pin_memory() for t in (z_p, m_p, logs_p, t_y, t_x)))
        dev_sets.append(tuple(t.to(dev) for t in (z_p, m_p, logs_p, t_y, t_x)))
    cpu_inputs = (z_p, m_p, logs_p, x_mask, y_mask)
    plans = [tts.AlignPlan(B, D, T, S, dev) for _ in range(NSETS)]   # separate outputs + workspace per set
    use_graph = not args.no_graph
    L.mas_take_launch_count()
    plans[0].run(*dev_sets[0])
    torch.cuda.synchronize()
    launches_per_step = L.mas_take_launch_count()
    if use_graph:
        for i in range(NSETS):
            plans[i].capture(0, *dev_sets[i])

    def step(i):
        k = i % NSETS
        if use_graph:
            plans[k].replay(0)
        else:
            plans[k].run(*dev_sets[k])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # REGIONS timed regions of exactly K steps each (barrier + synchronize on both sides of every region); the
    # median region is the reported one, so that the clock sampler sees more than two samples under load
    region_ms = []
    for _ in range(REGIONS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for i in range(args.steps):
            step(i)
        e1.record()
        barrier()
        region_ms.append(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms = sorted(region_ms)[len(region_ms) // 2]

    # ---- parity of what was just timed (outside the timed region): every buffer set's idx / durations against the
    #      CPU oracle -- (1) exactly the MAS optimum of the cost the GPU computes, (2) >= 99.99 % of cells on the
    #      reference expression's path with identical duration sums
    parity = None
    if rank == 0:
        import numpy as np
        from oracle import mas_oracle

        parity = {"checked_sets": NSETS, "exact_on_gpu_cost": True, "min_cell_agreement": 1.0, "duration_sums_identical": True,
                  "status_ok": True}
        for k in range(NSETS):
            z, m, l, ty, tx = dev_sets[k]
            hz, hm, hl, hty, htx = (t.cpu() for t in (z, m, l, ty, tx))
            idx = plans[k].idx.cpu().numpy()
            nc_gpu = tts.neg_cent(z, m, l).cpu().numpy()
            want = mas_oracle.maximum_path_c(nc_gpu, hty.numpy(), htx.numpy())
            parity["exact_on_gpu_cost"] &= bool(np.array_equal(np.where(want.sum(2) > 0, want.argmax(2), -1), idx))
            ref = mas_oracle.maximum_path_c(mas_oracle.neg_cent_torch(hz, hm, hl).numpy(), hty.numpy(), htx.numpy())
            got = plans[k].path.cpu().numpy().astype(np.int32)
            parity["min_cell_agreement"] = min(parity["min_cell_agreement"], float((got == ref).mean()))
            parity["duration_sums_identical"] &= bool(np.array_equal(plans[k].dur.cpu().numpy().sum(1), ref.sum((1, 2))))
            parity["status_ok"] &= bool((plans[k].status == 0).all())
        parity["ok"] = bool(parity["exact_on_gpu_cost"] and parity["min_cell_agreement"] >= 0.9999 and
                            parity["duration_sums_identical"] and parity["status_ok"])

    # ---- per-kernel timing on the same stream, each call sequence replayed from a CUDA graph so that
    #      host launch overhead is not in the number (dominant kernel -> roofline)
    kern = {}
    st = torch.cuda.current_stream(dev).cuda_stream
    cost_ws = torch.empty(L.mas_neg_cent_workspace_bytes(B, D, T, S), dtype=torch.uint8, device=dev)
    dp_ws = torch.empty(max(L.mas_maximum_path_workspace_bytes(B, T, S), 256), dtype=torch.uint8, device=dev)
    nc_sets = [torch.empty((B, T, S), dtype=torch.float32, device=dev) for _ in range(NSETS)]

    def cost_fn(i, stream):
        k = i % NSETS
        z, m, l, _, _ = dev_sets[k]
        rc = L.mas_neg_cent_f32(z.data_ptr(), m.data_ptr(), l.data_ptr(), nc_sets[k].data_ptr(), None,
                                cost_ws.data_ptr(), cost_ws.numel(), B, D, T, S, stream)
        assert rc == 0, rc

    def dp_fn(i, stream):
        k = i % NSETS
        _, _, _, ty, tx = dev_sets[k]
        p = plans[k]
        rc = L.mas_maximum_path_f32(nc_sets[k].data_ptr(), ty.data_ptr(), tx.data_ptr(), p.path.data_ptr(), 0,
                                    p.dur.data_ptr(), p.idx.data_ptr(), p.status.data_ptr(), dp_ws.data_ptr(),
                                    dp_ws.numel(), B, T, S, stream)
        assert rc == 0, rc

    def time_graph(fn, env=None):
        old = {}
        for k_, v_ in (env or {}).items():
            old[k_] = os.environ.get(k_)
            os.environ[k_] = v_
        _lib.reload_config()
        try:
            for i in range(NSETS):
                fn(i, st)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                cs = torch.cuda.current_stream(dev).cuda_stream
                for i in range(2 * NSETS):
                    fn(i, cs)
            g.replay()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a_, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a_.record()
                g.replay()
                b2.record()
                torch.cuda.synchronize()
                ts.append(a_.elapsed_time(b2) / (2 * NSETS))
            return sorted(ts)[len(ts) // 2]
        finally:
            for k_, v_ in old.items():
                if v_ is None:
                    os.environ.pop(k_, None)
                else:
                    os.environ[k_] = v_
            _lib.reload_config()

    kern["neg_cent_ms"] = time_graph(cost_fn)                      # prior preparation + contraction, standalone
    kern["prior_prep_ms"] = time_graph(cost_fn, {"MAS_STAGE": "1"})      # the prior preparation kernel alone
    kern["maximum_path_ms"] = time_graph(dp_fn)                    # standalone MAS on a resident cost plane
    kern["fused_kernel_ms"] = max(ms / args.steps - kern["prior_prep_ms"], 1e-6)   # the step is prior prep + fused kernel

    # ---- the other two forms of the same call, same shapes, same timing method (CUDA-graph replay, 3 rotating sets):
    #      VITS2 noise-scaled MAS (what cli.py:268-271 runs every training step) and compact outputs only (8f-3)
    noise_sets = [torch.randn((B, T, S), device=dev, generator=torch.Generator(device=dev).manual_seed(7 + i))
                  for i in range(NSETS)]
    lean = tts.AlignPlan(B, D, T, S, dev, want_path=False)

    def plan_fn(plan, noisy):
        def fn(i, stream):
            k = i % NSETS
            z, m, l, ty, tx = dev_sets[k]
            plan.run(z, m, l, ty, tx, noise_sets[k] if noisy else None, 0.01 if noisy else 0.0)   # (current stream)
        return fn

    extra = {}
    for name, plan, noisy, nbytes in (("noise_scaled_mas", plans[0], True, NOISE_BYTES),
                                      ("compact_outputs", lean, False, COMPACT_BYTES),
                                      ("noise_scaled_mas_compact", lean, True, NOISE_BYTES - 4 * T * S + 4 * (T + S))):
        L.mas_take_launch_count()
        plan_fn(plan, noisy)(0, st)
        torch.cuda.synchronize()
        n_launch = int(L.mas_take_launch_count())
        t_ms = time_graph(plan_fn(plan, noisy))
        extra[name] = {"ms_per_step": t_ms, "alignments_per_s": B / (t_ms * 1e-3), "launches_per_step": n_launch,
                       "algorithmic_bytes_per_step": nbytes * B, "status_ok": bool((plan.status == 0).all()),
                       "duration_sums_ok": bool((plan.dur.sum(1) == dev_sets[(2 * NSETS - 1) % NSETS][3]).all())}
    del noise_sets

    # ---- end to end through the public call with HOST buffers: pinned host -> H2D (copy stream, double
    #      buffered) -> align -> D2H of durations + compact path
    out_host = torch.empty((B, S + T), dtype=torch.int32).pin_memory()
    h2d = sum(t.numel() * t.element_size() for t in host_sets[0])
    d2h = out_host.numel() * 4
    stage = [tuple(torch.empty_like(t, device=dev) for t in host_sets[0]) for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    main_stream = torch.cuda.current_stream(dev)

    def e2e_step(i):
        hs, ds, p, k = host_sets[i % NSETS], stage[i % 2], plans[i % NSETS], i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[k])           # the step that last read this staging set is done
            for h, d_ in zip(hs, ds):
                d_.copy_(h, non_blocking=True)
            ev_ready[k].record(copy_stream)
        main_stream.wait_event(ev_ready[k])
        p.run(*ds)
        out_host[:, :S].copy_(p.dur, non_blocking=True)
        out_host[:, S:].copy_(p.idx, non_blocking=True)
        ev_free[k].record(main_stream)

    for k in range(2):
        ev_free[k].record(main_stream)
    for i in range(3):
        e2e_step(i)
    barrier()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(5, min(args.steps, 20))
    a.record()
    for i in range(e2e_steps):
        e2e_step(i)
    b_.record()
    barrier()
    e2e_ms = a.elapsed_time(b_)

    times = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()
    value = world * B * args.steps / (ms * 1e-3)
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, mean, threads = time_cpu(cpu_inputs, reps=CPU_BASELINE_REPS)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": cpu_kind(),
               "sample": f"{CPU_BASELINE_REPS} passes over one full B={B} batch ({mean * 1e3:.1f} ms each): torch CPU neg_cent on "
                         f"{threads} threads + serial Cython MAS as shipped; host has {os.cpu_count()} cpus"}

    if rank == 0:
        hbm, bf16, how = peaks()
        # dominant kernel of the step: the fused contraction + DP kernel.  Algorithmic bytes per launch
        # (SURVEY 8d): read z_p, m_p, logs_p once, write the dense fp32 path once = 2.228 MB per alignment.
        t_f = kern["fused_kernel_ms"] * 1e-3
        ach = FUSED_BYTES * B / t_f / 1e9
        traffic, traffic_src = None, None
        for tname in ("r2_fused_traffic.json", "r1_fused_traffic.json"):
            tpath = os.path.join(ROOT, "profiles", tname)
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
                traffic_src = f"profiles/{tname}: dram__bytes_read.sum + dram__bytes_write.sum of one committed `ncu --set full` capture of this kernel, NOT measured in this run"
                break
        roof = {"kernel": "mas_fused_pair_kernel", "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s",
                "frac": ach / hbm, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": how + " (copy bandwidth, MEASURED_PEAKS.json)",
                "duration_source": "step time (CUDA events, median region) minus the prior-preparation kernel timed alone",
                "algorithmic_bytes_per_launch": FUSED_BYTES * B,
                "tensor": {"achieved": COST_FLOPS * B / t_f / 1e12, "peak": bf16, "unit": "TFLOP/s (algorithmic fp32 "
                           "flops; the split-bf16 scheme issues 3x as many on the tensor pipe)"}}
        step_s = ms * 1e-3 / args.steps
        fused = {"bound": "hbm", "achieved": FUSED_BYTES * B / step_s / 1e9, "peak": hbm, "unit": "GB/s"}
        fused["frac"] = fused["achieved"] / hbm
        for e_ in extra.values():
            e_["roofline_frac"] = e_["algorithmic_bytes_per_step"] / (e_["ms_per_step"] * 1e-3) / 1e9 / hbm
        emit(({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "parallelism": f"dp{world} (batch-sharded, no data-path collective)",
                       "l2": f"{NSETS} rotating input/output buffer sets (~{NSETS * 190} MB) > 126 MB L2",
                       "launch": "cuda-graph replay" if use_graph else "C-ABI call per step"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps,
                    "note": "AlignPlan.run on pinned HOST buffers: H2D of z_p/m_p/logs_p/lengths on a copy stream (double "
                            "buffered), D2H of the COMPACT result (durations + idx; the dense path stays on the device "
                            "where its consumers are). PCIe-bound; at N>1 all GPUs share one host's DRAM/PCIe root"},
            "gpu_launches": int(launches_per_step * args.steps),
            "parity_checked": bool(parity and parity["ok"]), "parity": parity,
            "region_ms": region_ms, "regions": REGIONS,
            "clocks": clocks, "roofline": roof, "roofline_whole_step": fused, "kernels_ms": kern,
            "extra": extra,
            "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(obj) -> None:
    """The one JSON line, on the process's ORIGINAL stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, line)


def main():
    global _RESULT_FD
    args = parse()
    # stdout carries exactly one JSON line: native libraries that print to fd 1 (NCCL's version banner under
    # torchrun) are sent to stderr for the rest of the run, and the result goes out on a copy of the real stdout
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
