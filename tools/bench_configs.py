"""Device time of the hot path on every BASELINE.json config (CUDA-graph replay, CUDA events, one GPU):
standalone maximum_path, the alignment call without noise (fused kernel where the shape allows it), with VITS2
noise, and the compact-output forms -- each with its roofline fraction on the SURVEY.md 8(d) algorithmic bytes and
the reference's CPU path timed beside it on the host cores.

  python tools/bench_configs.py [c1 c2 ...] [--json out.json] [--no-cpu]

Inputs rotate over 2 sets; sizes are the configs' own.  One JSON document (all configs) goes to --json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import torch_tts_b200 as tts
from torch_tts_b200 import synthetic, _lib

dev = torch.device("cuda:0")
L = _lib.lib()
D = 192
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured"
    return 6650.0, "fallback"


def graph_time(fn, n_sets, reps=5):
    for i in range(n_sets):
        fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(2 * n_sets):
            fn(i)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (2 * n_sets) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]          # median, us per call


def cpu_reference(name, z, m, l, t_x, t_y, S, T, reps):
    """the reference's CPU path on the host: torch CPU neg_cent + the compiled Cython kernel (serial as shipped)"""
    from oracle import mas_oracle

    kind = "reference" if mas_oracle.ref_core() is not None else "port"
    mp = mas_oracle.ref_maximum_path_c if kind == "reference" else mas_oracle.maximum_path_c
    nc = mas_oracle.neg_cent_torch(z, m, l)
    vals = nc.numpy()
    t0 = time.perf_counter()
    for _ in range(reps):
        mp(vals, t_y.numpy(), t_x.numpy())
    t_mas = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.no_grad():
            mas_oracle.neg_cent_torch(z, m, l)
    t_nc = (time.perf_counter() - t0) / reps
    return {"kind": kind, "maximum_path_c_ms": t_mas * 1e3, "neg_cent_torch_ms": t_nc * 1e3,
            "torch_threads": torch.get_num_threads(), "host_cpus": os.cpu_count(), "passes": reps}


def main():
    args = sys.argv[1:]
    out_path = None
    if "--json" in args:
        i = args.index("--json")
        out_path = args[i + 1]
        del args[i:i + 2]
    no_cpu = "--no-cpu" in args
    args = [a for a in args if a != "--no-cpu"]
    names = args or ["c1", "c2", "c3", "c4", "c5"]
    peak, how = hbm_peak()
    doc = {"hbm_peak_gbs": peak, "peak_source": how, "device": torch.cuda.get_device_name(0),
           "env": {k: v for k, v in os.environ.items() if k.startswith("MAS_")}, "configs": {}}
    for name in names:
        B, S, T, ragged = synthetic.CONFIGS[name]
        t_x, t_y = synthetic.config_lengths(name)
        ty, tx = t_y.to(dev), t_x.to(dev)
        n_sets = 2
        ins, host0 = [], None
        for i in range(n_sets):
            z, m, l, _, _ = synthetic.prior_inputs(B, S, T, t_x, t_y, D, seed=i)
            if i == 0:
                host0 = (z, m, l)
            ins.append((z.to(dev), m.to(dev), l.to(dev)))
        ncs = [torch.randn((B, T, S), device=dev) * 50 - 470 for _ in range(n_sets)]
        noise = torch.randn((B, T, S), device=dev)
        path = torch.empty((B, T, S), device=dev)
        dur = torch.empty((B, S), dtype=torch.int32, device=dev)
        idx = torch.empty((B, T), dtype=torch.int32, device=dev)
        status = torch.empty(B, dtype=torch.int32, device=dev)
        ws = torch.empty(max(L.mas_fused_align_workspace_bytes(B, D, T, S, 1), L.mas_maximum_path_workspace_bytes(B, T, S), 256),
                         dtype=torch.uint8, device=dev)

        def st():
            return torch.cuda.current_stream().cuda_stream

        def mp(i, dense=True):
            rc = L.mas_maximum_path_f32(ncs[i % n_sets].data_ptr(), ty.data_ptr(), tx.data_ptr(),
                                        path.data_ptr() if dense else None, 0, dur.data_ptr(),
                                        idx.data_ptr(), status.data_ptr(), ws.data_ptr(), ws.numel(), B, T, S, st())
            assert rc == 0, rc

        def al(i, nz=None, scale=0.0, dense=True):
            z, m, l = ins[i % n_sets]
            rc = L.mas_fused_align_f32(z.data_ptr(), m.data_ptr(), l.data_ptr(), ty.data_ptr(), tx.data_ptr(),
                                       None if nz is None else nz.data_ptr(), scale, path.data_ptr() if dense else None, 0,
                                       dur.data_ptr(), idx.data_ptr(), status.data_ptr(), None, ws.data_ptr(), ws.numel(),
                                       B, D, T, S, st())
            assert rc == 0, rc

        mas_b = 2 * 4 * T * S                         # read neg_cent, write path
        fus_b = 4 * D * T + 8 * D * S + 4 * T * S     # read z_p, m_p, logs_p, write path
        nz_b = fus_b + 4 * T * S                      # + the noise draw
        rows = {}

        def row(key, t_us, bytes_per_utt, launches):
            ach = bytes_per_utt * B / t_us / 1e3
            rows[key] = {"us": round(t_us, 2), "alignments_per_s": round(B / t_us * 1e6), "algorithmic_MB_per_alignment":
                         round(bytes_per_utt / 1e6, 4), "achieved_GBs": round(ach, 1), "roofline_frac": round(ach / peak, 4),
                         "launches": launches}

        def count(fn):
            L.mas_take_launch_count()
            fn(0)
            torch.cuda.synchronize()
            return int(L.mas_take_launch_count())

        row("maximum_path", graph_time(mp, n_sets), mas_b, count(mp))
        row("align", graph_time(lambda i: al(i), n_sets), fus_b, count(lambda i: al(i)))
        ok_plain = bool((status == 0).all()) and bool((dur.sum(1) == ty).all())
        row("align_noise", graph_time(lambda i: al(i, noise, 0.01), n_sets), nz_b, count(lambda i: al(i, noise, 0.01)))
        ok_noise = bool((status == 0).all()) and bool((dur.sum(1) == ty).all())
        compact = L.mas_b200_abi_version() >= 2
        if compact:
            row("maximum_path_compact", graph_time(lambda i: mp(i, False), n_sets), 4 * T * S, count(lambda i: mp(i, False)))
            row("align_compact", graph_time(lambda i: al(i, dense=False), n_sets), fus_b - 4 * T * S,
                count(lambda i: al(i, dense=False)))
            row("align_noise_compact", graph_time(lambda i: al(i, noise, 0.01, False), n_sets), nz_b - 4 * T * S,
                count(lambda i: al(i, noise, 0.01, False)))
        entry = {"B": B, "S": S, "T": T, "D": D, "ragged": ragged, "results_ok": ok_plain and ok_noise, "gpu": rows}
        if not no_cpu:
            reps = 3 if B * T * S <= 64 * 1024 * 256 else 1
            entry["cpu_reference"] = cpu_reference(name, *host0, t_x, t_y, S, T, reps)
            c = entry["cpu_reference"]
            c["alignments_per_s_mas_only"] = round(B / (c["maximum_path_c_ms"] * 1e-3))
            c["alignments_per_s_cost_plus_mas"] = round(B / ((c["maximum_path_c_ms"] + c["neg_cent_torch_ms"]) * 1e-3))
        doc["configs"][name] = entry
        g = rows
        print(f"{name}: B={B} S={S} T={T} ragged={ragged} | maximum_path {g['maximum_path']['us']:7.1f} us "
              f"({g['maximum_path']['roofline_frac']:.2f}) | align {g['align']['us']:7.1f} us ({g['align']['roofline_frac']:.2f}, "
              f"{g['align']['launches']} launches) | align+noise {g['align_noise']['us']:7.1f} us "
              f"({g['align_noise']['roofline_frac']:.2f}, {g['align_noise']['launches']} launches)"
              + (f" | compact: {g['maximum_path_compact']['us']:.1f} / {g['align_compact']['us']:.1f} / "
                 f"{g['align_noise_compact']['us']:.1f} us" if compact else "") + f" | ok={entry['results_ok']}", flush=True)
        del ins, ncs, noise, path, ws
        torch.cuda.empty_cache()
    if out_path:
        with open(out_path, "w") as f:
            json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
