"""BASELINE.json configs[4] as written: multispeaker train_ms shapes, B=512 (ragged, S=256, T=1024, D=192),
batch-sharded across the ranks of one box, WITH the NCCL all-gather of the compact paths + durations inside the
timed region (train_ms.py:36-67, 231: one process per GPU, every rank aligns its own shard).

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_c5_sharded.py \
      [--steps K] [--noise] [--json OUT]

Strong scaling: the global batch is fixed at 512, every rank aligns 512 / N utterances.  Device-timed (CUDA events,
barrier + synchronize on both sides, MAX over ranks).  Rank 0 also checks the gathered result of the last step
against the CPU oracle on a sample of utterances (exact MAS optimum of the GPU's cost is covered by the tests; here:
>= 99.99 % agreement with the reference expression's path and identical duration sums) and prints one JSON line.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic
from torch_tts_b200.sharded import gather_compact

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=30)
ap.add_argument("--warmup", type=int, default=5)
ap.add_argument("--noise", action="store_true")
ap.add_argument("--json", default=None)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, S, T, _ = synthetic.CONFIGS["c5"]
D = synthetic.D_PRIOR
t_x, t_y = synthetic.config_lengths("c5", seed=5)
# rank r takes utterances r, r + N, ... of the length-sorted batch (DistributedBucketSampler strides the ranks the
# same way, data_utils.py:514): every rank gets the same mix of long and short utterances
ids = torch.arange(rank, B, world)
perm = torch.cat([torch.arange(r, B, world) for r in range(world)])     # order of the gathered result
n = ids.numel()
lo, hi = int(sum(len(range(r, B, world)) for r in range(rank))), int(sum(len(range(r, B, world)) for r in range(rank + 1)))
t_x_r, t_y_r = t_x[ids], t_y[ids]
z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(n, S, T, t_x_r, t_y_r, D, seed=1000 + rank)
NSETS = 3
sets = []
for i in range(NSETS):
    zz = z_p.roll(i, 0) if i else z_p          # distinct buffers so that consecutive steps do not hit L2
    sets.append((zz.to(dev).clone(), m_p.to(dev).clone(), logs_p.to(dev).clone()))
ty, tx = t_y_r.to(dev), t_x_r.to(dev)
noise = [torch.randn((n, T, S), device=dev) for _ in range(NSETS)] if args.noise else None
plans = [tts.AlignPlan(n, D, T, S, dev, with_noise=args.noise, want_path=True) for _ in range(NSETS)]


def step(i):
    k = i % NSETS
    z, m, l = sets[k]
    # set 0 carries the un-rolled inputs: the parity check below uses the last step run on it
    plans[k].run(z, m, l, ty, tx, noise[k] if noise else None, 0.01 if noise else 0.0)
    if world > 1:
        return gather_compact(plans[k].idx, plans[k].dur, batch=B, uniform=(B % world == 0))
    return plans[k].idx, plans[k].dur


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


for i in range(args.warmup):
    step(i)
barrier()
regions = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    regions.append(e0.elapsed_time(e1))
ms = sorted(regions)[len(regions) // 2]
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
# the un-timed variant without the gather, for the gather's share
barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.steps):
    k = i % NSETS
    z, m, l = sets[k]
    plans[k].run(z, m, l, ty, tx, noise[k] if noise else None, 0.01 if noise else 0.0)
e1.record()
barrier()
t2 = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t2, op=dist.ReduceOp.MAX)
ms_nogather = float(t2.item())

g_idx, g_dur = step(0)     # set 0: un-rolled inputs
torch.cuda.synchronize()
ok, agree = None, None
if not args.noise:
    # every rank checks a sample of ITS OWN utterances against the CPU oracle (inputs are rank-local)
    from oracle import mas_oracle
    sel = list(range(0, n, max(1, n // 4)))[:4]
    nc = mas_oracle.neg_cent_torch(z_p[sel], m_p[sel], logs_p[sel])
    ref = mas_oracle.maximum_path_c(nc.numpy(), t_y_r[sel].numpy(), t_x_r[sel].numpy())
    got = tts.expand_path(g_idx[lo:hi][sel].contiguous(), S).cpu().numpy().astype(np.int32)
    agree = float((got == ref).mean())
    ok = bool(agree >= 0.9999 and np.array_equal(g_dur.sum(1).cpu().numpy(), t_y[perm].numpy()))
    flag = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())
else:
    ok = bool(np.array_equal(g_dur.sum(1).cpu().numpy(), t_y[perm].numpy()))
if rank == 0:
    per = ms / args.steps
    bytes_per_alignment = 4 * D * T + 8 * D * S + 4 * T * S + (4 * T * S if args.noise else 0)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    out = {"config": "c5: B=512 ragged, S=256, T=1024, D=192, batch-sharded (strong scaling), NCCL all-gather of the compact "
                     "idx [512,1024] + durations [512,256] inside the timed region",
           "n_gpus": world, "utterances_per_rank": n, "noise_scaled": bool(args.noise), "steps": args.steps,
           "ms_per_step": per, "ms_per_step_without_gather": ms_nogather / args.steps,
           "alignments_per_s": B / (per * 1e-3),
           "roofline_frac_of_aggregate_hbm": bytes_per_alignment * B / (per * 1e-3) / 1e9 / (peaks["hbm_gbs"] * world),
           "gathered_result_ok": ok, "sampled_path_agreement": agree, "regions_ms": regions}
    line = json.dumps(out)
    print(line)
    if args.json:
        open(args.json, "w").write(line + "\n")
if world > 1:
    dist.destroy_process_group()
