"""torchrun --nproc-per-node N tools/run_sharded_check.py: every rank aligns its shard of a B=64 batch on its own
GPU, the compact results are all-gathered over NCCL and compared with the CPU oracle on rank 0."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import torch_tts_b200 as tts
from torch_tts_b200 import synthetic
from oracle import mas_oracle
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S, T = 64, 128, 512
t_x, t_y = synthetic.ragged_lengths(B, S, T, 7)
z_p, m_p, logs_p, x_mask, y_mask = synthetic.prior_inputs(B, S, T, t_x, t_y, seed=7)
lo, hi = tts.shard_bounds(B, rank, world)
sl = slice(lo, hi)
attn, w, g_idx, g_dur = tts.align_sharded(z_p[sl].to(dev), m_p[sl].to(dev), logs_p[sl].to(dev), x_mask[sl].to(dev),
                                          y_mask[sl].to(dev), gather=True, global_batch=B)
torch.cuda.synchronize()
if rank == 0:
    nc = mas_oracle.neg_cent_torch(z_p, m_p, logs_p)
    ref = mas_oracle.maximum_path_c(nc.numpy(), t_y.numpy(), t_x.numpy())
    got = tts.expand_path(g_idx, S).cpu().numpy().astype(np.int32)
    agree = (got == ref).mean()
    assert agree >= 0.9999, agree
    assert np.array_equal(g_dur.sum(1).cpu().numpy(), t_y.numpy())
    print(f"sharded check ok: world={world}, path agreement {agree:.6f}, duration sums identical")
dist.destroy_process_group()
