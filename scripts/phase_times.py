"""Per-phase device times of the row-sharded contrastive step (CUDA events around every engine call and collective).
torchrun --nproc-per-node N scripts/phase_times.py [B D]   (rank 0 prints; eager, so launch gaps are included in 'step')"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from distillclip_b200 import contrastive as ct

b, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 768)
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
group = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    group = dist.group.WORLD
gen = torch.Generator(device="cuda").manual_seed(2022)
n = b // world
si, st, ti, tt = bench.make_clip(dict(batch=b, dim=d), "cuda", gen, n, rank * n)
records = collections.defaultdict(list)
pending = []


def timed(name, fn):
    def wrapper(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        pending.append((name, e0, e1))
        return out
    return wrapper


eng = ct.CudaEngine()
for m in ("inv_norms", "row_stats", "col_finish", "losses", "coef", "transpose_norm", "row_acc", "col_acc_from_g", "col_acc_scatter", "finish_grads"):
    setattr(eng, m, timed(m, getattr(eng, m)))
ct._all_gather_rows = timed("all_gather_rows", ct._all_gather_rows)
ct._all_gather_cols = timed("all_gather_cols(stats)", ct._all_gather_cols)
ct._reduce_scatter_rows = timed("reduce_scatter(grad)", ct._reduce_scatter_rows)
if world > 1:
    dist.all_reduce = timed("all_reduce(col/sums)", dist.all_reduce)
up = torch.tensor([0.5, 0.5], device="cuda")
steps = []
for it in range(6):
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s0.record()
    out, saved = ct.contrastive_forward(eng, si, st, ti, tt, 2.0, group)
    ct.contrastive_backward(eng, saved, up)
    s1.record()
    torch.cuda.synchronize()
    if it >= 2:
        steps.append(s0.elapsed_time(s1))
        for name, e0, e1 in pending:
            records[name].append(e0.elapsed_time(e1))
    pending.clear()
if rank == 0:
    k = len(steps)
    print(f"world {world}  B={b} D={d}: step {sum(steps) / k:.3f} ms (eager)")
    tot = 0.0
    for name, v in records.items():
        per_step = sum(v) / k
        tot += per_step
        print(f"  {name:26s} {per_step:7.3f} ms  ({len(v) // k} calls)")
    print(f"  {'sum of phases':26s} {tot:7.3f} ms")
if world > 1:
    dist.destroy_process_group()
