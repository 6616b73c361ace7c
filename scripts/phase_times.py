"""Per-phase device times of the row-sharded contrastive pipeline (CUDA events around every engine call and exchange step).
torchrun --nproc-per-node N scripts/phase_times.py [B D]   (rank 0 prints; eager, ranks aligned by a barrier before each step,
so launch gaps are included in 'step'; chunked mode runs its tile launches on the main stream here so that they can be timed)"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from distillclip_b200 import contrastive as ct, pipeline as pl

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


eng = ct._ENGINE
xc = pl.exchange_for(group)
for m in ("prep", "fwd_chunk", "post1", "post2", "pair_bwd", "g_tiles", "row_acc_from_g", "col_acc_from_g", "col_acc_scatter", "finish2"):
    setattr(eng, m, timed(m, getattr(eng, m)))
for m in ("start_gather", "exchange_slots", "after_scatter", "wait_all"):
    if hasattr(xc, m):
        setattr(xc, m, timed("xc." + m, getattr(xc, m)))
xc.tile_streams = lambda: None                      # tiles on the main stream: their events then bracket the waits as well
one = torch.ones((), device="cuda")
for it in range(25):
    if group is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out, saved = pl.pipeline_forward(eng, xc, si, st, ti, tt, 2.0, (0.5, 0.5, 1.0, 1.0))
    em = torch.cuda.Event(enable_timing=True)
    em.record()
    pl.pipeline_backward(eng, saved, (one, None, None))
    e1.record()
    torch.cuda.synchronize()
    if it >= 5:
        records["step"].append(e0.elapsed_time(e1))
        records["forward"].append(e0.elapsed_time(em))
        records["backward"].append(em.elapsed_time(e1))
        agg = collections.defaultdict(float)
        for name, a, c in pending:
            agg[name] += a.elapsed_time(c)
        for k, v in agg.items():
            records[k].append(v)
    pending.clear()
if rank == 0:
    print(f"B={b} D={d} world={world} exchange={type(xc).__name__} single_chunk={getattr(saved['set'], 'single_chunk', None)}")
    for k, v in records.items():
        v = sorted(v)
        print(f"  {k:22s} median {v[len(v)//2]*1e3:9.1f} us   min {v[0]*1e3:9.1f} us")
if group is not None:
    dist.barrier()
    dist.destroy_process_group()
