"""Kernel-only timing of the tower launch (image / text stage shapes) with and without the TMA attention kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distillclip_b200 import ops
for name in ("image_stage", "text_stage"):
    cfg = bench.WORKLOADS[name]
    gen = torch.Generator(device="cuda").manual_seed(2022)
    stu, tea = bench.make_tower(cfg, "cuda", gen), bench.make_tower(cfg, "cuda", gen)
    el = bench.tower_elements(cfg)
    for only_attn in (True, False):
        for tma in (True, False):
            ops.USE_ATTN_TMA = tma
            entries = [(ops.KIND_ATTN_KL, 4, stu["attention_probs"], tea["attention_probs"], [True] * 4, 1.0)]
            nbytes = 6 * el["attention_probs_kl"]
            if not only_attn:
                entries.append((ops.KIND_MSE, 4, stu["representations"], tea["representations"], [True] * 4, 1.0))
                nbytes += 6 * el["hidden_rep_mse"]
            w = [1.0] * len(entries)
            bufs = ops.launch_tower(entries, w, w)
            ms = bench.time_kernel(lambda: ops.launch_tower(entries, w, w, out=bufs), 20, "cuda")
            print(f"{name:12s} attn_only={only_attn!s:5s} tma={tma!s:5s}: {ms*1e3:8.1f} us  {nbytes/ms/1e6:7.1f} GB/s")
