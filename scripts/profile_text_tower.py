"""Tower launch for ncu (text stage by default): 2 launches with the default tile (2 position groups per thread for 2-byte
loads), 2 with one group per thread, 2 with aligned 16-byte vectors."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distillclip_b200 import ops
cfg = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "text_stage"]
gen = torch.Generator(device="cuda").manual_seed(2022)
stu, tea = bench.make_tower(cfg, "cuda", gen), bench.make_tower(cfg, "cuda", gen)
entries = [(ops.KIND_ATTN_KL, 4, stu["attention_probs"], tea["attention_probs"], [True] * 4, 1.0),
           (ops.KIND_MSE, 4, stu["representations"], tea["representations"], [True] * 4, 1.0)]
if "embedding" in stu:
    entries.append((ops.KIND_MSE, 1, [stu["embedding"]], [tea["embedding"]], [True], 1.0))
w = [1.0] * len(entries)
for env in ({}, {"DCB_ATTN_GPT": "1"}, {"DCB_ATTN_ALIGNED": "1"}):
    for k in ("DCB_ATTN_NO_ALIGNED", "DCB_ATTN_GPT", "DCB_ATTN_ALIGNED", "DCB_ATTN_STAGED"):
        os.environ.pop(k, None)
    os.environ.update(env)
    bufs = ops.launch_tower(entries, w, w)
    ops.launch_tower(entries, w, w, out=bufs)
    torch.cuda.synchronize()
print("ok")
