"""Kernel-only timing of the attention-map KL kernel and of the whole tower launch at the image / text stage shapes:
per-thread loads (1 or 2 groups per thread) against aligned 16-byte vectors with in-register realignment."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distillclip_b200 import ops
pk = bench.peaks()
for name in ("image_stage", "text_stage"):
    cfg = bench.WORKLOADS[name]
    gen = torch.Generator(device="cuda").manual_seed(2022)
    stu, tea = bench.make_tower(cfg, "cuda", gen), bench.make_tower(cfg, "cuda", gen)
    els = bench.tower_elements(cfg)
    el = els["attention_probs_kl"]
    s_a, t_a = stu["attention_probs"], tea["attention_probs"]
    for label, env in (("per-thread gpt=1", {"DCB_ATTN_GPT": "1"}), ("per-thread gpt=2", {"DCB_ATTN_GPT": "2"}), ("per-thread gpt=4", {"DCB_ATTN_GPT": "4"}),
                       ("aligned 16 B", {"DCB_ATTN_ALIGNED": "1"}), ("staged W=512", {"DCB_ATTN_STAGED": "1"}),
                       ("staged W=256", {"DCB_ATTN_STAGED": "1", "DCB_ATTN_STAGE_W": "256"})):
        for k in ("DCB_ATTN_NO_ALIGNED", "DCB_ATTN_GPT", "DCB_ATTN_ALIGNED", "DCB_ATTN_STAGED", "DCB_ATTN_STAGE_W"):
            os.environ.pop(k, None)
        os.environ.update(env)
        p_a, _, g_a = ops.launch_attn_kl(s_a, t_a, len(s_a), 1.0, [True] * len(s_a))
        ms = bench.time_kernel(lambda: ops.launch_attn_kl(s_a, t_a, len(s_a), 1.0, [True] * len(s_a), out=(p_a, g_a)), 20, "cuda")
        print(f"{name:12s} attn_kl {label:18s}: {ms*1e3:8.1f} us  {6*el/ms/1e6:7.1f} GB/s  frac {6*el/ms/1e6/pk['hbm']:.3f}", flush=True)
        entries = [(ops.KIND_ATTN_KL, 4, s_a, t_a, [True] * 4, 1.0), (ops.KIND_MSE, 4, stu["representations"], tea["representations"], [True] * 4, 1.0)]
        if "embedding" in stu:
            entries.append((ops.KIND_MSE, 1, [stu["embedding"]], [tea["embedding"]], [True], 1.0))
        w = [1.0] * len(entries)
        bufs = ops.launch_tower(entries, w, w)
        ms = bench.time_kernel(lambda: ops.launch_tower(entries, w, w, out=bufs), 20, "cuda")
        tot = 6 * sum(els.values())
        print(f"{name:12s} tower   {label:18s}: {ms*1e3:8.1f} us  {tot/ms/1e6:7.1f} GB/s  frac {tot/ms/1e6/pk['hbm']:.3f}", flush=True)
