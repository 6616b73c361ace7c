"""Is operand supply of the contrastive kernels limited per SM (latency x bytes in flight) or chip-wide (L2 output)?
Runs the forward / pair-backward kernels with 37, 74 and 148 row blocks against the same 32768 columns: if the time per
tile stays put with fewer CTAs the limit is per SM, if it drops it is the shared L2."""
import os
import sys

os.environ["DCB_DEBUG_SPLITS"] = "1"       # one CTA (pair) per row block: the number of busy SMs = the number of row blocks

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from distillclip_b200 import contrastive as ct

cols, d = 32768, int(sys.argv[1]) if len(sys.argv) > 1 else 768
gen = torch.Generator(device="cuda").manual_seed(2022)
si, st, ti, tt = bench.make_clip(dict(batch=cols, dim=d), "cuda", gen, cols, 0)
eng = ct.CudaEngine()
inv = eng.inv_norms([si, st, ti, tt])
up = torch.tensor([0.5, 0.5], device="cuda")


def timed(fn, n=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for blocks in (18, 37, 74, 148):
    rows = blocks * 128
    a_s, a_t = si[:rows], ti[:rows]
    fwd = timed(lambda: eng.row_stats(a_s, st, a_t, tt, inv[0][:rows], inv[1], inv[2][:rows], inv[3], 0, 2.0, with_cols=True))
    stats, _, col = eng.row_stats(a_s, st, a_t, tt, inv[0][:rows], inv[1], inv[2][:rows], inv[3], 0, 2.0, with_cols=True)
    coef_r, gmax_r = eng.coef(stats, cols, 2.0, True, up)
    coef_c, gmax_c = eng.coef(col, cols, 2.0, True, up)
    bt = eng.transpose_norm(st, inv[1])
    bwd = timed(lambda: eng.row_acc(a_s, st, a_t, tt, bt, inv[0][:rows], inv[1], inv[2][:rows], inv[3], coef_r, coef_c,
                                    gmax_r, gmax_c, 2.0))
    tiles = blocks * (cols // 128)
    print(f"row blocks {blocks:4d}: fwd {fwd:7.3f} ms ({fwd * 1e6 / tiles:7.1f} ns per 128x128 tile)   "
          f"pair bwd {bwd:7.3f} ms ({bwd * 1e6 / tiles:7.1f} ns per tile)")
