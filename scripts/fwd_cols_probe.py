"""Ceiling of the forward-epilogue optimisation: the similarity/statistics kernel with and without the column sums
(B = 32768, D = 768), timed alone with a cold L2."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distillclip_b200 import contrastive as ct

cfg = bench.WORKLOADS["sweep"]
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(2022)
si, st, ti, tt = bench.make_clip_global(cfg, dev, gen)
eng = ct._ENGINE
inv = eng.inv_norms([si, st, ti, tt])
b, d = si.shape
flush = bench.L2Flush(dev)
for cols in (True, False):
    fn = lambda: eng.row_stats(si, st, ti, tt, inv[0], inv[1], inv[2], inv[3], 0, 2.0, with_cols=cols)
    fn()
    ms = bench.time_kernel(fn, 10, dev, flush)
    print(f"with_cols={cols}: {ms:.3f} ms  {4.0 * b * b * d / ms / 1e9:.0f} TF (incl. the combine/colreduce launches)")
