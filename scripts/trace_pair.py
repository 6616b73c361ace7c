"""Pipeline timeline of the pair backward kernel's first cluster (clock64 stamps): where a tile's time goes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from distillclip_b200 import contrastive as ct

b, d = int(sys.argv[1]), int(sys.argv[2])
cfg = dict(batch=b, dim=d)
gen = torch.Generator(device="cuda").manual_seed(2022)
si, st, ti, tt = bench.make_clip(cfg, "cuda", gen, b, 0)
eng = ct.CudaEngine()
out, saved = ct.contrastive_forward(eng, si, st, ti, tt, 2.0, None)
up = torch.tensor([0.5, 0.5], device="cuda")
for it in range(3):
    eng.trace_pair = torch.zeros(2, 64, 24, dtype=torch.int64, device="cuda")
    ct.contrastive_backward(eng, saved, up, want_img=True, want_txt=False)
    torch.cuda.synchronize()
tr = eng.trace_pair.cpu()
for cta in range(2):
    t = tr[cta]
    base = int(t[0, 8]) if int(t[0, 8]) else int(t[0, 0])
    print(f"--- CTA {cta} (clk relative to first event; MMA: stwait,stissued,gfullwait,gradissued | EPI: start,staged,stfull,loaded,computed,gempty,done)")
    for i in range(2, 14):
        row = [int(x) - base if int(x) else -1 for x in t[i]]
        print(i, "MMA", row[0:4], "EPI", row[8:15])
    if cta == 0:
        d_st = [int(t[i + 1, 0] - t[i, 0]) for i in range(4, 30)]
        print("tile period (MMA st-wait to st-wait):", sum(d_st) / len(d_st))
        print("S/T issue phase:", sum(int(t[i, 1] - t[i, 0]) for i in range(4, 30)) / 26, " of which waiting for operands:",
              sum(int(t[i, 4]) for i in range(4, 30)) / 26, " gfull wait:", sum(int(t[i, 2] - t[i + 1, 1]) for i in range(4, 30)) / 26,
              " grad issue:", sum(int(t[i, 3] - t[i, 2]) for i in range(4, 30)) / 26)
    d_ep = [int(t[i, 14] - t[i, 8]) for i in range(4, 30)]
    print("epilogue per tile:", sum(d_ep) / len(d_ep), " staged:", sum(int(t[i, 9] - t[i, 8]) for i in range(4, 30)) / 26,
          " wait stfull:", sum(int(t[i, 10] - t[i, 9]) for i in range(4, 30)) / 26, " tmem ld:", sum(int(t[i, 11] - t[i, 10]) for i in range(4, 30)) / 26,
          " compute:", sum(int(t[i, 12] - t[i, 11]) for i in range(4, 30)) / 26, " wait gempty:", sum(int(t[i, 13] - t[i, 12]) for i in range(4, 30)) / 26,
          " write+arrive:", sum(int(t[i, 14] - t[i, 13]) for i in range(4, 30)) / 26)
    print("   fine: after_sync fence:", sum(int(t[i, 16] - t[i, 10]) for i in range(4, 30)) / 26, " ld+wait:", sum(int(t[i, 17] - t[i, 16]) for i in range(4, 30)) / 26,
          " before_sync+arrive:", sum(int(t[i, 11] - t[i, 17]) for i in range(4, 30)) / 26, " smem stores:", sum(int(t[i, 18] - t[i, 13]) for i in range(4, 30)) / 26,
          " proxy fence:", sum(int(t[i, 19] - t[i, 18]) for i in range(4, 30)) / 26, " syncwarp+arrive:", sum(int(t[i, 14] - t[i, 19]) for i in range(4, 30)) / 26)
