import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distillclip_b200 import ops
b, h, n = 5, 12, int(sys.argv[2])
gen = torch.Generator().manual_seed(1)
s = torch.softmax(torch.randn(b, h, n, n, generator=gen), -1).to(torch.bfloat16).cuda()
t = torch.softmax(torch.randn(b, h, n, n, generator=gen), -1).to(torch.bfloat16).cuda()
mode = sys.argv[1]
need = [mode == "grad"]
entries = [(ops.KIND_ATTN_KL, 1, [s], [t], need, 1.0)]
try:
    res, grads, _ = ops.launch_tower(entries, [1.0], [1.0])
    torch.cuda.synchronize()
    print(mode, "tma:", float(res[0]))
except Exception as e:
    print(mode, "FAILED", str(e)[:200])
    sys.exit(0)
ops.USE_ATTN_TMA = False
res2, grads2, _ = ops.launch_tower(entries, [1.0], [1.0])
torch.cuda.synchronize()
print("ldg:", float(res2[0]))
if need[0]:
    print("grad diff", float((grads[0][0].float() - grads2[0][0].float()).abs().max()), float(grads2[0][0].float().abs().max()))
