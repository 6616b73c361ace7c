"""ONE process, TWO GPUs: the G^T GEMM with its fused reduce-scatter epilogue (clip_gt_gemm_kernel, scatter mode) running on
cuda:0 and storing the rows owned by 'rank 1' straight into a buffer on cuda:1 (peer access) -- so that ncu (which must not
wrap a multi-rank command) can count the NVLink bytes of the kernel:
    ncu --metrics nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,gpu__time_duration.sum -k regex:clip_gt_gemm ...
Algorithmic bytes over the link: (rows owned by the peer) x D x 4 (fp32 partial sums) x k_split."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from distillclip_b200 import contrastive as ct, pipeline as pl

b, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32768, 768)
world = 2
rows = b // world
assert torch.cuda.device_count() >= 2
torch.cuda.set_device(0)
cudart = ctypes.CDLL("libcudart.so.12")
rc = cudart.cudaDeviceEnablePeerAccess(1, 0)
assert rc in (0, 704), rc          # 704 = already enabled
eng = ct._ENGINE
k_split = eng.gt_splits(rows, b, d, scatter=True)
gen = torch.Generator(device="cuda:0").manual_seed(1)
g = eng.alloc_g(rows, b, torch.device("cuda:0"))
g[:, :b] = (torch.randn(rows, b, device="cuda:0", generator=gen) * 64).to(torch.float16)
at = (torch.randn(d, rows, device="cuda:0", generator=gen) / 8).to(torch.float16)
own = torch.zeros(world * k_split, rows, d, dtype=torch.float32, device="cuda:0")
peer = torch.zeros(world * k_split, rows, d, dtype=torch.float32, device="cuda:1")
for _ in range(3):
    eng.col_acc_scatter(g, at, rows, b, d, [own, pl.PeerRef(peer.data_ptr())], 0)
torch.cuda.synchronize()
ref = g[:, :b].double().t() @ at.double().t()                       # [b, d]
got = torch.cat([own[:k_split].double().sum(0), peer[:k_split].to("cuda:0").double().sum(0)])
err = float((got - ref).abs().max() / ref.abs().max())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.col_acc_scatter(g, at, rows, b, d, [own, pl.PeerRef(peer.data_ptr())], 0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
link_bytes = rows * d * 4 * k_split
print(f"rows={rows} cols={b} dim={d} k_split={k_split}: max rel err {err:.2e}; {ms*1e3:.1f} us per launch; "
      f"algorithmic NVLink bytes per launch {link_bytes/1e6:.1f} MB -> {link_bytes/ms/1e6:.0f} GB/s over the link")
assert err < 1e-4
