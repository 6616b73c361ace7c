"""Full-size parity at the BASELINE.json configurations the benchmark is quoted on ("What's weak" 1):

* configs[4]  B = 32768, D = 768 fused InfoNCE + logit KL against the chunked float64 oracle (oracle/chunked_fp64.py, run
  on the GPU in torch float64 -- test infrastructure): both losses to 1e-4, fp32 gradients of 512 sampled rows per side to 1e-3.
* configs[1] / configs[2]  image and text stage at their real shapes against the reference's op sequence in float64
  (oracle/torch_port.py on CUDA float64 tensors + autograd): values to 1e-4, fp32 kernel gradients to 1e-3, and the bf16
  gradients the module API hands back to the storage-rounding bound.
"""
import numpy as np
import pytest
import torch

from oracle import chunked_fp64 as ck
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu

LOSS_RTOL, GRAD_RTOL, GRAD_BF16_STORAGE_RTOL = 1e-4, 1e-3, 4e-3


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


def _clip_inputs(b, d, seed, device):
    gen = torch.Generator(device=device).manual_seed(seed)
    ti = torch.randn(b, d, device=device, generator=gen)
    tt = ti * 0.6 + 0.8 * torch.randn(b, d, device=device, generator=gen)
    si = ti + 0.5 * torch.randn(b, d, device=device, generator=gen)
    st = tt + 0.5 * torch.randn(b, d, device=device, generator=gen)
    return [x.to(torch.bfloat16).contiguous() for x in (si, st, ti, tt)]


@pytest.mark.parametrize("flow", ["pipeline", "round1_flow"])
@pytest.mark.parametrize("b,d,T", [(32768, 768, 2.0), (4096, 512, 2.0)])
def test_fused_contrastive_full_size_vs_chunked_fp64(cuda_device, b, d, T, flow):
    """BASELINE configs[4] and [3] (reference model/_loss.py:118-153 on the B x B logits, never materialised here), through
    the pipeline (the default: distillclip_b200/pipeline.py) and through the round-1 flow (its fallback)."""
    from distillclip_b200 import contrastive as ct, pipeline as pl
    si, st, ti, tt = _clip_inputs(b, d, 2022, cuda_device)
    w_hard, w_soft = 0.5, 0.5
    gen = torch.Generator().manual_seed(5)
    rows_i = sorted(torch.randperm(b, generator=gen)[:512].tolist())
    rows_t = sorted(torch.randperm(b, generator=gen)[:512].tolist())
    ref = ck.contrastive_chunked(si, st, ti, tt, temperature=T, weights={"hard": w_hard, "soft": w_soft},
                                 sample_img=rows_i, sample_txt=rows_t, chunk=2048)
    if flow == "pipeline":
        out, saved = pl.pipeline_forward(ct._ENGINE, pl.LocalExchange(), si, st, ti, tt, T, (w_hard, w_soft, 1.0, 1.0))
        one = torch.ones((), dtype=torch.float32, device=cuda_device)
        gi, gt = pl.pipeline_backward(ct._ENGINE, saved, (one, None, None), grad_dtype=torch.float32)
        assert float(out[4]) == pytest.approx(w_hard * ref["hard"] + w_soft * ref["soft"], rel=LOSS_RTOL)
    else:
        eng = ct.CudaEngine()
        out, saved = ct.contrastive_forward(eng, si, st, ti, tt, T, None)
        up = torch.tensor([w_hard, w_soft], dtype=torch.float32, device=cuda_device)
        gi, gt = ct.contrastive_backward(eng, saved, up, grad_dtype=torch.float32)
    torch.cuda.synchronize()
    assert float(out[0]) == pytest.approx(ref["hard"], rel=LOSS_RTOL)
    assert float(out[1]) == pytest.approx(ref["soft"], rel=LOSS_RTOL)
    assert _rel(gi[rows_i], ref["d_img"]) <= GRAD_RTOL
    assert _rel(gt[rows_t], ref["d_txt"]) <= GRAD_RTOL
    # per-row bound as well: no sampled row may be off by more than 5x the budget (catches a single bad tile / rank slice)
    per_row = (gi[rows_i].double() - ref["d_img"]).norm(dim=1) / ref["d_img"].norm(dim=1)
    assert float(per_row.max()) <= 5 * GRAD_RTOL
    per_row = (gt[rows_t].double() - ref["d_txt"]).norm(dim=1) / ref["d_txt"].norm(dim=1)
    assert float(per_row.max()) <= 5 * GRAD_RTOL


STAGES = {
    "image": dict(batch=256, tokens=50, heads=12, width=768, layers=4, names=["attention_probs_kl", "hidden_rep_mse"]),
    "text": dict(batch=512, tokens=77, heads=8, width=512, layers=4, names=["attention_probs_kl", "hidden_rep_mse", "embedding_mse"]),
}


def _stage_inputs(cfg, device, seed):
    gen = torch.Generator(device=device).manual_seed(seed)
    b, n, h, w, layers = cfg["batch"], cfg["tokens"], cfg["heads"], cfg["width"], cfg["layers"]
    d = dict(attention_probs=[torch.softmax(torch.randn(b, h, n, n, device=device, generator=gen), -1).to(torch.bfloat16) for _ in range(layers)],
             representations=[torch.randn(b, n, w, device=device, generator=gen).to(torch.bfloat16) for _ in range(layers)])
    if "embedding_mse" in cfg["names"]:
        d["embedding"] = torch.randn(b, n, w, device=device, generator=gen).to(torch.bfloat16)
    return d


@pytest.mark.parametrize("stage", ["image", "text"])
def test_stage_full_size_vs_float64_port(cuda_device, stage):
    """BASELINE configs[1] (image) / configs[2] (text) through LossCalculator (reference model/_loss.py:155-202 as
    DistillModel.training_step calls it, model/distil_model.py:100) against the float64 port of the same op sequence."""
    from distillclip_b200 import ops
    from distillclip_b200.model import LossCalculator, TextTransformerOutput, VisionTransformerOutput
    cfg = STAGES[stage]
    names = cfg["names"]
    stu, tea = _stage_inputs(cfg, cuda_device, 11), _stage_inputs(cfg, cuda_device, 12)
    # ---- float64 reference (values + gradients by autograd through the reference's own op sequence)
    s64 = {k: ([x.double().requires_grad_(True) for x in v] if isinstance(v, list) else v.double().requires_grad_(True)) for k, v in stu.items()}
    t64 = {k: ([x.double() for x in v] if isinstance(v, list) else v.double()) for k, v in tea.items()}
    scale = {n: 1 for n in names}
    percent = {n: 1 / len(names) for n in names}
    ref_loss, ref_res = tp.one_tower(names, scale, percent, None, s64, t64)
    ref_loss.backward()
    ref_grads = {k: ([x.grad for x in v] if isinstance(v, list) else v.grad) for k, v in s64.items()}
    # ---- module API (bf16 gradients)
    cls = VisionTransformerOutput if stage == "image" else TextTransformerOutput
    leaves = {k: ([x.clone().requires_grad_(True) for x in v] if isinstance(v, list) else v.clone().requires_grad_(True)) for k, v in stu.items()}
    calc = LossCalculator(names)
    loss, res = calc(cls(**leaves), cls(**tea), stage)
    loss.backward()
    torch.cuda.synchronize()
    assert float(loss) == pytest.approx(float(ref_loss), rel=LOSS_RTOL)
    for k in names:
        assert float(res[k]) == pytest.approx(float(ref_res[k]), rel=LOSS_RTOL), k
    for k, v in leaves.items():
        for got, want in zip(v if isinstance(v, list) else [v], ref_grads[k] if isinstance(v, list) else [ref_grads[k]]):
            assert _rel(got.grad, want) <= GRAD_BF16_STORAGE_RTOL, k
            # what the reference itself hands a bf16 leaf is its gradient rounded to bf16: the API output matches THAT to 1e-3
            assert _rel(got.grad, want.to(torch.bfloat16)) <= GRAD_RTOL, k
    # ---- the same launch with fp32 gradient output: the north-star tolerance proper
    fields = {"hidden_rep_mse": (ops.KIND_MSE, "representations"), "attention_probs_kl": (ops.KIND_ATTN_KL, "attention_probs"),
              "embedding_mse": (ops.KIND_MSE, "embedding")}
    entries = []
    for nm in names:
        kind, field = fields[nm]
        sv, tv = stu[field], tea[field]
        sv, tv = (sv if isinstance(sv, list) else [sv]), (tv if isinstance(tv, list) else [tv])
        entries.append((kind, len(sv), sv, tv, [True] * len(sv), percent[nm]))
    out, grads, _ = ops.launch_tower(entries, [1.0] * len(names), [percent[n] for n in names], grad_dtype=torch.float32)
    torch.cuda.synchronize()
    assert float(out[-1]) == pytest.approx(float(ref_loss), rel=LOSS_RTOL)
    for nm, gl in zip(names, grads):
        want = ref_grads[fields[nm][1]]
        for got, w in zip(gl, want if isinstance(want, list) else [want]):
            assert _rel(got, w) <= GRAD_RTOL, nm
            assert np.isfinite(float(got.abs().max()))
