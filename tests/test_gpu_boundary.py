"""Boundary robustness under the reference's real settings: fp16 AMP with a GradScaler
(reference config/final_config/image.yaml:69 `precision: 16`, model/distil_model.py:97-112 training_step under Lightning
AMP), mixed fp16/fp32 tower tensors, upstream gradients the forward did not expect, retain_graph, partial
set_scale / set_percent dicts, teacher embeddings of another width, lazy logits with unsupported inputs."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-4


def _stage(b=6, h=4, n=10, w=32, layers=2, seed=3):
    gen = torch.Generator().manual_seed(seed)
    mk_a = lambda: torch.softmax(torch.randn(b, h, n, n, generator=gen), -1)
    mk_h = lambda: torch.randn(b, n, w, generator=gen)
    return dict(attention_probs=[mk_a() for _ in range(layers)], representations=[mk_h() for _ in range(layers)],
                last_representation=torch.randn(b, w, generator=gen))


def _oracle(stu, tea, names):
    vals, grads = {}, {}
    f = lambda x: x.double().numpy()
    if "attention_probs_kl" in names:
        vals["attention_probs_kl"], grads["attention_probs"] = cf.attention_probs_kl([f(x) for x in stu["attention_probs"]], [f(x) for x in tea["attention_probs"]])
    if "hidden_rep_mse" in names:
        vals["hidden_rep_mse"], grads["representations"] = cf.hidden_mse([f(x) for x in stu["representations"]], [f(x) for x in tea["representations"]])
    if "out_l1" in names:
        v, g = cf.out_l1(f(stu["last_representation"]), f(tea["last_representation"]))
        vals["out_l1"], grads["last_representation"] = v, [g]
    return vals, grads


def test_unexpected_upstream_gradient_is_recomputed_exactly(cuda_device):
    """(3 * loss).backward(): the forward assumed 1.0; backward detects the mismatch on the device and recomputes from the
    inputs -- the result is bit-identical to a forward that assumed 3.0 from the start (one rounding, not two)."""
    from distillclip_b200 import ops
    from distillclip_b200.model import HiddenMSE
    d = _stage(b=64, n=50, w=96, layers=3)
    stu = [x.to(torch.bfloat16).cuda().requires_grad_(True) for x in d["representations"]]
    tea = [x.to(torch.bfloat16).cuda() for x in _stage(b=64, n=50, w=96, layers=3, seed=4)["representations"]]
    (3.0 * HiddenMSE()(stu, tea)).backward()
    got = [x.grad.clone() for x in stu]
    old = ops.EXPECTED_GRAD_SCALE
    try:
        ops.EXPECTED_GRAD_SCALE = 3.0
        stu2 = [x.detach().clone().requires_grad_(True) for x in stu]
        (3.0 * HiddenMSE()(stu2, tea)).backward()
    finally:
        ops.EXPECTED_GRAD_SCALE = old
    for a, b in zip(got, stu2):
        assert torch.equal(a, b.grad)
    want = [(3.0 * 2.0 * (s.detach().float() - t.float()) / (s.numel() * 3)).to(torch.bfloat16) for s, t in zip(stu, tea)]
    for a, w in zip(got, want):
        assert rel_l2(a.float().cpu().numpy(), w.float().cpu().numpy()) <= 4e-3


def test_retain_graph_second_backward(cuda_device):
    """A second backward of the same graph recomputes into fresh buffers: .grad accumulates to exactly twice the gradient."""
    from distillclip_b200.model import LossCalculator, VisionTransformerOutput
    d, t = _stage(), _stage(seed=9)
    names = ["attention_probs_kl", "hidden_rep_mse", "out_l1"]
    mk = lambda src, grad: VisionTransformerOutput(
        attention_probs=[x.to(torch.bfloat16).cuda().requires_grad_(grad) for x in src["attention_probs"]],
        representations=[x.to(torch.bfloat16).cuda().requires_grad_(grad) for x in src["representations"]],
        last_representation=src["last_representation"].to(torch.bfloat16).cuda().requires_grad_(grad))
    stu, tea = mk(d, True), mk(t, False)
    loss, res = LossCalculator(names)(stu, tea, "image")
    loss.backward(retain_graph=True)
    first = [x.grad.clone() for x in [*stu.attention_probs, *stu.representations, stu.last_representation]]
    loss.backward(retain_graph=True)
    res["hidden_rep_mse"].backward()                       # and a backward from a dict entry alone
    leaves = [*stu.attention_probs, *stu.representations, stu.last_representation]
    for i, (g1, x) in enumerate(zip(first, leaves)):
        extra = 3.0 if 2 <= i < 4 else 0.0                 # d(res)/d(rep) = grad / percent (1/3): three more units
        want = g1.float() * (2.0 + extra)
        assert rel_l2(x.grad.float().cpu().numpy(), want.cpu().numpy()) <= 1.2e-2       # sums of bf16-rounded terms


@pytest.mark.parametrize("use_scaler_hint", [False, True])
def test_fp16_autocast_with_grad_scaler_and_mixed_dtypes(cuda_device, use_scaler_hint):
    """LossCalculator inside torch.autocast(fp16) with a GradScaler at 65536: attention maps arrive in fp32 (autocast runs
    softmax in fp32), hidden states in fp16.  hidden-MSE gradients are ~1e-6 before scaling -- below fp16's normal range:
    they must come out right after unscaling whether or not the scaler was handed to LossCalculator."""
    from distillclip_b200.model import LossCalculator, VisionTransformerOutput
    d, t = _stage(b=16, n=50, w=96), _stage(b=16, n=50, w=96, seed=9)
    names = ["attention_probs_kl", "hidden_rep_mse", "out_l1"]
    rep16 = [x.to(torch.float16) for x in d["representations"]]
    last16 = d["last_representation"].to(torch.float16)
    stu = VisionTransformerOutput(attention_probs=[x.cuda().requires_grad_(True) for x in d["attention_probs"]],           # fp32
                                  representations=[x.cuda().requires_grad_(True) for x in rep16],                          # fp16
                                  last_representation=last16.cuda().requires_grad_(True))
    tea = VisionTransformerOutput(attention_probs=[x.cuda() for x in t["attention_probs"]],
                                  representations=[x.to(torch.float16).cuda() for x in t["representations"]],
                                  last_representation=t["last_representation"].to(torch.float16).cuda())
    calc = LossCalculator(names)
    scaler = torch.amp.GradScaler("cuda", init_scale=65536.0)
    if use_scaler_hint:
        calc.grad_scaler = scaler
    with torch.autocast("cuda", dtype=torch.float16):
        loss, res = calc(stu, tea, "image")
    scaler.scale(loss).backward()
    torch.cuda.synchronize()
    src = dict(attention_probs=d["attention_probs"], representations=[x.float() for x in rep16], last_representation=last16.float())
    tsrc = dict(attention_probs=t["attention_probs"], representations=[x.to(torch.float16).float() for x in t["representations"]],
                last_representation=t["last_representation"].to(torch.float16).float())
    vals, grads = _oracle(src, tsrc, names)
    assert float(loss.detach()) == pytest.approx(sum(vals.values()) / 3, rel=LOSS_RTOL)
    for key, leaves in (("attention_probs", stu.attention_probs), ("representations", stu.representations),
                        ("last_representation", [stu.last_representation])):
        for x, g in zip(leaves, grads[key]):
            assert x.grad.dtype == x.dtype
            got = x.grad.double().cpu().numpy() / 65536.0
            assert np.isfinite(got).all()
            tol = 1e-3 if x.dtype == torch.float32 else 2e-3          # fp16 storage: 11 bits at the SCALED magnitude
            assert rel_l2(got, g / 3) <= tol, key


def test_partial_scale_dict_after_set_scale(cuda_device):
    """set_scale with a dict that misses a name: the reference's loop over loss_scale leaves that term out of the total and
    unscaled in the dict (model/_loss.py:195-200); tower-only kinds (out_l1) must work on this path too."""
    from distillclip_b200.model import LossCalculator, VisionTransformerOutput
    d, t = _stage(), _stage(seed=9)
    names = ["hidden_rep_mse", "out_l1"]
    mk = lambda src, grad: VisionTransformerOutput(
        representations=[x.to(torch.bfloat16).cuda().requires_grad_(grad) for x in src["representations"]],
        last_representation=src["last_representation"].to(torch.bfloat16).cuda().requires_grad_(grad))
    stu, tea = mk(d, True), mk(t, False)
    calc = LossCalculator(names)
    calc.set_scale({"out_l1": 2.0})
    loss, res = calc(stu, tea, "image")
    loss.backward()
    f = lambda x: x.to(torch.bfloat16).double().numpy()
    l1, _ = cf.out_l1(f(d["last_representation"]), f(t["last_representation"]))
    hm, _ = cf.hidden_mse([f(x) for x in d["representations"]], [f(x) for x in t["representations"]])
    assert float(loss.detach()) == pytest.approx(0.5 * 2.0 * l1, rel=LOSS_RTOL)
    assert float(res["out_l1"]) == pytest.approx(2.0 * l1, rel=LOSS_RTOL) and float(res["hidden_rep_mse"]) == pytest.approx(hm, rel=LOSS_RTOL)
    assert all(x.grad is None for x in stu.representations) and stu.last_representation.grad is not None


def test_contrastive_fallbacks_and_lazy_logit_errors(cuda_device):
    from distillclip_b200._lib import DistillClipB200Error
    from distillclip_b200.model import CLIPOutput, LossCalculator, TextTransformerOutput, VisionTransformerOutput
    gen = torch.Generator().manual_seed(1)
    b, d = 64, 32
    mk = lambda dim, dt: torch.randn(b, dim, generator=gen).to(dt).cuda()

    def out(img, txt, logits):
        o = CLIPOutput(visual_output=VisionTransformerOutput(last_representation=img), text_output=TextTransformerOutput(last_representation=txt))
        if logits:
            a, c = img.float(), txt.float()
            lg = (a / a.norm(dim=1, keepdim=True)) @ (c / c.norm(dim=1, keepdim=True)).t()
            o.i2t_logits, o.t2i_logits = lg, lg.T
        return o
    si, st = mk(d, torch.bfloat16).requires_grad_(True), mk(d, torch.bfloat16).requires_grad_(True)
    ti, tt = mk(2 * d, torch.bfloat16), mk(2 * d, torch.bfloat16)          # teacher twice as wide: legal in the reference
    calc = LossCalculator(["hard_label", "soft_label"], temperature=2.0)
    loss, res = calc(out(si, st, True), out(ti, tt, True), "all")          # falls back to the logits modules
    loss.backward()
    s_ref, _ = cf.clip_logits(si.detach().float().cpu().numpy(), st.detach().float().cpu().numpy())
    t_ref, _ = cf.clip_logits(ti.float().cpu().numpy(), tt.float().cpu().numpy())
    want = 0.5 * (cf.soft_label(s_ref, t_ref, 2.0)[0] + cf.soft_label(s_ref.T, t_ref.T, 2.0)[0])
    assert float(res["soft_label"]) == pytest.approx(want, rel=LOSS_RTOL)
    with pytest.raises(DistillClipB200Error, match="carries no logits"):
        calc(out(si, st, False), out(ti, tt, False), "all")
    s32 = mk(d, torch.float32)
    with pytest.raises(DistillClipB200Error, match="carries no logits"):   # fp32 embeddings + lazy logits
        LossCalculator(["hard_label"])(out(s32, s32.clone(), False), out(s32, s32, False), "all")
