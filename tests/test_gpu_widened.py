"""SURVEY.md section 8f widening on the GPU: the losses of the three shipped recipes (out_l1, out_cos, cos_diff) and the
attention-map / attention-score MSE siblings, against the reference-generated golden fixtures and the oracle.
Tolerances as in test_gpu_streaming.py (loss 1e-4, fp32 gradients 1e-3, bf16-stored gradients 4e-3)."""
import numpy as np
import pytest
import torch

from conftest import golden, numbered, rel_l2
from oracle import closed_form as cf

pytestmark = pytest.mark.gpu
LOSS_RTOL, GRAD_RTOL, GRAD_BF16_STORAGE_RTOL = 1e-4, 1e-3, 4e-3


def dev(x, dtype=torch.bfloat16, grad=False):
    return torch.tensor(np.asarray(x), device="cuda").to(dtype).requires_grad_(grad)


@pytest.mark.parametrize("name,cls", [("out_l1", "OutL1Loss"), ("out_l1_with_ties", "OutL1Loss"), ("out_cos", "OutCosLoss")])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_out_losses_golden(cuda_device, name, cls, dtype):
    import distillclip_b200.model as m
    g = golden(name)
    s, t = dev(g["stu0"], dtype, True), dev(g["tea0"], dtype)
    loss = getattr(m, cls)()(s, t)
    loss.backward()
    assert float(loss.detach()) == pytest.approx(float(g["loss_f64"]), rel=LOSS_RTOL)
    tol = GRAD_RTOL if dtype == torch.float32 else GRAD_BF16_STORAGE_RTOL
    assert rel_l2(s.grad.float().cpu().numpy(), g["grad0_f64"]) <= tol
    if name == "out_l1_with_ties":      # sign(0) = 0 exactly where student == teacher
        ties = g["stu0"] == g["tea0"]
        assert ties.any() and float(s.grad.float().cpu().numpy()[ties].__abs__().max()) == 0.0


@pytest.mark.parametrize("name,cls", [("attn_probs_mse", "AttentionProbsMSE"), ("attn_score_mse", "AttentionScoreMSE")])
def test_attention_mean_mse_golden(cuda_device, name, cls):
    import distillclip_b200.model as m
    g = golden(name)
    stu, tea = [dev(x, torch.float32, True) for x in numbered(g, "stu")], [dev(x, torch.float32) for x in numbered(g, "tea")]
    loss = getattr(m, cls)()(stu, tea)
    loss.backward()
    assert float(loss.detach()) == pytest.approx(float(g["loss_f64"]), rel=LOSS_RTOL)
    for i, s in enumerate(stu):
        ref = g[f"grad{i}_f64"]
        if np.all(ref == 0):
            assert s.grad is None or float(s.grad.abs().max()) == 0.0
        else:
            assert rel_l2(s.grad.cpu().numpy(), ref) <= GRAD_RTOL


@pytest.mark.parametrize("b,h,n", [(4, 12, 50), (3, 8, 77)])
def test_attention_mean_mse_random_bf16(cuda_device, b, h, n):
    from distillclip_b200.model import AttentionProbsMSE
    gen = torch.Generator().manual_seed(3)
    stu = [torch.softmax(torch.randn(b, h, n, n, generator=gen), -1).to(torch.bfloat16) for _ in range(2)]
    tea = [torch.softmax(torch.randn(b, h, n, n, generator=gen), -1).to(torch.bfloat16) for _ in range(2)]
    ref_loss, ref_grads = cf.attention_mean_mse([s.float().numpy() for s in stu], [t.float().numpy() for t in tea])
    ds = [s.cuda().requires_grad_(True) for s in stu]
    loss = AttentionProbsMSE()(ds, [t.cuda() for t in tea])
    loss.backward()
    assert float(loss.detach()) == pytest.approx(ref_loss, rel=LOSS_RTOL)
    for s, r in zip(ds, ref_grads):
        assert rel_l2(s.grad.float().cpu().numpy(), r) <= GRAD_BF16_STORAGE_RTOL


@pytest.mark.parametrize("name", ["cos_diff_n17", "cos_diff_n64"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cos_diff_golden(cuda_device, name, dtype):
    """CLIPCosDiff on materialised logits and on the `.T` view; the off-diagonal index construction (get_neg_element)
    is exact: gradients are compared entry by entry."""
    from distillclip_b200.model import CLIPCosDiff
    g = golden(name)
    s, t = dev(g["stu"], dtype, True), dev(g["tea"], dtype)      # fixture values are bf16-exact
    loss = CLIPCosDiff()(s, t)
    loss.backward()
    assert float(loss.detach()) == pytest.approx(float(g["loss_f64"]), rel=LOSS_RTOL)
    assert np.allclose(s.grad.float().cpu().numpy(), g["grad_f64"], rtol=GRAD_BF16_STORAGE_RTOL, atol=0)
    s.grad = None
    loss = CLIPCosDiff()(s.T, t.T)
    loss.backward()
    assert float(loss.detach()) == pytest.approx(float(g["loss_T_f64"]), rel=LOSS_RTOL)
    assert np.allclose(s.grad.float().cpu().numpy(), g["grad_T_f64"], rtol=GRAD_BF16_STORAGE_RTOL, atol=0)


def _tower(g, prefix, cls, grad):
    kw = {}
    for f in ("last_representation", "embedding"):
        kw[f] = dev(g[f"{prefix}.{f}"], grad=grad)
    for f in ("attention_probs", "representations"):
        kw[f] = [dev(x, grad=grad) for x in numbered(g, f"{prefix}.{f}.")]
    return cls(**kw)


def _leaves(t):
    return [t.last_representation, *t.attention_probs, *t.representations, t.embedding]


def _check_calc(g, loss, res, leaves):
    assert float(loss.detach()) == pytest.approx(float(g["loss_f64"]), rel=LOSS_RTOL)
    for k, v in res.items():
        assert float(v.detach()) == pytest.approx(float(g[f"res.{k}_f64"]), rel=LOSS_RTOL), k
    assert set(res) == {k[4:-4] for k in g if k.startswith("res.") and k.endswith("_f64")}
    for i, leaf in enumerate(leaves):
        key = f"grad{i}_f64"
        if key in g:
            assert rel_l2(leaf.grad.float().cpu().numpy(), g[key]) <= GRAD_BF16_STORAGE_RTOL, key
        else:
            assert leaf.grad is None


@pytest.mark.parametrize("name,kwargs", [
    ("calc_shipped_image", dict(loss_name=["out_l1", "out_cos"])),
    ("calc_attn_mse_mix", dict(loss_name=["attention_probs_mse", "attention_probs_kl", "hidden_rep_mse", "out_l1"],
                               loss_scale={"attention_probs_mse": 3.0}))])
def test_shipped_one_tower_recipes(cuda_device, name, kwargs):
    """config/final_config/image.yaml and text.yaml ship loss_name = ['out_l1', 'out_cos']."""
    from distillclip_b200.model import LossCalculator, VisionTransformerOutput
    g = golden(name)
    stu, tea = _tower(g, "stu", VisionTransformerOutput, True), _tower(g, "tea", VisionTransformerOutput, False)
    loss, res = LossCalculator(**kwargs)(stu, tea, "image")
    loss.backward()
    _check_calc(g, loss, res, _leaves(stu))


@pytest.mark.parametrize("route", ["lazy_logits", "fused_with_logits_present", "logits_modules"])
def test_shipped_lclip_recipe(cuda_device, route):
    """config/final_config/l_clip.yaml ships ['out_l1', 'out_cos', 'cos_diff'].  lazy_logits: the CLIPOutput carries NO logits
    (LazyLogitsCLIP) and cos_diff comes from the embeddings inside the fused tcgen05 kernels (no B x B matrix
    anywhere); logits_modules: CLIPCosDiff on the caller's logits (CLIPModel.forward, reference clip_model.py:36-44)."""
    from distillclip_b200.model import (CLIPOutput, LossCalculator, TextTransformerOutput, VisionTransformerOutput)
    g = golden("calc_shipped_lclip")
    sv, stx = _tower(g, "stu.visual", VisionTransformerOutput, True), _tower(g, "stu.text", TextTransformerOutput, True)
    tv, ttx = _tower(g, "tea.visual", VisionTransformerOutput, False), _tower(g, "tea.text", TextTransformerOutput, False)

    def clip_out(v, x):
        if route == "lazy_logits":
            return CLIPOutput(visual_output=v, text_output=x)
        a, b = v.last_representation.float(), x.last_representation.float()
        lg = (a / a.norm(dim=1, keepdim=True)) @ (b / b.norm(dim=1, keepdim=True)).t()
        return CLIPOutput(visual_output=v, text_output=x, i2t_logits=lg, t2i_logits=lg.T)
    calc = LossCalculator(["out_l1", "out_cos", "cos_diff"])
    calc.fused_contrastive = route != "logits_modules"
    loss, res = calc(clip_out(sv, stx), clip_out(tv, ttx), "all")
    loss.backward()
    _check_calc(g, loss, res, _leaves(sv) + _leaves(stx))


# ---- second batch: out_kl / out_ce / logits_mse ---------------------------------------------------------------------
@pytest.mark.parametrize("name,make", [("out_kl_t2", lambda m: m.OutKLLoss(2.0)), ("out_kl_t05_wide", lambda m: m.OutKLLoss(0.5)),
                                       ("out_ce", lambda m: m.OutCELoss()), ("out_ce_wide", lambda m: m.OutCELoss()),
                                       ("logits_mse", lambda m: m.LogitsMSE())])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_row_softmax_and_logits_mse_golden(cuda_device, name, make, dtype):
    import distillclip_b200.model as m
    g = golden(name)
    s, t = dev(g["stu0"], dtype, True), dev(g["tea0"], dtype)
    loss = make(m)(s, t)
    (3.0 * loss).backward()              # a non-unit upstream gradient
    assert float(loss.detach()) == pytest.approx(float(g["loss_f64"]), rel=LOSS_RTOL)
    tol = GRAD_RTOL if dtype == torch.float32 else GRAD_BF16_STORAGE_RTOL
    assert rel_l2(s.grad.float().cpu().numpy() / 3.0, g["grad0_f64"]) <= tol


@pytest.mark.parametrize("mode,temperature", [(0, 1.0), (0, 4.0), (1, None)])
def test_row_softmax_random_bf16_vs_oracle(cuda_device, mode, temperature):
    """CLIP-sized pooled outputs (B=256, D=512) in bf16 against the f64 closed form on the same bf16-rounded inputs."""
    from distillclip_b200 import ops
    gen = torch.Generator().manual_seed(5)
    s = (torch.randn(256, 512, generator=gen) * 1.5).to(torch.bfloat16)
    t = (torch.randn(256, 512, generator=gen) * 1.5).to(torch.bfloat16)
    sd = s.cuda().requires_grad_(True)
    loss = ops.row_softmax_loss(sd, t.cuda(), temperature, mode)
    loss.backward()
    ref, gref = cf.out_kl(s.float().numpy(), t.float().numpy(), temperature) if mode == 0 else \
        cf.out_ce(s.float().numpy(), t.float().numpy())
    assert float(loss.detach()) == pytest.approx(float(ref), rel=LOSS_RTOL)
    assert rel_l2(sd.grad.float().cpu().numpy(), gref) <= GRAD_BF16_STORAGE_RTOL


def test_out_kl_ce_one_tower_recipe(cuda_device):
    from distillclip_b200.model import LossCalculator, VisionTransformerOutput
    g = golden("calc_out_kl_ce_image")
    stu, tea = _tower(g, "stu", VisionTransformerOutput, True), _tower(g, "tea", VisionTransformerOutput, False)
    loss, res = LossCalculator(["out_ce", "out_kl", "hidden_rep_mse"], temperature=4.0)(stu, tea, "image")
    assert list(res) == ["out_ce", "out_kl", "hidden_rep_mse"]
    loss.backward()
    _check_calc(g, loss, res, _leaves(stu))


@pytest.mark.parametrize("fused", [True, False, "lazy"])
def test_out_kl_ce_logits_mse_two_tower_recipe(cuda_device, fused):
    from distillclip_b200.model import (CLIPOutput, LossCalculator, TextTransformerOutput, VisionTransformerOutput)
    g = golden("calc_out_kl_ce_logits_mse")
    sv, stx = _tower(g, "stu.visual", VisionTransformerOutput, True), _tower(g, "stu.text", TextTransformerOutput, True)
    tv, ttx = _tower(g, "tea.visual", VisionTransformerOutput, False), _tower(g, "tea.text", TextTransformerOutput, False)

    def clip_out(v, x):
        if fused == "lazy":                  # no logits at all: logits_mse and hard_label both from the embeddings
            return CLIPOutput(visual_output=v, text_output=x)
        a, b = v.last_representation.float(), x.last_representation.float()
        lg = (a / a.norm(dim=1, keepdim=True)) @ (b / b.norm(dim=1, keepdim=True)).t()
        return CLIPOutput(visual_output=v, text_output=x, i2t_logits=lg, t2i_logits=lg.T)
    calc = LossCalculator(["out_kl", "out_ce", "logits_mse", "hard_label"], temperature=2.0, loss_scale={"out_kl": 0.1})
    calc.fused_contrastive = bool(fused)
    loss, res = calc(clip_out(sv, stx), clip_out(tv, ttx), "all")
    loss.backward()
    _check_calc(g, loss, res, _leaves(sv) + _leaves(stx))


# ---- third batch: LastValueMapKL -----------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["value_map_kl_h3_n7", "value_map_kl_h12_n10", "value_map_kl_scores"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_last_value_map_kl_golden(cuda_device, name, dtype):
    import distillclip_b200.model as m
    g = golden(name)
    s, t = dev(g["stu0"], dtype, True), dev(g["tea0"], dtype)
    loss = m.LastValueMapKL()(s, t)
    (2.0 * loss).backward()
    assert float(loss.detach()) == pytest.approx(float(g["loss_f64"]), rel=LOSS_RTOL)
    tol = GRAD_RTOL if dtype == torch.float32 else GRAD_BF16_STORAGE_RTOL
    assert rel_l2(s.grad.float().cpu().numpy() / 2.0, g["grad0_f64"]) <= tol


def test_last_value_map_kl_full_size(cuda_device):
    """Image-tower size (B=64, H=12, N=50; odd and even vector paths) against the f64 closed form, plus KL(t, t) = 0."""
    import distillclip_b200.model as m
    gen = torch.Generator().manual_seed(3)
    for n in (50, 7):
        s = torch.softmax(torch.randn(64, 12, n, n, generator=gen), -1).to(torch.bfloat16)
        t = torch.softmax(torch.randn(64, 12, n, n, generator=gen), -1).to(torch.bfloat16)
        sd = s.cuda().requires_grad_(True)
        loss = m.LastValueMapKL()(sd, t.cuda())
        loss.backward()
        ref, gref = cf.last_value_map_kl(s.float().numpy(), t.float().numpy())
        assert float(loss.detach()) == pytest.approx(float(ref), rel=LOSS_RTOL)
        assert rel_l2(sd.grad.float().cpu().numpy(), gref) <= GRAD_BF16_STORAGE_RTOL
        td = t.cuda().requires_grad_(True)
        zero = m.LastValueMapKL()(td, t.cuda())
        zero.backward()
        assert abs(float(zero.detach())) <= 1e-6 * float(ref) and float(td.grad.abs().max()) == 0.0


def test_last_value_map_kl_in_loss_calculator(cuda_device):
    from distillclip_b200.model import LossCalculator, VisionTransformerOutput
    gen = torch.Generator().manual_seed(4)
    mk = lambda *sh: torch.randn(*sh, generator=gen).to(torch.bfloat16)
    stu = VisionTransformerOutput(last_representation=mk(6, 32).cuda().requires_grad_(True), value_map=mk(6, 4, 9, 9).cuda().requires_grad_(True))
    tea = VisionTransformerOutput(last_representation=mk(6, 32).cuda(), value_map=mk(6, 4, 9, 9).cuda())
    calc = LossCalculator(["out_l1", "last_value_map_kl"], loss_scale={"last_value_map_kl": 0.5})
    assert calc.get_control_output().need_value_map
    loss, res = calc(stu, tea, "image")
    loss.backward()
    kl, gkl = cf.last_value_map_kl(stu.value_map.detach().float().cpu().numpy(), tea.value_map.float().cpu().numpy())
    l1, _ = cf.out_l1(stu.last_representation.detach().float().cpu().numpy(), tea.last_representation.float().cpu().numpy())
    assert float(res["last_value_map_kl"].detach()) == pytest.approx(0.5 * kl, rel=LOSS_RTOL)
    assert float(loss.detach()) == pytest.approx(0.5 * (0.5 * kl) + 0.5 * l1, rel=LOSS_RTOL)
    assert rel_l2(stu.value_map.grad.float().cpu().numpy(), 0.25 * gkl) <= GRAD_BF16_STORAGE_RTOL


# ---- validation metrics from the embeddings ---------------------------------------------------------
@pytest.mark.parametrize("name", ["retrieval_b200_d64", "retrieval_b77_d40"])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_retrieval_metrics_golden(cuda_device, name, dtype):
    """norm_and_logits + top-k accuracy + diagonal scores without the B x B logits; fp32 embeddings go through the
    bf16 hi/lo split.  Accuracies are counts / B: exact unless two logits of a row tie within rounding."""
    from distillclip_b200.metrics import retrieval_metrics
    g = golden(name)
    b = g["img"].shape[0]
    res = retrieval_metrics(dev(g["img"], dtype), dev(g["txt"], dtype), prefix="stu_")
    assert set(res) == {f"stu_acc_top{k}" for k in (1, 3, 5, 10, 20, 50)} | {"stu_softmax_mean_score", "stu_mean_score"}
    for k in (1, 3, 5, 10, 20, 50):
        assert abs(float(res[f"stu_acc_top{k}"]) - float(g[f"acc_top{k}_f64"])) <= 1.0 / b + 1e-6, k
    assert float(res["stu_softmax_mean_score"]) == pytest.approx(float(g["softmax_mean_score_f64"]), rel=LOSS_RTOL)
    assert float(res["stu_mean_score"]) == pytest.approx(float(g["mean_score_f64"]), rel=LOSS_RTOL)


def test_retrieval_ranks_exact_and_fp32_split(cuda_device):
    """Per-row ranks against the f64 oracle at B=1000, D=512: bf16 inputs are exact products on the tensor cores, so ranks
    may differ only where two logits are within fp32 accumulation error; the fp32 path must resolve differences bf16
    rounding of the inputs would hide (embeddings that differ only below bf16 precision)."""
    from distillclip_b200 import contrastive as ct
    from distillclip_b200.metrics import retrieval_metrics
    gen = torch.Generator().manual_seed(8)
    img = torch.randn(1000, 512, generator=gen).to(torch.bfloat16)
    txt = (img.float() + 4.0 * torch.randn(1000, 512, generator=gen)).to(torch.bfloat16)
    eng = ct.CudaEngine()
    a, b = img.cuda(), txt.cuda()
    inv = eng.inv_norms([a, b])
    stats, _ = eng.row_stats(a, b, None, None, inv[0], inv[1], None, None, 0, None)
    ranks = eng.rank_counts(a, b, inv[0], inv[1], 0, stats[4].contiguous()).cpu().numpy()
    s_ref, _ = cf.clip_logits(img.float().numpy(), txt.float().numpy())
    ref = (s_ref > np.diag(s_ref)[:, None]).sum(1)
    assert (ranks != ref).mean() <= 0.01 and np.abs(ranks - ref).max() <= 2
    # fp32: perturb below bf16 resolution; the oracle on the fp32 values is the reference
    img32 = img.float() * (1 + 1e-3 * torch.randn(1000, 512, generator=gen))
    txt32 = txt.float() * (1 + 1e-3 * torch.randn(1000, 512, generator=gen))
    res = retrieval_metrics(img32.cuda(), txt32.cuda())
    want = cf.retrieval_metrics(img32.numpy(), txt32.numpy())
    for k in (1, 5, 50):
        assert abs(float(res[f"acc_top{k}"]) - want[f"acc_top{k}"]) <= 2e-3, k
    assert float(res["mean_score"]) == pytest.approx(want["mean_score"], rel=2e-5)
    assert float(res["softmax_mean_score"]) == pytest.approx(want["softmax_mean_score"], rel=2e-5)
