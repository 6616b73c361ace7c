"""Generate the golden fixtures in this directory by running the REFERENCE's own modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own; these fixtures are the
pins for both the CPU oracle (`oracle/`) and the CUDA path.  Inputs are drawn in bf16 (seeded),
stored as bf16-exact float32, and pushed through the reference upcast to float32 and to float64
(SURVEY.md F11).  Outputs: loss values and student gradients from reference autograd.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

REF = os.environ.get("DISTILLCLIP_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
with contextlib.redirect_stdout(io.StringIO()):
    from model._loss import LossCalculator                      # noqa: E402
    from model.loss_component import (AttentionProbsKL, AttentionProbsMSE, AttentionScoreMSE, CLIPCosDiff,  # noqa: E402
                                      EmbedMSELoss, HardLabel, HiddenMSE, LastValueMapKL, LogitsMSE, OutCELoss, OutCosLoss, OutKLLoss,
                                      OutL1Loss, SoftLabel)
    from model.component.clip_model import CLIPModel             # noqa: E402
    from model.component.output import (CLIPOutput, ControlOutput, TextTransformerOutput,  # noqa: E402
                                        VisionTransformerOutput)

HERE = os.path.dirname(os.path.abspath(__file__))
G = torch.Generator().manual_seed(2022)


def bf16(*shape, scale=1.0):
    return (torch.randn(*shape, generator=G) * scale).to(torch.bfloat16)


def probs(b, h, n, causal=False):
    x = torch.randn(b, h, n, n, generator=G)
    if causal:
        x = x + torch.full((n, n), float("-inf")).triu_(1)
    return torch.softmax(x, dim=-1).to(torch.bfloat16)


def run(fn, stu, tea, dtype):
    """stu/tea: lists of bf16 tensors. Returns loss (python float) and grads wrt stu (dtype)."""
    s = [x.to(dtype).requires_grad_(True) for x in stu]
    t = [x.to(dtype) for x in tea]
    loss = fn(s, t)
    loss.backward()
    return loss.detach().numpy(), [(x.grad if x.grad is not None else torch.zeros_like(x)).numpy() for x in s]


def save(name, **arrs):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrs)
    print("wrote", name, {k: np.shape(v) for k, v in arrs.items()})


def pack_io(prefix, tensors):
    return {f"{prefix}{i}": t.float().numpy() for i, t in enumerate(tensors)}


# ---- a4 attention-map KL -------------------------------------------------------------------
def attn_case(name, stu, tea):
    mod = AttentionProbsKL()
    out = {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        loss, grads = run(lambda s, t: mod(s, t), stu, tea, dt)
        out[f"loss_{tag}"] = loss
        for i, g in enumerate(grads):
            out[f"grad{i}_{tag}"] = g
    save(name, **pack_io("stu", stu), **pack_io("tea", tea), **out)


attn_case("attn_kl_heads4v2", [probs(3, 4, 7) for _ in range(2)], [probs(3, 2, 7) for _ in range(2)])
attn_case("attn_kl_even", [probs(2, 3, 6) for _ in range(3)], [probs(2, 3, 6) for _ in range(3)])
attn_case("attn_kl_tea_causal", [probs(2, 2, 5)], [probs(2, 2, 5, causal=True)])
attn_case("attn_kl_both_causal_nan", [probs(2, 2, 5, causal=True)], [probs(2, 2, 5, causal=True)])
# zip truncation: 3 student layers vs 2 teacher layers, divisor stays len(stu)=3
attn_case("attn_kl_zip_trunc", [probs(2, 2, 5) for _ in range(3)], [probs(2, 2, 5) for _ in range(2)])


# ---- a5/a6 hidden / embedding MSE -------------------------------------------------------------
def mse_case(name, mod, stu, tea, as_list=True):
    out = {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        if as_list:
            loss, grads = run(lambda s, t: mod(s, t), stu, tea, dt)
        else:
            loss, grads = run(lambda s, t: mod(s[0], t[0]), stu, tea, dt)
        out[f"loss_{tag}"] = loss
        for i, g in enumerate(grads):
            out[f"grad{i}_{tag}"] = g
    save(name, **pack_io("stu", stu), **pack_io("tea", tea), **out)


mse_case("hidden_mse_l3", HiddenMSE(), [bf16(3, 7, 16) for _ in range(3)], [bf16(3, 7, 16) for _ in range(3)])
mse_case("hidden_mse_odd", HiddenMSE(), [bf16(5, 3, 13) for _ in range(2)], [bf16(5, 3, 13) for _ in range(2)])
mse_case("hidden_mse_zip_trunc", HiddenMSE(), [bf16(2, 3, 8) for _ in range(3)], [bf16(2, 3, 8) for _ in range(2)])
mse_case("embed_mse", EmbedMSELoss(), [bf16(3, 7, 16)], [bf16(3, 7, 16)], as_list=False)


# ---- a1/a2/a3 contrastive + logit KL through the reference CLIPModel ------------------------------
class _Tower(torch.nn.Module):
    """Stand-in encoder: returns the given embedding as `last_representation`."""

    def __init__(self, kind):
        super().__init__()
        self.kind = kind

    def forward(self, x, control_output):
        cls = VisionTransformerOutput if self.kind == "image" else TextTransformerOutput
        return cls(last_representation=x)


def clip_case(name, b, d, temperature):
    si, st, ti, tt = bf16(b, d), bf16(b, d), bf16(b, d), bf16(b, d)
    # correlate student with teacher and image with text a little, like a part-trained model
    tt = (ti.float() * 0.6 + 0.8 * tt.float()).to(torch.bfloat16)
    si = (ti.float() + 0.5 * si.float()).to(torch.bfloat16)
    st = (tt.float() + 0.5 * st.float()).to(torch.bfloat16)
    model = CLIPModel(True, _Tower("image"), _Tower("text"))
    hard, soft = HardLabel(), SoftLabel(temperature)
    out = {"temperature": np.float64(temperature)}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        a = si.to(dt).requires_grad_(True)
        bb = st.to(dt).requires_grad_(True)
        s_out = model(bb, a)                       # forward(text, image)
        with torch.no_grad():
            t_out = model(tt.to(dt), ti.to(dt))
        h = 0.5 * (hard(s_out.i2t_logits) + hard(s_out.t2i_logits))
        k = 0.5 * (soft(s_out.i2t_logits, t_out.i2t_logits) + soft(s_out.t2i_logits, t_out.t2i_logits))
        gh = torch.autograd.grad(h, [a, bb], retain_graph=True)
        gk = torch.autograd.grad(k, [a, bb], retain_graph=True)
        out[f"hard_{tag}"] = h.detach().numpy()
        out[f"soft_{tag}"] = k.detach().numpy()
        out[f"hard_i2t_{tag}"] = hard(s_out.i2t_logits).detach().numpy()
        out[f"hard_t2i_{tag}"] = hard(s_out.t2i_logits).detach().numpy()
        out[f"soft_i2t_{tag}"] = soft(s_out.i2t_logits, t_out.i2t_logits).detach().numpy()
        out[f"soft_t2i_{tag}"] = soft(s_out.t2i_logits, t_out.t2i_logits).detach().numpy()
        out[f"dhard_img_{tag}"], out[f"dhard_txt_{tag}"] = gh[0].numpy(), gh[1].numpy()
        out[f"dsoft_img_{tag}"], out[f"dsoft_txt_{tag}"] = gk[0].numpy(), gk[1].numpy()
        out[f"i2t_logits_{tag}"] = s_out.i2t_logits.detach().numpy()
        out[f"tea_i2t_logits_{tag}"] = t_out.i2t_logits.detach().numpy()
        # gradients of the per-module losses w.r.t. materialised logits (HardLabel / SoftLabel API)
        lg = s_out.i2t_logits.detach().clone().requires_grad_(True)
        out[f"dhard_dlogits_{tag}"] = torch.autograd.grad(hard(lg), lg)[0].numpy()
        lg = s_out.i2t_logits.detach().clone().requires_grad_(True)
        out[f"dsoft_dlogits_{tag}"] = torch.autograd.grad(soft(lg, t_out.i2t_logits), lg)[0].numpy()
    out["labels"] = torch.arange(b).numpy()          # hard_label.py:11, int64
    save(name, stu_img=si.float().numpy(), stu_txt=st.float().numpy(),
         tea_img=ti.float().numpy(), tea_txt=tt.float().numpy(), **out)


clip_case("clip_b24_d32_t2", 24, 32, 2.0)
clip_case("clip_b40_d64_t4", 40, 64, 4.0)
clip_case("clip_b130_d72_t1", 130, 72, 1.0)


# ---- section 8f widening: out_l1 / out_cos / attention mean MSE / cos_diff ---------------------------------------
mse_case("out_l1", OutL1Loss(), [bf16(9, 40)], [bf16(9, 40)], as_list=False)
mse_case("out_l1_with_ties", OutL1Loss(), [bf16(4, 16).round()], [bf16(4, 16).round()], as_list=False)   # exact zeros: sign(0) = 0
mse_case("out_cos", OutCosLoss(), [bf16(9, 40)], [bf16(9, 40)], as_list=False)
mse_case("attn_probs_mse", AttentionProbsMSE(), [probs(3, 4, 7) for _ in range(2)], [probs(3, 2, 7) for _ in range(2)])
mse_case("attn_score_mse", AttentionScoreMSE(), [bf16(2, 3, 6, 6) for _ in range(3)], [bf16(2, 3, 6, 6) for _ in range(2)])


def cos_diff_case(name, n):
    s = (bf16(n, n, scale=0.3)).clamp(-1, 1)
    t = (s.float() + 0.2 * bf16(n, n).float()).to(torch.bfloat16)
    mod = CLIPCosDiff()
    out = {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        a = s.to(dt).requires_grad_(True)
        loss = mod(a, t.to(dt))
        loss.backward()
        out[f"loss_{tag}"] = loss.detach().numpy()
        out[f"grad_{tag}"] = a.grad.numpy()
        b = s.to(dt).requires_grad_(True)
        loss_t = mod(b.T, t.to(dt).T)
        loss_t.backward()
        out[f"loss_T_{tag}"] = loss_t.detach().numpy()
        out[f"grad_T_{tag}"] = b.grad.numpy()
    save(name, stu=s.float().numpy(), tea=t.float().numpy(), **out)


cos_diff_case("cos_diff_n17", 17)
cos_diff_case("cos_diff_n64", 64)


# ---- a7-a10 LossCalculator end to end --------------------------------------------------------------
def tower(kind, b, heads, n, w, layers, tea_heads=None):
    cls = VisionTransformerOutput if kind == "image" else TextTransformerOutput
    return cls(last_representation=bf16(b, 32),
               attention_probs=[probs(b, tea_heads or heads, n) for _ in range(layers)],
               representations=[bf16(b, n, w) for _ in range(layers)],
               embedding=bf16(b, n, w))


FIELDS = ("last_representation", "attention_probs", "representations", "embedding")


def tower_to(t, dtype, grad):
    kw = {}
    for f in FIELDS:
        v = getattr(t, f)
        if isinstance(v, list):
            kw[f] = [x.to(dtype).requires_grad_(grad) for x in v]
        else:
            kw[f] = v.to(dtype).requires_grad_(grad)
    return type(t)(**kw)


def flat(t):
    out = []
    for f in FIELDS:
        v = getattr(t, f)
        out += v if isinstance(v, list) else [v]
    return out


def tower_arrays(prefix, t):
    d = {}
    for f in FIELDS:
        v = getattr(t, f)
        if isinstance(v, list):
            for i, x in enumerate(v):
                d[f"{prefix}.{f}.{i}"] = x.float().numpy()
        else:
            d[f"{prefix}.{f}"] = v.float().numpy()
    return d


def calc_case(name, kwargs, model_type, stu, tea):
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        calc = LossCalculator(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in kwargs.items()})
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        if model_type == "all":
            sv, stx = tower_to(stu[0], dt, True), tower_to(stu[1], dt, True)
            tv, ttx = tower_to(tea[0], dt, False), tower_to(tea[1], dt, False)
            model = CLIPModel(True, _Tower("image"), _Tower("text"))

            def fwd(v, x):
                a = v.last_representation / v.last_representation.norm(dim=1, keepdim=True)
                b_ = x.last_representation / x.last_representation.norm(dim=1, keepdim=True)
                lg = a @ b_.t()                       # clip_model.py:37-44
                return CLIPOutput(visual_output=v, text_output=x, i2t_logits=lg, t2i_logits=lg.T)
            s_out, t_out = fwd(sv, stx), fwd(tv, ttx)
            leaves = flat(sv) + flat(stx)
        else:
            s_out, t_out = tower_to(stu, dt, True), tower_to(tea, dt, False)
            leaves = flat(s_out)
        loss, res = calc(s_out, t_out, model_type)
        grads = torch.autograd.grad(loss, leaves, allow_unused=True)
        out[f"loss_{tag}"] = loss.detach().numpy()
        for k, v in res.items():
            out[f"res.{k}_{tag}"] = v.detach().numpy() if torch.is_tensor(v) else np.float64(v)
        for i, g in enumerate(grads):
            if g is not None:
                out[f"grad{i}_{tag}"] = g.numpy()
    out["percent_keys"] = np.array(list(calc.percent.keys()))
    out["percent_vals"] = np.array([float(v) for v in calc.percent.values()])
    out["scale_keys"] = np.array(list(calc.loss_scale.keys()))
    out["scale_vals"] = np.array([float(v) for v in calc.loss_scale.values()])
    if model_type == "all":
        arrs = {**tower_arrays("stu.visual", stu[0]), **tower_arrays("stu.text", stu[1]),
                **tower_arrays("tea.visual", tea[0]), **tower_arrays("tea.text", tea[1])}
    else:
        arrs = {**tower_arrays("stu", stu), **tower_arrays("tea", tea)}
    save(name, **arrs, **out)


calc_case("calc_image_stage",
          dict(loss_name=["attention_probs_kl", "hidden_rep_mse"], loss_scale={"attention_probs_kl": 0.5},
               percent={"attention_probs_kl": 0.3}),
          "image", tower("image", 4, 3, 6, 24, 2), tower("image", 4, 3, 6, 24, 2))
calc_case("calc_text_stage",
          dict(loss_name=["attention_probs_kl", "hidden_rep_mse", "embedding_mse"],
               loss_scale={"hidden_rep_mse": 2.0}),
          "text", tower("text", 3, 2, 7, 16, 2), tower("text", 3, 2, 7, 16, 2, tea_heads=4))
calc_case("calc_lclip_stage",
          dict(loss_name=["hard_label", "soft_label", "hidden_rep_mse"], temperature=2.0,
               loss_scale={"soft_label": 0.25},
               percent={"hard_label": 0.5, "soft_label": 0.25, "hidden_rep_mse": 0.25}),
          "all",
          (tower("image", 20, 3, 6, 24, 2), tower("text", 20, 2, 7, 16, 2)),
          (tower("image", 20, 3, 6, 24, 2), tower("text", 20, 2, 7, 16, 2)))
calc_case("calc_lclip_logits_only",
          dict(loss_name=["hard_label", "soft_label"], temperature=3.0),
          "all",
          (tower("image", 36, 3, 6, 24, 1), tower("text", 36, 2, 7, 16, 1)),
          (tower("image", 36, 3, 6, 24, 1), tower("text", 36, 2, 7, 16, 1)))

# the three shipped recipes (config/final_config/{image,text,l_clip}.yaml: loss_name lists)
calc_case("calc_shipped_image", dict(loss_name=["out_l1", "out_cos"]), "image",
          tower("image", 6, 3, 6, 24, 2), tower("image", 6, 3, 6, 24, 2))
calc_case("calc_shipped_lclip", dict(loss_name=["out_l1", "out_cos", "cos_diff"]), "all",
          (tower("image", 20, 3, 6, 24, 1), tower("text", 20, 2, 7, 16, 1)),
          (tower("image", 20, 3, 6, 24, 1), tower("text", 20, 2, 7, 16, 1)))
calc_case("calc_attn_mse_mix",
          dict(loss_name=["attention_probs_mse", "attention_probs_kl", "hidden_rep_mse", "out_l1"],
               loss_scale={"attention_probs_mse": 3.0}),
          "image", tower("image", 3, 4, 6, 24, 2), tower("image", 3, 2, 6, 24, 2))

# ---- section 8f widening, second batch: out_kl / out_ce / logits_mse (drawn after every earlier case so those stay put) ---
mse_case("out_kl_t2", OutKLLoss(2.0), [bf16(9, 40)], [bf16(9, 40)], as_list=False)
mse_case("out_kl_t05_wide", OutKLLoss(0.5), [bf16(5, 300, scale=2.0)], [bf16(5, 300, scale=2.0)], as_list=False)
mse_case("out_ce", OutCELoss(), [bf16(9, 40)], [bf16(9, 40)], as_list=False)
mse_case("out_ce_wide", OutCELoss(), [bf16(5, 300, scale=3.0)], [bf16(5, 300, scale=3.0)], as_list=False)
mse_case("logits_mse", LogitsMSE(), [bf16(13, 13, scale=0.3)], [bf16(13, 13, scale=0.3)], as_list=False)
calc_case("calc_out_kl_ce_logits_mse",
          dict(loss_name=["out_kl", "out_ce", "logits_mse", "hard_label"], temperature=2.0, loss_scale={"out_kl": 0.1}),
          "all",
          (tower("image", 20, 3, 6, 24, 1), tower("text", 20, 2, 7, 16, 1)),
          (tower("image", 20, 3, 6, 24, 1), tower("text", 20, 2, 7, 16, 1)))
calc_case("calc_out_kl_ce_image", dict(loss_name=["out_ce", "out_kl", "hidden_rep_mse"], temperature=4.0), "image",
          tower("image", 6, 3, 6, 24, 2), tower("image", 6, 3, 6, 24, 2))

# ---- third batch: LastValueMapKL (softmax over the HEAD axis of the value-relation maps, as the encoder emits them) -----
mse_case("value_map_kl_h3_n7", LastValueMapKL(), [probs(3, 3, 7)], [probs(3, 3, 7)], as_list=False)
mse_case("value_map_kl_h12_n10", LastValueMapKL(), [probs(2, 12, 10)], [probs(2, 12, 10)], as_list=False)
mse_case("value_map_kl_scores", LastValueMapKL(), [bf16(2, 8, 5, 5, scale=3.0)], [bf16(2, 8, 5, 5, scale=3.0)], as_list=False)

# ---- validation metrics (dual_distill_model.py:204-224, 271-275).  The module itself cannot be imported here (it pulls in
# pytorch_lightning and torchmetrics, neither is installed): norm_and_logits and log_diag_score are restated line by line,
# torchmetrics' multiclass top-k accuracy as "label among torch.topk(logits, k)".
def retrieval_case(name, b, d, noise):
    img = bf16(b, d)
    txt = (img.float() + noise * bf16(b, d).float()).to(torch.bfloat16)
    out = {}
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        a, c = img.to(dt), txt.to(dt)
        a = a / a.norm(dim=1, keepdim=True)
        c = c / c.norm(dim=1, keepdim=True)
        logits = a @ c.t()
        label = torch.arange(b)
        for k in (1, 3, 5, 10, 20, 50):
            out[f"acc_top{k}_{tag}"] = (logits.topk(min(k, b), dim=1).indices == label[:, None]).any(1).to(dt).mean().numpy()
        out[f"softmax_mean_score_{tag}"] = torch.diagonal(torch.softmax(logits, dim=1)).mean().numpy()
        out[f"mean_score_{tag}"] = torch.diagonal(logits).mean().numpy()
    save(name, img=img.float().numpy(), txt=txt.float().numpy(), **out)


retrieval_case("retrieval_b200_d64", 200, 64, 3.0)
retrieval_case("retrieval_b77_d40", 77, 40, 1.5)

# ---- host-logic facts (flags, errors) -------------------------------------------------------------
facts = {}
with contextlib.redirect_stdout(io.StringIO()):
    c = LossCalculator(["embedding_mse", "hidden_rep_mse", "attention_probs_kl", "attention_probs_mse"])
co = c.get_control_output()
facts["control_all4"] = np.array([co.need_emb, co.need_attn_score, co.need_value_map, co.need_attn_prob, co.need_rep])
with contextlib.redirect_stdout(io.StringIO()):
    c = LossCalculator(["attention_probs_kl"])
co = c.get_control_output()
# F8: the typo sets a non-existent attribute, need_attn_prob stays False
facts["control_kl_only"] = np.array([co.need_emb, co.need_attn_score, co.need_value_map, co.need_attn_prob, co.need_rep])
facts["control_kl_only_has_typo_attr"] = np.array(hasattr(co, "attention_probs_mse"))
with contextlib.redirect_stdout(io.StringIO()):
    c = LossCalculator(["hard_label", "soft_label"], temperature=1.0, percent={"hard_label": 0.7})
facts["percent_fill_keys"] = np.array(list(c.percent.keys()))
facts["percent_fill_vals"] = np.array(list(c.percent.values()), dtype=np.float64)
try:
    with contextlib.redirect_stdout(io.StringIO()):
        LossCalculator(["nope"])
    facts["invalid_name_error"] = np.array("none")
except ValueError as e:
    facts["invalid_name_error"] = np.array(str(e))
try:
    with contextlib.redirect_stdout(io.StringIO()):
        LossCalculator(["hard_label", "hidden_rep_mse"], percent={"hard_label": 1.0})
    facts["neg_percent_error"] = np.array("none")
except ValueError as e:
    facts["neg_percent_error"] = np.array("ValueError")
try:
    AttentionProbsKL()([], [])
    facts["empty_list_error"] = np.array("none")
except ZeroDivisionError:
    facts["empty_list_error"] = np.array("ZeroDivisionError")
# fill-in rule quirk: default = (1-sum)/len(GIVEN), so 3 names with 1 given breaks the sum assert
try:
    with contextlib.redirect_stdout(io.StringIO()):
        LossCalculator(["hard_label", "soft_label", "hidden_rep_mse"], temperature=1.0, percent={"hard_label": 0.5})
    facts["fill_3names_1given"] = np.array("none")
except AssertionError:
    facts["fill_3names_1given"] = np.array("AssertionError")
save("host_facts", **facts)
print("torch", torch.__version__)
