"""The CPU oracle (oracle/closed_form.py, oracle/torch_port.py) against the golden fixtures that were
produced by the reference's own modules (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import golden, numbered, rel_l2
from oracle import closed_form as cf
from oracle import torch_port as tp

ATTN = ["attn_kl_heads4v2", "attn_kl_even", "attn_kl_tea_causal", "attn_kl_zip_trunc"]
MSE = ["hidden_mse_l3", "hidden_mse_odd", "hidden_mse_zip_trunc"]
CLIP = ["clip_b24_d32_t2", "clip_b40_d64_t4", "clip_b130_d72_t1"]


@pytest.mark.parametrize("name", ATTN)
def test_attn_kl_closed_form(name):
    g = golden(name)
    stu, tea = numbered(g, "stu"), numbered(g, "tea")
    loss, grads = cf.attention_probs_kl(stu, tea)
    assert abs(loss - g["loss_f64"]) <= 1e-10 * abs(g["loss_f64"])
    for i, gr in enumerate(grads):
        assert rel_l2(gr, g[f"grad{i}_f64"]) <= 1e-10 or np.all(g[f"grad{i}_f64"] == 0) and np.all(gr == 0)


def test_attn_kl_nan_on_coincident_zeros():
    g = golden("attn_kl_both_causal_nan")
    assert np.isnan(g["loss_f32"]) and np.isnan(g["loss_f64"])      # reference behaviour (F10)
    loss, grads = cf.attention_probs_kl(numbered(g, "stu"), numbered(g, "tea"))
    assert np.isnan(loss)
    assert np.array_equal(np.isnan(grads[0]), np.isnan(g["grad0_f64"]))


@pytest.mark.parametrize("name", ATTN)
def test_attn_kl_torch_port(name):
    g = golden(name)
    stu = [torch.tensor(x, requires_grad=True) for x in numbered(g, "stu")]
    tea = [torch.tensor(x) for x in numbered(g, "tea")]
    loss = tp.attention_probs_kl(stu, tea)
    loss.backward()
    assert loss.item() == pytest.approx(float(g["loss_f32"]), rel=1e-6)
    for i, s in enumerate(stu):
        if s.grad is not None:
            assert rel_l2(s.grad.numpy(), g[f"grad{i}_f32"]) <= 1e-6


@pytest.mark.parametrize("name", MSE)
def test_hidden_mse(name):
    g = golden(name)
    stu, tea = numbered(g, "stu"), numbered(g, "tea")
    loss, grads = cf.hidden_mse(stu, tea)
    assert abs(loss - g["loss_f64"]) <= 1e-12 * abs(g["loss_f64"])
    for i, gr in enumerate(grads):
        assert np.allclose(gr, g[f"grad{i}_f64"], rtol=1e-12, atol=0)
    ts = [torch.tensor(x, requires_grad=True) for x in stu]
    l2 = tp.hidden_mse(ts, [torch.tensor(x) for x in tea])
    assert l2.item() == pytest.approx(float(g["loss_f32"]), rel=1e-6)


def test_embed_mse():
    g = golden("embed_mse")
    loss, grad = cf.embed_mse(g["stu0"], g["tea0"])
    assert abs(loss - g["loss_f64"]) <= 1e-12 * abs(g["loss_f64"])
    assert np.allclose(grad, g["grad0_f64"], rtol=1e-12, atol=0)
    assert tp.embed_mse(torch.tensor(g["stu0"]), torch.tensor(g["tea0"])).item() == pytest.approx(
        float(g["loss_f32"]), rel=1e-6)


def test_empty_lists_raise_like_reference():
    assert str(golden("host_facts")["empty_list_error"]) == "ZeroDivisionError"
    with pytest.raises(ZeroDivisionError):
        cf.attention_probs_kl([], [])
    with pytest.raises(ZeroDivisionError):
        tp.attention_probs_kl([], [])
    with pytest.raises(ZeroDivisionError):
        cf.hidden_mse([], [])


@pytest.mark.parametrize("name", CLIP)
def test_contrastive_closed_form(name):
    g = golden(name)
    T = float(g["temperature"])
    s, st = cf.clip_logits(g["stu_img"], g["stu_txt"])
    assert np.allclose(s, g["i2t_logits_f64"], rtol=0, atol=1e-12)
    assert np.array_equal(cf.labels(s.shape[0]), g["labels"]) and g["labels"].dtype == np.int64
    h = cf.contrastive_from_embeddings(g["stu_img"], g["stu_txt"], g["tea_img"], g["tea_txt"], T, w_hard=1.0)
    k = cf.contrastive_from_embeddings(g["stu_img"], g["stu_txt"], g["tea_img"], g["tea_txt"], T, w_soft=1.0)
    for key in ("hard", "hard_i2t", "hard_t2i", "soft", "soft_i2t", "soft_t2i"):
        assert h[key] == pytest.approx(float(g[key + "_f64"]), rel=1e-10), key
    assert rel_l2(h["d_img"], g["dhard_img_f64"]) <= 1e-10
    assert rel_l2(h["d_txt"], g["dhard_txt_f64"]) <= 1e-10
    assert rel_l2(k["d_img"], g["dsoft_img_f64"]) <= 1e-9
    assert rel_l2(k["d_txt"], g["dsoft_txt_f64"]) <= 1e-9
    # per-module (materialised logits) API
    l, dl = cf.hard_label(g["i2t_logits_f64"])
    assert l == pytest.approx(float(g["hard_i2t_f64"]), rel=1e-12)
    assert rel_l2(dl, g["dhard_dlogits_f64"]) <= 1e-12
    l, dl = cf.soft_label(g["i2t_logits_f64"], g["tea_i2t_logits_f64"], T)
    assert l == pytest.approx(float(g["soft_i2t_f64"]), rel=1e-9)
    assert rel_l2(dl, g["dsoft_dlogits_f64"]) <= 1e-9


@pytest.mark.parametrize("name", CLIP)
def test_contrastive_torch_port(name):
    g = golden(name)
    T = float(g["temperature"])
    a = torch.tensor(g["stu_img"], requires_grad=True)
    b = torch.tensor(g["stu_txt"], requires_grad=True)
    s, st = tp.clip_logits(a, b)
    t, tt = tp.clip_logits(torch.tensor(g["tea_img"]), torch.tensor(g["tea_txt"]))
    h = 0.5 * (tp.hard_label(s) + tp.hard_label(st))
    k = 0.5 * (tp.soft_label(s, t, T) + tp.soft_label(st, tt, T))
    assert h.item() == pytest.approx(float(g["hard_f32"]), rel=1e-6)
    assert k.item() == pytest.approx(float(g["soft_f32"]), rel=1e-5)
    gi, gt = torch.autograd.grad(h, [a, b])
    assert rel_l2(gi.numpy(), g["dhard_img_f32"]) <= 1e-5


def test_weight_rules():
    f = golden("host_facts")
    scale, percent = cf.resolve_weights(["hard_label", "soft_label"], None, {"hard_label": 0.7})
    assert list(percent.keys()) == [str(k) for k in f["percent_fill_keys"]]
    assert np.allclose(list(percent.values()), f["percent_fill_vals"], rtol=0, atol=1e-15)
    assert str(f["neg_percent_error"]) == "ValueError"
    with pytest.raises(ValueError):
        cf.resolve_weights(["hard_label", "hidden_rep_mse"], None, {"hard_label": 1.0})
    assert str(f["fill_3names_1given"]) == "AssertionError"
    with pytest.raises(AssertionError):
        cf.resolve_weights(["hard_label", "soft_label", "hidden_rep_mse"], None, {"hard_label": 0.5})


def _tower_from(g, prefix):
    d = {}
    for f in ("last_representation", "embedding"):
        d[f] = g[f"{prefix}.{f}"]
    for f in ("attention_probs", "representations"):
        d[f] = numbered(g, f"{prefix}.{f}.")
    return d


@pytest.mark.parametrize("name,kind", [("calc_image_stage", "one"), ("calc_text_stage", "one"),
                                       ("calc_lclip_stage", "two"), ("calc_lclip_logits_only", "two")])
def test_calculator_port(name, kind):
    g = golden(name)
    names = [str(k) for k in g["scale_keys"]]
    scale = dict(zip(names, g["scale_vals"].tolist()))
    percent = dict(zip([str(k) for k in g["percent_keys"]], g["percent_vals"].tolist()))
    temperature = {"calc_lclip_stage": 2.0, "calc_lclip_logits_only": 3.0}.get(name)

    def to_t(d, grad):
        out = {}
        for k, v in d.items():
            out[k] = [torch.tensor(x, requires_grad=grad) for x in v] if isinstance(v, list) \
                else torch.tensor(v, requires_grad=grad)
        return out
    if kind == "one":
        stu, tea = to_t(_tower_from(g, "stu"), True), to_t(_tower_from(g, "tea"), False)
        loss, res = tp.one_tower(names, scale, percent, temperature, stu, tea)
    else:
        stu = {"visual": to_t(_tower_from(g, "stu.visual"), True), "text": to_t(_tower_from(g, "stu.text"), True)}
        tea = {"visual": to_t(_tower_from(g, "tea.visual"), False), "text": to_t(_tower_from(g, "tea.text"), False)}
        loss, res = tp.two_tower(names, scale, percent, temperature, stu, tea)
    assert loss.item() == pytest.approx(float(g["loss_f32"]), rel=2e-6)
    for k, v in res.items():
        assert float(v.detach()) == pytest.approx(float(g[f"res.{k}_f32"]), rel=2e-6), k
    assert set(res.keys()) == {k[4:-4] for k in g if k.startswith("res.") and k.endswith("_f32")}


# ---- section 8f widening -------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,fn", [("out_l1", "out_l1"), ("out_l1_with_ties", "out_l1"), ("out_cos", "out_cos")])
def test_out_losses(name, fn):
    g = golden(name)
    loss, grad = getattr(cf, fn)(g["stu0"], g["tea0"])
    assert abs(loss - g["loss_f64"]) <= 1e-12 * max(abs(g["loss_f64"]), 1e-30)
    assert np.allclose(grad, g["grad0_f64"], rtol=1e-10, atol=1e-18)
    t = getattr(tp, fn)(torch.tensor(g["stu0"]), torch.tensor(g["tea0"]))
    assert t.item() == pytest.approx(float(g["loss_f32"]), rel=1e-6)


@pytest.mark.parametrize("name", ["attn_probs_mse", "attn_score_mse"])
def test_attention_mean_mse(name):
    g = golden(name)
    stu, tea = numbered(g, "stu"), numbered(g, "tea")
    loss, grads = cf.attention_mean_mse(stu, tea)
    assert abs(loss - g["loss_f64"]) <= 1e-12 * abs(g["loss_f64"])
    for i, gr in enumerate(grads):
        assert np.allclose(gr, g[f"grad{i}_f64"], rtol=1e-10, atol=1e-20)
    t = tp.attention_mean_mse([torch.tensor(x) for x in stu], [torch.tensor(x) for x in tea])
    assert t.item() == pytest.approx(float(g["loss_f32"]), rel=1e-6)


@pytest.mark.parametrize("name", ["cos_diff_n17", "cos_diff_n64"])
def test_cos_diff(name):
    g = golden(name)
    loss, grad = cf.cos_diff(g["stu"], g["tea"])
    assert loss == pytest.approx(float(g["loss_f64"]), rel=1e-12)
    assert np.allclose(grad, g["grad_f64"], rtol=1e-12, atol=0)
    loss_t, grad_t = cf.cos_diff(g["stu"].T, g["tea"].T)
    assert loss_t == pytest.approx(float(g["loss_T_f64"]), rel=1e-12)
    assert np.allclose(grad_t.T, g["grad_T_f64"], rtol=1e-12, atol=0)
    assert tp.cos_diff(torch.tensor(g["stu"]), torch.tensor(g["tea"])).item() == pytest.approx(float(g["loss_f32"]), rel=1e-6)


@pytest.mark.parametrize("name,fn,temperature", [("out_kl_t2", "out_kl", 2.0), ("out_kl_t05_wide", "out_kl", 0.5),
                                                 ("out_ce", "out_ce", None), ("out_ce_wide", "out_ce", None),
                                                 ("logits_mse", "logits_mse", None),
                                                 ("value_map_kl_h3_n7", "last_value_map_kl", None),
                                                 ("value_map_kl_h12_n10", "last_value_map_kl", None),
                                                 ("value_map_kl_scores", "last_value_map_kl", None)])
def test_row_softmax_and_logits_mse(name, fn, temperature):
    g = golden(name)
    args = (temperature,) if temperature else ()
    loss, grad = getattr(cf, fn)(g["stu0"], g["tea0"], *args)
    assert loss == pytest.approx(float(g["loss_f64"]), rel=1e-11)
    assert np.allclose(grad, g["grad0_f64"], rtol=1e-9, atol=1e-16)
    t = getattr(tp, fn)(torch.tensor(g["stu0"]), torch.tensor(g["tea0"]), *args)
    assert t.item() == pytest.approx(float(g["loss_f32"]), rel=1e-6)


@pytest.mark.parametrize("name,kind,temperature", [("calc_shipped_image", "one", None), ("calc_shipped_lclip", "two", None),
                                                   ("calc_attn_mse_mix", "one", None), ("calc_out_kl_ce_image", "one", 4.0),
                                                   ("calc_out_kl_ce_logits_mse", "two", 2.0)])
def test_calculator_port_shipped_recipes(name, kind, temperature):
    """The loss lists of config/final_config/{image,text,l_clip}.yaml through the torch port."""
    g = golden(name)
    names = [str(k) for k in g["scale_keys"]]
    scale = dict(zip(names, g["scale_vals"].tolist()))
    percent = dict(zip([str(k) for k in g["percent_keys"]], g["percent_vals"].tolist()))

    def to_t(d):
        return {k: [torch.tensor(x) for x in v] if isinstance(v, list) else torch.tensor(v) for k, v in d.items()}
    if kind == "one":
        loss, res = tp.one_tower(names, scale, percent, temperature, to_t(_tower_from(g, "stu")), to_t(_tower_from(g, "tea")))
    else:
        stu = {"visual": to_t(_tower_from(g, "stu.visual")), "text": to_t(_tower_from(g, "stu.text"))}
        tea = {"visual": to_t(_tower_from(g, "tea.visual")), "text": to_t(_tower_from(g, "tea.text"))}
        loss, res = tp.two_tower(names, scale, percent, temperature, stu, tea)
    assert float(loss) == pytest.approx(float(g["loss_f32"]), rel=2e-6)
    for k, v in res.items():
        assert float(v) == pytest.approx(float(g[f"res.{k}_f32"]), rel=2e-6), k


@pytest.mark.parametrize("name", ["retrieval_b200_d64", "retrieval_b77_d40"])
def test_retrieval_metrics(name):
    """Validation metrics (dual_distill_model.py:204-224): the oracle's rank formulation (label in the top k iff fewer than
    k logits of its row are strictly larger) against torch.topk membership."""
    g = golden(name)
    res = cf.retrieval_metrics(g["img"], g["txt"])
    for k, v in res.items():
        assert v == pytest.approx(float(g[f"{k}_f64"]), rel=1e-12, abs=1e-15), k


@pytest.mark.parametrize("name", CLIP)
@pytest.mark.parametrize("chunk", [7, 64])
def test_chunked_fp64_oracle_matches_golden_and_closed_form(name, chunk):
    """oracle/chunked_fp64.py (the full-size checker: literal formulas, row chunks, two passes) against the reference-made
    fixtures and the dense float64 oracle, including the cos_diff / logits_mse terms and sampled-row gradients."""
    from oracle import chunked_fp64 as ck
    g = golden(name)
    T = float(g["temperature"])
    emb = [torch.tensor(g[k]).to(torch.bfloat16).float() for k in ("stu_img", "stu_txt", "tea_img", "tea_txt")]
    b = emb[0].shape[0]
    rows_i, rows_t = list(range(0, b, 3)), list(range(1, b, 4))
    w = {"hard": 0.7, "soft": 0.3, "cos_diff": 0.4, "logits_mse": 1.5}
    out = ck.contrastive_chunked(*emb, temperature=T, weights=w, sample_img=rows_i, sample_txt=rows_t, chunk=chunk)
    assert out["hard"] == pytest.approx(float(g["hard_f64"]), rel=1e-10)
    assert out["soft"] == pytest.approx(float(g["soft_f64"]), rel=1e-9)
    np_emb = [x.numpy() for x in emb]
    ref = cf.contrastive_from_embeddings(*np_emb, T, w_hard=w["hard"], w_soft=w["soft"])
    s, _ = cf.clip_logits(np_emb[0], np_emb[1])
    t, _ = cf.clip_logits(np_emb[2], np_emb[3])
    cd, g_cd = cf.cos_diff(s, t)
    cd2, g_cd2 = cf.cos_diff(s.T, t.T)
    lm, g_lm = cf.logits_mse(s, t)
    assert out["cos_diff"] == pytest.approx(0.5 * (cd + cd2), rel=1e-12)
    assert out["logits_mse"] == pytest.approx(lm, rel=1e-12)
    dl = ref["d_logits"] + w["cos_diff"] * 0.5 * (g_cd + g_cd2.T) + w["logits_mse"] * g_lm
    d_img, d_txt = cf.clip_logits_backward(np_emb[0], np_emb[1], dl)
    assert rel_l2(out["d_img"].numpy(), d_img[rows_i]) <= 1e-10
    assert rel_l2(out["d_txt"].numpy(), d_txt[rows_t]) <= 1e-10
