"""CPU oracle, torch flavour: the reference's loss path restated as plain functions.
TEST / BASELINE INFRASTRUCTURE ONLY -- never imported by the product package.

Why it exists next to `closed_form.py`: the reference's arithmetic lives in PyTorch ATen
(`KLDivLoss`, `MSELoss`, `CrossEntropyLoss`, `softmax`, `@`), and the
reference tree itself cannot travel to the GPU box.  This port issues the same ATen op
sequence as the reference modules, so (a) autograd provides gradients at sizes where the
float64 numpy oracle is too slow, and (b) `bench.py --impl reference` / `cpu_baseline` can time
"the reference's CPU implementation of the path" on the box's host cores.

Pinned against the reference itself by `tests/golden/make_golden.py` (run in the build
container, where /root/reference is mounted) and `tests/test_oracle_golden.py`.

Citations are relative to /root/reference.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

IMAGE_TEXT_LOSS = ("hard_label", "soft_label", "logits_mse", "fine_grain", "cos_diff")


def hard_label(logits: torch.Tensor) -> torch.Tensor:
    # loss_component/hard_label.py:10-12
    target = torch.arange(logits.shape[0], device=logits.device)
    return F.cross_entropy(logits, target, reduction="mean")


def soft_label(stu_logits, tea_logits, temperature) -> torch.Tensor:
    # loss_component/soft_label.py:11-16  (softmax().log(), not log_softmax)
    log_q = F.softmax(stu_logits / temperature, dim=1).log()
    p = F.softmax(tea_logits / temperature, dim=1)
    return F.kl_div(log_q, p, reduction="sum") * temperature ** 2


def clip_logits(img: torch.Tensor, txt: torch.Tensor):
    # component/clip_model.py:36-44
    a = img / img.norm(dim=1, keepdim=True)
    b = txt / txt.norm(dim=1, keepdim=True)
    s = a @ b.t()
    return s, s.T


def _head_mean(x: torch.Tensor) -> torch.Tensor:
    return torch.sum(x, dim=1) / x.shape[1]


def attention_probs_kl(stu: Sequence[torch.Tensor], tea: Sequence[torch.Tensor]):
    # loss_component/attention_probs_kl.py:10-22
    acc = 0
    for k, (s, t) in enumerate(zip(stu, tea)):
        term = F.kl_div(_head_mean(s).log(), _head_mean(t), reduction="sum")
        acc = term if k == 0 else acc + term
    acc /= len(stu)          # ZeroDivisionError on empty lists, as in the reference (F8)
    return acc


def hidden_mse(stu: Sequence[torch.Tensor], tea: Sequence[torch.Tensor]):
    # loss_component/hidden_mse.py:9-17
    acc = 0
    for k, (s, t) in enumerate(zip(stu, tea)):
        term = F.mse_loss(s, t)
        acc = term if k == 0 else acc + term
    acc /= len(stu)
    return acc


def embed_mse(stu: torch.Tensor, tea: torch.Tensor):
    # loss_component/embed_mse.py:9-10
    return F.mse_loss(stu, tea)


def out_l1(stu, tea):
    # loss_component/out_l1.py:9-10
    return F.l1_loss(stu, tea)


def out_cos(stu, tea):
    # loss_component/out_cos.py:10-11
    return F.cosine_embedding_loss(stu, tea, torch.ones(len(stu), device=stu.device))


def attention_mean_mse(stu: Sequence[torch.Tensor], tea: Sequence[torch.Tensor]):
    # loss_component/attention_probs_mse.py:10-22 and attention_score_mse.py:10-22
    acc = 0
    for k, (s, t) in enumerate(zip(stu, tea)):
        term = F.mse_loss(_head_mean(s), _head_mean(t))
        acc = term if k == 0 else acc + term
    acc /= len(stu)
    return acc


def out_kl(stu, tea, temperature):
    # loss_component/out_kl.py:12-16
    return F.kl_div(F.log_softmax(stu / temperature, dim=1), F.softmax(tea / temperature, dim=1),
                    reduction="sum") * temperature ** 2


def out_ce(stu, tea):
    # loss_component/out_ce.py:9-13
    return F.cross_entropy(stu, tea.softmax(dim=1))


def last_value_map_kl(stu, tea):
    # loss_component/last_value_map_kl.py:10-14
    return F.kl_div(F.softmax(stu, dim=1).log(), F.softmax(tea, dim=1), reduction="sum")


def logits_mse(stu_logits, tea_logits):
    # loss_component/logits_mse.py:9-10
    return F.mse_loss(stu_logits, tea_logits)


def _neg_elements(x):
    n = x.shape[0]
    return x.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten()      # clip_cos_diff.py:5-8


def cos_diff(stu_logits, tea_logits):
    # loss_component/clip_cos_diff.py:16-23
    pos = torch.mean(torch.relu(torch.diagonal(tea_logits) - torch.diagonal(stu_logits)))
    neg = torch.mean(torch.relu(_neg_elements(stu_logits) - _neg_elements(tea_logits)))
    return neg + pos


def one_tower(names, scale, percent, temperature, stu: Dict, tea: Dict):
    """_loss.py:155-202 for the in-scope names.  `stu`/`tea` are dicts with keys
    last_representation / attention_probs / representations / embedding."""
    res = {}
    for n in names:
        if n == "embedding_mse":
            res[n] = embed_mse(stu["embedding"], tea["embedding"])
        elif n == "hidden_rep_mse":
            res[n] = hidden_mse(stu["representations"], tea["representations"])
        elif n == "attention_probs_kl":
            res[n] = attention_probs_kl(stu["attention_probs"], tea["attention_probs"])
        elif n == "attention_probs_mse":
            res[n] = attention_mean_mse(stu["attention_probs"], tea["attention_probs"])
        elif n == "attention_score_mse":
            res[n] = attention_mean_mse(stu["attention_scores"], tea["attention_scores"])
        elif n == "out_l1":
            res[n] = out_l1(stu["last_representation"], tea["last_representation"])
        elif n == "out_cos":
            res[n] = out_cos(stu["last_representation"], tea["last_representation"])
        elif n == "out_kl":
            assert temperature, "You should give the temperature for the kl loss"
            res[n] = out_kl(stu["last_representation"], tea["last_representation"], temperature)
        elif n == "out_ce":
            res[n] = out_ce(stu["last_representation"], tea["last_representation"])
        elif n == "last_value_map_kl":
            res[n] = last_value_map_kl(stu["value_map"], tea["value_map"])
    loss = 0
    for n, sc in scale.items():
        if n in IMAGE_TEXT_LOSS:
            continue
        res[n] = res[n] * sc
        loss += res[n] * percent[n]
    return loss, res


def two_tower(names, scale, percent, temperature, stu: Dict, tea: Dict):
    """_loss.py:118-153.  `stu`/`tea`: {'visual': {...}, 'text': {...}} tower dicts as above;
    logits are produced from last_representation exactly as clip_model.py:36-44 does."""
    res = {}
    il, ires = one_tower(names, scale, percent, temperature, stu["visual"], tea["visual"])
    tl, tres = one_tower(names, scale, percent, temperature, stu["text"], tea["text"])
    for k, v in ires.items():
        res["image_" + k] = v
    for k, v in tres.items():
        res["text_" + k] = v
    s_i2t, s_t2i = clip_logits(stu["visual"]["last_representation"], stu["text"]["last_representation"])
    if "soft_label" in names or "cos_diff" in names or "logits_mse" in names:
        t_i2t, t_t2i = clip_logits(tea["visual"]["last_representation"], tea["text"]["last_representation"])
    for n in names:
        if n == "hard_label":
            res[n] = 0.5 * (hard_label(s_i2t) + hard_label(s_t2i))
        elif n == "soft_label":
            assert temperature
            res[n] = 0.5 * (soft_label(s_i2t, t_i2t, temperature) + soft_label(s_t2i, t_t2i, temperature))
        elif n == "cos_diff":
            res[n] = 0.5 * (cos_diff(s_i2t, t_i2t) + cos_diff(s_t2i, t_t2i))
        elif n == "logits_mse":
            res[n] = 0.5 * (logits_mse(s_i2t, t_i2t) + logits_mse(s_t2i, t_t2i))
    loss = 0.5 * (il + tl)
    for n, sc in scale.items():
        if n in IMAGE_TEXT_LOSS:
            res[n] = res[n] * sc
            loss += res[n] * percent[n]
    return loss, res


# --------------------------------------------------------------------------------------
# Synthetic workloads (SURVEY.md section 8d distributions; seed 2022 echoes main.py:24)
# --------------------------------------------------------------------------------------
def synth_attention(b, h, n, layers, gen, dtype=torch.bfloat16, causal=False):
    out = []
    for _ in range(layers):
        x = torch.randn(b, h, n, n, generator=gen)
        if causal:
            x = x + torch.full((n, n), float("-inf")).triu_(1)
        out.append(torch.softmax(x, dim=-1).to(dtype))
    return out


def synth_hidden(b, n, w, layers, gen, dtype=torch.bfloat16):
    return [torch.randn(b, n, w, generator=gen).to(dtype) for _ in range(layers)]


def synth_embeddings(b, d, gen, dtype=torch.bfloat16):
    """teacher ~ N(0,1); student = teacher + 0.5 N(0,1); un-normalised."""
    ti = torch.randn(b, d, generator=gen)
    tt = ti * 0.6 + 0.8 * torch.randn(b, d, generator=gen)   # correlated image/text pairs
    si = ti + 0.5 * torch.randn(b, d, generator=gen)
    st = tt + 0.5 * torch.randn(b, d, generator=gen)
    return [x.to(dtype) for x in (si, st, ti, tt)]


def stage_step_cpu(names, stu: Dict, tea: Dict, temperature=None, two=False, threads: Optional[int] = None):
    """One fwd+bwd of the reference path on fp32 CPU copies (F11: upcast first). Returns loss float."""
    if threads:
        torch.set_num_threads(threads)
    scale = {n: 1 for n in names}
    percent = {n: 1 / len(names) for n in names}
    fn = two_tower if two else one_tower
    loss, _ = fn(names, scale, percent, temperature, stu, tea)
    loss.backward()
    return float(loss.detach())
