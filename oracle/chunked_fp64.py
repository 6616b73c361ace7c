"""Oracle, chunked float64 flavour: the two-tower logit losses of the reference evaluated from embeddings at sizes where
the B x B matrices do not fit (B = 32768: 8.6 GB per fp64 matrix) -- row chunks of the logits in torch float64.
TEST / BASELINE INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's parity check (as the
checker, never as the thing measured); the product package never imports it.

What it restates (citations relative to /root/reference):
  * model/component/clip_model.py:36-44   a_hat = a/||a||, S = a_hat_img a_hat_txt^T, t2i = S.T
  * model/loss_component/hard_label.py:10-12   CrossEntropy(mean)(S, arange(B))
  * model/loss_component/soft_label.py:11-16   KLDiv(sum)(softmax(S/T).log(), softmax(Tt/T)) * T^2
  * model/loss_component/clip_cos_diff.py:5-23 mean relu(Tt_ii - S_ii) + mean_{i != j} relu(S_ij - Tt_ij)
  * model/loss_component/logits_mse.py:9-10    mean (S - Tt)^2
  * model/_loss.py:130-145                     0.5 * (loss(i2t) + loss(t2i)) for each of them
The formulas are the literal ones (log-softmax per row / per column of the chunk, p_t (log p_t - log p_s)); nothing of
the kernels' algebra (shift by 1, second-order KL form, Kahan sums) is used.  Two passes over the row chunks: the first
collects the row and column log-sum-exps, the second the column-direction loss terms.  Gradients are produced for a SAMPLE
of image rows and text rows (full rows / columns of dL/dS are rebuilt from the stored log-sum-exps), which is what a
full-size parity test needs.  Pinned by tests/test_oracle_golden.py against oracle/closed_form.py, which is itself pinned
to fixtures generated from the reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

F64 = torch.float64


def _normalise(x: torch.Tensor):
    x = x.to(F64)
    r = 1.0 / x.norm(dim=1, keepdim=True)
    return x * r, r


def _dlogits(s, t, rows, cols, st, temperature, w):
    """dL/dS for the sub-matrix [rows, cols] (index tensors) of the logits, L = sum_k w_k loss_k with every loss already
    the 0.5 (i2t + t2i) mean of _loss.py:130-145.  `s`, `t`: that sub-matrix of the student / teacher logits."""
    b = st["b"]
    g = torch.zeros_like(s)
    diag = rows[:, None] == cols[None, :]
    if w.get("hard"):
        p_row = torch.exp(s - st["lse1_row"][rows][:, None])
        p_col = torch.exp(s - st["lse1_col"][cols][None, :])
        g += w["hard"] * 0.5 / b * (p_row + p_col - 2.0 * diag.to(F64))
    if w.get("soft"):
        T = temperature
        ps_r = torch.exp(s / T - st["lseT_s_row"][rows][:, None])
        pt_r = torch.exp(t / T - st["lseT_t_row"][rows][:, None])
        ps_c = torch.exp(s / T - st["lseT_s_col"][cols][None, :])
        pt_c = torch.exp(t / T - st["lseT_t_col"][cols][None, :])
        g += w["soft"] * 0.5 * T * ((ps_r - pt_r) + (ps_c - pt_c))
    if w.get("cos_diff"):
        # identical in both directions (the same element set): 0.5 (i2t + t2i) = one of them.  relu'(0) = 0 (ATen)
        off = (s > t).to(F64) / (b * (b - 1.0))
        on = -(t > s).to(F64) / b
        g += w["cos_diff"] * torch.where(diag, on, off)
    if w.get("logits_mse"):
        g += w["logits_mse"] * 2.0 * (s - t) / (float(b) * float(b))
    return g


def contrastive_chunked(stu_img, stu_txt, tea_img=None, tea_txt=None, temperature: Optional[float] = None,
                        weights: Optional[Dict[str, float]] = None, sample_img: Optional[Sequence[int]] = None,
                        sample_txt: Optional[Sequence[int]] = None, chunk: int = 2048, device=None) -> Dict:
    """-> dict(hard, soft, cos_diff, logits_mse [python floats, the un-weighted 0.5 (i2t + t2i) values],
               d_img [len(sample_img), D], d_txt [len(sample_txt), D] float64 tensors: gradients of
               sum_k weights[k] * loss_k w.r.t. the UN-normalised student rows in the samples).
    Inputs: [B, D] tensors of any float dtype (upcast to float64 first); teacher optional (hard only)."""
    dev = torch.device(device) if device is not None else stu_img.device
    w = dict(weights or {})
    has_t = tea_img is not None
    a_s, ra = _normalise(stu_img.to(dev))
    b_s, rb = _normalise(stu_txt.to(dev))
    if has_t:
        a_t, _ = _normalise(tea_img.to(dev))
        b_t, _ = _normalise(tea_txt.to(dev))
        T = float(temperature) if temperature else 1.0
    b = a_s.shape[0]
    ar = torch.arange(b, device=dev)
    neg_inf = float("-inf")
    st = {"b": b, "lse1_row": torch.empty(b, dtype=F64, device=dev), "lse1_col": torch.full((b,), neg_inf, dtype=F64, device=dev)}
    if has_t:
        for k in ("lseT_s_row", "lseT_t_row"):
            st[k] = torch.empty(b, dtype=F64, device=dev)
        for k in ("lseT_s_col", "lseT_t_col"):
            st[k] = torch.full((b,), neg_inf, dtype=F64, device=dev)
    acc = {k: torch.zeros((), dtype=F64, device=dev) for k in ("hard_row", "hard_col", "soft_row", "soft_col", "pos", "neg", "mse")}
    diag_s = torch.empty(b, dtype=F64, device=dev)
    # ---- pass 1: row terms, column log-sum-exps (running logaddexp over the chunks), element-wise losses
    for r0 in range(0, b, chunk):
        rows = ar[r0:r0 + chunk]
        s = a_s[rows] @ b_s.t()
        ds = s[torch.arange(len(rows), device=dev), rows]
        diag_s[rows] = ds
        lse = torch.logsumexp(s, dim=1)
        st["lse1_row"][rows] = lse
        acc["hard_row"] += (lse - ds).sum()
        st["lse1_col"] = torch.logaddexp(st["lse1_col"], torch.logsumexp(s, dim=0))
        if has_t:
            t = a_t[rows] @ b_t.t()
            ls, lt = torch.log_softmax(s / T, dim=1), torch.log_softmax(t / T, dim=1)
            st["lseT_s_row"][rows] = torch.logsumexp(s / T, dim=1)
            st["lseT_t_row"][rows] = torch.logsumexp(t / T, dim=1)
            acc["soft_row"] += (lt.exp() * (lt - ls)).sum()
            st["lseT_s_col"] = torch.logaddexp(st["lseT_s_col"], torch.logsumexp(s / T, dim=0))
            st["lseT_t_col"] = torch.logaddexp(st["lseT_t_col"], torch.logsumexp(t / T, dim=0))
            dt = t[torch.arange(len(rows), device=dev), rows]
            acc["pos"] += torch.relu(dt - ds).sum()
            acc["neg"] += torch.relu(s - t).sum() - torch.relu(ds - dt).sum()
            acc["mse"] += ((s - t) ** 2).sum()
    acc["hard_col"] = (st["lse1_col"] - diag_s).sum()
    # ---- pass 2: column-direction KL with the complete column log-sum-exps
    if has_t:
        for r0 in range(0, b, chunk):
            rows = ar[r0:r0 + chunk]
            s = a_s[rows] @ b_s.t()
            t = a_t[rows] @ b_t.t()
            ls = s / T - st["lseT_s_col"][None, :]
            lt = t / T - st["lseT_t_col"][None, :]
            acc["soft_col"] += (lt.exp() * (lt - ls)).sum()
    out = {"hard": float(0.5 * (acc["hard_row"] + acc["hard_col"]) / b)}
    if has_t:
        out["soft"] = float(0.5 * T * T * (acc["soft_row"] + acc["soft_col"]))
        out["cos_diff"] = float(acc["pos"] / b + acc["neg"] / (b * (b - 1.0))) if b > 1 else float("nan")
        out["logits_mse"] = float(acc["mse"] / (float(b) * float(b)))
    # ---- gradients of the sampled rows
    if sample_img is not None and len(sample_img):
        idx = torch.as_tensor(list(sample_img), device=dev, dtype=torch.long)
        s = a_s[idx] @ b_s.t()
        t = a_t[idx] @ b_t.t() if has_t else None
        g = _dlogits(s, t, idx, ar, st, T if has_t else None, w) @ b_s
        ah = a_s[idx]
        out["d_img"] = ra[idx] * (g - ah * (ah * g).sum(1, keepdim=True))
    if sample_txt is not None and len(sample_txt):
        idx = torch.as_tensor(list(sample_txt), device=dev, dtype=torch.long)
        g = torch.zeros(len(idx), a_s.shape[1], dtype=F64, device=dev)
        for r0 in range(0, b, 4 * chunk):
            rows = ar[r0:r0 + 4 * chunk]
            s = a_s[rows] @ b_s[idx].t()
            t = a_t[rows] @ b_t[idx].t() if has_t else None
            g += _dlogits(s, t, rows, idx, st, T if has_t else None, w).t() @ a_s[rows]
        bh = b_s[idx]
        out["d_txt"] = rb[idx] * (g - bh * (bh * g).sum(1, keepdim=True))
    return out
