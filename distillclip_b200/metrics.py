"""Validation metrics of the dual-tower stage from the embeddings, without materialising the B x B logits.

Mirrors what `DualDistillModel.validation_step / validation_epoch_end` log for one pair of embedding matrices
(reference model/dual_distill_model.py:129-187): `norm_and_logits` (:271-275) followed by
  * `log_acc`        (:220-224)  top-k accuracy of logits against label arange(B) for k in k_list
  * `log_diag_score` (:204-211)  mean of diag(softmax(logits, 1)) and mean of diag(logits)
The similarity tiles come from the fused forward kernel (hard-label instantiation: A_i = sum_j exp(S_ij - 1) and S_ii),
a second pass over the same tiles counts rank_i = #{j : S_ij > S_ii}; the label is in the top k iff rank_i < k.
fp32 embeddings (the reference calls `.float()` first) are split into bf16 high and low parts and run as ONE GEMM over
the concatenated dimension [hi | lo | hi] x [hi | hi | lo] (a_hi.b_hi + a_lo.b_hi + a_hi.b_lo: ~16 mantissa bits).
With a process group every rank scores its own image rows against all text rows and the sums are all-reduced (the
reference all-gathers and lets every rank compute everything).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from . import contrastive as ct
from . import ops

K_LIST = (1, 3, 5, 10, 20, 50)            # reference dual_distill_model.py:87


def _split_fp32(x: torch.Tensor, first_lo: bool):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo, hi] if first_lo else [hi, hi, lo], dim=1).contiguous()


def retrieval_metrics(img: torch.Tensor, txt: torch.Tensor, k_list: Sequence[int] = K_LIST, group=None,
                      prefix: str = "", engine=None) -> Dict[str, torch.Tensor]:
    """-> {f'{prefix}acc_top{k}', f'{prefix}softmax_mean_score', f'{prefix}mean_score'} as 0-dim fp32 CUDA tensors for
    logits = normalise(img) @ normalise(txt).T of the GLOBAL batch (rows of all ranks in rank order)."""
    if engine is None:                       # product path: CUDA only; tests pass a CPU engine double for the sharding logic
        ops._require_cuda(img, "image embeddings")
        ops._require_cuda(txt, "text embeddings")
        engine = ct._ENGINE
    if img.dim() != 2 or img.shape != txt.shape:
        raise ValueError(f"expected equal [B, D] embeddings, got {tuple(img.shape)} and {tuple(txt.shape)}")
    img, txt = img.detach(), txt.detach()
    rank, world = ct._shard_info(group)
    b_local = img.shape[0]
    if img.dtype == torch.float32:
        a_inv = (1.0 / img.norm(dim=1)).contiguous()
        b_inv = (1.0 / txt.norm(dim=1)).contiguous()
        a, b = _split_fp32(img, True), _split_fp32(txt, False)
    else:
        if img.is_cuda:
            ops.dtype_code(img)
        a, b = img.contiguous(), txt.to(img.dtype).contiguous()
        a_inv, b_inv = engine.inv_norms([a, b])
    if a.shape[1] % 8:
        raise ValueError("embedding dimension must be a multiple of 8")
    b_all, b_inv_all = ct._all_gather_rows(b, group, world), ct._all_gather_rows(b_inv, group, world)
    offset = rank * b_local
    stats, _ = engine.row_stats(a, b_all, None, None, a_inv, b_inv_all, None, None, offset, None)
    diag = stats[4].contiguous()
    ranks = engine.rank_counts(a, b_all, a_inv, b_inv_all, offset, diag)
    ks = torch.tensor(list(k_list), dtype=torch.float32, device=img.device)
    sums = torch.cat([(ranks[None, :] < ks[:, None]).sum(1).to(torch.float64),
                      (torch.exp(diag.double() - 1.0) / stats[0].double()).sum().reshape(1), diag.double().sum().reshape(1)])
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(sums, group=group)
    out = (sums / float(b_local * world)).to(torch.float32)
    res = {f"{prefix}acc_top{k}": out[i] for i, k in enumerate(k_list)}
    res[f"{prefix}softmax_mean_score"] = out[len(k_list)]
    res[f"{prefix}mean_score"] = out[len(k_list) + 1]
    return res
