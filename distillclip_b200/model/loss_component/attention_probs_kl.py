from torch import nn

from ... import ops


class AttentionProbsKL(nn.Module):
    """KL(sum) between head-averaged teacher and student attention maps, averaged over layers --
    reference attention_probs_kl.py:10-22 (student/teacher head counts may differ; empty lists raise
    ZeroDivisionError; coincident zeros give NaN, exactly like the reference)."""

    def forward(self, stu_attn_probs, tea_attn_probs):
        return ops.stream_loss(ops.KIND_ATTN_KL, stu_attn_probs, tea_attn_probs)
