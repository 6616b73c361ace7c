from torch import nn

from ... import ops


class AttentionProbsMSE(nn.Module):
    """MSE(mean) between head-averaged student and teacher attention maps, averaged over layers -- reference
    attention_probs_mse.py:10-22 (same list semantics as AttentionProbsKL: zip truncation, len(stu) divisor)."""

    def forward(self, stu_attn_probs, tea_attn_probs):
        return ops.stream_loss(ops.KIND_ATTN_MSE, stu_attn_probs, tea_attn_probs)
