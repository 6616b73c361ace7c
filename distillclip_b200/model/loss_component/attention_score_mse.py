from torch import nn

from ... import ops


class AttentionScoreMSE(nn.Module):
    """MSE(mean) between head-averaged student and teacher pre-softmax attention scores, averaged over layers --
    reference attention_score_mse.py:10-22."""

    def forward(self, stu_attn_score, tea_attn_score):
        return ops.stream_loss(ops.KIND_ATTN_MSE, stu_attn_score, tea_attn_score)
