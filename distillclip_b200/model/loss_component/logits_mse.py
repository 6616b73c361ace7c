from torch import nn

from ... import ops


class LogitsMSE(nn.Module):
    """nn.MSELoss()(stu_logits, tea_logits) on materialised [B, B] logits -- reference logits_mse.py:9-10 (a `.T` view is
    made contiguous once; the value is transpose-invariant)."""

    def forward(self, stu_logits, tea_logits):
        return ops.stream_loss(ops.KIND_MSE, [stu_logits], [tea_logits])
