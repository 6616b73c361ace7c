from torch import nn

from ... import ops


class OutKLLoss(nn.Module):
    """KLDiv(sum)(log_softmax(stu/T, dim=1), softmax(tea/T, dim=1)) * T^2 on the pooled outputs [B, D] --
    reference out_kl.py:6-16 (constructor argument `t` = temperature)."""

    def __init__(self, t):
        super().__init__()
        self.temperature = t

    def forward(self, stu_out, tea_out):
        return ops.row_softmax_loss(stu_out, tea_out, self.temperature, 0)
