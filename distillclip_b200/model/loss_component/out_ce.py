from torch import nn

from ... import ops


class OutCELoss(nn.Module):
    """CrossEntropy(mean)(stu_out, softmax(tea_out, dim=1)) with soft targets, [B, D] -- reference out_ce.py:9-13."""

    def forward(self, stu_out, tea_out):
        return ops.row_softmax_loss(stu_out, tea_out, None, 1)
