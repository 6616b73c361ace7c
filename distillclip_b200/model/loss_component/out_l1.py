from torch import nn

from ... import ops


class OutL1Loss(nn.Module):
    """nn.L1Loss(mean)(student output, teacher output) -- reference out_l1.py:9-10 (shipped in all three final configs)."""

    def forward(self, stu_out, tea_out):
        return ops.stream_loss(ops.KIND_L1, [stu_out], [tea_out])
