from torch import nn

from ... import ops


class OutCosLoss(nn.Module):
    """nn.CosineEmbeddingLoss()(stu, tea, ones): mean over the batch of 1 - cos(stu_i, tea_i) -- reference out_cos.py:10-11
    (shipped in all three final configs).  Inputs are [B, D]."""

    def forward(self, stu_out, tea_out):
        return ops.stream_loss(ops.KIND_COS, [stu_out], [tea_out])
