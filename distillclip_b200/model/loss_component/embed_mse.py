from torch import nn

from ... import ops


class EmbedMSELoss(nn.Module):
    """MSE(student embedding, teacher embedding) -- reference embed_mse.py:9-10."""

    def forward(self, stu_embedding, tea_embedding):
        return ops.stream_loss(ops.KIND_MSE, [stu_embedding], [tea_embedding])
