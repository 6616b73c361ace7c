from torch import nn

from ... import contrastive


class SoftLabel(nn.Module):
    """T^2 * KL(sum)(softmax(tea/T) || softmax(stu/T)) on materialised logits -- reference soft_label.py:11-16."""

    def __init__(self, temperature):
        super().__init__()
        self.temperature = temperature

    def forward(self, stu_logits, tea_logits):
        return contrastive.soft_label_from_logits(stu_logits, tea_logits, self.temperature)
