from torch import nn

from ... import contrastive


class CLIPCosDiff(nn.Module):
    """mean relu(tea_ii - stu_ii) + mean_{i != j} relu(stu_ij - tea_ij) on materialised [B, B] logits -- reference
    clip_cos_diff.py:12-23 (`get_neg_element` = every off-diagonal entry); shipped in l_clip.yaml."""

    def forward(self, stu_logits, tea_logits):
        return contrastive.cos_diff_from_logits(stu_logits, tea_logits)
