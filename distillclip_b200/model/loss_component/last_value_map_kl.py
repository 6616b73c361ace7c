from torch import nn

from ... import ops


class LastValueMapKL(nn.Module):
    """KLDiv(sum)(softmax(stu_value_map, dim=1).log(), softmax(tea_value_map, dim=1)) -- reference
    last_value_map_kl.py:10-14; the softmax runs over the head axis of the [B, H, N, N] value-relation maps."""

    def forward(self, stu_value_map, tea_value_map):
        return ops.value_map_kl(stu_value_map, tea_value_map)
