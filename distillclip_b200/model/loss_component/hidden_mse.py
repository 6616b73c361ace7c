from torch import nn

from ... import ops


class HiddenMSE(nn.Module):
    """Mean over layers of MSE(student hidden, teacher hidden) -- reference hidden_mse.py:9-17.

    Same constructor / forward signature; owns no parameters or buffers.  All layers run in one
    CUDA launch that also writes the student gradients (csrc/mse_stream.cu)."""

    def forward(self, stu_hidden, tea_hidden):
        return ops.stream_loss(ops.KIND_MSE, stu_hidden, tea_hidden)
