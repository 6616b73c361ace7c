"""`LazyLogitsCLIP` around a caller-owned CLIPModel stand-in with the reference's interface (reference
model/component/clip_model.py:7-62: `encode_image`, `encode_text`, `forward(text, image, control_output)`): same tower
outputs, no B x B matrix, everything else forwarded."""
import pytest
import torch
from torch import nn

from distillclip_b200.model import (CLIPOutput, ControlOutput, LazyLogitsCLIP, TextTransformerOutput, VisionTransformerOutput)


class _Tower(nn.Module):
    def __init__(self, cls):
        super().__init__()
        self.cls = cls
        self.seen = []

    def forward(self, x, control_output):
        assert isinstance(control_output, ControlOutput)
        self.seen.append(control_output)
        return self.cls(last_representation=x)


class _CallerCLIP(nn.Module):
    """What the caller already has: builds the logits in forward, like the reference class."""

    def __init__(self, only_last_rep=False):
        super().__init__()
        self.image_encoder, self.text_encoder = _Tower(VisionTransformerOutput), _Tower(TextTransformerOutput)
        self.only_last_rep = only_last_rep
        self.forward_calls = 0

    def encode_image(self, image, control_output=None):
        out = self.image_encoder(image, control_output or ControlOutput())
        return out.last_representation if self.only_last_rep else out

    def encode_text(self, text, control_output=None):
        out = self.text_encoder(text, control_output or ControlOutput())
        return out.last_representation if self.only_last_rep else out

    def forward(self, text, image, control_output=None):
        self.forward_calls += 1
        i, t = self.encode_image(image, control_output), self.encode_text(text, control_output)
        return i, t, i @ t.t()

    def hyper_para(self):
        return {"image_width": 8}


def test_lazy_wrapper_builds_no_logits_and_passes_outputs_through():
    img, txt = torch.randn(6, 8), torch.randn(6, 8)
    inner = _CallerCLIP()
    co = ControlOutput()
    out = LazyLogitsCLIP(inner)(txt, img, co)
    assert isinstance(out, CLIPOutput)
    assert out.i2t_logits is None and out.t2i_logits is None
    assert out.visual_output.last_representation is img and out.text_output.last_representation is txt
    assert inner.forward_calls == 0                       # the caller's forward (and its matmul) never ran
    assert inner.image_encoder.seen == [co] and inner.text_encoder.seen == [co]


def test_lazy_wrapper_forwards_everything_else():
    inner = _CallerCLIP()
    lazy = LazyLogitsCLIP(inner)
    assert lazy.hyper_para() == {"image_width": 8}
    assert lazy.image_encoder is inner.image_encoder
    assert set(lazy.state_dict()) == {"clip_model." + k for k in inner.state_dict()}
    with pytest.raises(TypeError):
        LazyLogitsCLIP(nn.Linear(2, 2))


def test_only_last_rep_delegates_to_the_wrapped_forward():
    img, txt = torch.randn(5, 8), torch.randn(5, 8)
    inner = _CallerCLIP(only_last_rep=True)
    f_i, f_t, logits = LazyLogitsCLIP(inner)(txt, img)
    assert inner.forward_calls == 1 and torch.equal(logits, img @ txt.t())
