"""The logit-producer mirror (reference model/component/clip_model.py:31-49): same outputs as the reference formula,
and no B x B matrix when lazy."""
import numpy as np
import torch
from torch import nn

from conftest import golden
from distillclip_b200.model import CLIPModel, CLIPOutput, ControlOutput, TextTransformerOutput, VisionTransformerOutput


class _Tower(nn.Module):
    def __init__(self, cls):
        super().__init__()
        self.cls = cls

    def forward(self, x, control_output):
        assert isinstance(control_output, ControlOutput)
        return self.cls(last_representation=x)


def test_forward_matches_reference_logits():
    g = golden("clip_b24_d32_t2")            # i2t_logits_f32 was produced by the reference CLIPModel.forward
    img, txt = torch.tensor(g["stu_img"]), torch.tensor(g["stu_txt"])
    model = CLIPModel(True, _Tower(VisionTransformerOutput), _Tower(TextTransformerOutput))
    out = model(txt, img)
    assert isinstance(out, CLIPOutput)
    assert np.allclose(out.i2t_logits.numpy(), g["i2t_logits_f32"], rtol=0, atol=1e-6)
    assert out.t2i_logits.data_ptr() == out.i2t_logits.data_ptr() and out.t2i_logits.shape == out.i2t_logits.T.shape
    assert out.visual_output.last_representation is img and out.text_output.last_representation is txt


def test_lazy_forward_builds_no_logits():
    img, txt = torch.randn(6, 8), torch.randn(6, 8)
    out = CLIPModel(True, _Tower(VisionTransformerOutput), _Tower(TextTransformerOutput), lazy_logits=True)(txt, img)
    assert out.i2t_logits is None and out.t2i_logits is None
    assert out.visual_output.last_representation is img


def test_only_last_rep_path():
    img, txt = torch.randn(5, 8), torch.randn(5, 8)
    f_i, f_t, logits = CLIPModel(False, _Tower(VisionTransformerOutput), _Tower(TextTransformerOutput), only_last_rep=True)(txt, img)
    assert torch.allclose(f_i.norm(dim=1), torch.ones(5)) and torch.allclose(logits, f_i @ f_t.t())
