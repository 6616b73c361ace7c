"""Logit producer: mirror of the reference `CLIPModel` (reference model/component/clip_model.py:7-62) whose forward can
skip the B x B logit matrix.

The reference's `forward` L2-normalises the two pooled embeddings and materialises `image_feature @ text_feature.t()`
(`clip_model.py:36-44`) -- 4.3 GB in fp32 at B = 32768 -- only for `HardLabel` / `SoftLabel` to reduce it again.  The fused
kernels behind `LossCalculator` start from `last_representation`, so with `lazy_logits=True` this class returns the same
`CLIPOutput` with `i2t_logits = t2i_logits = None` and the matrix never exists.  With the default
`lazy_logits=False` it behaves exactly like the reference (needed for `cos_diff`, which reads the logits).

The encoders are the caller's modules (out of scope here); this class only composes them.
"""
from typing import Optional

from torch import nn

from .output import CLIPOutput, ControlOutput


class CLIPModel(nn.Module):
    def __init__(self, is_student: bool, image_encoder: nn.Module, text_encoder: nn.Module,
                 norm=False, only_last_rep=False, lazy_logits: bool = False):
        super().__init__()
        self.image_encoder = image_encoder
        self.text_encoder = text_encoder
        self.is_student = is_student
        self.norm = norm
        self.only_last_rep = only_last_rep
        self.lazy_logits = lazy_logits

    def encode_image(self, image, control_output: ControlOutput = None):
        if control_output is None:
            control_output = ControlOutput()
        if self.only_last_rep:
            return self.image_encoder(image, control_output).last_representation
        return self.image_encoder(image, control_output)

    def encode_text(self, text, control_output: ControlOutput = None):
        if control_output is None:
            control_output = ControlOutput()
        if self.only_last_rep:
            return self.text_encoder(text, control_output).last_representation
        return self.text_encoder(text, control_output)

    def forward(self, text, image, control_output: Optional[ControlOutput] = None):
        if control_output is None:
            control_output = ControlOutput()
        image_output = self.encode_image(image, control_output)
        text_output = self.encode_text(text, control_output)
        if not self.only_last_rep:
            if self.lazy_logits:        # the fused loss kernels normalise and contract tile by tile
                return CLIPOutput(visual_output=image_output, text_output=text_output)
            image_feature = image_output.last_representation / image_output.last_representation.norm(dim=1, keepdim=True)
            text_feature = text_output.last_representation / text_output.last_representation.norm(dim=1, keepdim=True)
            logits = image_feature @ text_feature.t()
            return CLIPOutput(visual_output=image_output, text_output=text_output, i2t_logits=logits, t2i_logits=logits.T)
        image_feature = image_output / image_output.norm(dim=1, keepdim=True)
        text_feature = text_output / text_output.norm(dim=1, keepdim=True)
        return image_feature, text_feature, image_feature @ text_feature.t()

    def init_layers_with_teacher(self, text_layer_map, image_layer_map, teacher_state_dict=None, init_type=None):
        self.image_encoder.init_layers_with_teacher(image_layer_map, teacher_state_dict, init_type)
        self.text_encoder.init_layers_with_teacher(text_layer_map, teacher_state_dict, init_type)

    def hyper_para(self):
        res = {}
        for k, v in self.image_encoder.hyper_para().items():
            res['image_' + k] = v
        for k, v in self.text_encoder.hyper_para().items():
            res['text_' + k] = v
        return res
