"""`LazyLogitsCLIP`: stops the caller's `CLIPModel` from materialising the B x B logit matrix.

The reference's `CLIPModel.forward` (reference model/component/clip_model.py:31-49) runs both encoders, L2-normalises the
pooled embeddings and writes `image_feature @ text_feature.t()` -- 4.3 GB in fp32 at B = 32768 -- only for the logit
losses to reduce it again.  Every logit loss of this package (`hard_label`, `soft_label`, `cos_diff`, `logits_mse`) starts
from `last_representation` inside the fused tcgen05 kernels, so the matrix is not needed.  This wrapper takes the
caller's own `CLIPModel` instance (encoders, weights and checkpoints stay the caller's -- nothing of that class is
re-implemented here), calls its `encode_image` / `encode_text` (reference clip_model.py:17-29) and returns a `CLIPOutput`
whose `i2t_logits` / `t2i_logits` are None.  Everything else is forwarded to the wrapped module.
"""
from typing import Optional

from torch import nn

from .output import CLIPOutput, ControlOutput


class LazyLogitsCLIP(nn.Module):
    def __init__(self, clip_model: nn.Module):
        super().__init__()
        for name in ("encode_image", "encode_text"):
            if not callable(getattr(clip_model, name, None)):
                raise TypeError(f"LazyLogitsCLIP wraps a CLIPModel-like module; {type(clip_model).__name__} has no {name}()")
        self.clip_model = clip_model

    def forward(self, text, image, control_output: Optional[ControlOutput] = None):
        if getattr(self.clip_model, "only_last_rep", False):
            # the (features, features, logits) tuple of reference clip_model.py:46-49 feeds the validation metrics; callers
            # that want those without the matrix use distillclip_b200.metrics.retrieval_metrics on the two features
            return self.clip_model(text, image, control_output)
        if control_output is None:
            control_output = ControlOutput()
        return CLIPOutput(visual_output=self.clip_model.encode_image(image, control_output),
                          text_output=self.clip_model.encode_text(text, control_output))

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(super().__getattr__("clip_model"), name)
