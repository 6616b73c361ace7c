"""Wire format between encoders and the loss stack.

Field names follow the reference dataclasses (reference model/component/output.py:7-35,63-68) because the
loss stack reads them by name: `last_representation`, `attention_probs`, `representations`, `embedding`,
`i2t_logits`, `t2i_logits`, ...  Only the types that cross the loss boundary are defined here; the
encoder-internal ones (AttentionOutput, TransformerLayerOutput, ...) belong to the encoders, which are
out of scope.

Reference quirk, not reproduced: `value_map` / `embedding` / `last_layer_output` default to the tuple
`(None,)` there because of trailing commas; here they default to None.
"""
from dataclasses import dataclass
from typing import List, Optional

import torch

__all__ = ["ControlOutput", "VisionTransformerOutput", "TextTransformerOutput", "CLIPOutput"]


@dataclass
class ControlOutput:
    """Which optional tensors the encoders should return (requested by LossCalculator.get_control_output)."""
    need_emb: bool = False
    need_attn_score: bool = False
    need_value_map: bool = False
    need_attn_prob: bool = False
    need_rep: bool = False


@dataclass
class _TowerOutput:
    last_representation: Optional[torch.Tensor] = None      # [B, D] pooled embedding (un-normalised)
    last_layer_output: Optional[torch.Tensor] = None        # [B, N, W]
    attention_scores: Optional[List[torch.Tensor]] = None   # L x [B, H, N, N] pre-softmax
    attention_probs: Optional[List[torch.Tensor]] = None    # L x [B, H, N, N]
    representations: Optional[List[torch.Tensor]] = None    # L x [B, N, W]
    value_map: Optional[torch.Tensor] = None
    embedding: Optional[torch.Tensor] = None                # [B, N, W]


@dataclass
class VisionTransformerOutput(_TowerOutput):
    pass


@dataclass
class TextTransformerOutput(_TowerOutput):
    pass


@dataclass
class CLIPOutput:
    visual_output: Optional[VisionTransformerOutput] = None
    text_output: Optional[TextTransformerOutput] = None
    i2t_logits: Optional[torch.Tensor] = None
    t2i_logits: Optional[torch.Tensor] = None
