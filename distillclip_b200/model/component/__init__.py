from .clip_model import LazyLogitsCLIP
from .output import CLIPOutput, ControlOutput, TextTransformerOutput, VisionTransformerOutput

__all__ = ["LazyLogitsCLIP", "CLIPOutput", "ControlOutput", "TextTransformerOutput", "VisionTransformerOutput"]
