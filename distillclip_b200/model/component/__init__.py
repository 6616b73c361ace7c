from .clip_model import CLIPModel
from .output import CLIPOutput, ControlOutput, TextTransformerOutput, VisionTransformerOutput

__all__ = ["CLIPModel", "CLIPOutput", "ControlOutput", "TextTransformerOutput", "VisionTransformerOutput"]
