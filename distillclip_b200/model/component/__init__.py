from .output import CLIPOutput, ControlOutput, TextTransformerOutput, VisionTransformerOutput

__all__ = ["CLIPOutput", "ControlOutput", "TextTransformerOutput", "VisionTransformerOutput"]
