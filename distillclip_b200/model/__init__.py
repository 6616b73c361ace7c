"""Mirror of the reference's `model` package for the loss hot path (same import paths below `model`)."""
from ._loss import IMAGE_TEXT_LOSS, LOSSNAME, LossCalculator
from .component.clip_model import LazyLogitsCLIP
from .component.output import CLIPOutput, ControlOutput, TextTransformerOutput, VisionTransformerOutput
from .loss_component import (AttentionProbsKL, AttentionProbsMSE, AttentionScoreMSE, CLIPCosDiff, EmbedMSELoss, HardLabel,
                             HiddenMSE, LastValueMapKL, LogitsMSE, OutCELoss, OutCosLoss, OutKLLoss, OutL1Loss, SoftLabel)

__all__ = ["LossCalculator", "LOSSNAME", "IMAGE_TEXT_LOSS", "LazyLogitsCLIP", "CLIPOutput", "ControlOutput",
           "TextTransformerOutput", "VisionTransformerOutput", "AttentionProbsKL", "AttentionProbsMSE", "AttentionScoreMSE",
           "CLIPCosDiff", "EmbedMSELoss", "HardLabel", "HiddenMSE", "LastValueMapKL", "LogitsMSE", "OutCELoss", "OutCosLoss", "OutKLLoss", "OutL1Loss",
           "SoftLabel"]
