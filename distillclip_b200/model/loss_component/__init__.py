from .attention_probs_kl import AttentionProbsKL
from .attention_probs_mse import AttentionProbsMSE
from .attention_score_mse import AttentionScoreMSE
from .clip_cos_diff import CLIPCosDiff
from .embed_mse import EmbedMSELoss
from .hard_label import HardLabel
from .hidden_mse import HiddenMSE
from .last_value_map_kl import LastValueMapKL
from .logits_mse import LogitsMSE
from .out_ce import OutCELoss
from .out_cos import OutCosLoss
from .out_kl import OutKLLoss
from .out_l1 import OutL1Loss
from .soft_label import SoftLabel

__all__ = ["AttentionProbsKL", "AttentionProbsMSE", "AttentionScoreMSE", "CLIPCosDiff", "EmbedMSELoss", "HardLabel",
           "HiddenMSE", "LastValueMapKL", "LogitsMSE", "OutCELoss", "OutCosLoss", "OutKLLoss", "OutL1Loss", "SoftLabel"]
