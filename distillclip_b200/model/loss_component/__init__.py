from .attention_probs_kl import AttentionProbsKL
from .embed_mse import EmbedMSELoss
from .hard_label import HardLabel
from .hidden_mse import HiddenMSE
from .soft_label import SoftLabel

__all__ = ["AttentionProbsKL", "EmbedMSELoss", "HardLabel", "HiddenMSE", "SoftLabel"]
