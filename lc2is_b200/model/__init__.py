from .text_patch import TextToPatch  # noqa: F401
from .loss import AuxiliaryLoss, ContrastiveLoss, NPairLoss  # noqa: F401
