from .cross_entropy import CrossEntropyLoss  # noqa: F401
from .ppyolo_loss import PPYoloELoss, pad_targets_host  # noqa: F401
