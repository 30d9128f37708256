"""CrossEntropyLoss returning (loss, loss_item) like the reference's label_smoothing_cross_entropy_loss.py:86-111.
Classification logits are tiny ([N, num_classes] fp32): this stays a torch op (not on the section-8 kernel list)."""
import torch
from torch import nn

from ...common.registry import register_loss


@register_loss(name="CrossEntropyLoss", deprecated_name="cross_entropy")
class CrossEntropyLoss(nn.CrossEntropyLoss):
    def __init__(self, weight=None, ignore_index=-100, reduction="mean", smooth_eps=None, smooth_dist=None, from_logits=True):
        super().__init__(weight=weight, ignore_index=ignore_index, reduction=reduction, label_smoothing=float(smooth_eps or 0.0))
        if smooth_dist is not None or not from_logits:
            raise NotImplementedError("smooth_dist / from_logits=False are not implemented")

    def forward(self, input, target):
        loss = super().forward(input.float(), target)
        return loss, loss.unsqueeze(0).detach()
