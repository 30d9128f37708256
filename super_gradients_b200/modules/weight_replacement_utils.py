"""Input-channel surgery on a convolution (reference: modules/weight_replacement_utils.py:9-65), the leaf of every
`replace_input_channels` (ConvBNAct / Conv, ResNet.conv1, the detectors' backbones; `models.get(..., num_input_channels=...)` after
loading a checkpoint trained on 3 channels)."""
from typing import Callable, Optional

import torch
from torch import nn

__all__ = ["replace_conv2d_input_channels", "replace_conv2d_input_channels_with_random_weights"]


def replace_conv2d_input_channels(conv: nn.Conv2d, in_channels: int, fn: Optional[Callable[[nn.Conv2d, int], nn.Conv2d]] = None) -> nn.Module:
    return fn(conv, in_channels) if fn else replace_conv2d_input_channels_with_random_weights(conv=conv, in_channels=in_channels)


def replace_conv2d_input_channels_with_random_weights(conv: nn.Conv2d, in_channels: int) -> nn.Conv2d:
    """Same hyper-parameters, device and dtype; the filters keep their first min(old, new) input channels, extra channels (and the
    bias) are drawn from a normal distribution with the old filters' (bias') mean and standard deviation."""
    if in_channels % conv.groups != 0:
        raise ValueError(f"Incompatible number of input channels ({in_channels}) with the number of groups ({conv.groups})."
                         f"The number of input channels must be divisible by the number of groups.")  # fmt: skip
    new = nn.Conv2d(in_channels, conv.out_channels, kernel_size=conv.kernel_size, stride=conv.stride, padding=conv.padding, dilation=conv.dilation, groups=conv.groups,
                    bias=conv.bias is not None, device=conv.weight.device, dtype=conv.weight.dtype)  # fmt: skip
    with torch.no_grad():
        keep = min(in_channels, conv.in_channels)
        if in_channels > conv.in_channels:
            nn.init.normal_(new.weight[:, keep:], mean=conv.weight.mean().item(), std=conv.weight.std().item())
        new.weight[:, :keep] = conv.weight[:, :keep]
        if conv.bias is not None:
            nn.init.normal_(new.bias, mean=conv.bias.mean().item(), std=conv.bias.std().item())
    return new
