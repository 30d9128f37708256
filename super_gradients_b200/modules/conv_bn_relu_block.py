"""ConvBNReLU: the ReLU specialisation of ConvBNAct with the reference's keyword surface and state-dict keys
(reference: modules/conv_bn_relu_block.py:8-60), so recipes and checkpoints written for it keep working."""
from typing import Tuple, Union

from torch import nn

from .conv_bn_act_block import ConvBNAct

_IntOrPair = Union[int, Tuple[int, int]]


class ConvBNReLU(ConvBNAct):
    def __init__(self, in_channels: int, out_channels: int, kernel_size: _IntOrPair, stride: _IntOrPair = 1, padding: _IntOrPair = 0, dilation: _IntOrPair = 1,
                 groups: int = 1, bias: bool = True, padding_mode: str = "zeros", use_normalization: bool = True, eps: float = 1e-5, momentum: float = 0.1,
                 affine: bool = True, track_running_stats: bool = True, device=None, dtype=None, use_activation: bool = True, inplace: bool = False):  # fmt: skip
        conv = dict(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=stride, padding=padding, dilation=dilation, groups=groups,
                    bias=bias, padding_mode=padding_mode, device=device, dtype=dtype)  # fmt: skip
        norm = dict(use_normalization=use_normalization, eps=eps, momentum=momentum, affine=affine, track_running_stats=track_running_stats)
        act = dict(activation_type=nn.ReLU if use_activation else None, activation_kwargs={"inplace": inplace} if inplace else None)
        super().__init__(**conv, **norm, **act)
