"""SPP context module (reference: training/models/detection_models/csp_darknet53.py:135-157)."""
from typing import Tuple, Type

from torch import nn

from .... import functional as SF
from ....common.factories import resolve_activation
from ....common.registry import register_detection_module
from ....modules import BaseDetectionModule, Conv


@register_detection_module()
class SPP(BaseDetectionModule):
    def __init__(self, in_channels, output_channels, k: Tuple, activation_type: Type[nn.Module]):
        super().__init__(in_channels)
        activation_type = resolve_activation(activation_type)
        self._output_channels = output_channels
        hidden_channels = in_channels // 2
        self.cv1 = Conv(in_channels, hidden_channels, 1, 1, activation_type)
        self.cv2 = Conv(hidden_channels * (len(k) + 1), output_channels, 1, 1, activation_type)
        self.m = nn.ModuleList([nn.MaxPool2d(kernel_size=x, stride=1, padding=x // 2) for x in k])  # parameter-free markers
        self._ks = tuple(k)

    def forward(self, x):
        x = self.cv1(x)
        return self.cv2(SF.concat([x] + [SF.max_pool(x, k, 1, k // 2) for k in self._ks]))

    @property
    def out_channels(self):
        return self._output_channels
