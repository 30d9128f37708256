"""YoloNASPANNeckWithC2 (reference: training/models/detection_models/yolo_nas/panneck.py:11-64): two up stages (C5 -> P4-ish
-> P3) followed by two down stages (P3 -> P4 -> P5); each stage is built by the detection-module factory from its arch
params with the input channels of its actual producers injected."""
from typing import List, Tuple

from torch import Tensor

from .....common.factories import DetectionModulesFactory
from .....common.registry import register_detection_module
from .....modules import BaseDetectionModule


@register_detection_module("YoloNASPANNeckWithC2")
class YoloNASPANNeckWithC2(BaseDetectionModule):
    def __init__(self, in_channels: List[int], neck1, neck2, neck3, neck4):
        super().__init__(in_channels)
        c2, c3, c4, c5 = in_channels
        factory = DetectionModulesFactory()

        def build(params, sources):
            return factory.get(factory.insert_module_param(params, "in_channels", sources))

        self.neck1 = build(neck1, [c5, c4, c3])  # up: (skip-ready intermediate, output)
        up1_inter, up1_out = self.neck1.out_channels
        self.neck2 = build(neck2, [up1_out, c3, c2])
        up2_inter, up2_out = self.neck2.out_channels
        self.neck3 = build(neck3, [up2_out, up2_inter])  # down
        self.neck4 = build(neck4, [self.neck3.out_channels, up1_inter])
        self._out_channels = [up2_out, self.neck3.out_channels, self.neck4.out_channels]

    @property
    def out_channels(self):
        return self._out_channels

    def forward(self, inputs: Tuple[Tensor, Tensor, Tensor, Tensor]) -> Tuple[Tensor, Tensor, Tensor]:
        c2, c3, c4, c5 = inputs
        inter1, top = self.neck1([c5, c4, c3])
        inter2, p3 = self.neck2([top, c3, c2])
        p4 = self.neck3([p3, inter2])
        p5 = self.neck4([p4, inter1])
        return p3, p4, p5
