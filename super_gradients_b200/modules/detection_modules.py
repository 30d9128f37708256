"""NStageBackbone (reference: modules/detection_modules.py:36-111)."""
from typing import List

from ..common.factories import DetectionModulesFactory
from ..common.registry import register_detection_module
from .base_modules import BaseDetectionModule


@register_detection_module()
class NStageBackbone(BaseDetectionModule):
    """stem -> N stages -> context module; returns the outputs of the layers named in `out_layers`."""

    def __init__(self, in_channels: int, out_layers: List[str], stem, stages, context_module):
        super().__init__(in_channels)
        factory = DetectionModulesFactory()
        self.num_stages = len(stages)
        self.stem = factory.get(factory.insert_module_param(stem, "in_channels", in_channels))
        prev_channels = self.stem.out_channels
        for i in range(self.num_stages):
            new_stage = factory.get(factory.insert_module_param(stages[i], "in_channels", prev_channels))
            setattr(self, f"stage{i + 1}", new_stage)
            prev_channels = new_stage.out_channels
        if context_module is not None:
            self.context_module = factory.get(factory.insert_module_param(context_module, "in_channels", prev_channels))
        else:
            self.context_module = None
        self.out_layers = out_layers
        self._out_channels = [getattr(self, layer).out_channels for layer in self.out_layers]
        self._all_layers = ["stem"] + [f"stage{i}" for i in range(1, self.num_stages + 1)] + (["context_module"] if self.context_module is not None else [])

    @property
    def out_channels(self):
        return self._out_channels

    def forward(self, x):
        outputs = []
        for layer in self._all_layers:
            x = getattr(self, layer)(x)
            if layer in self.out_layers:
                outputs.append(x)
        return outputs

    def replace_input_channels(self, in_channels: int, compute_new_weights_fn=None):
        if not hasattr(self.stem, "replace_input_channels"):
            raise NotImplementedError(f"`{self.stem.__class__.__name__}` does not support `replace_input_channels`")
        self.stem.replace_input_channels(in_channels=in_channels, compute_new_weights_fn=compute_new_weights_fn)

    def get_input_channels(self) -> int:
        return self.stem.get_input_channels()
