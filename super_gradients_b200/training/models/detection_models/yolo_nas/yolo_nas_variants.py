"""YoloNAS S / M / L (reference: training/models/detection_models/yolo_nas/yolo_nas_variants.py:75-212) and the
export-time pre-NMS top-k decoding module (:24-72)."""
import copy
from typing import Any, Optional, Tuple

import torch
from torch import Tensor, nn

from .....common.registry import register_model
from ....utils import HpmStruct, get_param
from ...arch_params_factory import get_arch_params
from ..customizable_detector import CustomizableDetector
from ..pp_yolo_e import PPYoloEPostPredictionCallback


class YoloNASDecodingModule(nn.Module):
    """max-class confidence -> top-k -> gather (yolo_nas_variants.py:53-72)."""

    def __init__(self, num_pre_nms_predictions: int = 1000):
        super().__init__()
        self.num_pre_nms_predictions = num_pre_nms_predictions

    def get_num_pre_nms_predictions(self) -> int:
        return self.num_pre_nms_predictions

    def forward(self, inputs: Tuple[Tuple[Tensor, Tensor], Tuple[Tensor, ...]]):
        pred_bboxes, pred_scores = inputs[0] if not torch.is_tensor(inputs[0]) else inputs
        nms_top_k = self.num_pre_nms_predictions
        pred_cls_conf, _ = torch.max(pred_scores, dim=2)
        idx = torch.topk(pred_cls_conf, dim=1, k=nms_top_k, largest=True, sorted=True).indices
        return torch.gather(pred_bboxes, 1, idx.unsqueeze(-1).expand(-1, -1, 4)), torch.gather(pred_scores, 1, idx.unsqueeze(-1).expand(-1, -1, pred_scores.size(2)))


class YoloNAS(CustomizableDetector):
    def __init__(self, backbone, heads, neck=None, num_classes: int = None, bn_eps: Optional[float] = None, bn_momentum: Optional[float] = None, inplace_act: Optional[bool] = True, in_channels: int = 3):
        super().__init__(backbone, heads, neck, num_classes, bn_eps, bn_momentum, inplace_act, in_channels)

    def get_post_prediction_callback(self, *, conf: float, iou: float, nms_top_k: int, max_predictions: int, multi_label_per_box: bool, class_agnostic_nms: bool) -> PPYoloEPostPredictionCallback:
        return PPYoloEPostPredictionCallback(score_threshold=conf, nms_threshold=iou, nms_top_k=nms_top_k, max_predictions=max_predictions, multi_label_per_box=multi_label_per_box, class_agnostic_nms=class_agnostic_nms)

    def get_decoding_module(self, num_pre_nms_predictions: int, **kwargs) -> nn.Module:
        return YoloNASDecodingModule(num_pre_nms_predictions)

    def get_input_shape_steps(self) -> Tuple[int, int]:
        return 32, 32

    def get_minimum_input_shape_size(self) -> Tuple[int, int]:
        return 32, 32

    @property
    def num_classes(self):
        return self.heads.num_classes


def _variant(arch_name: str):
    class _YoloNASVariant(YoloNAS):
        def __init__(self, arch_params: Any):
            default_arch_params = get_arch_params(arch_name)
            merged = HpmStruct(**copy.deepcopy(default_arch_params))
            merged.override(**(arch_params.to_dict() if hasattr(arch_params, "to_dict") else dict(arch_params or {})))
            super().__init__(
                backbone=merged.backbone,
                neck=merged.neck,
                heads=merged.heads,
                num_classes=get_param(merged, "num_classes", None),
                in_channels=get_param(merged, "in_channels", 3),
                bn_momentum=get_param(merged, "bn_momentum", None),
                bn_eps=get_param(merged, "bn_eps", None),
                inplace_act=get_param(merged, "inplace_act", None),
            )

    return _YoloNASVariant


YoloNAS_S = register_model("yolo_nas_s")(type("YoloNAS_S", (_variant("yolo_nas_s_arch_params"),), {}))
YoloNAS_M = register_model("yolo_nas_m")(type("YoloNAS_M", (_variant("yolo_nas_m_arch_params"),), {}))
YoloNAS_L = register_model("yolo_nas_l")(type("YoloNAS_L", (_variant("yolo_nas_l_arch_params"),), {}))
