"""YoloXPostPredictionCallback (reference: training/models/detection_models/yolo_base.py:74-140) over the batched NMS kernel.
Only the iterative NMS it defaults to is on the path; matrix NMS (YoloX-only, detection_utils.py:337-390) is not implemented."""
from typing import List, Optional, Tuple, Union

import torch
from torch import Tensor

from ...utils.detection_utils import non_max_suppression
from .pp_yolo_e.post_prediction_callback import DetectionPostPredictionCallback


class YoloXPostPredictionCallback(DetectionPostPredictionCallback):
    def __init__(self, conf: float = 0.001, iou: float = 0.6, classes: List[int] = None, nms_type: str = "iterative", max_predictions: int = 300, with_confidence: bool = True,
                 class_agnostic_nms: bool = False, multi_label_per_box: bool = True):  # fmt: skip
        super().__init__()
        if str(getattr(nms_type, "value", nms_type)).lower() != "iterative":
            raise NotImplementedError("matrix NMS is not implemented (only NMS_Type.ITERATIVE is on the hot path)")
        self.conf, self.iou, self.classes, self.max_pred = conf, iou, classes, max_predictions
        self.with_confidence, self.class_agnostic_nms, self.multi_label_per_box = with_confidence, class_agnostic_nms, multi_label_per_box

    @torch.no_grad()
    def forward(self, x: Union[Tensor, Tuple[Tensor, List[Tensor]]], device: str = None) -> List[Optional[Tensor]]:
        """x (or x[0]): [B, A, 5 + C] rows (cx, cy, w, h, objectness, class scores) -> per image [n <= max_predictions, 6]."""
        if isinstance(x, (tuple, list)):
            x = x[0]
        res = non_max_suppression(x, conf_thres=self.conf, iou_thres=self.iou, with_confidence=self.with_confidence, multi_label_per_box=self.multi_label_per_box,
                                  class_agnostic_nms=self.class_agnostic_nms)  # fmt: skip
        return [im[: self.max_pred] if (im is not None and im.shape[0] > self.max_pred) else im for im in res]
