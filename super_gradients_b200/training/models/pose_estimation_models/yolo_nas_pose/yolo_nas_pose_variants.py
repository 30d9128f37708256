"""YoloNASPose N / S / M / L (reference: pose_estimation_models/yolo_nas_pose/yolo_nas_pose_variants.py:93-406): the
YOLO-NAS backbone and neck with the pose heads; tensor-input predict() = forward + YoloNASPosePostPredictionCallback."""
import copy
from typing import Any, List, Optional, Tuple

import torch

from .....common.registry import register_model
from .....module_interfaces import PoseEstimationPredictions
from ....utils import HpmStruct, get_param
from ...arch_params_factory import get_arch_params
from ...detection_models.customizable_detector import CustomizableDetector
from .yolo_nas_pose_post_prediction_callback import YoloNASPosePostPredictionCallback


class YoloNASPoseDecodingModule(torch.nn.Module):
    """Export-time pre-NMS selection (reference: yolo_nas_pose_variants.py:27-90): the `num_pre_nms_predictions` most
    confident poses per image, in descending confidence -> (boxes [B, k, 4], scores [B, k, 1], joints [B, k, J, 3] = (x, y,
    joint confidence)).  Plain torch ops: this module exists for the ONNX / TensorRT export graph, not for the predict() path
    (which runs the batched NMS kernel on all anchors)."""

    def __init__(self, num_pre_nms_predictions: int = 1000):
        super().__init__()
        self.num_pre_nms_predictions = num_pre_nms_predictions

    def get_num_pre_nms_predictions(self) -> int:
        return self.num_pre_nms_predictions

    def forward(self, inputs):
        boxes, conf, coords, joint_scores = inputs if torch.jit.is_tracing() else inputs[0]
        idx = torch.topk(conf, dim=1, k=self.num_pre_nms_predictions, largest=True, sorted=True).indices  # [B, k, 1]
        joints = torch.cat([coords, joint_scores.unsqueeze(3)], dim=3)
        take = lambda t: torch.gather(t, 1, idx.reshape(idx.shape[0], -1, *([1] * (t.dim() - 2))).expand(-1, -1, *t.shape[2:]))  # noqa: E731
        return take(boxes), take(conf), take(joints)


class YoloNASPose(CustomizableDetector):
    def __init__(self, backbone, heads, neck=None, num_classes: int = None, bn_eps: Optional[float] = None, bn_momentum: Optional[float] = None, inplace_act: Optional[bool] = True, in_channels: int = 3):
        super().__init__(backbone, heads, neck, num_classes, bn_eps, bn_momentum, inplace_act, in_channels)
        self._edge_links = None
        self._edge_colors = None
        self._keypoint_colors = None
        self._default_nms_conf = None
        self._default_nms_iou = None
        self._default_pre_nms_max_predictions = None
        self._default_post_nms_max_predictions = None

    def get_decoding_module(self, num_pre_nms_predictions: int, **kwargs) -> torch.nn.Module:
        return YoloNASPoseDecodingModule(num_pre_nms_predictions)

    def get_post_prediction_callback(self, conf: float, iou: float, pre_nms_max_predictions=1000, post_nms_max_predictions=300) -> YoloNASPosePostPredictionCallback:
        return YoloNASPosePostPredictionCallback(pose_confidence_threshold=conf, nms_iou_threshold=iou, pre_nms_max_predictions=pre_nms_max_predictions, post_nms_max_predictions=post_nms_max_predictions)

    def set_dataset_processing_params(self, edge_links=None, edge_colors=None, keypoint_colors=None, image_processor=None, conf: Optional[float] = None, iou: Optional[float] = 0.7,
                                      pre_nms_max_predictions=300, post_nms_max_predictions=100) -> None:  # fmt: skip
        self._edge_links = edge_links or self._edge_links
        self._edge_colors = edge_colors or self._edge_colors
        self._keypoint_colors = keypoint_colors or self._keypoint_colors
        self._image_processor = image_processor or self._image_processor
        self._default_nms_conf = conf or self._default_nms_conf
        self._default_nms_iou = iou or self._default_nms_iou
        self._default_pre_nms_max_predictions = pre_nms_max_predictions or self._default_pre_nms_max_predictions
        self._default_post_nms_max_predictions = post_nms_max_predictions or self._default_post_nms_max_predictions

    @torch.no_grad()
    def predict(self, images: torch.Tensor, iou: Optional[float] = None, conf: Optional[float] = None, pre_nms_max_predictions: Optional[int] = None,
                post_nms_max_predictions: Optional[int] = None, batch_size: int = 32, fuse_model: bool = True) -> List[PoseEstimationPredictions]:  # fmt: skip
        """images: a pre-processed tensor [B, C, H, W], or raw images (one uint8 H x W x C array or a list of them, any sizes) that
        go through the model's image processor (default: the YOLO-NAS-POSE COCO chain) as one fused GPU launch per image, with
        poses and boxes returned in original-image pixels.  One PoseEstimationPredictions per image."""
        if not torch.is_tensor(images):
            import numpy as np

            from ....processing import default_yolo_nas_pose_coco_processing_params

            raw = [images] if isinstance(images, np.ndarray) else list(images)
            processor = self._image_processor or default_yolo_nas_pose_coco_processing_params()["image_processor"]
            device = next(self.parameters()).device
            out: List[PoseEstimationPredictions] = []
            cb = self.get_post_prediction_callback(
                conf=conf or self._default_nms_conf or 0.5, iou=iou or self._default_nms_iou or 0.7,
                pre_nms_max_predictions=pre_nms_max_predictions or self._default_pre_nms_max_predictions or 300,
                post_nms_max_predictions=post_nms_max_predictions or self._default_post_nms_max_predictions or 100,
            )  # fmt: skip
            was_training = self.training
            self.eval()
            for i in range(0, len(raw), batch_size):
                # one fused pre-processing launch per image, ONE model / decode / NMS pass per batch, the padding / rescaling of every
                # image undone by batched launches, ONE device -> host copy per result tensor; the per-image objects are host views
                # (the reference's pipeline returns host arrays as well: pipelines.py:445-452)
                batch, geos = processor.preprocess_batch(raw[i : i + batch_size], device)
                with torch.no_grad():
                    res = self(batch)
                    rows, poses, _idx, count = cb.forward_batched(res if self.heads.inference_mode is False else (res, None))
                    shift, scale = processor.batch_shift_scale(geos, device)
                    boxes = (rows[..., :4] + shift.repeat(1, 2)[:, None, :]) * scale.repeat(1, 2)[:, None, :]
                    poses = torch.cat([(poses[..., :2] + shift[:, None, None, :]) * scale[:, None, None, :], poses[..., 2:]], dim=-1)
                    boxes_h, poses_h, scores_h, counts = boxes.cpu(), poses.cpu(), rows[..., 4].cpu(), count.tolist()
                out += [PoseEstimationPredictions(poses=poses_h[b, :n], scores=scores_h[b, :n], bboxes_xyxy=boxes_h[b, :n]) for b, n in enumerate(counts)]
            self.train(was_training)
            return out
        cb = self.get_post_prediction_callback(
            conf=conf or self._default_nms_conf or 0.5, iou=iou or self._default_nms_iou or 0.7,
            pre_nms_max_predictions=pre_nms_max_predictions or self._default_pre_nms_max_predictions or 300,
            post_nms_max_predictions=post_nms_max_predictions or self._default_post_nms_max_predictions or 100,
        )  # fmt: skip
        was_training = self.training
        self.eval()
        out: List[PoseEstimationPredictions] = []
        for i in range(0, images.shape[0], batch_size):
            res = self(images[i : i + batch_size])
            out += cb(res if self.heads.inference_mode is False else (res, None))
        self.train(was_training)
        return out

    def get_input_shape_steps(self) -> Tuple[int, int]:
        return 32, 32

    def get_minimum_input_shape_size(self) -> Tuple[int, int]:
        return 32, 32

    @property
    def num_classes(self):
        return self.heads.num_classes


def _variant(arch_name: str):
    class _YoloNASPoseVariant(YoloNASPose):
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

    return _YoloNASPoseVariant


YoloNASPose_N = register_model("yolo_nas_pose_n")(type("YoloNASPose_N", (_variant("yolo_nas_pose_n_arch_params"),), {}))
YoloNASPose_S = register_model("yolo_nas_pose_s")(type("YoloNASPose_S", (_variant("yolo_nas_pose_s_arch_params"),), {}))
YoloNASPose_M = register_model("yolo_nas_pose_m")(type("YoloNASPose_M", (_variant("yolo_nas_pose_m_arch_params"),), {}))
YoloNASPose_L = register_model("yolo_nas_pose_l")(type("YoloNASPose_L", (_variant("yolo_nas_pose_l_arch_params"),), {}))
