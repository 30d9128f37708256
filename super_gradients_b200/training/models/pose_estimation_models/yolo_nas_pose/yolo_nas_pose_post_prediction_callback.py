"""YoloNASPosePostPredictionCallback (reference: training/models/pose_estimation_models/yolo_nas_pose/
yolo_nas_pose_post_prediction_callback.py:10-94): confidence threshold (>=) -> top-k (sorted) -> class-agnostic NMS ->
gather boxes / scores / poses, for the whole batch in ONE launch of csrc/nms.cu (the same kernel that serves
PPYoloEPostPredictionCallback, in its single-score, class-agnostic mode) instead of a Python loop over images calling
torchvision.ops.boxes.nms."""
from typing import List, Tuple

import torch
from torch import Tensor

from ..... import kernels as K
from .....module_interfaces import AbstractPoseEstimationPostPredictionCallback, PoseEstimationPredictions


class YoloNASPosePostPredictionCallback(AbstractPoseEstimationPostPredictionCallback):
    def __init__(self, pose_confidence_threshold: float, nms_iou_threshold: float, pre_nms_max_predictions: int, post_nms_max_predictions: int):
        if post_nms_max_predictions > pre_nms_max_predictions:
            raise ValueError("post_nms_max_predictions must be less than pre_nms_max_predictions")
        super().__init__()
        self.pose_confidence_threshold = pose_confidence_threshold
        self.nms_iou_threshold = nms_iou_threshold
        self.pre_nms_max_predictions = pre_nms_max_predictions
        self.post_nms_max_predictions = post_nms_max_predictions

    @torch.no_grad()
    def forward_batched(self, outputs) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """Device-resident result: rows [B, P, 6] (x1, y1, x2, y2, conf, 0), poses [B, P, K, 3], anchor index [B, P]
        (-1 past the count) and count [B], P = post_nms_max_predictions."""
        pred_bboxes_xyxy, pred_bboxes_conf, pred_pose_coords, pred_pose_scores = outputs[0]
        if pred_bboxes_conf.dim() == 2:
            pred_bboxes_conf = pred_bboxes_conf.unsqueeze(-1)
        rows, idx, count = K.batched_nms(
            pred_bboxes_xyxy, pred_bboxes_conf, self.pose_confidence_threshold, self.nms_iou_threshold, self.pre_nms_max_predictions,
            self.post_nms_max_predictions, multi_label=False, class_agnostic=True, thr_inclusive=True,
        )  # fmt: skip
        safe = idx.clamp_min(0).long()
        coords = torch.gather(pred_pose_coords.float(), 1, safe[:, :, None, None].expand(-1, -1, pred_pose_coords.shape[2], 2))
        jscore = torch.gather(pred_pose_scores.float(), 1, safe[:, :, None].expand(-1, -1, pred_pose_scores.shape[2]))
        poses = torch.cat([coords, jscore.unsqueeze(-1)], dim=-1)
        return rows, poses, idx, count

    @torch.no_grad()
    def __call__(self, outputs) -> List[PoseEstimationPredictions]:
        rows, poses, _idx, count = self.forward_batched(outputs)
        counts = count.tolist()  # the one device->host read of the post-processing
        return [PoseEstimationPredictions(poses=poses[b, :n], scores=rows[b, :n, 4], bboxes_xyxy=rows[b, :n, :4]) for b, n in enumerate(counts)]
