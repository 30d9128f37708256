"""The producer of YoloNASPoseLoss's target triple (reference:
training/datasets/pose_estimation_datasets/yolo_nas_pose_collate_fn.py:14-123): samples -> (images [B, 3, H, W],
(boxes [N, 1+4] xyxy, joints [N, J, 1+3], is_crowd [N, 1+1]) each with the sample index prepended, extras)."""
from typing import Dict, List, Tuple

import numpy as np
import torch
from torch import Tensor
from torch.utils.data.dataloader import default_collate

from ....common.registry import register_collate_function

__all__ = ["YoloNASPoseCollateFN", "undo_flat_collate_tensors_with_batch_index", "flat_collate_tensors_with_batch_index"]


def flat_collate_tensors_with_batch_index(labels_batch: List[Tensor]) -> Tensor:
    """[n_i, ..., d] tensors -> [sum n_i, ..., 1 + d] with the sample index as the first element of the last dimension."""
    rows = [torch.cat((labels.new_full(labels.shape[:-1] + (1,), i), labels), dim=-1) for i, labels in enumerate(labels_batch)]
    return torch.cat(rows, 0)


def undo_flat_collate_tensors_with_batch_index(flat_tensor: Tensor, batch_size: int) -> List[Tensor]:
    index = flat_tensor[(slice(None),) + (0,) * (flat_tensor.ndim - 1)]
    return [flat_tensor[index == i][..., 1:] for i in range(batch_size)]


@register_collate_function()
class YoloNASPoseCollateFN:
    """Samples are objects with `image` (HWC numpy), `mask`, `bboxes_xywh` [n, 4], `joints` [n, J, 3], `is_crowd` [n] or None and
    `additional_samples` (the reference's PoseEstimationSample)."""

    def __init__(self, set_image_to_none: bool = True):
        self.set_image_to_none = set_image_to_none

    def __call__(self, batch) -> Tuple[Tensor, Tuple[Tensor, Tensor, Tensor], Dict]:
        images, boxes, joints, crowd = [], [], [], []
        for sample in batch:
            b, j, c = self._get_targets(sample)
            boxes.append(b)
            joints.append(j)
            crowd.append(c)
            sample.image = torch.from_numpy(np.transpose(sample.image, [2, 0, 1]))
            sample.mask = torch.from_numpy(sample.mask)
            images.append(sample.image)
            if self.set_image_to_none:
                sample.image = sample.mask = None
            sample.additional_samples = None
        return default_collate(images), (flat_collate_tensors_with_batch_index(boxes), flat_collate_tensors_with_batch_index(joints), flat_collate_tensors_with_batch_index(crowd)), {"gt_samples": batch}  # fmt: skip

    def _get_targets(self, sample) -> Tuple[Tensor, Tensor, Tensor]:
        if sample.image.shape[:2] != sample.mask.shape[:2]:
            raise ValueError(f"Image and mask should have the same shape {sample.image.shape[:2]} != {sample.mask.shape[:2]}")
        xywh = np.asarray(sample.bboxes_xywh)
        xyxy = np.concatenate([xywh[..., :2], xywh[..., :2] + xywh[..., 2:4]], axis=-1)
        is_crowd = np.zeros(len(xyxy)) if sample.is_crowd is None else sample.is_crowd
        return torch.from_numpy(xyxy), torch.from_numpy(sample.joints), torch.from_numpy(is_crowd.astype(int).reshape((-1, 1)))
