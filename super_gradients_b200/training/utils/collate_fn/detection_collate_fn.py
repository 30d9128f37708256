"""The producer of the flat detection target tensor the loss / metric rows consume (reference:
training/utils/collate_fn/detection_collate_fn.py:10-49): per-sample (image, targets [n, 5+]) -> (images [B, C, H, W] float,
targets [N, 6+] with the sample index prepended)."""
from typing import List, Tuple, Union

import numpy as np
import torch

from ....common.registry import register_collate_function


class DatasetItemsException(Exception):
    def __init__(self, data_sample, collate_type, expected_item_names):
        n = len(data_sample) if hasattr(data_sample, "__len__") else "?"
        super().__init__(f"`{collate_type.__name__}` only supports Datasets that return a tuple {expected_item_names}, but got a tuple of len={n}")


@register_collate_function()
class DetectionCollateFN:
    def __init__(self):
        self.expected_item_names = ("image", "targets")

    def __call__(self, data) -> Tuple[torch.Tensor, torch.Tensor]:
        try:
            images_batch, labels_batch = list(zip(*data))
        except (ValueError, TypeError):
            raise DatasetItemsException(data_sample=data[0], collate_type=type(self), expected_item_names=self.expected_item_names)
        return self._format_images(images_batch), self._format_targets(labels_batch)

    @staticmethod
    def _format_images(images_batch: List[Union[torch.Tensor, np.ndarray]]) -> torch.Tensor:
        stack = torch.stack([torch.as_tensor(img) for img in images_batch], 0)
        return torch.moveaxis(stack, -1, 1).float() if stack.shape[3] == 3 else stack  # HWC samples -> float NCHW

    @staticmethod
    def _format_targets(labels_batch: List[Union[torch.Tensor, np.ndarray]]) -> torch.Tensor:
        rows = []
        for i, labels in enumerate(labels_batch):
            labels = torch.as_tensor(labels)
            rows.append(torch.cat((labels.new_full((labels.shape[0], 1), i), labels), dim=-1))
        return torch.cat(rows, 0)
