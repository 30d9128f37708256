"""Reference: module_interfaces/pose_estimation_post_prediction_callback.py:12-37."""
import abc
import dataclasses
from typing import Any, List, Optional, Union

import numpy as np
from torch import Tensor


@dataclasses.dataclass
class PoseEstimationPredictions:
    """Pose predictions of one image: poses [N, K, 3] = (x, y, joint score), scores [N], bboxes_xyxy [N, 4] (or None)."""

    poses: Union[Tensor, np.ndarray]
    scores: Union[Tensor, np.ndarray]
    bboxes_xyxy: Optional[Union[Tensor, np.ndarray]]


class AbstractPoseEstimationPostPredictionCallback(abc.ABC):
    @abc.abstractmethod
    def __call__(self, predictions: Any) -> List[PoseEstimationPredictions]:
        ...
