"""Result / callback interfaces shared by the prediction paths (reference: module_interfaces/)."""
from .pose_estimation_post_prediction_callback import AbstractPoseEstimationPostPredictionCallback, PoseEstimationPredictions  # noqa: F401
