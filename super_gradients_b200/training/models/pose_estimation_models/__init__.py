from .yolo_nas_pose import YoloNASPosePostPredictionCallback  # noqa: F401
