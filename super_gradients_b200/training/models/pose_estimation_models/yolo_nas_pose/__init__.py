from .yolo_nas_pose_post_prediction_callback import YoloNASPosePostPredictionCallback  # noqa: F401
