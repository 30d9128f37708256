from .yolo_nas_pose_dfl_head import YoloNASPoseDFLHead  # noqa: F401
from .yolo_nas_pose_ndfl_heads import YoloNASPoseNDFLHeads  # noqa: F401
from .yolo_nas_pose_post_prediction_callback import YoloNASPosePostPredictionCallback  # noqa: F401
from .yolo_nas_pose_variants import YoloNASPose, YoloNASPose_L, YoloNASPose_M, YoloNASPose_N, YoloNASPose_S  # noqa: F401
