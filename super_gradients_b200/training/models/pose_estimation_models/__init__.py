from .yolo_nas_pose import YoloNASPose, YoloNASPose_L, YoloNASPose_M, YoloNASPose_N, YoloNASPose_S, YoloNASPoseDFLHead, YoloNASPoseNDFLHeads, YoloNASPosePostPredictionCallback  # noqa: F401
