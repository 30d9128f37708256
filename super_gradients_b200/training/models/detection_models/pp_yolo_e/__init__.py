from .post_prediction_callback import DetectionPostPredictionCallback, PPYoloEPostPredictionCallback  # noqa: F401
