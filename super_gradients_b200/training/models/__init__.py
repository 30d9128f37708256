from .classification_models import resnet  # noqa: F401  (registers resnet18/34/50/101, resnet18_cifar)
from .detection_models import csp_darknet53, yolo_nas  # noqa: F401  (registers detection modules + yolo_nas_s/m/l)
from .detection_models.customizable_detector import CustomizableDetector  # noqa: F401
from . import pose_estimation_models  # noqa: F401  (registers the pose heads + yolo_nas_pose_n/s/m/l)
from .model_factory import get  # noqa: F401
from ...modules import detection_modules  # noqa: F401  (registers NStageBackbone)
