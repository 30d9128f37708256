from .dfl_heads import NDFLHeads, YoloNASDFLHead  # noqa: F401
from .panneck import YoloNASPANNeckWithC2  # noqa: F401
from .yolo_nas_variants import YoloNAS, YoloNAS_L, YoloNAS_M, YoloNAS_S, YoloNASDecodingModule  # noqa: F401
from .yolo_stages import YoloNASBottleneck, YoloNASCSPLayer, YoloNASDownStage, YoloNASStage, YoloNASStem, YoloNASUpStage  # noqa: F401
