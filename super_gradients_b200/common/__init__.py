from .registry import (  # noqa: F401
    ALL_DETECTION_MODULES,
    ARCHITECTURES,
    LOSSES,
    register_detection_module,
    register_loss,
    register_model,
)
