from .base_modules import BaseDetectionModule  # noqa: F401
from .conv_bn_act_block import Conv, ConvBNAct  # noqa: F401
from .conv_bn_relu_block import ConvBNReLU  # noqa: F401
from .qarepvgg_block import QARepVGGBlock  # noqa: F401
from .skip_connections import Residual  # noqa: F401
from .utils import autopad, width_multiplier  # noqa: F401
