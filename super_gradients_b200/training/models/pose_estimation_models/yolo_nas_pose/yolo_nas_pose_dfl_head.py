"""YoloNASPoseDFLHead with the reference's names, constructor and state-dict keys
(training/models/pose_estimation_models/yolo_nas_pose/yolo_nas_pose_dfl_head.py:15-172): single-class detection head
(stem, cls / reg towers, 1x1 predictions) plus the keypoint tower on one pyramid level.  Every convolution is a fused
libsgb200 GEMM; the head returns bf16 NHWC maps and leaves decoding to YoloNASPoseNDFLHeads."""
import math
from functools import partial
from typing import Tuple

import torch
from torch import Tensor, nn

from ..... import functional as SF
from .....common.registry import register_detection_module
from .....modules import BaseDetectionModule, ConvBNReLU, QARepVGGBlock
from .....modules.utils import width_multiplier


class _PlainConv1x1(nn.Conv2d):
    """nn.Conv2d(k=1, bias=False) of the shared-stem variant, executed by the fused conv kernel."""

    def __init__(self, cin, cout):
        super().__init__(cin, cout, kernel_size=1, stride=1, padding=0, bias=False)
        self._cache = SF.WeightCache()

    def forward(self, x):
        return SF.conv_bias(x, self.weight, None, stride=1, pad=0, cache=self._cache)


@register_detection_module()
class YoloNASPoseDFLHead(BaseDetectionModule):
    def __init__(self, in_channels: int, bbox_inter_channels: int, pose_inter_channels: int, pose_regression_blocks: int, shared_stem: bool, pose_conf_in_class_head: bool,
                 pose_block_use_repvgg: bool, width_mult: float, first_conv_group_size: int, num_classes: int, stride: int, reg_max: int, cls_dropout_rate: float = 0.0,
                 reg_dropout_rate: float = 0.0):  # fmt: skip
        super().__init__(in_channels)
        bbox_inter_channels = width_multiplier(bbox_inter_channels, width_mult, 8)
        pose_inter_channels = width_multiplier(pose_inter_channels, width_mult, 8)
        if first_conv_group_size == 0:
            groups = 0
        elif first_conv_group_size == -1:
            groups = 1
        else:
            groups = bbox_inter_channels // first_conv_group_size
        if groups > 1:
            raise NotImplementedError("grouped first convolutions are not used by the shipped YOLO-NAS-POSE recipes and are not implemented")
        if cls_dropout_rate > 0 or reg_dropout_rate > 0:
            raise NotImplementedError("head dropout is not used by the shipped YOLO-NAS-POSE recipes and is not implemented")
        self.num_classes = num_classes
        self.shared_stem = shared_stem
        self.pose_conf_in_class_head = pose_conf_in_class_head
        if self.shared_stem:
            max_input = max(bbox_inter_channels, pose_inter_channels)
            self.stem = ConvBNReLU(in_channels, max_input, kernel_size=1, stride=1, padding=0, bias=False)
            self.pose_stem = _PlainConv1x1(max_input, pose_inter_channels) if max_input != pose_inter_channels else nn.Identity()
            self.bbox_stem = _PlainConv1x1(max_input, bbox_inter_channels) if max_input != bbox_inter_channels else nn.Identity()
        else:
            self.stem = nn.Identity()
            self.pose_stem = ConvBNReLU(in_channels, pose_inter_channels, kernel_size=1, stride=1, padding=0, bias=False)
            self.bbox_stem = ConvBNReLU(in_channels, bbox_inter_channels, kernel_size=1, stride=1, padding=0, bias=False)
        first_cls_conv = [ConvBNReLU(bbox_inter_channels, bbox_inter_channels, kernel_size=3, stride=1, padding=1, groups=groups, bias=False)] if groups else []
        self.cls_convs = nn.Sequential(*first_cls_conv, ConvBNReLU(bbox_inter_channels, bbox_inter_channels, kernel_size=3, stride=1, padding=1, bias=False))
        first_reg_conv = [ConvBNReLU(bbox_inter_channels, bbox_inter_channels, kernel_size=3, stride=1, padding=1, groups=groups, bias=False)] if groups else []
        self.reg_convs = nn.Sequential(*first_reg_conv, ConvBNReLU(bbox_inter_channels, bbox_inter_channels, kernel_size=3, stride=1, padding=1, bias=False))
        if pose_block_use_repvgg:
            pose_block = partial(QARepVGGBlock, use_alpha=True)
        else:
            pose_block = partial(ConvBNReLU, kernel_size=3, stride=1, padding=1, bias=False)
        self.pose_convs = nn.Sequential(*[pose_block(pose_inter_channels, pose_inter_channels) for _ in range(pose_regression_blocks)])
        self.reg_pred = nn.Conv2d(bbox_inter_channels, 4 * (reg_max + 1), 1, 1, 0)
        if self.pose_conf_in_class_head:
            self.cls_pred = nn.Conv2d(bbox_inter_channels, 1 + self.num_classes, 1, 1, 0)
            self.pose_pred = nn.Conv2d(pose_inter_channels, 2 * self.num_classes, 1, 1, 0)  # each keypoint is x, y
        else:
            self.cls_pred = nn.Conv2d(bbox_inter_channels, 1, 1, 1, 0)
            self.pose_pred = nn.Conv2d(pose_inter_channels, 3 * self.num_classes, 1, 1, 0)  # each keypoint is x, y, confidence
        self.cls_dropout_rate = nn.Identity()
        self.reg_dropout_rate = nn.Identity()
        self.stride = stride
        self.prior_prob = 1e-2
        self._initialize_biases()
        self._cls_cache, self._reg_cache, self._pose_cache = SF.WeightCache(), SF.WeightCache(), SF.WeightCache()

    def replace_num_classes(self, num_classes: int, compute_new_weights_fn=None):
        dev = self.cls_pred.weight.device
        if self.pose_conf_in_class_head:
            self.cls_pred = nn.Conv2d(self.cls_pred.in_channels, 1 + num_classes, 1, 1, 0).to(dev)
            self.pose_pred = nn.Conv2d(self.pose_pred.in_channels, 2 * num_classes, 1, 1, 0).to(dev)
        else:
            self.pose_pred = nn.Conv2d(self.pose_pred.in_channels, 3 * num_classes, 1, 1, 0).to(dev)
        self.num_classes = num_classes
        self._initialize_biases()
        self._cls_cache, self._pose_cache = SF.WeightCache(), SF.WeightCache()

    @property
    def out_channels(self):
        return None

    def forward(self, x) -> Tuple[Tensor, Tensor, Tensor]:
        """Returns bf16 NHWC maps (reg_output [B, 4*(reg_max+1), H, W], cls_output, pose_output).  With
        pose_conf_in_class_head (all shipped variants) cls_output is [B, 1 + J, H, W] (channel 0 = person logit, 1..J = joint
        logits) and pose_output [B, 2J, H, W] (channel 2j = x offset, 2j+1 = y offset of joint j)."""
        x = self.stem(x)
        pose_features = self.pose_stem(x)
        bbox_features = self.bbox_stem(x)
        cls_feat = self.cls_convs(bbox_features)
        cls_output = SF.conv_bias(cls_feat, self.cls_pred.weight, self.cls_pred.bias, stride=1, pad=0, cache=self._cls_cache)
        reg_feat = self.reg_convs(bbox_features)
        reg_output = SF.conv_bias(reg_feat, self.reg_pred.weight, self.reg_pred.bias, stride=1, pad=0, cache=self._reg_cache)
        pose_feat = self.pose_convs(pose_features)
        pose_output = SF.conv_bias(pose_feat, self.pose_pred.weight, self.pose_pred.bias, stride=1, pad=0, cache=self._pose_cache)
        return reg_output, cls_output, pose_output

    def _initialize_biases(self):
        prior_bias = -math.log((1 - self.prior_prob) / self.prior_prob)
        torch.nn.init.constant_(self.cls_pred.bias, prior_bias)
