"""YoloNASPoseNDFLHeads (reference: pose_estimation_models/yolo_nas_pose/yolo_nas_pose_ndfl_heads.py:23-242): runs the
per-level heads and decodes boxes (DFL softmax integral), person score, keypoints and joint scores with two launches per
level (sgb_dfl_decode, sgb_pose_keypoint_decode) into the reference's [B, L, *] tensors; in training the raw outputs carry
gradients back into the per-level maps (functional._PoseDecode) for YoloNASPoseLoss (training/losses/yolo_nas_pose_loss.py)."""
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from ..... import functional as SF
from .....common.factories import DetectionModulesFactory
from .....common.registry import register_detection_module
from .....modules import BaseDetectionModule
from ...detection_models.yolo_nas.dfl_heads import generate_anchors_for_grid_cell


@register_detection_module()
class YoloNASPoseNDFLHeads(BaseDetectionModule):
    def __init__(self, num_classes: int, in_channels: Tuple[int, int, int], heads_list, grid_cell_scale: float = 5.0, grid_cell_offset: float = 0.5, reg_max: int = 16,
                 inference_mode: bool = False, eval_size: Optional[Tuple[int, int]] = None, width_mult: float = 1.0, pose_offset_multiplier: float = 1.0,
                 compensate_grid_cell_offset: bool = True):  # fmt: skip
        in_channels = [max(round(c * width_mult), 1) for c in in_channels]
        super().__init__(in_channels)
        self.in_channels = tuple(in_channels)
        self.num_classes = num_classes
        self.grid_cell_scale = grid_cell_scale
        self.grid_cell_offset = grid_cell_offset
        self.reg_max = reg_max
        self.eval_size = eval_size
        self.pose_offset_multiplier = pose_offset_multiplier
        self.compensate_grid_cell_offset = compensate_grid_cell_offset
        self.inference_mode = inference_mode
        proj = torch.linspace(0, self.reg_max, self.reg_max + 1).reshape([1, self.reg_max + 1, 1, 1])
        self.register_buffer("proj_conv", proj, persistent=False)
        factory = DetectionModulesFactory()
        for i in range(len(heads_list)):
            heads_list[i] = factory.insert_module_param(heads_list[i], "num_classes", num_classes)
            heads_list[i] = factory.insert_module_param(heads_list[i], "reg_max", reg_max)
        self.num_heads = len(heads_list)
        fpn_strides: List[int] = []
        for i in range(self.num_heads):
            new_head = factory.get(factory.insert_module_param(heads_list[i], "in_channels", in_channels[i]))
            fpn_strides.append(new_head.stride)
            setattr(self, f"head{i + 1}", new_head)
        self.fpn_strides = tuple(fpn_strides)
        self._anchor_cache = {}

    def replace_num_classes(self, num_classes: int, compute_new_weights_fn=None):
        for i in range(self.num_heads):
            getattr(self, f"head{i + 1}").replace_num_classes(num_classes, compute_new_weights_fn)
        self.num_classes = num_classes

    @property
    def out_channels(self):
        return None

    def _anchors(self, shapes, device):
        key = (tuple(shapes), str(device))
        if key not in self._anchor_cache:
            self._anchor_cache[key] = generate_anchors_for_grid_cell(shapes, self.fpn_strides, self.grid_cell_scale, self.grid_cell_offset, device)
        return self._anchor_cache[key]

    def forward(self, feats: Tuple[Tensor, ...]):
        """Returns decoded (pred_bboxes [B, L, 4], pred_scores [B, L, 1], pred_pose_coords [B, L, J, 2], pred_pose_scores
        [B, L, J]) in inference_mode, else (decoded, raw) with raw = (cls_logits, reg_distri, pose_coords, pose_logits,
        anchors, anchor_points, num_anchors_list, stride_tensor) like the reference."""
        feats = feats[: self.num_heads]
        regs, clss, poses = [], [], []
        for i, feat in enumerate(feats):
            head = getattr(self, f"head{i + 1}")
            if not head.pose_conf_in_class_head:
                raise NotImplementedError("pose_conf_in_class_head=False is not used by the shipped YOLO-NAS-POSE recipes and is not implemented")
            reg, cls, pose = head(feat)
            regs.append(reg)
            clss.append(cls)
            poses.append(pose)
        pb, ps, pc, pj, cl, rd, pl = SF.pose_decode(regs, clss, poses, self.fpn_strides, self.num_classes, self.reg_max, self.grid_cell_offset, self.pose_offset_multiplier,
                                                    self.compensate_grid_cell_offset)  # fmt: skip
        decoded = pb, ps, pc.detach(), pj  # the reference decodes from detached clones (:201-202)
        if self.inference_mode:
            return decoded
        shapes = [(f.shape[2], f.shape[3]) for f in feats]
        anchors, anchor_points, num_anchors_list, stride_tensor = self._anchors(shapes, pb.device)
        return decoded, (cl, rd, pc, pl, anchors, anchor_points, num_anchors_list, stride_tensor)
