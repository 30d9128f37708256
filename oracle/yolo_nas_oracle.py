"""Whole-graph CPU oracle of YOLO-NAS (TEST INFRASTRUCTURE -- see oracle/sg_oracle.py for the rules).

A functional fp32 restatement of CustomizableDetector.forward for the YOLO-NAS family, driven by the arch-params dict
and a reference-format state dict (same keys as the reference's `model.state_dict()`):
  NStageBackbone.forward            modules/detection_modules.py:83-90
  YoloNASStem / Stage / CSPLayer / Bottleneck / UpStage / DownStage   yolo_nas/yolo_stages.py:61-63,144-150,234-235,319-332,390-395
  SPP                               detection_models/csp_darknet53.py:135-157
  YoloNASPANNeckWithC2.forward      yolo_nas/panneck.py:56-64
  YoloNASDFLHead / NDFLHeads        yolo_nas/dfl_heads.py:73-83,199-245
Pinned by tests/test_oracle_golden.py::test_tiny_yolo_nas_whole_graph against tests/golden/tiny_yolo_nas.pt
(outputs, loss and gradients produced by the unmodified reference).
"""
import math
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import sg_oracle as O


def width_multiplier(original, factor, divisor=None):
    if divisor is None:
        return int(original * factor)
    return math.ceil(int(original * factor) / divisor) * divisor


class YoloNASOracle:
    def __init__(self, arch: dict, state: Dict[str, torch.Tensor], training: bool = True):
        self.arch, self.p, self.training = arch, state, training
        self.eps = float(arch.get("bn_eps") or 1e-5)
        self.mom = float(arch.get("bn_momentum") or 0.1)

    # ---- leaves
    def _qarep(self, x, prefix, stride, residual):
        return O.qarepvgg_forward(x, self.p, prefix, stride, residual, "relu", self.training, self.eps, self.mom)

    def _conv(self, x, prefix, k, stride, conv="conv", bn="bn"):
        return O.conv_bn_act(x, self.p, prefix, stride, k // 2, "relu", self.training, self.eps, self.mom, conv, bn)

    def _csp(self, x, prefix, num_blocks, block, concat_intermediates):
        x1 = self._conv(x, prefix + "conv1.", 1, 1)
        outs = [x1]
        for i in range(num_blocks):
            bp = f"{prefix}bottlenecks.{i}."
            cur = outs[-1]
            if block == "qarep":
                y = self._qarep(self._qarep(cur, bp + "cv1.", 1, True), bp + "cv2.", 1, True)
            else:
                y = self._conv(self._conv(cur, bp + "cv1.", 3, 1), bp + "cv2.", 3, 1)
            alpha = self.p.get(bp + "alpha", 1.0)
            outs.append(O.q(alpha * cur + y))
        x1s = outs if concat_intermediates else [outs[-1]]
        x2 = self._conv(x, prefix + "conv2.", 1, 1)
        return self._conv(torch.cat((*x1s, x2), 1), prefix + "conv3.", 1, 1)

    # ---- backbone / neck / heads
    def backbone(self, x):
        bb = self.arch["backbone"]["NStageBackbone"]
        outs = {}
        x = self._qarep(x, "backbone.stem.conv.", 2, False)
        outs["stem"] = x
        for i, st in enumerate(bb["stages"]):
            a = st["YoloNASStage"]
            pre = f"backbone.stage{i + 1}."
            x = self._qarep(x, pre + "downsample.", 2, False)
            x = self._csp(x, pre + "blocks.", a["num_blocks"], "qarep", a.get("concat_intermediates", False))
            outs[f"stage{i + 1}"] = x
        if bb.get("context_module"):
            ks = bb["context_module"]["SPP"]["k"]
            x = O.spp(x, self.p, "backbone.context_module.", ks, "relu", self.training, self.eps, self.mom)
            outs["context_module"] = x
        return [outs[k] for k in bb["out_layers"]]

    def _up_stage(self, prefix, a, inputs):
        x, s1, s2 = inputs
        s1 = self._conv(s1, prefix + "reduce_skip1.", 1, 1)
        s2 = self._conv(s2, prefix + "reduce_skip2.", 1, 1)
        s2 = self._conv(s2, prefix + "downsample.", 3, 2)
        x_inter = self._conv(x, prefix + "conv.", 1, 1)
        up = O.q(F.conv_transpose2d(x_inter, O.qw(self.p[prefix + "upsample.weight"]), self.p[prefix + "upsample.bias"], stride=2))
        x = self._conv(torch.cat([up, s1, s2], 1), prefix + "reduce_after_concat.", 1, 1)
        nb = a["num_blocks"]
        nb = max(round(nb * a.get("depth_mult", 1)), 1) if nb > 1 else nb
        return x_inter, self._csp(x, prefix + "blocks.", nb, "qarep", a.get("concat_intermediates", False))

    def _down_stage(self, prefix, a, inputs):
        x, skip = inputs
        x = self._conv(x, prefix + "conv.", 3, 2)
        nb = a["num_blocks"]
        nb = max(round(nb * a.get("depth_mult", 1)), 1) if nb > 1 else nb
        return self._csp(torch.cat([x, skip], 1), prefix + "blocks.", nb, "conv", a.get("concat_intermediates", False))

    def neck(self, feats):
        c2, c3, c4, c5 = feats
        nk = self.arch["neck"]["YoloNASPANNeckWithC2"]
        n1i, x = self._up_stage("neck.neck1.", nk["neck1"]["YoloNASUpStage"], [c5, c4, c3])
        n2i, p3 = self._up_stage("neck.neck2.", nk["neck2"]["YoloNASUpStage"], [x, c3, c2])
        p4 = self._down_stage("neck.neck3.", nk["neck3"]["YoloNASDownStage"], [p3, n2i])
        p5 = self._down_stage("neck.neck4.", nk["neck4"]["YoloNASDownStage"], [p4, n1i])
        return p3, p4, p5

    def pose_heads(self, feats):
        """YoloNASPoseNDFLHeads over YoloNASPoseDFLHead levels (pose_estimation_models/yolo_nas_pose/
        yolo_nas_pose_dfl_head.py:131-167, yolo_nas_pose_ndfl_heads.py:126-206); separate stems, joint logits in the class head."""
        hd = self.arch["heads"]["YoloNASPoseNDFLHeads"]
        J = hd["num_classes"]
        regs, clss, prs, pls, strides = [], [], [], [], []
        cba = lambda t, pre, k: O.conv_bn_act(t, self.p, pre, 1, k // 2, "relu", self.training, self.eps, self.mom)  # noqa: E731
        for i, (f, h) in enumerate(zip(feats, hd["heads_list"])):
            a = h["YoloNASPoseDFLHead"]
            assert not a["shared_stem"] and a["pose_conf_in_class_head"] and not a["pose_block_use_repvgg"] and a["first_conv_group_size"] == 0
            pre = f"heads.head{i + 1}."
            pose_f = cba(f, pre + "pose_stem.seq.", 1)
            bbox_f = cba(f, pre + "bbox_stem.seq.", 1)
            c = cba(bbox_f, pre + "cls_convs.0.seq.", 3)
            r = cba(bbox_f, pre + "reg_convs.0.seq.", 3)
            for b in range(a["pose_regression_blocks"]):
                pose_f = cba(pose_f, pre + f"pose_convs.{b}.seq.", 3)
            cls_out = O.q(F.conv2d(c, O.qw(self.p[pre + "cls_pred.weight"]), self.p[pre + "cls_pred.bias"]))
            regs.append(O.q(F.conv2d(r, O.qw(self.p[pre + "reg_pred.weight"]), self.p[pre + "reg_pred.bias"])))
            pose_out = O.q(F.conv2d(pose_f, O.qw(self.p[pre + "pose_pred.weight"]), self.p[pre + "pose_pred.bias"]))
            clss.append(cls_out[:, 0:1])
            pls.append(cls_out[:, 1:])
            prs.append(pose_out.reshape(pose_out.shape[0], J, 2, pose_out.shape[2], pose_out.shape[3]))
            strides.append(a["stride"])
        return O.pose_ndfl_decode(regs, clss, prs, pls, strides, reg_max=hd.get("reg_max", 16), pose_offset_multiplier=hd.get("pose_offset_multiplier", 1.0),
                                  compensate_grid_cell_offset=hd.get("compensate_grid_cell_offset", True))  # fmt: skip

    def heads(self, feats):
        if "YoloNASPoseNDFLHeads" in self.arch["heads"]:
            return self.pose_heads(feats)
        hd = self.arch["heads"]["NDFLHeads"]
        regs, clss, strides = [], [], []
        for i, (f, h) in enumerate(zip(feats, hd["heads_list"])):
            pre = f"heads.head{i + 1}."
            x = O.conv_bn_act(f, self.p, pre + "stem.seq.", 1, 0, "relu", self.training, self.eps, self.mom)
            c = O.conv_bn_act(x, self.p, pre + "cls_convs.0.seq.", 1, 1, "relu", self.training, self.eps, self.mom)
            r = O.conv_bn_act(x, self.p, pre + "reg_convs.0.seq.", 1, 1, "relu", self.training, self.eps, self.mom)
            clss.append(O.q(F.conv2d(c, O.qw(self.p[pre + "cls_pred.weight"]), self.p[pre + "cls_pred.bias"])))
            regs.append(O.q(F.conv2d(r, O.qw(self.p[pre + "reg_pred.weight"]), self.p[pre + "reg_pred.bias"])))
            strides.append(h["YoloNASDFLHead"]["stride"])
        return O.ndfl_decode(regs, clss, strides, reg_max=hd.get("reg_max", 16))

    def forward(self, x):
        return self.heads(self.neck(self.backbone(O.q(x))))


def train_step(arch: dict, state: Dict[str, torch.Tensor], x: torch.Tensor, targets: torch.Tensor, num_classes: int, live: List[str]):
    """fp32 CPU forward + PPYoloELoss + backward; returns (loss, items, {param: grad}).  `live` = trainable keys."""
    p = dict(state)
    for k in live:
        p[k] = p[k].detach().clone().requires_grad_(True)
    outs = YoloNASOracle(arch, p, training=True).forward(x)
    loss, items = O.ppyoloe_loss(outs[1], targets, num_classes)
    loss.backward()
    return loss.detach(), items, {k: p[k].grad for k in live if p[k].grad is not None}


def random_state(shapes: Dict[str, tuple], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random reference-format state dict from a {key: shape} table (tests/golden/state_keys.pt), for timing runs that
    must not touch the product package: conv weights ~ N(0, fan_in^-1/2), BN weight / running_var = 1, rest = 0."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, shp in shapes.items():
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var") or (k.endswith(".weight") and len(shp) == 1):
            out[k] = torch.ones(shp)
        elif len(shp) >= 2:
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            out[k] = torch.randn(shp, generator=g) * fan_in**-0.5
        elif k.endswith("alpha"):
            out[k] = torch.ones(shp)
        else:
            out[k] = torch.zeros(shp)
    return out
