"""Debug helper (not a test): layer-by-layer comparison of the GPU model against the whole-graph CPU oracle."""
import copy
import sys

import torch

sys.path.insert(0, ".")
from oracle.yolo_nas_oracle import YoloNASOracle  # noqa: E402
from super_gradients_b200 import functional as SF  # noqa: E402
from super_gradients_b200.training.models.detection_models.yolo_nas import YoloNAS  # noqa: E402


def l2rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


g = torch.load("tests/golden/tiny_yolo_nas.pt", weights_only=False)
ap = copy.deepcopy(g["arch"])
m = YoloNAS(backbone=ap["backbone"], neck=ap["neck"], heads=ap["heads"], num_classes=4, bn_eps=1e-3, bn_momentum=0.03, inplace_act=True, in_channels=3)
m.load_state_dict(g["sd0"], strict=False)
m.cuda().train()
orc = YoloNASOracle(g["arch"], {k: v.clone() for k, v in g["sd0"].items()}, training=True)
x = g["x"]
# backbone, layer by layer
xo = x
xg = SF.to_nhwc(x.cuda())
bb = m.backbone
o = orc._qarep(xo, "backbone.stem.conv.", 2, False)
gq = bb.stem(xg)
print("stem", l2rel(gq, o))
outs_o = orc.backbone(x)
outs_g = bb(SF.to_nhwc(x.cuda()))
for i, (a, b) in enumerate(zip(outs_g, outs_o)):
    print("backbone out", i, tuple(b.shape), l2rel(a, b))
# feed the ORACLE's backbone outputs into both necks to isolate the neck
orc2 = YoloNASOracle(g["arch"], {k: v.clone() for k, v in g["sd0"].items()}, training=True)
feats_o = orc2.backbone(x)
pn_o = orc2.neck(feats_o)
pn_g = m.neck([SF.to_nhwc(f.cuda().bfloat16().contiguous(memory_format=torch.channels_last)) for f in feats_o])
for i, (a, b) in enumerate(zip(pn_g, pn_o)):
    print("neck out (oracle feats in)", i, tuple(b.shape), l2rel(a, b))
(pb_o, ps_o), raw_o = orc2.heads(pn_o)
(pb_g, ps_g), raw_g = m.heads([SF.to_nhwc(f.cuda().bfloat16().contiguous(memory_format=torch.channels_last)) for f in pn_o])
print("heads (oracle feats in): cls", l2rel(raw_g[0], raw_o[0]), "reg", l2rel(raw_g[1], raw_o[1]), "boxes", l2rel(pb_g, pb_o))
# individual pieces of stage 3 (concat_intermediates) and the SPP
st = bb.stage1
xin = outs_o[0] * 0 + torch.randn_like(outs_o[0])
