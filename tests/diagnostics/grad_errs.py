"""Debug helper: per-parameter gradient error of the tiny YOLO-NAS train step vs the bf16-emulating CPU oracle."""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from oracle import sg_oracle as O  # noqa: E402
from oracle.yolo_nas_oracle import YoloNASOracle  # noqa: E402
from super_gradients_b200.training.losses import PPYoloELoss  # noqa: E402
from super_gradients_b200.training.models.detection_models.yolo_nas import YoloNAS  # noqa: E402


def l2rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


g = torch.load("tests/golden/tiny_yolo_nas.pt", weights_only=False)
ap = copy.deepcopy(g["arch"])
m = YoloNAS(backbone=ap["backbone"], neck=ap["neck"], heads=ap["heads"], num_classes=4, bn_eps=1e-3, bn_momentum=0.03, inplace_act=True, in_channels=3)
m.load_state_dict(g["sd0"], strict=False)
m.cuda().train()
(pb, ps), raw = m(g["x"].cuda())
crit = PPYoloELoss(num_classes=4, use_static_assigner=False)
loss, items = crit(((pb, ps), raw), g["targets"])
loss.backward()
params = dict(m.named_parameters())
live = [k for k in g["param_names"] if "rbr_reparam" not in k]
with O.bf16_emulation():
    pe = {k: v.clone() for k, v in g["sd0"].items()}
    for k in live:
        pe[k].requires_grad_(True)
    (pbe, pse), rawe = YoloNASOracle(g["arch"], pe, training=True).forward(g["x"])
    losse, itemse = O.ppyoloe_loss(rawe, g["targets"], 4)
    losse.backward()
print("loss", float(loss), float(losse), float(g["loss"]))
print("cls", l2rel(raw[0], rawe[0]), "reg", l2rel(raw[1], rawe[1]))
rows = []
for k in live:
    if pe[k].grad is None:
        continue
    ref32 = g["grads"].get(k)
    rows.append((l2rel(params[k].grad, pe[k].grad), float(params[k].grad.norm()), float(pe[k].grad.norm()), l2rel(params[k].grad, ref32) if ref32 is not None else -1, l2rel(pe[k].grad, ref32) if ref32 is not None else -1, k))
print("err(gpu,emul)  |gpu|  |emul|  err(gpu,fp32ref)  err(emul,fp32ref)  name")
for r in rows:
    print(f"{r[0]:8.3f} {r[1]:10.3e} {r[2]:10.3e} {r[3]:8.3f} {r[4]:8.3f}  {r[5]}")
