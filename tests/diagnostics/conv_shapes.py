"""Lists every convolution call (fprop / dgrad / wgrad) of one training step of a model, in launch order, with its shape --
run on the CPU stand-in of the kernel wrappers (tests/cpu_backend.py), so the list can be matched against a GPU timeline dump
(tools/timeline.py --dump) by order.  Test infrastructure (imports the stand-in, which imports oracle/).

    python tests/diagnostics/conv_shapes.py [--model yolo_nas_s] [--size 640] > shapes.txt
"""
import argparse
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cpu_backend  # noqa: E402

import bench  # noqa: E402
from super_gradients_b200 import kernels as K  # noqa: E402
from super_gradients_b200.training import models  # noqa: E402
from super_gradients_b200.training.losses import PPYoloELoss  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="yolo_nas_s")
    ap.add_argument("--size", type=int, default=640)
    args = ap.parse_args()
    mp = pytest.MonkeyPatch()
    cpu_backend.install_training(mp)
    log = []
    for name in ("conv_fprop", "conv_dgrad", "conv_wgrad"):
        orig = getattr(K, name)

        def wrap(*a, _orig=orig, _name=name, **kw):
            if _name == "conv_fprop":
                x, _w, kout, r, s, stride = a[0], a[1], a[2], a[3], a[4], a[5]
                log.append((_name, tuple(x.shape[1:]), kout, r, stride))
            elif _name == "conv_dgrad":
                dy, _w, xs, r, s, stride = a[0], a[1], a[2], a[3], a[4], a[5]
                log.append((_name, tuple(xs[1:]), dy.shape[1], r, stride))
            else:
                x, dy, r, s, stride = a[0], a[1], a[2], a[3], a[4]
                log.append((_name, tuple(x.shape[1:]), dy.shape[1], r, stride))
            return _orig(*a, **kw)

        mp.setattr(K, name, wrap)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    model = models.get(args.model, num_classes=bench.NCLS).train()
    crit = PPYoloELoss(num_classes=bench.NCLS, use_static_assigner=False)
    x = torch.randn(1, 3, args.size, args.size)
    t = torch.tensor([[0, 3, 0.5, 0.5, 0.3, 0.3], [0, 7, 0.3, 0.6, 0.2, 0.2]]) * torch.tensor([1, 1, args.size, args.size, args.size, args.size])
    out = model(x)
    loss = crit(out, t)
    loss = loss[0] if isinstance(loss, (tuple, list)) else loss
    loss.backward()
    for i, (name, xs, k, r, stride) in enumerate(log):
        c, h, w = xs
        print(f"{i:4d} {name:11s} C={c:4d} K={k:4d} {r}x{r} s{stride} in {h}x{w}")
    mp.undo()


if __name__ == "__main__":
    main()
