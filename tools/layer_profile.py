"""Per-layer timing of one YOLO-NAS-S train step (eager, CUDA events around every C-ABI call).

    python tools/layer_profile.py [batch] > gpurun_out/layers.txt

Groups launches by (entry point, problem shape) and prints count, total time, algorithmic minimum HBM bytes and the
bandwidth / tensor throughput that corresponds to.  Eager per-call timings include a few us of launch overhead.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from super_gradients_b200 import kernels as K  # noqa: E402
from super_gradients_b200.training import models  # noqa: E402
from super_gradients_b200.training.losses import PPYoloELoss, pad_targets_host  # noqa: E402
from super_gradients_b200.training.sg_trainer import TrainStep, setup_device  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = setup_device()
    torch.manual_seed(0)
    model = models.get("yolo_nas_s", num_classes=bench.NCLS).to(dev).train()
    crit = PPYoloELoss(num_classes=bench.NCLS, use_static_assigner=False)
    step = TrainStep(model, crit, "AdamW", {"weight_decay": 1e-5}, zero_wd_on_bias_and_bn=True, ema=True)
    x, t = bench.synth_batch(batch, 0)
    tt = tuple(a.to(dev) for a in pad_targets_host(t, batch, bench.NBOX))
    x = x.to(dev)
    step.set_hyper_params(2e-4, 0.9997)
    for _ in range(3):
        step._step_eager(x, tt)
    torch.cuda.synchronize()
    K.PROFILE.clear()
    K.PROFILE_ON[0] = True
    reps = 3
    for _ in range(reps):
        step._step_eager(x, tt)
    torch.cuda.synchronize()
    K.PROFILE_ON[0] = False
    agg = {}
    for name, a, b, tag in K.PROFILE:
        e = agg.setdefault((name, tag), [0, 0.0])
        e[0] += 1
        e[1] += a.elapsed_time(b)
    rows = []
    for (name, tag), (cnt, ms) in agg.items():
        cnt //= reps
        us = ms * 1e3 / reps
        byts = flops = 0
        if name.startswith("sgb_conv_") and len(tag) == 7:
            N, H, W, C, Kc, R, s = tag
            P, Q = (H + s - 1) // s, (W + s - 1) // s
            xb, yb = N * H * W * C * 2, N * P * Q * Kc * 2
            byts = cnt * (xb + yb)
            flops = cnt * 2.0 * N * P * Q * Kc * C * R * R
        rows.append((us, name, tag, cnt, byts, flops))
    rows.sort(reverse=True)
    tot = sum(r[0] for r in rows)
    print(f"total {tot / 1e3:.2f} ms over {sum(r[3] for r in rows)} calls (eager, batch {batch})")
    print(f"{'us':>9} {'share':>6} {'cnt':>4}  {'GB/s':>7} {'TF/s':>6}  call (N,H,W,C,K,R,stride)")
    for us, name, tag, cnt, byts, flops in rows[:70]:
        gbs = byts / (us * 1e-6) / 1e9 if byts else 0
        tfs = flops / (us * 1e-6) / 1e12 if flops else 0
        print(f"{us:9.1f} {100 * us / tot:5.1f}% {cnt:4d}  {gbs:7.0f} {tfs:6.1f}  {name} {tag}")
    per = {}
    for us, name, *_ in rows:
        per[name] = per.get(name, 0) + us
    print()
    for k, v in sorted(per.items(), key=lambda kv: -kv[1]):
        print(f"{v / 1e3:8.3f} ms  {k}")


if __name__ == "__main__":
    main()
