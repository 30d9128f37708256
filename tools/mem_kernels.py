"""The memory-bound kernels of the hot path at benchmark size, one table: DFL + IoU loss (forward + backward, one kernel), the
task-aligned assigner's kernels, batched NMS for the detection callback (config 2: B = 32, 8400 anchors x 80 classes) and for the
pose callback (config 5: B = 64, 8400 anchors x 1 score).  CUDA events around every C-ABI call (kernels.PROFILE), L2 flushed between
repetitions, algorithmic bytes from SURVEY.md section 8(d) (9.94 MB / image for the loss: logits read + gradient written in fp32;
2.82 MB / image for the NMS: boxes + scores read).

    python tools/mem_kernels.py > gpurun_out/mem_kernels.txt                     # timing table
    ncu --set full -k regex:"nms_kernel|loss_kernel|tal_" -c 12 python tools/mem_kernels.py --reps 1   # dram__bytes for the same launches
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from super_gradients_b200 import kernels as K  # noqa: E402
from super_gradients_b200.training.losses import PPYoloELoss  # noqa: E402
from super_gradients_b200.training.models.detection_models.pp_yolo_e.post_prediction_callback import PPYoloEPostPredictionCallback  # noqa: E402
from super_gradients_b200.training.models.pose_estimation_models import YoloNASPosePostPredictionCallback  # noqa: E402


def anchors(img=640, strides=(8, 16, 32)):
    pts, st = [], []
    for s in strides:
        n = img // s
        ys, xs = torch.meshgrid(torch.arange(n), torch.arange(n), indexing="ij")
        pts.append(torch.stack([(xs.flatten() + 0.5) * s, (ys.flatten() + 0.5) * s], 1).float())
        st.append(torch.full((n * n, 1), float(s)))
    return torch.cat(pts), torch.cat(st), [(img // s) ** 2 for s in strides]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--breakdown", action="store_true", help="also list the kernels of one NMS call with their durations")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    B, C, L = 32, 80, 8400
    pts, st, nums = anchors()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator().manual_seed(1)
    # head outputs of a lightly trained detector: class logits around the -4.6 prior with a few confident anchors, DFL logits N(0, 1)
    cl = (torch.randn(B, L, C, generator=g) * 1.5 - 4.6).to(dev).requires_grad_(True)
    rd = torch.randn(B, L, 68, generator=g).to(dev).requires_grad_(True)
    _x, targets = bench.synth_batch(B, 3)
    crit = PPYoloELoss(num_classes=C, use_static_assigner=False)
    raw = (cl, rd, None, pts.to(dev), nums, st.to(dev))

    def loss_step():
        cl.grad = rd.grad = None
        with torch.no_grad():
            proj = torch.arange(17, device=dev, dtype=torch.float32)
            d = (torch.softmax(rd.detach().reshape(B, L, 4, 17), -1) * proj).sum(-1)
            p = pts.to(dev) / st.to(dev)
            pb = torch.cat([p - d[..., :2], p + d[..., 2:]], -1) * st.to(dev)
            ps = torch.sigmoid(cl.detach())
        loss, _ = crit(((pb, ps), raw), targets)
        loss.backward()
        return pb, ps

    pb, ps = loss_step()
    det_cb = PPYoloEPostPredictionCallback(score_threshold=0.03, nms_threshold=0.65, nms_top_k=1000, max_predictions=300)  # the COCO validation recipe
    Bp = 64
    pboxes = torch.rand(Bp, L, 4, generator=g) * 300
    pboxes[..., 2:] += pboxes[..., :2] + 20
    pconf = torch.rand(Bp, L, 1, generator=g) ** 8  # a few hundred anchors above 0.5 per image
    pcoords, pjs = torch.rand(Bp, L, 17, 2, generator=g) * 640, torch.rand(Bp, L, 17, generator=g)
    pose_in = tuple(t.to(dev) for t in (pboxes, pconf, pcoords, pjs))
    pose_cb = YoloNASPosePostPredictionCallback(pose_confidence_threshold=0.5, nms_iou_threshold=0.7, pre_nms_max_predictions=300, post_nms_max_predictions=100)

    rows = {}

    def record(tag):
        for name, a, b, _t in K.PROFILE:
            rows.setdefault((tag, name), []).append(a.elapsed_time(b) * 1e3)
        K.PROFILE.clear()

    for rep in range(args.reps + 2):
        on = rep >= 2
        for tag, fn in (("loss B=32", loss_step), ("det nms B=32", lambda: det_cb.forward_batched((pb, ps))), ("pose nms B=64", lambda: pose_cb.forward_batched((pose_in, None)))):
            flush.zero_()
            torch.cuda.synchronize()
            K.PROFILE.clear()
            K.PROFILE_ON[0] = on
            out = fn()
            torch.cuda.synchronize()
            K.PROFILE_ON[0] = False
            if on:
                record(tag)
    if args.breakdown:  # kernels of one detection / pose NMS call (CUPTI records through torch.profiler)
        from torch.profiler import ProfilerActivity, profile

        for tag, fn in (("det nms B=32", lambda: det_cb.forward_batched((pb, ps))), ("pose nms B=64", lambda: pose_cb.forward_batched((pose_in, None)))):
            flush.zero_()
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                fn()
                torch.cuda.synchronize()
            print(f"-- kernels of one `{tag}` call")
            for e in sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start):
                print(f"   {e.time_range.end - e.time_range.start:9.1f} us  {e.name[:90]}")
    kept_det = float(det_cb.forward_batched((pb, ps))[2].float().mean())
    kept_pose = float(pose_cb.forward_batched((pose_in, None))[3].float().mean())
    hbm = bench.peaks()[1]
    alg = {
        ("loss B=32", "sgb_dfl_iou_loss_fwd_bwd"): B * 2 * 4 * L * (C + 68),  # logits read once, gradients written once (fp32)
        ("loss B=32", "sgb_tal_assign"): B * 4 * L * (C + 68),               # the assigner reads the logits once
        ("det nms B=32", "sgb_batched_nms"): B * 4 * L * (4 + C),
        ("pose nms B=64", "sgb_batched_nms"): Bp * 4 * L * (4 + 1),
    }
    print(f"memory-bound kernels at benchmark size ({args.reps} repetitions, L2 flushed before each; HBM peak {hbm:.0f} GB/s)")
    print(f"mean detections kept per image: detection {kept_det:.1f}, pose {kept_pose:.1f}")
    print(f"{'case':16s} {'C-ABI call':30s} {'us (median)':>12s} {'alg. MB':>9s} {'GB/s':>8s} {'of HBM':>7s}")
    for (tag, name), v in sorted(rows.items()):
        v.sort()
        us = v[len(v) // 2]
        byts = alg.get((tag, name))
        extra = f"{byts / 1e6:9.1f} {byts / us / 1e3:8.0f} {byts / us / 1e3 / hbm * 100:6.1f}%" if byts else ""
        print(f"{tag:16s} {name:30s} {us:12.1f} {extra}")
    _ = out


if __name__ == "__main__":
    main()
