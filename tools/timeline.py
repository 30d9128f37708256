"""In-situ kernel timeline of the CUDA-graph train step (CUPTI activity records through torch.profiler -- there is no nsys in
the image).  Unlike an ncu launch list (serialised, cold caches) these are the kernels' start / end times inside the real
replay, so the output separates the time kernels execute from the idle gaps between dependent launches.

    python tools/timeline.py [--batch 32] [--steps 3] [--model yolo_nas_s] > gpurun_out/timeline.txt

Prints: step wall time, sum of kernel durations, idle time (no kernel running), overlap, then per-kernel-name totals sorted by
time with average duration and the average gap that precedes the kernel.  Works under torchrun (each rank prints its own
table to gpurun_out/timeline_rank<r>.txt) so it doubles as the multi-GPU all-reduce / skew trace.
"""
import argparse
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from super_gradients_b200.training import models  # noqa: E402
from super_gradients_b200.training.losses import PPYoloELoss, pad_targets_host  # noqa: E402
from super_gradients_b200.training.sg_trainer import TrainStep, setup_device  # noqa: E402


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*$", "", name)
    return name[:70]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--model", default="yolo_nas_s")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--top", type=int, default=60)
    ap.add_argument("--aten", action="store_true", help="eager step with Python stacks: which call sites launch the remaining ATen kernels")
    ap.add_argument("--dump", default=None, help="also write every launch of the LAST timed step in start order (start us, duration us, name) to this file")
    args = ap.parse_args()
    import torch.distributed as dist

    dev = setup_device()
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    torch.manual_seed(0)
    model = models.get(args.model, num_classes=bench.NCLS).to(dev).train()
    crit = PPYoloELoss(num_classes=bench.NCLS, use_static_assigner=False)
    step = TrainStep(model, crit, "AdamW", {"weight_decay": 1e-5}, zero_wd_on_bias_and_bn=True, ema=True)
    xs, ts = [], []
    for i in range(2):
        x, t = bench.synth_batch(args.batch, 1000 * rank + i)
        xs.append(x.to(dev))
        ts.append(tuple(a.to(dev) for a in pad_targets_host(t, args.batch, bench.NBOX)))
    step.set_hyper_params(2e-4, 0.9997)
    step.run(xs[0], ts[0])
    if not args.no_graph:
        step.capture(xs[0], ts[0], warmup=2)
    for i in range(3):
        step.set_hyper_params(2e-4, 0.9997)
        step.run(xs[i % 2], ts[i % 2])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile

    if args.aten:
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
            step.set_hyper_params(2e-4, 0.9997)
            step._step_eager(xs[0], ts[0])
            torch.cuda.synchronize()
        rows = [e for e in prof.key_averages(group_by_stack_n=8) if e.key.startswith("aten::") and e.self_device_time_total > 0]
        rows.sort(key=lambda e: -e.self_device_time_total)
        for e in rows[: args.top]:
            stack = [fr for fr in e.stack if "super_gradients_b200" in fr or "bench.py" in fr][:3]
            print(f"{e.key:28s} n={e.count:4d} cuda {e.self_device_time_total / 1e3:8.3f} ms   " + " <- ".join(fr.split("/")[-1] for fr in stack))
        return
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            step.set_hyper_params(2e-4, 0.9997)
            step.run(xs[i % 2], ts[i % 2])
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs), key=lambda r: r[0])
    if not ks:
        print("no CUDA activity records captured")
        return
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    wall = (t1 - t0) / args.steps
    busy_sum = sum(k[1] - k[0] for k in ks) / args.steps
    # union of intervals -> time at least one kernel is running
    union, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
    for s, e, _ in ks[1:]:
        if s > cur_e:
            union += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    union += cur_e - cur_s
    union /= args.steps
    per = {}
    prev_end = ks[0][0]
    for s, e, n in ks:
        d = per.setdefault(short(n), [0, 0.0, 0.0])
        d[0] += 1
        d[1] += e - s
        d[2] += max(0.0, s - prev_end)
        prev_end = max(prev_end, e)
    out = []
    out.append(f"rank {rank}/{world} model {args.model} batch {args.batch} graph {not args.no_graph}: {len(ks) // args.steps} kernels/step")
    out.append(f"wall {wall / 1e3:.3f} ms/step   sum of kernel durations {busy_sum / 1e3:.3f} ms   GPU busy (union) {union / 1e3:.3f} ms   idle {(wall - union) / 1e3:.3f} ms   overlap {(busy_sum - union) / 1e3:.3f} ms")
    out.append(f"{'kernel':70s} {'n/step':>7s} {'ms/step':>8s} {'share':>6s} {'avg us':>7s} {'gap before us':>13s}")
    for n, (c, dur, gap) in sorted(per.items(), key=lambda kv: -kv[1][1])[: args.top]:
        out.append(f"{n:70s} {c / args.steps:7.1f} {dur / args.steps / 1e3:8.3f} {dur / args.steps / wall * 100:5.1f}% {dur / c:7.1f} {gap / c:13.2f}")
    text = "\n".join(out)
    if args.dump and rank == 0:
        cut = t0 + (t1 - t0) * (args.steps - 1) / args.steps
        last = [k for k in ks if k[0] >= cut]
        with open(args.dump, "w") as f:
            for s, e, n in last:
                f.write(f"{(s - last[0][0]):10.1f} {(e - s):8.1f} {short(n)}\n")
    if world > 1:
        os.makedirs("gpurun_out", exist_ok=True)
        open(f"gpurun_out/timeline_rank{rank}.txt", "w").write(text + "\n")
        if rank == 0:
            print(text)
    else:
        print(text)
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
