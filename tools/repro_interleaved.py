"""Reproducer for DESIGN.md section 8, item 1: two TrainSteps alive in one process, steps interleaved, model b re-synchronised
from model a before every step.  Prints, per step, whether repeated forwards of the SAME model on the SAME parameters agree
(`a-a`, `b-b`), whether the two models agree (`a-b`) and the loss / gradient agreement of the step itself.

    python tools/repro_interleaved.py                       # plain
    compute-sanitizer --tool memcheck  python tools/repro_interleaved.py
    compute-sanitizer --tool initcheck python tools/repro_interleaved.py
    compute-sanitizer --tool racecheck python tools/repro_interleaved.py   # slow: set STEPS=1
"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from super_gradients_b200 import functional as SF  # noqa: E402
from super_gradients_b200.training.losses import PPYoloELoss, pad_targets_host  # noqa: E402
from super_gradients_b200.training.models.detection_models.yolo_nas import YoloNAS  # noqa: E402
from super_gradients_b200.training.sg_trainer import TrainStep  # noqa: E402


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def make(g, batched):
    ap = copy.deepcopy(g["arch"])
    m = YoloNAS(backbone=ap["backbone"], neck=ap["neck"], heads=ap["heads"], num_classes=4, bn_eps=1e-3, bn_momentum=0.03, inplace_act=True, in_channels=3)
    m.load_state_dict({k: v.clone() for k, v in g["sd0"].items()}, strict=False)
    m = m.cuda().train()
    st = TrainStep(m, PPYoloELoss(num_classes=4, use_static_assigner=False), "SGD", {"weight_decay": 1e-5, "momentum": 0.9}, zero_wd_on_bias_and_bn=True, ema=True)
    st.batched_plumbing = batched
    return m, st


def trace_forward(m, x):
    """One no-grad forward with a hook on every module: [(qualified name, output clone)] in execution order."""
    rec, hooks = [], []

    def mk(name):
        def hook(_mod, _inp, out):
            if torch.is_tensor(out):
                rec.append((name, out.detach().clone()))
        return hook

    for name, mod in m.named_modules():
        if name:
            hooks.append(mod.register_forward_hook(mk(name)))
    with torch.no_grad():
        m(x)
    for h in hooks:
        h.remove()
    return rec


def first_divergence(ma, mb, x):
    """Runs both models (identical parameters) once and reports the first module, in execution order, whose outputs differ."""
    ra, rb = trace_forward(ma, x), trace_forward(mb, x)
    for (na, ta), (nb, tb) in zip(ra, rb):
        assert na == nb
        if ta.shape != tb.shape or not torch.equal(ta, tb):
            d = (ta.float() - tb.float()).abs()
            mod = dict(ma.named_modules())[na]
            return f"first differing module: {na} ({type(mod).__name__}) out {tuple(ta.shape)} max|d|={float(d.max()):.3e} at {int(d.argmax())} n_diff={int((d > 0).sum())}"
    return "all module outputs identical"


def main():
    g = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_yolo_nas.pt"), weights_only=False)
    x = g["x"].cuda()
    t = tuple(v.cuda() for v in pad_targets_host(g["targets"], x.shape[0], 16))
    ma, sa = make(g, os.environ.get("BATCHED_A", "0") == "1")
    mb, sb = make(g, False)
    for i in range(int(os.environ.get("STEPS", "4"))):
        sb.flat.params.copy_(sa.flat.params)
        sb.flat.buffers.copy_(sa.flat.buffers)
        for qa, qb in zip(sa.state, sb.state):
            qb.copy_(qa)
        SF.bump_weight_epoch()
        sa._filters_stale = True
        sa.set_hyper_params(1e-3, 0.99)
        sb.set_hyper_params(1e-3, 0.99)
        outs = []
        for m in (ma, mb, ma, mb):
            with torch.no_grad():
                (_pb, _ps), raw = m(x)
            outs.append(raw[0].clone())
        if rel(outs[0], outs[1]) > 0 and os.environ.get("LOCALISE", "1") == "1":
            print("   ", first_divergence(ma, mb, x), flush=True)
        la, _ = sa.forward_backward(x, t)
        lb, _ = sb.forward_backward(x, t)
        print(f"step {i}: forward cls a-b {rel(outs[0], outs[1]):.3e}  a-a {rel(outs[0], outs[2]):.3e}  b-b {rel(outs[1], outs[3]):.3e}   "
              f"loss a {float(la):.7f} b {float(lb):.7f}  grads rel {rel(sa.flat.grads, sb.flat.grads):.3e}", flush=True)  # fmt: skip
        sa.optimizer_step()
        sb.optimizer_step()
        sa.opt_steps += 1
        sb.opt_steps += 1


if __name__ == "__main__":
    main()
