"""TrainStep on the GPU: batched per-step plumbing (zero arena, batched filter re-layout, deferred gradient layout change,
foreach counters) against the per-layer launches, and CUDA-graph replay against eager execution."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(g, seed_offset=0):
    from super_gradients_b200.training.models.detection_models.yolo_nas import YoloNAS

    ap = copy.deepcopy(g["arch"])
    m = YoloNAS(backbone=ap["backbone"], neck=ap["neck"], heads=ap["heads"], num_classes=4, bn_eps=1e-3, bn_momentum=0.03, inplace_act=True, in_channels=3)
    m.load_state_dict({k: v.clone() for k, v in g["sd0"].items()}, strict=False)
    return m.to(DEV).train()


def _step(g, **kw):
    from super_gradients_b200.training.losses import PPYoloELoss
    from super_gradients_b200.training.sg_trainer import TrainStep

    m = _model(g)
    crit = PPYoloELoss(num_classes=4, use_static_assigner=False)
    # SGD: the update is linear in the gradient, so fp32 atomics-order noise stays noise (AdamW's m / sqrt(v) turns the
    # sign of a noise-level gradient into a full +-lr step, which makes two correct runs drift apart)
    st = TrainStep(m, crit, "SGD", {"weight_decay": 1e-5, "momentum": 0.9}, zero_wd_on_bias_and_bn=True, ema=True)
    for k, v in kw.items():
        setattr(st, k, v)
    return m, crit, st


def _targets(g):
    from super_gradients_b200.training.losses import pad_targets_host

    gb, gl, gv = pad_targets_host(g["targets"], g["x"].shape[0], 16)
    return gb.to(DEV), gl.to(DEV), gv.to(DEV)


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20))


def test_batched_plumbing_matches_per_layer_launches(golden):
    """Three steps with the batched plumbing, then the same three steps with one launch per layer, every step started from
    the recorded state of the first run.  The 4x4-map fixture amplifies fp32/fp64 atomics-order noise chaotically (see
    test_bf16_emulation_sensitivity: +-1 bf16 ulp already moves gradients by a median of 50 %), so the bounds below only
    separate "same computation" from a wrong work table / missing layer / stale filter (those give errors >= 1)."""
    from super_gradients_b200 import functional as SF

    g = golden("tiny_yolo_nas")
    x, t = g["x"].to(DEV), _targets(g)
    ma, _, sa = _step(g, batched_plumbing=True)
    trace = []
    for i in range(3):  # step 1 sizes the arena, steps 2-3 run from it with the cached work tables
        before = (sa.flat.params.clone(), sa.flat.buffers.clone(), [q.clone() for q in sa.state])
        sa.set_hyper_params(1e-3, 0.99)
        la, _ = sa.forward_backward(x, t)
        grads = sa.flat.grads.clone()
        sa.optimizer_step()
        sa.opt_steps += 1
        trace.append((before, float(la), grads, sa.flat.params.clone(), sa.flat.buffers.clone()))
    assert sa.arena.buf is not None and sa.ctx.weight_table is not None and sa.ctx.wgrad_table is not None  # really batched
    nbt_a = {k: int(v) for k, v in ma.state_dict().items() if k.endswith("num_batches_tracked")}
    del ma, sa
    mb, _, sb = _step(g, batched_plumbing=False)
    for i, (before, la, grads, params_after, buffers_after) in enumerate(trace):
        sb.flat.params.copy_(before[0])
        sb.flat.buffers.copy_(before[1])
        for q, r in zip(sb.state, before[2]):
            q.copy_(r)
        SF.bump_weight_epoch()
        sb.set_hyper_params(1e-3, 0.99)
        lb, _ = sb.forward_backward(x, t)
        assert abs(la - float(lb)) <= 1e-2 * abs(float(lb)), (i, la, float(lb))
        assert rel(grads, sb.flat.grads) < 0.6, (i, rel(grads, sb.flat.grads))
        sb.optimizer_step()
        sb.opt_steps += 1
        assert rel(params_after, sb.flat.params) < 1e-3, i
        assert rel(buffers_after, sb.flat.buffers) < 1e-2, i
    nbt_b = {k: int(v) for k, v in mb.state_dict().items() if k.endswith("num_batches_tracked")}
    assert nbt_a == nbt_b and set(nbt_a.values()) == {3}


def test_cuda_graph_replay_matches_eager(golden):
    g = golden("tiny_yolo_nas")
    x, t = g["x"].to(DEV), _targets(g)
    ma, _, sa = _step(g)
    mb, _, sb = _step(g)
    sb.set_hyper_params(1e-3, 0.99)
    sb.capture(x, t, warmup=2)
    # capture() restores parameters, optimizer moments, EMA, BatchNorm buffers and the step counter after its warm-up steps:
    # the captured twin starts where the eager twin starts
    assert sa.opt_steps == sb.opt_steps == 0
    assert torch.equal(sa.flat.params, sb.flat.params) and torch.equal(sa.flat.buffers, sb.flat.buffers)
    for i in range(3):
        sa.set_hyper_params(1e-3, 0.99)
        sb.set_hyper_params(1e-3, 0.99)
        la, _ = sa.run(x, t)
        lb, _ = sb.run(x, t)
        assert abs(float(la) - float(lb)) <= 2e-2 * abs(float(la)), (i, float(la), float(lb))
    assert rel(sb.flat.params, sa.flat.params) < 1e-3
    assert rel(sb.ema_params, sa.ema_params) < 1e-3
