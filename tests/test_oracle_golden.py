"""Pins oracle/sg_oracle.py (the CPU restatement) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_goldens.py) and against the installed torchvision for the third-party NMS arithmetic."""
import numpy as np
import pytest
import torch

from oracle import sg_oracle as O


def _canon(rows):
    if rows.shape[0] == 0:
        return rows
    order = np.lexsort((rows[:, 5], rows[:, 3], rows[:, 2], rows[:, 1], rows[:, 0], -rows[:, 4]))
    return rows[order]


def _clone(sd):
    return {k: v.clone() for k, v in sd.items()}


@pytest.mark.parametrize("case", ["s1_res", "s2"])
def test_qarepvgg_train_eval_fused(golden, case):
    g = golden("qarepvgg")[case]
    p = _clone(g["sd0"])
    x = g["x"].clone().requires_grad_(True)
    for k in ("branch_3x3.conv.weight", "branch_3x3.bn.weight", "branch_3x3.bn.bias", "branch_1x1.weight", "branch_1x1.bias", "post_bn.weight", "post_bn.bias"):
        p[k].requires_grad_(True)
    y = O.qarepvgg_forward(x, p, "", g["stride"], g["residual"], "relu", True, 1e-3, 0.03)
    torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=1e-5)
    y.backward(g["gy"])
    torch.testing.assert_close(x.grad, g["gx"], rtol=1e-4, atol=1e-5)
    for k, v in g["grads"].items():
        torch.testing.assert_close(p[k].grad, v, rtol=1e-4, atol=2e-5)
    for k in g["sd1"]:
        if "running" in k:
            torch.testing.assert_close(p[k], g["sd1"][k], rtol=1e-5, atol=1e-6)
    p1 = _clone(g["sd1"])
    with torch.no_grad():
        torch.testing.assert_close(O.qarepvgg_forward(g["x"], p1, "", g["stride"], g["residual"], "relu", False, 1e-3, 0.03), g["y_eval"], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(O.qarepvgg_forward_fused(g["x"], p1, "", g["stride"], g["residual"], "relu", 1e-3, full=False), g["y_partial"], rtol=1e-4, atol=1e-4)
        torch.testing.assert_close(O.qarepvgg_forward_fused(g["x"], p1, "", g["stride"], g["residual"], "relu", 1e-3, full=True), g["y_full"], rtol=1e-4, atol=1e-4)


def test_conv_blocks(golden):
    G = golden("conv_blocks")
    g = G["conv3x3_s2"]
    torch.testing.assert_close(O.conv_bn_act(g["x"], _clone(g["sd0"]), "", 2, 1, "relu", True, 1e-5, 0.1), g["y"], rtol=1e-5, atol=1e-5)
    g = G["conv1x1"]
    torch.testing.assert_close(O.conv_bn_act(g["x"], _clone(g["sd0"]), "", 1, 0, "relu", True, 1e-5, 0.1), g["y"], rtol=1e-5, atol=1e-5)
    g = G["convbnrelu"]
    torch.testing.assert_close(O.conv_bn_act(g["x"], _clone(g["sd0"]), "seq.", 1, 1, "relu", True, 1e-5, 0.1), g["y"], rtol=1e-5, atol=1e-5)
    g = G["bottleneck_s2"]
    torch.testing.assert_close(O.resnet_bottleneck(g["x"], _clone(g["sd0"]), "", 2, True, True), g["y"], rtol=1e-5, atol=1e-5)
    g = G["bottleneck_id"]
    torch.testing.assert_close(O.resnet_bottleneck(g["x"], _clone(g["sd0"]), "", 1, False, True), g["y"], rtol=1e-5, atol=1e-5)
    g = G["basic_s2"]
    torch.testing.assert_close(O.resnet_basic_block(g["x"], _clone(g["sd0"]), "", 2, True, True), g["y"], rtol=1e-5, atol=1e-5)
    g = G["spp"]
    torch.testing.assert_close(O.spp(g["x"], _clone(g["sd0"]), "", (5, 9, 13), "relu", True, 1e-5, 0.1), g["y"], rtol=1e-5, atol=1e-5)


def test_resnet50_oracle_matches_reference(golden):
    """oracle/resnet_oracle.py (the CPU arm of bench.py --config 4) == the unmodified reference's ResNet-50 for its own seeded
    initialisation: train-mode logits, loss, the recorded gradients and every parameter's gradient norm; eval-mode logits.  The state
    dict is the product constructor's (it consumes the RNG exactly as the reference's does -- the init fingerprints are pinned in
    test_host_logic.py); only nn.Module construction runs here, no kernel."""
    from oracle import resnet_oracle as R
    from super_gradients_b200.training import models

    g = golden("other_configs")["resnet50"]
    torch.manual_seed(0)
    m = models.get("resnet50", num_classes=1000)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    live = [k for k, _ in m.named_parameters()]
    x = g["x"].float()
    torch.testing.assert_close(R.resnet_forward("resnet50", {k: v.clone() for k, v in sd.items()}, x, True), g["train_logits"], rtol=2e-3, atol=2e-3)
    loss, grads = R.train_step("resnet50", sd, x, g["y"], live)  # updates the BatchNorm running statistics held in `sd` (once)
    torch.testing.assert_close(loss, g["loss"], rtol=1e-4, atol=1e-5)
    for k, ref in g["grads"].items():
        mine = grads[k] if grads[k].numel() < 20000 else grads[k].flatten()[:: grads[k].numel() // 10000]
        assert float((mine - ref).norm() / ref.norm()) < 2e-2, k
    worst = max(abs(float(grads[k].norm()) / v - 1) for k, v in g["grad_norms"].items() if v > 1e-6)
    assert worst < 2e-2, worst
    # eval: with the running statistics of that one training forward
    torch.testing.assert_close(R.resnet_forward("resnet50", sd, x, False), g["eval_logits"], rtol=2e-3, atol=2e-3)


def test_resnet_blocks_with_drop_path(golden):
    """Config 4 runs with droppath_prob 0.05: the oracle's blocks with the recorded per-image scale == the unmodified reference
    (outputs and, through autograd, input / parameter gradients)."""
    G = golden("droppath")
    for name, fn, args in (("bottleneck_s2", O.resnet_bottleneck, (2, True)), ("bottleneck_id", O.resnet_bottleneck, (1, False)), ("basic_s2", O.resnet_basic_block, (2, True))):
        g = G[name]
        p = _clone(g["sd0"])
        for k in g["grads"]:
            p[k].requires_grad_(True)
        x = g["x"].clone().requires_grad_(True)
        y = fn(x, p, "", args[0], args[1], True, sample_scale=g["scale"])
        torch.testing.assert_close(y, g["y"], rtol=1e-5, atol=1e-5)
        y.backward(g["gy"])
        torch.testing.assert_close(x.grad, g["gx"], rtol=1e-4, atol=1e-5)
        for k, v in g["grads"].items():
            torch.testing.assert_close(p[k].grad, v, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("case", ["regular", "ragged_with_empty", "no_targets"])
def test_loss_and_assigner(golden, case):
    G = golden("loss")
    g = G[case]
    cls = g["cls_logits"].clone().requires_grad_(True)
    reg = g["reg_distri"].clone().requires_grad_(True)
    raw = (cls, reg, G["anchors"], G["anchor_points"], G["nums"], G["stride_tensor"])
    loss, items, (al, ab, asc) = O.ppyoloe_loss(raw, g["targets"], 4, return_assignment=True)
    torch.testing.assert_close(loss, g["loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(items, g["items"], rtol=1e-5, atol=1e-6)
    assert torch.equal(al, g["assigned_labels"])
    torch.testing.assert_close(asc, g["assigned_scores"], rtol=1e-5, atol=1e-7)
    pos = al != 4
    torch.testing.assert_close(ab[pos], g["assigned_bboxes"][pos])
    loss.backward()
    torch.testing.assert_close(cls.grad, g["g_cls"], rtol=1e-4, atol=1e-7)
    torch.testing.assert_close(reg.grad, g["g_reg"], rtol=1e-4, atol=1e-7)
    gc, gb, pm = O.pad_targets(g["targets"], 3)
    assert torch.equal(gc, g["gt_class"]) and torch.equal(pm, g["pad_gt_mask"])
    torch.testing.assert_close(gb, g["gt_bbox"])


def test_box_losses(golden):
    g = golden("loss")["boxes"]
    p = g["p"].clone().requires_grad_(True)
    l = O.giou_loss(p, g["g"])
    torch.testing.assert_close(l, g["giou"], rtol=1e-5, atol=1e-6)
    l.sum().backward()
    torch.testing.assert_close(p.grad, g["g_giou"], rtol=1e-4, atol=1e-6)
    p.grad = None
    l = O.ciou_loss(p, g["g"])
    torch.testing.assert_close(l, g["ciou"], rtol=1e-5, atol=1e-6)
    l.sum().backward()
    torch.testing.assert_close(p.grad, g["g_ciou"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("case", ["multi_small", "multi_topk", "multi_vanilla", "single_label", "class_agnostic", "nothing_passes"])
def test_postprocess_matches_reference_callback(golden, case):
    g = golden("nms")[case]
    res, _ = O.ppyoloe_postprocess(g["boxes"], g["scores"], **g["params"])
    assert len(res) == len(g["result"])
    for mine, ref in zip(res, g["result"]):
        assert mine.shape == tuple(ref.shape), (mine.shape, ref.shape)
        # torch.topk / unstable sorts leave the order of EXACTLY tied scores implementation-defined in the reference
        # itself (torchvision documents the same for nms): canonicalise the order inside tie groups before comparing.
        np.testing.assert_array_equal(_canon(mine), _canon(ref.numpy()))


def test_nms_numpy_matches_installed_torchvision():
    import torchvision

    gen = torch.Generator().manual_seed(7)
    for n, c in [(0, 1), (1, 1), (57, 3), (700, 4), (1500, 5)]:
        xy = torch.rand(n, 2, generator=gen) * 100
        wh = torch.rand(n, 2, generator=gen) * 40 + 1
        boxes = torch.cat([xy, xy + wh], -1)
        scores = (torch.rand(n, generator=gen) * 20).round() / 20  # plenty of exact ties
        idxs = torch.randint(0, c, (n,), generator=gen)
        ref = torchvision.ops.nms(boxes, scores, 0.5)
        np.testing.assert_array_equal(O.nms_numpy(boxes.numpy(), scores.numpy(), 0.5), ref.numpy())
        ref = torchvision.ops.batched_nms(boxes, scores, idxs, 0.5)
        mine = O.batched_nms_numpy(boxes.numpy(), scores.numpy(), idxs.numpy(), 0.5)
        if boxes.numel() > 4000:  # vanilla path ends with an unstable sort: compare as (score-ordered) sets
            assert sorted(mine.tolist()) == sorted(ref.tolist())
            np.testing.assert_array_equal(scores.numpy()[mine], scores.numpy()[ref.numpy()])
        else:
            np.testing.assert_array_equal(mine, ref.numpy())


def test_decode_matches_tiny_model_outputs(golden):
    """ndfl_decode on the tiny model's raw outputs reproduces its decoded outputs (dfl_heads.py:199-245)."""
    g = golden("tiny_yolo_nas")
    B = g["x"].shape[0]
    # rebuild per-level NCHW head outputs from the raw [B, L, *] tensors
    shapes, strides = [(16, 16), (8, 8), (4, 4)], (8, 16, 32)
    regs, clss, a0 = [], [], 0
    for h, w in shapes:
        n = h * w
        regs.append(g["train_reg_distri"][:, a0 : a0 + n].permute(0, 2, 1).reshape(B, -1, h, w))
        clss.append(g["train_cls_logits"][:, a0 : a0 + n].permute(0, 2, 1).reshape(B, -1, h, w))
        a0 += n
    (pb, ps), raw = O.ndfl_decode(regs, clss, strides)
    torch.testing.assert_close(pb, g["train_pred_bboxes"], rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(ps, g["train_pred_scores"], rtol=1e-5, atol=1e-6)
    loss, items = O.ppyoloe_loss(raw, g["targets"], 4)
    torch.testing.assert_close(loss, g["loss"], rtol=1e-5, atol=1e-6)


def test_tiny_yolo_nas_whole_graph(golden):
    """oracle/yolo_nas_oracle.py reproduces the reference's whole-model train outputs, loss and gradients."""
    from oracle.yolo_nas_oracle import YoloNASOracle, train_step

    g = golden("tiny_yolo_nas")
    state = {k: v.clone() for k, v in g["sd0"].items()}
    (pb, ps), raw = YoloNASOracle(g["arch"], {k: v.clone() for k, v in state.items()}, training=True).forward(g["x"])
    torch.testing.assert_close(raw[0], g["train_cls_logits"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(raw[1], g["train_reg_distri"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(pb, g["train_pred_bboxes"], rtol=1e-4, atol=1e-3)
    live = [k for k in g["param_names"] if "rbr_reparam" not in k]
    loss, items, grads = train_step(g["arch"], state, g["x"], g["targets"], 4, live)
    torch.testing.assert_close(loss, g["loss"], rtol=1e-4, atol=1e-6)
    for k, v in g["grads"].items():
        torch.testing.assert_close(grads[k], v, rtol=2e-3, atol=1e-5)
    for k, (s, n) in g["grad_sums"].items():
        assert abs(float(grads[k].double().norm()) - n) <= 2e-3 * n + 1e-6, k
    # running statistics after the training forward, then eval-mode outputs
    for k, v in g["running1"].items():
        torch.testing.assert_close(state[k], v, rtol=1e-4, atol=1e-6)
    (eb, es), _ = YoloNASOracle(g["arch"], {**g["sd0"], **g["running1"]}, training=False).forward(g["x"])
    torch.testing.assert_close(es, g["eval_pred_scores"], rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(eb, g["eval_pred_bboxes"], rtol=1e-4, atol=1e-3)


def test_bf16_emulation_sensitivity(golden):
    """Justifies the whole-graph GPU tolerances: two bf16 emulations of the SAME graph that differ only in the
    accumulation precision of the sums (fp32 vs fp64) already diverge by ~0.7 % (cls), ~6 % (reg), ~0.6 % (boxes) in the
    forward outputs, and -- because the task-aligned assigner is a discrete top-k over those outputs and the deep maps are
    4x4 with train-mode BatchNorm -- by a MEDIAN of ~50 % in the per-parameter gradients (norms agree, directions do not).
    The whole-graph gradient comparison of tests/test_modules_gpu.py is therefore only a sanity bound; gradient parity is
    carried by the per-block tests (tight tolerances against this oracle) and the kernel tests."""
    from oracle.yolo_nas_oracle import YoloNASOracle

    g = golden("tiny_yolo_nas")
    rel = lambda a, b: float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12))  # noqa: E731
    live = [k for k in g["param_names"] if "rbr_reparam" not in k]

    def run(dt):
        with O.bf16_emulation():
            pe = {k: (v.detach().clone().to(dt) if v.dtype.is_floating_point else v.clone()) for k, v in g["sd0"].items()}
            for k in live:
                pe[k].requires_grad_(True)
            keep = O._r
            if dt == torch.float64:
                O._r = lambda t: t.bfloat16().to(t.dtype)
            try:
                (pb, _), raw = YoloNASOracle(g["arch"], pe, True).forward(g["x"].to(dt))
                loss, _ = O.ppyoloe_loss(raw, g["targets"], 4)
                loss.backward()
            finally:
                O._r = keep
        return pe, pb, raw

    p32, pb, raw = run(torch.float32)
    p64, pb2, raw2 = run(torch.float64)
    spread = dict(cls=rel(raw2[0], raw[0]), reg=rel(raw2[1], raw[1]), boxes=rel(pb2, pb))
    assert 1e-3 < spread["cls"] < 7.5e-3 and 1e-2 < spread["reg"] < 6.5e-2 and spread["boxes"] < 1e-2, spread
    errs = sorted(rel(p64[k].grad, p32[k].grad) for k in live if p32[k].grad is not None and p32[k].grad.norm() > 1e-6)
    norm_ratio = sorted(abs(float(torch.log(p64[k].grad.float().norm() / p32[k].grad.norm()))) for k in live if p32[k].grad is not None and p32[k].grad.norm() > 1e-6)
    assert 0.3 < errs[len(errs) // 2] < 0.65, errs[len(errs) // 2]      # measured 0.51
    assert norm_ratio[len(norm_ratio) // 2] < 0.1, norm_ratio[len(norm_ratio) // 2]


@pytest.mark.parametrize("case", ["regular", "topk", "few_joints", "nothing_passes"])
def test_pose_postprocess_oracle_matches_reference(golden, case):
    """oracle.yolo_nas_pose_postprocess == the unmodified YoloNASPosePostPredictionCallback (row N2): same instances in
    the same order (scores are continuous random numbers: no exact ties), bit-identical boxes / scores / poses."""
    g = golden("pose_nms")[case]
    res, idx = O.yolo_nas_pose_postprocess(g["boxes"], g["conf"], g["coords"], g["jscores"], **g["params"])
    assert len(res) == len(g["result"])
    for (poses, scores, boxes), kept, (rp, rs, rb) in zip(res, idx, g["result"]):
        np.testing.assert_array_equal(scores, rs.numpy())
        np.testing.assert_array_equal(boxes, rb.numpy())
        np.testing.assert_array_equal(poses, rp.numpy())
        assert poses.shape[0] == kept.shape[0] <= g["params"]["post_nms_max_predictions"]
    if case == "topk":
        assert any(int((c.reshape(-1) >= g["params"]["pose_confidence_threshold"]).sum()) > g["params"]["pre_nms_max_predictions"] for c in g["conf"])
    if case == "nothing_passes":
        assert all(r[0].shape[0] == 0 for r in res)


def test_pose_decode_oracle_matches_reference(golden):
    """Row L8: oracle.pose_ndfl_decode == YoloNASPoseNDFLHeads.forward of the unmodified reference (per-level head outputs
    captured from a real yolo_nas_pose_n forward)."""
    g = golden("pose")["decode"]
    lv = g["levels"]
    decoded, raw = O.pose_ndfl_decode([t[0] for t in lv], [t[1] for t in lv], [t[2] for t in lv], [t[3] for t in lv], g["strides"], g["reg_max"],
                                      g["cell_offset"], g["cell_scale"], g["pose_offset_multiplier"], g["compensate"])  # fmt: skip
    for mine, ref in zip(decoded, g["decoded"]):
        torch.testing.assert_close(mine, ref, rtol=1e-5, atol=1e-4)
    for i in (0, 1, 2, 3, 4, 5, 7):
        torch.testing.assert_close(raw[i], g["raw"][i], rtol=1e-5, atol=1e-4)
    assert list(raw[6]) == list(g["raw"][6])


@pytest.mark.parametrize("case", ["default", "oks_rescale_bce_giou", "recipe"])
def test_pose_loss_oracle_matches_reference(golden, case):
    """Row L7: oracle.yolo_nas_pose_loss (assigner with crowd handling / optional OKS weighting, focal or BCE classification,
    CIoU or GIoU, DFL, keypoint OKS regression + visibility classification) == YoloNASPoseLoss: value, components, gradients."""
    g = golden("pose")["loss_" + case]
    raw = list(g["raw"])
    leaves = [raw[i].clone().requires_grad_(True) for i in range(4)]
    kw = dict(g["kw"])
    loss, items = O.yolo_nas_pose_loss(
        tuple(leaves) + tuple(raw[4:]), g["targets"], g["sigmas"], classification_loss_type=kw.get("classification_loss_type", "focal"),
        regression_iou_loss_type=kw.get("regression_iou_loss_type", "ciou"), pose_classification_loss_type=kw.get("pose_classification_loss_type", "bce"),
        assigner_multiply_by_pose_oks=kw.get("assigner_multiply_by_pose_oks", False), rescale_pose_loss_with_assigned_score=kw.get("rescale_pose_loss_with_assigned_score", False),
        w_cls=kw.get("classification_loss_weight", 1.0), w_iou=kw.get("iou_loss_weight", 2.5), w_dfl=kw.get("dfl_loss_weight", 0.5), w_pose_cls=kw.get("pose_cls_loss_weight", 1.0),
        w_pose_reg=kw.get("pose_reg_loss_weight", 1.0),
    )  # fmt: skip
    torch.testing.assert_close(loss, g["loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(items, g["items"], rtol=1e-5, atol=1e-6)
    assert float(items[3]) > 0 and float(items[4]) > 0  # the keypoint terms are live in the fixture
    loss.backward()
    for leaf, ref in zip(leaves, g["grads"]):
        torch.testing.assert_close(leaf.grad, ref, rtol=1e-4, atol=1e-7)


def test_tiny_yolo_nas_pose_oracle_matches_reference(golden):
    """Whole-graph oracle with the pose heads (eval mode, fp32) == the reference YoloNASPose built from the same arch and
    state dict: decoded boxes / scores / keypoints, raw head outputs, and the post-prediction callback's instances."""
    from oracle.yolo_nas_oracle import YoloNASOracle

    g = golden("tiny_yolo_nas_pose")
    decoded, raw = YoloNASOracle(g["arch"], {k: v.clone() for k, v in g["sd0"].items()}, training=False).forward(g["x"])
    for mine, ref in zip(decoded, g["decoded"]):
        torch.testing.assert_close(mine, ref, rtol=1e-4, atol=1e-3)
    for i in (0, 1, 2, 3, 5, 7):
        torch.testing.assert_close(raw[i], g["raw"][i], rtol=1e-4, atol=1e-3)
    res, _ = O.yolo_nas_pose_postprocess(*g["decoded"], **g["cb"])
    assert len(res) == len(g["preds"]) and sum(r[0].shape[0] for r in res) > 0
    for (poses, scores, boxes), (rp, rs, rb) in zip(res, g["preds"]):
        np.testing.assert_array_equal(poses, rp.numpy())
        np.testing.assert_array_equal(scores, rs.numpy())
        np.testing.assert_array_equal(boxes, rb.numpy())


_POSE_KW = dict(classification_loss_weight="w_cls", iou_loss_weight="w_iou", dfl_loss_weight="w_dfl", pose_cls_loss_weight="w_pose_cls", pose_reg_loss_weight="w_pose_reg",
                bbox_assigner_topk="topk", bbox_assigned_alpha="alpha", bbox_assigned_beta="beta")  # fmt: skip


def pose_oracle_train_step(arch, sd0, x, targets, sigmas, kw):
    """Whole-graph oracle, train mode: forward, YoloNASPoseLoss restatement, backward.  Returns (loss, items, raw, params)."""
    from oracle.yolo_nas_oracle import YoloNASOracle

    p = {k: v.clone() for k, v in sd0.items()}
    for k, v in p.items():
        if v.dtype.is_floating_point and "running_" not in k:
            v.requires_grad_(True)
    _decoded, raw = YoloNASOracle(arch, p, training=True).forward(x)
    loss, items = O.yolo_nas_pose_loss(raw, targets, sigmas, **{_POSE_KW.get(k, k): v for k, v in kw.items()})
    loss.backward()
    return loss.detach(), items, raw, p


def test_tiny_yolo_nas_pose_train_oracle_matches_reference(golden):
    """Row L7 end to end in fp32: train-mode whole-graph oracle + pose-loss restatement == the unmodified reference on the
    same tiny model (loss, components, raw head outputs, updated BatchNorm statistics, every parameter's gradient)."""
    g0, g = golden("tiny_yolo_nas_pose"), golden("tiny_yolo_nas_pose_train")
    loss, items, raw, p = pose_oracle_train_step(g0["arch"], g0["sd0"], g["x"], g["targets"], g["sigmas"], g["kw"])
    torch.testing.assert_close(items, g["items"], rtol=2e-4, atol=1e-6)
    for mine, ref in zip(raw[:4], g["raw"]):
        torch.testing.assert_close(mine.detach(), ref, rtol=1e-4, atol=2e-3)
    for k, v in g["running1"].items():
        torch.testing.assert_close(p[k], v, rtol=1e-4, atol=1e-5)
    for k, ref in g["grads"].items():
        torch.testing.assert_close(p[k].grad, ref, rtol=2e-3, atol=2e-6 + 1e-3 * float(ref.abs().max()))
    live = [k for k in g["grad_sums"] if k in p]
    assert len(live) == len(g["grad_sums"])
    zero_ref = {k for k, v in g["grad_sums"].items() if tuple(v) == (0.0, 0.0)}
    zero_mine = {k for k in live if p[k].grad is None or float(p[k].grad.abs().sum()) == 0.0}
    assert zero_mine == zero_ref
    for k in live:
        if k not in zero_ref:
            # (biases in front of a train-mode BatchNorm have an identically zero gradient: both sides hold fp32 noise ~1e-6)
            n_ref = g["grad_sums"][k][1]
            assert abs(float(p[k].grad.double().norm()) - n_ref) <= 5e-3 * n_ref + 2e-5, k


@pytest.mark.parametrize("case", ["multi_conf", "multi_raw", "single", "agnostic", "one_empty_image", "nothing_passes"])
def test_yolox_nms_oracle_matches_reference(golden, case):
    """Row N3: the YoloX-format non_max_suppression restatement == the unmodified reference (rows, order, None for empty)."""
    g = golden("yolox_nms")[case]
    res = O.yolox_non_max_suppression(g["pred"], **g["kw"])
    assert len(res) == len(g["result"])
    for mine, ref in zip(res, g["result"]):
        assert (mine is None) == (ref is None)
        if ref is not None:
            np.testing.assert_array_equal(mine, ref.numpy())


def test_bench_cpu_port_is_the_reference_train_step(golden):
    """`bench.py --impl reference` and the `cpu_baseline` leg time the oracle PORT of the train step (the reference tree does not exist
    on the GPU box).  This pins that port at the benchmarked workload itself -- YOLO-NAS-S, the bench's synthetic batch, a 2-image
    640 x 640 sample, the state `bench.cpu_step_fn` starts from -- against the UNMODIFIED reference run once in the build container
    (tests/golden/make_goldens.py::golden_port_fidelity, via oracle/ref_shim.py): loss, its three components, checksums of the raw
    head outputs and every parameter's gradient norm.  Nothing is timed here."""
    import bench
    from oracle.yolo_nas_oracle import random_state, train_step

    g = golden("port_fidelity_2x640")
    table = golden("state_keys")
    state = random_state(table["yolo_nas_s"], seed=0)
    live = [k for k in table["yolo_nas_s/param_names"] if "rbr_reparam" not in k]
    x, t = bench.synth_batch(2, 123, 640)
    assert int(t.shape[0]) == g["n_targets"]
    loss, items, grads = train_step(bench._arch_yaml("yolo_nas_s"), state, x, t, bench.NCLS, live)
    assert abs(float(loss) - g["loss"]) <= 1e-4 * abs(g["loss"]), (float(loss), g["loss"])
    for mine, ref in zip(items.reshape(-1).tolist(), g["items"]):
        assert abs(mine - ref) <= 1e-4 * abs(ref) + 1e-6, (mine, ref)
    assert set(grads) == set(g["grad_norms"]), set(grads) ^ set(g["grad_norms"])
    # Parameters whose gradient is identically zero -- biases in front of a BatchNorm (branch_3x3.bn.bias / branch_1x1.bias of a
    # QARepVGG block behind post_bn, the up-sampling ConvTranspose's bias) and the regression branches of pyramid levels without a
    # positive anchor in this sample -- hold fp32 round-off (< 2e-4 against norms up to 3.6e3) or an exact 0 on both sides.
    scale = max(g["grad_norms"].values())
    noise = [k for k in grads if g["grad_norms"][k] < 1e-6 * scale]  # 3.6e-3; the largest round-off norm is 1.1e-3 (stem, 204 800 pixels per image)
    assert any(k.endswith("branch_1x1.bias") for k in noise) and all(float(grads[k].norm()) < 1e-6 * scale for k in noise)
    assert not any(k.endswith(".weight") and "bn" not in k for k in noise if g["grad_norms"][k] > 0)  # no filter is treated as noise
    worst = max(abs(float(v.norm()) - g["grad_norms"][k]) / g["grad_norms"][k] for k, v in grads.items() if k not in noise)
    assert worst < 5e-4, worst  # measured 5.4e-5


def test_bench_cpu_port_is_the_reference_resnet50_train_step(golden):
    """The same pin for config 4's CPU arm (`bench.cpu_step_fn`, kind train_cls): ResNet-50 at 224 x 224 on a 4-image sample of the
    bench's classification batch, from the port's own initial state, drop-path off on both sides (its masks are pinned by droppath.pt):
    loss, a checksum of the logits and every parameter's gradient norm against the unmodified reference."""
    import bench
    from oracle import resnet_oracle as R
    from oracle.yolo_nas_oracle import random_state

    g = golden("port_fidelity_2x640")["resnet50"]
    table = golden("state_keys")
    state = random_state(table["resnet50"], seed=0)
    live = list(table["resnet50/param_names"])
    x, y = bench.synth_cls_batch(4, 123, 224)
    loss, grads = R.train_step("resnet50", state, x, y, live, droppath_prob=0.0)
    assert abs(float(loss) - g["loss"]) <= 1e-4 * abs(g["loss"]), (float(loss), g["loss"])
    assert set(grads) == set(g["grad_norms"])
    scale = max(g["grad_norms"].values())
    noise = [k for k in grads if g["grad_norms"][k] < 1e-6 * scale]
    assert all(float(grads[k].norm()) < 1e-5 * scale for k in noise)
    worst = max(abs(float(v.norm()) - g["grad_norms"][k]) / g["grad_norms"][k] for k, v in grads.items() if k not in noise)
    assert worst < 2e-3, worst
