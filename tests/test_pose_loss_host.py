"""Row L7 without a GPU: the arithmetic of the CUDA pose-loss kernels (csrc/pose_loss_math.cuh, compiled for the host by
tests/host_pose_loss.py) under the product's YoloNASPoseLoss module, against (1) the loss value, components and gradients
recorded from the UNMODIFIED reference (tests/golden/pose.pt) and (2) the oracle on seeded random cases that cover crowd
instances, invisible joints, images without instances and an empty batch.  The kernels' parallel schedule (pose_loss.cu) is
covered by the `-m gpu` tests."""
import shutil

import numpy as np
import pytest
import torch

from oracle import sg_oracle as O

import cpu_backend

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")

CASES = ["loss_default", "loss_oks_rescale_bce_giou", "loss_recipe"]


def _module_forward_backward(kw, sigmas, raw, targets, monkeypatch):
    from super_gradients_b200.training.losses import YoloNASPoseLoss

    cpu_backend.install_training(monkeypatch)
    crit = YoloNASPoseLoss(oks_sigmas=sigmas, **kw)
    leaves = [t.detach().clone().requires_grad_(True) for t in raw[:4]]
    loss, items = crit((None, (*leaves, *raw[4:])), targets)
    loss.backward()
    return loss.detach(), items, [t.grad for t in leaves]


@pytest.mark.parametrize("case", CASES)
def test_pose_loss_kernel_math_matches_the_reference(golden, monkeypatch, case):
    g = golden("pose")[case]
    loss, items, grads = _module_forward_backward(g["kw"], g["sigmas"], g["raw"], g["targets"], monkeypatch)
    torch.testing.assert_close(items, g["items"], rtol=2e-4, atol=1e-6)
    torch.testing.assert_close(loss, g["loss"], rtol=2e-4, atol=1e-6)
    for name, a, b in zip(("cls_logits", "reg_distri", "pose_coords", "pose_logits"), grads, g["grads"]):
        assert a is not None and float(b.abs().max()) > 0, name
        torch.testing.assert_close(a, b, rtol=2e-3, atol=2e-7 + 1e-4 * float(b.abs().max()), msg=lambda m, name=name: f"{name}: {m}")


def _random_case(seed, B=3, J=5, reg_max=7, sizes=((6, 6), (3, 3)), strides=(8, 16), n_inst=(3, 0, 2), crowd_every=2):
    gen = torch.Generator().manual_seed(seed)
    anchors, anchor_points, nums, stride_tensor = O.anchors_for_levels(sizes, strides)
    L = anchor_points.shape[0]
    img = sizes[0][0] * strides[0]
    boxes, joints, crowd = [], [], []
    k = 0
    for b, n in enumerate(n_inst):
        for _ in range(n):
            c = torch.rand(2, generator=gen) * img * 0.6 + img * 0.2
            wh = torch.rand(2, generator=gen) * img * 0.4 + img * 0.15
            x1y1, x2y2 = (c - wh / 2).clamp(0, img - 2), (c + wh / 2).clamp(2, img)
            boxes.append(torch.tensor([b, *x1y1.tolist(), *x2y2.tolist()]))
            jj = torch.rand(J, 2, generator=gen) * (x2y2 - x1y1) + x1y1
            vis = (torch.rand(J, generator=gen) > 0.3).float() * (1 + (torch.rand(J, generator=gen) > 0.5).float())
            joints.append(torch.cat([torch.full((J, 1), float(b)), jj, vis[:, None]], 1))
            crowd.append(torch.tensor([float(b), 1.0 if (k % crowd_every == crowd_every - 1) else 0.0]))
            k += 1
    if boxes:
        targets = (torch.stack(boxes), torch.stack(joints), torch.stack(crowd))
    else:
        targets = (torch.zeros(0, 5), torch.zeros(0, J, 4), torch.zeros(0, 2))
    cls_logits = torch.randn(B, L, 1, generator=gen) * 1.5 - 1.0
    reg_distri = torch.randn(B, L, 4 * (reg_max + 1), generator=gen) * 1.5
    pose_coords = torch.rand(B, L, J, 2, generator=gen) * img
    # a few predictions close to the targets so that OKS terms are not all ~0
    if boxes:
        for i in range(len(boxes)):
            b = int(targets[0][i, 0])
            sel = torch.randint(0, L, (6,), generator=gen)
            pose_coords[b, sel] = targets[1][i, :, 1:3] + torch.randn(6, J, 2, generator=gen) * 3.0
    pose_logits = torch.randn(B, L, J, generator=gen)
    raw = (cls_logits, reg_distri, pose_coords, pose_logits, anchors, anchor_points, nums, stride_tensor)
    sigmas = (torch.rand(J, generator=gen) * 0.08 + 0.025).tolist()
    return raw, targets, sigmas


KWS = [
    {},
    dict(classification_loss_type="bce", regression_iou_loss_type="giou", pose_classification_loss_type="focal"),
    dict(assigner_multiply_by_pose_oks=True, rescale_pose_loss_with_assigned_score=True),
    dict(assigner_multiply_by_pose_oks=True, pose_classification_loss_type="focal", dfl_loss_weight=0.01, pose_reg_loss_weight=34.0, bbox_assigner_topk=5),
]


def _oracle(raw, targets, sigmas, kw):
    names = dict(classification_loss_weight="w_cls", iou_loss_weight="w_iou", dfl_loss_weight="w_dfl", pose_cls_loss_weight="w_pose_cls", pose_reg_loss_weight="w_pose_reg",
                 bbox_assigner_topk="topk", bbox_assigned_alpha="alpha", bbox_assigned_beta="beta")  # fmt: skip
    okw = {names.get(k, k): v for k, v in kw.items()}
    leaves = [t.detach().clone().requires_grad_(True) for t in raw[:4]]
    loss, items = O.yolo_nas_pose_loss((*leaves, *raw[4:]), targets, sigmas, **okw)
    loss.backward()
    return loss.detach(), items, [t.grad if t.grad is not None else torch.zeros_like(t) for t in leaves]


@pytest.mark.parametrize("kw_i", range(len(KWS)))
@pytest.mark.parametrize("seed,n_inst", [(0, (3, 0, 2)), (1, (1, 4, 1)), (2, (0, 0, 5))])
def test_pose_loss_kernel_math_matches_the_oracle(monkeypatch, kw_i, seed, n_inst):
    kw = KWS[kw_i]
    raw, targets, sigmas = _random_case(seed, n_inst=n_inst)
    loss, items, grads = _module_forward_backward(kw, sigmas, raw, targets, monkeypatch)
    loss_e, items_e, grads_e = _oracle(raw, targets, sigmas, kw)
    torch.testing.assert_close(items, items_e, rtol=3e-4, atol=1e-6)
    for name, a, b in zip(("cls_logits", "reg_distri", "pose_coords", "pose_logits"), grads, grads_e):
        torch.testing.assert_close(a, b, rtol=3e-3, atol=3e-7 + 1e-4 * float(b.abs().max()), msg=lambda m, name=name: f"{name}: {m}")


def test_pose_loss_empty_batch(monkeypatch):
    raw, targets, sigmas = _random_case(5, n_inst=(0, 0, 0))
    loss, items, grads = _module_forward_backward({}, sigmas, raw, targets, monkeypatch)
    loss_e, items_e, grads_e = _oracle(raw, targets, sigmas, {})
    torch.testing.assert_close(items, items_e, rtol=3e-4, atol=1e-6)
    assert float(items[1]) == 0.0 and float(items[3]) == 0.0 and float(grads[1].abs().max()) == 0.0
    torch.testing.assert_close(grads[0], grads_e[0], rtol=3e-3, atol=1e-7)


def test_pad_pose_targets_matches_the_oracle_unpacking():
    from super_gradients_b200.training.losses import pad_pose_targets_host

    _, targets, _ = _random_case(3, n_inst=(2, 0, 3))
    gb, gp, gc, gv = pad_pose_targets_host(targets, 3, 6)
    _lab, eb, epad, epose, ecrowd = O.pose_unpack_targets(*targets, batch_size=3)
    n = eb.shape[1]
    np.testing.assert_array_equal(gb[:, :n].numpy(), eb.numpy())
    np.testing.assert_array_equal(gp[:, :n].numpy(), epose.numpy())
    np.testing.assert_array_equal(gv[:, :n].numpy(), epad[..., 0].numpy().astype(np.uint8))
    np.testing.assert_array_equal(gc[:, :n].numpy(), ecrowd[..., 0].numpy().astype(np.uint8))
    assert int(gv[:, n:].sum()) == 0
    with pytest.raises(ValueError):
        pad_pose_targets_host(targets, 3, 2)
