"""The Python glue above kernels.py, run on the CPU through tests/cpu_backend.py (a stand-in for the kernel wrappers):
module wiring, head decoding, post-prediction callbacks and predict() of the detection and pose models against the
whole-graph oracle.  The CUDA kernels themselves are NOT exercised here (see the `-m gpu` tests)."""
import copy

import numpy as np
import pytest
import torch

from oracle import sg_oracle as O
from oracle.yolo_nas_oracle import YoloNASOracle

import cpu_backend


def l2rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def _load(m, sd):
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("rbr_reparam" in k for k in missing)
    return m.eval()


def test_yolo_nas_eval_glue_matches_oracle(golden, monkeypatch):
    from super_gradients_b200.training.models.detection_models.yolo_nas import YoloNAS

    cpu_backend.install(monkeypatch)
    g = golden("tiny_yolo_nas")
    ap = copy.deepcopy(g["arch"])
    m = YoloNAS(backbone=ap["backbone"], neck=ap["neck"], heads=ap["heads"], num_classes=4, bn_eps=1e-3, bn_momentum=0.03, inplace_act=True, in_channels=3)
    sd = {**g["sd0"], **g["running1"]}
    _load(m, sd)
    with torch.no_grad():
        (eb, es), raw = m(g["x"])
    with O.bf16_emulation():
        (ebe, ese), rawe = YoloNASOracle(g["arch"], {k: v.clone() for k, v in sd.items()}, training=False).forward(g["x"])
    assert l2rel(es, ese) < 2e-2 and l2rel(eb, ebe) < 2e-2 and l2rel(raw[1], rawe[1]) < 3e-2
    assert l2rel(es, g["eval_pred_scores"]) < 0.1 and l2rel(eb, g["eval_pred_bboxes"]) < 0.1  # the unmodified reference (fp32)
    torch.testing.assert_close(raw[3], rawe[3])  # anchor points


def test_yolo_nas_pose_eval_predict_glue_matches_oracle(golden, monkeypatch):
    from super_gradients_b200.training.models.pose_estimation_models import YoloNASPose, YoloNASPosePostPredictionCallback

    cpu_backend.install(monkeypatch)
    g = golden("tiny_yolo_nas_pose")
    ap = copy.deepcopy(g["arch"])
    m = YoloNASPose(backbone=ap["backbone"], neck=ap["neck"], heads=ap["heads"], num_classes=5, bn_eps=1e-3, bn_momentum=0.03, inplace_act=True, in_channels=3)
    _load(m, g["sd0"])
    with torch.no_grad():
        decoded, raw = m(g["x"])
    with O.bf16_emulation():
        dec_e, raw_e = YoloNASOracle(g["arch"], {k: v.clone() for k, v in g["sd0"].items()}, training=False).forward(g["x"])
    for name, a, b in zip(("boxes", "scores", "pose_coords", "pose_scores"), decoded, dec_e):
        assert tuple(a.shape) == tuple(b.shape), name
        assert l2rel(a, b) < 2e-2, (name, l2rel(a, b))
    for i in (0, 1, 3):
        assert l2rel(raw[i], raw_e[i]) < 3e-2, (i, l2rel(raw[i], raw_e[i]))
    for i in (4, 5, 7):
        torch.testing.assert_close(raw[i], raw_e[i])
    assert list(raw[6]) == list(raw_e[6])
    for name, a, b in zip(("boxes", "scores", "pose_coords", "pose_scores"), decoded, g["decoded"]):  # unmodified reference, fp32
        assert l2rel(a, b) < 0.1, (name, l2rel(a, b))
    # callback + predict() on the model's own outputs == the oracle post-processing of the same tensors
    cb = YoloNASPosePostPredictionCallback(**g["cb"])
    preds = cb((decoded, raw))
    ref, _ = O.yolo_nas_pose_postprocess(*decoded, **g["cb"])
    assert sum(r[0].shape[0] for r in ref) > 0
    for pr, (rposes, rscores, rboxes) in zip(preds, ref):
        np.testing.assert_array_equal(pr.scores.numpy(), rscores)
        np.testing.assert_array_equal(pr.bboxes_xyxy.numpy(), rboxes)
        np.testing.assert_array_equal(pr.poses.numpy(), rposes)
    out = m.predict(g["x"], conf=g["cb"]["pose_confidence_threshold"], iou=g["cb"]["nms_iou_threshold"], pre_nms_max_predictions=100, post_nms_max_predictions=20)
    assert [int(o.scores.shape[0]) for o in out] == [r[0].shape[0] for r in ref]
    feats = [torch.randn(1, c, 4, 4).bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True) for c in m.heads.in_channels]
    with pytest.raises(NotImplementedError, match="training"):  # pose training is not built yet: it must say so, not run something else
        m.heads(feats)
