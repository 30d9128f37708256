"""YoloX-format post-processing (row N3; reference: training/utils/detection_utils.py:63-86, 279-334): objectness filter,
objectness x class score, multi- or single-label candidates, class-aware or class-agnostic NMS -- on the same one-CTA-per-image
kernel as the PP-YOLOE / YOLO-NAS callback (csrc/nms.cu), i.e. for the whole batch in one launch instead of a Python loop over
images around torchvision.  The kernel keeps its IoU bit-matrix in shared memory, so an image may have at most 1024 candidates
above the confidence threshold; more raise (the reference has no such limit)."""
from typing import List, Optional

import torch
from torch import Tensor

from ... import kernels as K
from ...lib import SgbError

MAX_CANDIDATES = 1024


def convert_cxcywh_bbox_to_xyxy(input_bbox: Tensor) -> Tensor:
    """[..., (cx, cy, w, h)] -> [..., (x1, y1, x2, y2)], 2-d (one image) or 3-d (a batch) like the reference."""
    cx, cy, w, h = input_bbox[..., 0], input_bbox[..., 1], input_bbox[..., 2], input_bbox[..., 3]
    return torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], -1)


@torch.no_grad()
def non_max_suppression(prediction: Tensor, conf_thres: float = 0.1, iou_thres: float = 0.6, multi_label_per_box: bool = True, with_confidence: bool = False,
                        class_agnostic_nms: bool = False) -> List[Optional[Tensor]]:  # fmt: skip
    """prediction [B, A, 5 + C] = (cx, cy, w, h, objectness, class scores).  Returns, per image, [n, 6] rows
    (x1, y1, x2, y2, confidence, class) in descending confidence, or None when nothing survives (as the reference does)."""
    K.require_cuda(prediction, "prediction")
    pred = prediction.float()
    obj = pred[..., 4]
    cls = pred[..., 5:] * obj.unsqueeze(-1) if with_confidence else pred[..., 5:]
    # anchors whose objectness fails the filter can never become candidates: their class scores drop below any threshold
    scores = torch.where((obj > conf_thres).unsqueeze(-1), cls, torch.full_like(cls, -1.0)).contiguous()
    passing = scores > conf_thres
    n_cand = (passing.sum((1, 2)) if multi_label_per_box else passing.any(-1).sum(1)).max()
    if int(n_cand) > MAX_CANDIDATES:  # one host read; this is the predict / validation path
        raise SgbError(f"non_max_suppression: {int(n_cand)} candidates above conf_thres={conf_thres} in one image; the NMS kernel holds at most {MAX_CANDIDATES}")
    boxes = convert_cxcywh_bbox_to_xyxy(pred[..., :4]).contiguous()
    rows, _idx, count = K.batched_nms(boxes, scores, conf_thres, iou_thres, MAX_CANDIDATES, MAX_CANDIDATES, multi_label=multi_label_per_box, class_agnostic=class_agnostic_nms,
                                      thr_inclusive=False)  # fmt: skip
    counts = count.tolist()
    return [rows[b, : counts[b]] if counts[b] else None for b in range(rows.shape[0])]
