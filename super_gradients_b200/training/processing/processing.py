"""predict() pre/post-processing on the GPU.  Same class names, constructor arguments and order of
operations as the reference's training/processing/processing.py, but a ComposeProcessing does not run five numpy / cv2 passes per
image on the host: it folds its chain into ONE kernel launch per image (csrc/preprocess.cu: OpenCV's fixed-point INTER_LINEAR
resize, constant padding, channel reversal, standardisation, mean / std, bf16 NHWC store straight into the batch tensor) and undoes
padding and rescaling on the prediction tensors with the reference's float32 arithmetic.

Supported chain (what the YOLO-NAS / YOLO-NAS-POSE defaults use, processing.py:960-980, 1060-1075), each step optional, in this
order: ReverseImageChannels, one *Rescale, one *Padding, StandardizeImage, NormalizeImage, ImagePermute((2, 0, 1)).
Anything else raises NotImplementedError."""
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from ... import kernels as K


@dataclass
class ImageGeometry:
    """Per-image record of what the chain did (the reference's RescaleMetadata + DetectionPadToSizeMetadata)."""

    original_shape: Tuple[int, int]
    scale_factor_h: float
    scale_factor_w: float
    resized_shape: Tuple[int, int]
    pad_top: int
    pad_left: int


class Processing:
    resizes_image = False


class ReverseImageChannels(Processing):
    pass


class StandardizeImage(Processing):
    def __init__(self, max_value: float = 255.0):
        self.max_value = float(max_value)


class NormalizeImage(Processing):
    def __init__(self, mean: List[float], std: List[float]):
        self.mean = np.array(mean, dtype=np.float32).reshape(-1)
        self.std = np.array(std, dtype=np.float32).reshape(-1)


class ImagePermute(Processing):
    def __init__(self, permutation: Tuple[int, int, int] = (2, 0, 1)):
        self.permutation = tuple(permutation)
        if self.permutation != (2, 0, 1):
            raise NotImplementedError("only the HWC -> CHW permutation (2, 0, 1) is supported")


class _Rescale(Processing):
    resizes_image = True
    keep_aspect = False

    def __init__(self, output_shape: Tuple[int, int]):
        self.output_shape = tuple(output_shape)

    def target(self, height: int, width: int):
        """-> (new_h, new_w, scale_h, scale_w) exactly as the reference computes them (processing.py:516-550)."""
        if self.keep_aspect:
            scale = min(self.output_shape[0] / height, self.output_shape[1] / width)
            if scale != 1.0:
                return round(height * scale), round(width * scale), scale, scale
            return height, width, scale, scale
        return self.output_shape[0], self.output_shape[1], self.output_shape[0] / height, self.output_shape[1] / width


class DetectionRescale(_Rescale):
    pass


class DetectionLongestMaxSizeRescale(_Rescale):
    keep_aspect = True


class KeypointsLongestMaxSizeRescale(DetectionLongestMaxSizeRescale):
    pass


class _Padding(Processing):
    resizes_image = True
    center = False

    def __init__(self, output_shape: Tuple[int, int], pad_value: int):
        self.output_shape = tuple(output_shape)
        self.pad_value = pad_value

    def top_left(self, height: int, width: int) -> Tuple[int, int]:
        pad_h, pad_w = self.output_shape[0] - height, self.output_shape[1] - width
        if pad_h < 0 or pad_w < 0:
            raise ValueError(f"image {height}x{width} is larger than the padded shape {self.output_shape}")
        return (pad_h // 2, pad_w // 2) if self.center else (0, 0)


class DetectionCenterPadding(_Padding):
    center = True


class DetectionBottomRightPadding(_Padding):
    pass


class KeypointsBottomRightPadding(DetectionBottomRightPadding):
    pass


class ComposeProcessing(Processing):
    def __init__(self, processings: Sequence[Processing]):
        self.processings = list(processings)
        order = [ReverseImageChannels, _Rescale, _Padding, StandardizeImage, NormalizeImage, ImagePermute]
        pos = -1
        self.reverse = self.rescale = self.padding = self.standardize = self.normalize = None
        for p in self.processings:
            idx = next((i for i, t in enumerate(order) if isinstance(p, t)), None)
            if idx is None or idx <= pos:
                raise NotImplementedError(f"unsupported processing chain at {type(p).__name__}: supported order is "
                                          "[ReverseImageChannels] [Rescale] [Padding] [StandardizeImage] [NormalizeImage] [ImagePermute]")  # fmt: skip
            pos = idx
            name = ("reverse", "rescale", "padding", "standardize", "normalize", "permute")[idx]
            setattr(self, name, p)
        if self.padding is None and self.rescale is not None and self.rescale.keep_aspect:
            raise NotImplementedError("an aspect-preserving rescale needs a padding step to give the batch one shape")

    @property
    def resizes_image(self) -> bool:
        return self.rescale is not None or self.padding is not None

    def geometry(self, height: int, width: int) -> Tuple[ImageGeometry, Tuple[int, int]]:
        nh, nw, sh, sw = self.rescale.target(height, width) if self.rescale is not None else (height, width, 1.0, 1.0)
        if self.padding is not None:
            top, left = self.padding.top_left(nh, nw)
            canvas = self.padding.output_shape
        else:
            top, left, canvas = 0, 0, (nh, nw)
        return ImageGeometry((height, width), sh, sw, (nh, nw), top, left), canvas

    def preprocess_batch(self, images: Sequence[np.ndarray], device) -> Tuple[torch.Tensor, List[ImageGeometry]]:
        """images: uint8 H x W x C arrays (any sizes).  Returns the model input -- bf16 NHWC [B, 16, H, W] with channels >= C
        zero, which the model mirrors consume as is -- and the per-image geometry for postprocess_*()."""
        geos, canvases = zip(*(self.geometry(im.shape[0], im.shape[1]) for im in images))
        if len(set(canvases)) != 1:
            raise ValueError(f"the images of a batch must map to one input shape, got {sorted(set(canvases))}")
        oh, ow = canvases[0]
        batch = K.empty_nhwc(len(images), 16, oh, ow, device)
        for b, (im, g) in enumerate(zip(images, geos)):
            if im.dtype != np.uint8 or im.ndim != 3:
                raise ValueError("predict() images must be uint8 H x W x C arrays")
            src = torch.from_numpy(np.ascontiguousarray(im)).to(device, non_blocking=True)
            K.preprocess_u8(src, batch[b : b + 1], g.resized_shape, (g.pad_top, g.pad_left), pad_value=self.padding.pad_value if self.padding is not None else 0.0,
                            max_value=self.standardize.max_value if self.standardize is not None else 0.0, reverse_channels=self.reverse is not None,
                            mean=self.normalize.mean if self.normalize is not None else None, std=self.normalize.std if self.normalize is not None else None)  # fmt: skip
        return batch, list(geos)

    @staticmethod
    def batch_shift_scale(geos: Sequence[ImageGeometry], device) -> Tuple[torch.Tensor, torch.Tensor]:
        """Per-image (shift [B, 2] = (-pad_left, -pad_top), scale [B, 2] = float32(1 / scale_w), float32(1 / scale_h)): the operands
        of postprocess_boxes / postprocess_keypoints for a whole batch, so that undoing the padding / rescaling of every image is
        a handful of batched launches ((v + shift) * scale in float32: the same two roundings per coordinate as the per-image form)."""
        shift = torch.tensor([[-g.pad_left, -g.pad_top] for g in geos], dtype=torch.float32)
        scale = torch.tensor([[float(np.float32(1 / g.scale_factor_w)), float(np.float32(1 / g.scale_factor_h))] for g in geos], dtype=torch.float32)
        return shift.to(device, non_blocking=True), scale.to(device, non_blocking=True)

    @staticmethod
    def postprocess_boxes(boxes_xyxy: torch.Tensor, g: ImageGeometry) -> torch.Tensor:
        """[n, >= 4] rows whose first four columns are xyxy in model-input pixels -> original-image pixels: shift by the padding,
        then multiply by float32(1 / scale) (_shift_bboxes_xyxy, _rescale_bboxes: transforms/utils.py:47-62, 155-166)."""
        out = boxes_xyxy.float().clone()
        out[:, [0, 2]] += -g.pad_left
        out[:, [1, 3]] += -g.pad_top
        sx, sy = np.float32(1 / g.scale_factor_w), np.float32(1 / g.scale_factor_h)
        out[:, :4] *= torch.tensor([sx, sy, sx, sy], dtype=torch.float32, device=out.device)
        return out

    @staticmethod
    def postprocess_keypoints(poses: torch.Tensor, g: ImageGeometry) -> torch.Tensor:
        """[n, J, >= 2] keypoints (x, y, ...) -> original-image pixels (_shift_keypoints, _rescale_keypoints)."""
        out = poses.float().clone()
        out[..., 0] += -g.pad_left
        out[..., 1] += -g.pad_top
        out[..., 0] *= float(np.float32(1 / g.scale_factor_w))
        out[..., 1] *= float(np.float32(1 / g.scale_factor_h))
        return out


def default_yolo_nas_coco_processing_params() -> dict:
    """processing.py:960-980 (class names are dataset metadata and not part of this mirror)."""
    return dict(image_processor=ComposeProcessing([DetectionLongestMaxSizeRescale(output_shape=(636, 636)), DetectionCenterPadding(output_shape=(640, 640), pad_value=114),
                                                   StandardizeImage(max_value=255.0), ImagePermute(permutation=(2, 0, 1))]), iou=0.7, conf=0.25)  # fmt: skip


def default_yolo_nas_pose_coco_processing_params() -> dict:
    """processing.py:1060-1085."""
    return dict(image_processor=ComposeProcessing([ReverseImageChannels(), KeypointsLongestMaxSizeRescale(output_shape=(640, 640)),
                                                   KeypointsBottomRightPadding(output_shape=(640, 640), pad_value=127), StandardizeImage(max_value=255.0),
                                                   ImagePermute(permutation=(2, 0, 1))]), conf=0.5)  # fmt: skip
