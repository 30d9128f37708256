"""Row (f)-N3 without a GPU: (1) the oracle's fixed-point bilinear restatement == the installed OpenCV, (2) the oracle's chain ==
the unmodified reference's ComposeProcessing (hash of the bf16 model input, box / keypoint post-processing), (3) the product's
ComposeProcessing running on the HOST BUILD of the CUDA kernel's arithmetic (csrc/preprocess_math.cuh) == the same fixtures,
bit for bit.  The kernel's launch (preprocess.cu) is covered by the `-m gpu` tests."""
import hashlib
import shutil

import numpy as np
import pytest
import torch

from oracle import sg_oracle as O

import cpu_backend

CHAINS = ["yolo_nas_default", "pose_default", "stretch_normalize"]


def _image(case):
    h, w = case["image_shape"]
    return np.random.RandomState(case["image_seed"]).randint(0, 256, (h, w, 3)).astype(np.uint8)


def _sha(t_bf16):
    return hashlib.sha256(t_bf16.contiguous().view(torch.int16).numpy().tobytes()).hexdigest()


def test_resize_restatement_matches_installed_opencv():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.RandomState(3)
    for (h, w, dh, dw) in [(480, 640, 477, 636), (427, 640, 424, 636), (375, 500, 477, 636), (720, 1280, 358, 636), (100, 100, 200, 200), (64, 48, 17, 13), (600, 800, 300, 400),
                           (33, 47, 640, 480), (5, 7, 3, 2), (2, 2, 9, 9), (1, 5, 4, 10)]:  # fmt: skip
        img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
        np.testing.assert_array_equal(O.resize_linear_u8(img, dh, dw), cv2.resize(img, dsize=(dw, dh), interpolation=cv2.INTER_LINEAR), err_msg=str((h, w, dh, dw)))


@pytest.mark.parametrize("chain", CHAINS)
def test_processing_oracle_matches_reference(golden, chain):
    g = golden("processing")[chain]
    for case in g["cases"]:
        pre, meta = O.preprocess_image(_image(case), **g["kw"])
        t = torch.from_numpy(pre).to(torch.bfloat16)
        assert tuple(t.shape) == tuple(case["pre_shape"])
        torch.testing.assert_close(t[:, ::37, ::41].float(), case["pre_sample"].float(), rtol=0, atol=0)
        assert _sha(t) == case["pre_sha256"], case["image_shape"]
        np.testing.assert_array_equal(O.postprocess_boxes(case["boxes"].numpy(), meta), case["boxes_post"].numpy())


def _product_chain(chain):
    from super_gradients_b200.training import processing as P

    return {
        "yolo_nas_default": lambda: P.default_yolo_nas_coco_processing_params()["image_processor"],
        "pose_default": lambda: P.default_yolo_nas_pose_coco_processing_params()["image_processor"],
        "stretch_normalize": lambda: P.ComposeProcessing([P.DetectionRescale(output_shape=(96, 160)), P.StandardizeImage(max_value=255.0),
                                                          P.NormalizeImage(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]), P.ImagePermute(permutation=(2, 0, 1))]),
    }[chain]()  # fmt: skip


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
@pytest.mark.parametrize("chain", CHAINS)
def test_product_processing_on_the_kernel_arithmetic_matches_reference(golden, monkeypatch, chain):
    cpu_backend.install(monkeypatch)
    g = golden("processing")[chain]
    cp = _product_chain(chain)
    for case in g["cases"]:
        if case["image_shape"][0] * case["image_shape"][1] > 700 * 700 and chain != "yolo_nas_default":
            continue  # the serial host build is slow; one chain covers the large image
        batch, geos = cp.preprocess_batch([_image(case)], "cpu")
        assert tuple(batch.shape) == (1, 16) + tuple(case["pre_shape"][1:]) and batch.dtype == torch.bfloat16
        assert float(batch[:, 3:].abs().max()) == 0.0  # padding channels
        t = batch[0, :3].contiguous()
        torch.testing.assert_close(t[:, ::37, ::41].float(), case["pre_sample"].float(), rtol=0, atol=0)
        assert _sha(t) == case["pre_sha256"], case["image_shape"]
        torch.testing.assert_close(cp.postprocess_boxes(case["boxes"], geos[0]), case["boxes_post"], rtol=0, atol=0)
        if "poses" in case:
            torch.testing.assert_close(cp.postprocess_keypoints(case["poses"], geos[0]), case["poses_post"], rtol=0, atol=0)


def test_unsupported_chains_raise():
    from super_gradients_b200.training import processing as P

    with pytest.raises(NotImplementedError):
        P.ComposeProcessing([P.StandardizeImage(), P.DetectionRescale((8, 8))])  # wrong order
    with pytest.raises(NotImplementedError):
        P.ComposeProcessing([P.DetectionLongestMaxSizeRescale((8, 8)), P.StandardizeImage()])  # ragged batch shapes
    with pytest.raises(NotImplementedError):
        P.ImagePermute((1, 2, 0))
