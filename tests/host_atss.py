"""Test infrastructure: builds tests/host_kernels/atss_host.cpp (serial host driver around the product header
super_gradients_b200/csrc/atss_math.cuh) with g++ and exposes it with the signature of kernels.atss_assign."""
import ctypes
import os
import subprocess
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = {}


def _handle():
    if "h" not in _LIB:
        d = tempfile.mkdtemp(prefix="sgb_atss_host_")
        so = os.path.join(d, "atss_host.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "host_kernels", "atss_host.cpp"), "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "super_gradients_b200", "csrc"), "-o", so], check=True)  # fmt: skip
        _LIB["h"] = ctypes.CDLL(so)
    return _LIB["h"]


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def atss_assign(desc, reg_distri, anchors, anchor_points, stride_tensor, level_sizes, gt_boxes, gt_labels, gt_valid, sums):
    B, L = desc.B, desc.L
    reg_distri, anchors, anchor_points, stride_tensor = (t.detach().contiguous().float() for t in (reg_distri, anchors, anchor_points, stride_tensor))
    al = torch.empty((B, L), dtype=torch.int32)
    ab = torch.empty((B, L, 4), dtype=torch.float32)
    asc = torch.empty((B, L), dtype=torch.float32)
    lv = (ctypes.c_int32 * len(level_sizes))(*[int(v) for v in level_sizes])
    rc = _handle().atss_assign_host(ctypes.byref(desc), _p(reg_distri), _p(anchors), _p(anchor_points), _p(stride_tensor), lv, len(level_sizes), _p(gt_boxes.contiguous()),
                                    _p(gt_labels.contiguous()), _p(gt_valid.contiguous()), _p(al), _p(ab), _p(asc), _p(sums))  # fmt: skip
    assert rc == 0, rc
    return al, ab, asc
