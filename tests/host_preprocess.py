"""Test infrastructure: builds tests/host_kernels/preprocess_host.cpp (serial host driver around the product header
super_gradients_b200/csrc/preprocess_math.cuh) with g++ and exposes it with the signature of kernels.preprocess_u8."""
import ctypes
import os
import subprocess
import tempfile

import torch

from super_gradients_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = {}


def _handle():
    if "h" not in _LIB:
        d = tempfile.mkdtemp(prefix="sgb_prep_host_")
        so = os.path.join(d, "preprocess_host.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "host_kernels", "preprocess_host.cpp"), "-I", os.path.join(ROOT, "include"),
                        "-I", os.path.join(ROOT, "super_gradients_b200", "csrc"), "-o", so], check=True)  # fmt: skip
        _LIB["h"] = ctypes.CDLL(so)
    return _LIB["h"]


def preprocess_u8(src, out_slot, dst_hw, pad_tl, pad_value=114.0, max_value=255.0, reverse_channels=False, mean=None, std=None):
    slot = out_slot if out_slot.dim() == 4 else out_slot.unsqueeze(0)
    d = L.PreprocDesc()
    d.src_h, d.src_w, d.src_c = src.shape
    d.src_pitch = src.shape[1] * src.shape[2]
    d.dst_h, d.dst_w = int(dst_hw[0]), int(dst_hw[1])
    d.out_h, d.out_w = slot.shape[2], slot.shape[3]
    d.pad_top, d.pad_left = int(pad_tl[0]), int(pad_tl[1])
    d.out_pitch = slot.shape[1]
    d.reverse_channels = 1 if reverse_channels else 0
    d.pad_value = float(pad_value)
    d.max_value = float(max_value) if max_value else 0.0
    d.normalize = 1 if mean is not None else 0
    for i in range(4):
        d.mean[i] = float(mean[i]) if mean is not None and i < len(mean) else 0.0
        d.std[i] = float(std[i]) if std is not None and i < len(std) else 1.0
    buf = torch.empty((d.out_h, d.out_w, d.out_pitch), dtype=torch.int16)
    s = src.contiguous()
    rc = _handle().preprocess_host(ctypes.byref(d), ctypes.c_void_p(s.data_ptr()), ctypes.c_void_p(buf.data_ptr()))
    assert rc == 0
    slot[0].copy_(buf.view(torch.bfloat16).permute(2, 0, 1))
    return out_slot
