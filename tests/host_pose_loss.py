"""Test infrastructure: builds tests/host_kernels/pose_loss_host.cpp (a serial host driver around the product header
super_gradients_b200/csrc/pose_loss_math.cuh) with g++ and calls it through ctypes."""
import ctypes
import os
import subprocess

import torch

from super_gradients_b200 import lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = {}


def build(tmp_dir: str):
    key = str(tmp_dir)
    if key not in _LIB:
        so = os.path.join(key, "pose_loss_host.so")
        cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "host_kernels", "pose_loss_host.cpp"), "-I", os.path.join(ROOT, "include"),
               "-I", os.path.join(ROOT, "super_gradients_b200", "csrc"), "-o", so]  # fmt: skip
        subprocess.run(cmd, check=True)
        _LIB[key] = ctypes.CDLL(so)
    return _LIB[key]


def run(handle, d: "L.PoseLossDesc", cls_logits, reg_distri, pose_coords, pose_logits, anchor_points, stride_tensor, gt_boxes, gt_poses, gt_crowd, gt_valid, sigmas, grad_scale=1.0):
    """-> dict(assigned_gt, assigned_score, sums, grads (cls, reg, pose, pose_logits), items)."""
    f32 = lambda t: t.detach().contiguous().float()  # noqa: E731
    cl, rd, pc, pl = f32(cls_logits).reshape(d.B, d.L), f32(reg_distri), f32(pose_coords), f32(pose_logits)
    ap, st = f32(anchor_points), f32(stride_tensor).reshape(-1)
    gb, gp, gc, gv, sg = f32(gt_boxes), f32(gt_poses), gt_crowd.contiguous().to(torch.uint8), gt_valid.contiguous().to(torch.uint8), f32(sigmas)
    agt = torch.empty((d.B, d.L), dtype=torch.int32)
    asc = torch.empty((d.B, d.L), dtype=torch.float32)
    sums = torch.zeros(8, dtype=torch.float64)
    gcl, grd, gpc, gpl = torch.empty_like(cl), torch.empty_like(rd), torch.empty_like(pc), torch.empty_like(pl)
    items = torch.empty(6, dtype=torch.float32)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    rc = handle.pose_loss_host(ctypes.byref(d), p(cl), p(rd), p(pc), p(pl), p(ap), p(st), p(gb), p(gp), p(gc), p(gv), p(sg), ctypes.c_float(grad_scale), p(agt), p(asc), p(sums),
                               p(gcl), p(grd), p(gpc), p(gpl), p(items))  # fmt: skip
    assert rc == 0
    return dict(assigned_gt=agt, assigned_score=asc, sums=sums, grads=(gcl.reshape(d.B, d.L, 1), grd, gpc, gpl), items=items)
