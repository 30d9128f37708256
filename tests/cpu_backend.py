"""Test infrastructure: a CPU stand-in for the INFERENCE subset of super_gradients_b200.kernels (fp32 torch math, bf16
storage at the same points as the CUDA kernels).  It replaces the kernel wrappers, not the product: with it installed the
Python glue above kernels.py (functional.py, the module mirrors, head decoding, post-prediction callbacks, predict()) runs
on a machine without a GPU, so wiring mistakes (argument order, channel slices, anchor bases, head plumbing) are caught
by the CPU suite.  It says nothing about the CUDA kernels themselves -- those are covered by the `-m gpu` parity tests.

Only tests may import this module.  Anything outside the subset raises NotImplementedError.
"""
import torch
import torch.nn.functional as F

from oracle import sg_oracle as O
from super_gradients_b200 import kernels as K
from super_gradients_b200 import lib as L


def _bf16(t):
    return t.to(torch.bfloat16)


def _act(y, act):
    code = K.act_code(act)
    if code == K.ACT_RELU:
        return torch.relu(y)
    if code == K.ACT_SILU:
        return F.silu(y)
    return y


def conv_fprop(x, w_krsc, Kout, R, S, stride, pad, *, scale=None, shift=None, residual=None, stats=None, act=K.ACT_NONE, out=None, out_f32=False):
    n, c, h, w = x.shape
    wt = w_krsc[..., :c].permute(0, 3, 1, 2).float()  # KRSC -> OIHW, padding channels dropped with the activation's
    y = F.conv2d(x.float(), wt, stride=stride, padding=pad)
    if scale is not None:
        y = y * scale.float().view(1, -1, 1, 1)
    if shift is not None:
        y = y + shift.float().view(1, -1, 1, 1)
    if residual is not None:
        y = y + residual.float()
    y = _act(y, act)
    if out is None:
        out = torch.empty(y.shape, dtype=torch.float32, memory_format=torch.channels_last) if out_f32 else K.empty_nhwc(n, Kout, y.shape[2], y.shape[3], x.device)
    out.copy_(y if out_f32 else _bf16(y))
    if stats is not None:
        s = out.double()
        stats[0, 0] += s.sum((0, 2, 3))
        stats[0, 1] += (s * s).sum((0, 2, 3))
    return out


def weight_prepare(w_oihw, c_pad=None, want_crsk=True, scale=None, add_identity=False, out=None):
    Kk, C, R, S = w_oihw.shape
    c_pad = c_pad or ((C + 7) // 8) * 8
    w = w_oihw.detach().float() * (float(scale) if scale is not None else 1.0)
    if add_identity:
        w = w.clone()
        idx = torch.arange(min(Kk, C))
        w[idx, idx, R // 2, S // 2] += 1.0
    krsc = torch.zeros((Kk, R, S, c_pad), dtype=torch.bfloat16)
    krsc[..., :C] = _bf16(w.permute(0, 2, 3, 1))
    crsk = None
    if want_crsk and c_pad == C:
        kp = ((Kk + 7) // 8) * 8
        crsk = torch.zeros((C, R, S, kp), dtype=torch.bfloat16)
        crsk[..., :Kk] = _bf16(w.permute(1, 2, 3, 0))
    return krsc, crsk


def convt2x2_fprop(x_small, w_up, bias, C_up):
    n, Kin, P, Q = x_small.shape
    w = w_up.float().reshape(2, 2, C_up, Kin).permute(3, 2, 0, 1)  # [(dh, dw, co)][ci] -> ConvTranspose2d's [ci][co][dh][dw]
    y = F.conv_transpose2d(x_small.float(), w, bias.float() if bias is not None else None, stride=2)
    out = K.empty_nhwc(n, C_up, 2 * P, 2 * Q, x_small.device)
    out.copy_(_bf16(y))
    return out


def nchw_f32_to_nhwc_bf16(x, c_align=8):
    n, c, h, w = x.shape
    ca = ((c + c_align - 1) // c_align) * c_align
    out = torch.zeros((n, ca, h, w), dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    out[:, :c] = _bf16(x.float())
    return out


def nhwc_bf16_to_nchw_f32(x):
    return x.float().contiguous()


def bn_act_infer(x, gamma, beta, running_mean, running_var, eps, act, residual=None):
    rstd = torch.rsqrt(running_var.float() + eps)
    g = gamma.float() if gamma is not None else torch.ones_like(rstd)
    b = beta.float() if beta is not None else torch.zeros_like(rstd)
    y = x.float() * (g * rstd).view(1, -1, 1, 1) + (b - running_mean.float() * g * rstd).view(1, -1, 1, 1)
    if residual is not None:
        y = y + residual.float()
    out = K.empty_nhwc(*x.shape, x.device)
    out.copy_(_bf16(_act(y, act)))
    return out


def maxpool_fwd(x, k, stride, pad, want_idx=True, out=None):
    y = F.max_pool2d(x.float(), k, stride, pad)
    if out is None:
        out = K.empty_nhwc(x.shape[0], x.shape[1], y.shape[2], y.shape[3], x.device)
    out.copy_(_bf16(y))
    return out, None


def axpby(x1, a, x2=None, b=0.0, out=None):
    y = float(a) * x1.float() + (float(b) * x2.float() if x2 is not None else 0.0)
    if out is None:
        out = K.empty_nhwc(*x1.shape, x1.device)
    out.copy_(_bf16(y))
    return out


def scale_add(x1, a_dev, x2=None, out=None):
    y = float(a_dev.reshape(-1)[0]) * x1.float() + (x2.float() if x2 is not None else 0.0)
    if out is None:
        out = K.empty_nhwc(*x1.shape, x1.device)
    out.copy_(_bf16(y))
    return out


def dfl_decode(reg, cls, L_total, anchor_base, ncls, reg_max, stride, cell_offset, pred_bboxes, pred_scores, cls_logits=None, reg_distri=None):
    n, _, hf, wf = reg.shape
    hw, nb = hf * wf, reg_max + 1
    r = reg.float()[:, : 4 * nb].permute(0, 2, 3, 1).reshape(n, hw, 4, nb)
    d = (torch.softmax(r, -1) * torch.arange(nb, dtype=torch.float32)).sum(-1)
    ys, xs = torch.meshgrid(torch.arange(hf, dtype=torch.float32) + cell_offset, torch.arange(wf, dtype=torch.float32) + cell_offset, indexing="ij")
    pts = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1)
    rows = slice(anchor_base, anchor_base + hw)
    pred_bboxes[:, rows] = torch.cat([pts - d[..., :2], pts + d[..., 2:]], -1) * stride
    c = cls.float()[:, :ncls].permute(0, 2, 3, 1).reshape(n, hw, ncls)
    pred_scores[:, rows] = torch.sigmoid(c)
    if cls_logits is not None:
        cls_logits[:, rows] = c
    if reg_distri is not None:
        reg_distri[:, rows] = r.reshape(n, hw, 4 * nb)


def pose_keypoint_decode(pose, logit, logit_off, L_total, anchor_base, J, stride, cell_offset, offset_multiplier, compensate, pose_coords, pose_scores, pose_logits=None):
    n, _, hf, wf = pose.shape
    hw = hf * wf
    off = pose.float()[:, : 2 * J].permute(0, 2, 3, 1).reshape(n, hw, J, 2)  # channel 2j + {x, y}
    ys, xs = torch.meshgrid(torch.arange(hf, dtype=torch.float32) + cell_offset, torch.arange(wf, dtype=torch.float32) + cell_offset, indexing="ij")
    pts = torch.stack([xs.reshape(-1), ys.reshape(-1)], -1).view(1, hw, 1, 2)
    rows = slice(anchor_base, anchor_base + hw)
    pose_coords[:, rows] = (off * offset_multiplier + (pts - (cell_offset if compensate else 0.0))) * stride
    lg = logit.float()[:, logit_off : logit_off + J].permute(0, 2, 3, 1).reshape(n, hw, J)
    pose_scores[:, rows] = torch.sigmoid(lg)
    if pose_logits is not None:
        pose_logits[:, rows] = lg


def batched_nms(boxes, scores, score_thr, iou_thr, top_k, max_out, multi_label=True, class_agnostic=False, thr_inclusive=None):
    """Same contract as the CUDA wrapper, implemented with the oracle's post-processing (incl. its tie rule)."""
    B, Lc, C = scores.shape
    inclusive = (not multi_label) if thr_inclusive is None else bool(thr_inclusive)
    if multi_label or not inclusive:
        raise NotImplementedError("CPU stand-in: only the single-label inclusive mode is implemented")
    out = torch.zeros((B, max_out, 6))
    oidx = torch.full((B, max_out), -1, dtype=torch.int32)
    cnt = torch.zeros((B,), dtype=torch.int32)
    rows, idx = O.ppyoloe_postprocess(boxes, scores, score_thr, iou_thr, top_k, max_out, multi_label_per_box=False, class_agnostic_nms=class_agnostic)
    for b, (r, i) in enumerate(zip(rows, idx)):
        n = r.shape[0]
        cnt[b] = n
        out[b, :n] = torch.from_numpy(r)
        oidx[b, :n] = torch.from_numpy(i).int()
    return out, oidx, cnt


_SUBSET = dict(conv_fprop=conv_fprop, weight_prepare=weight_prepare, convt2x2_fprop=convt2x2_fprop, nchw_f32_to_nhwc_bf16=nchw_f32_to_nhwc_bf16,
               nhwc_bf16_to_nchw_f32=nhwc_bf16_to_nchw_f32, bn_act_infer=bn_act_infer, maxpool_fwd=maxpool_fwd, axpby=axpby, scale_add=scale_add,
               dfl_decode=dfl_decode, pose_keypoint_decode=pose_keypoint_decode, batched_nms=batched_nms)  # fmt: skip


def install(monkeypatch):
    """Routes the inference subset of kernels.py to the CPU stand-ins and makes every other kernel call fail loudly."""

    def refuse(name, *args):
        raise NotImplementedError(f"CPU stand-in: {name} is outside the inference subset")

    monkeypatch.setattr(K, "require_cuda", lambda t, name="tensor": None)
    monkeypatch.setattr(L, "call", refuse)
    for name, fn in _SUBSET.items():
        monkeypatch.setattr(K, name, fn)
