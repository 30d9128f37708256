"""Test infrastructure: a CPU stand-in for super_gradients_b200.kernels (fp32 torch math, bf16 storage at the same points
as the CUDA kernels): `install()` covers the inference subset, `install_training()` adds the backward / BatchNorm /
QARepVGG / loss / optimizer wrappers so that a whole TrainStep runs.  It replaces the kernel wrappers, not the product: with it installed the
Python glue above kernels.py (functional.py, the module mirrors, head decoding, post-prediction callbacks, predict()) runs
on a machine without a GPU, so wiring mistakes (argument order, channel slices, anchor bases, head plumbing) are caught
by the CPU suite.  It says nothing about the CUDA kernels themselves -- those are covered by the `-m gpu` parity tests.

The training stand-ins are written from the layer definitions (torch autograd on the textbook formulation), not from the
CUDA kernels' algebra; they exist to exercise the step plumbing (flat state, step arena, deferred weight gradients,
batched work tables, filter caches) deterministically.

Only tests may import this module.  Anything outside the installed subset raises NotImplementedError.
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


def stem_patches(x, R, stride, pad, c_out):
    n, c, h, w = x.shape
    P, Q = (h + 2 * pad - R) // stride + 1, (w + 2 * pad - R) // stride + 1
    cols = torch.nn.functional.unfold(x.float(), R, padding=pad, stride=stride).reshape(n, c, R * R, P, Q)  # [n, c, (r, s), P, Q]
    out = torch.zeros((n, c_out, P, Q), dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    out[:, : c * R * R] = _bf16(cols.permute(0, 2, 1, 3, 4).reshape(n, R * R * c, P, Q))  # channel = (r * R + s) * C + c
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
    y, idx = F.max_pool2d(x.float(), k, stride, pad, return_indices=True)
    if out is None:
        out = K.empty_nhwc(x.shape[0], x.shape[1], y.shape[2], y.shape[3], x.device)
    out.copy_(_bf16(y))
    return out, (idx if want_idx else None)


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


def scale_add_dot(x1, a_dev, xd, x2=None, out=None):
    dot = K.zeros((x1.shape[1],), torch.float64, x1.device)
    dot += (x1.double() * xd.double()).sum((0, 2, 3))
    return scale_add(x1, a_dev, x2, out), dot


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
    """Same contract as the CUDA wrapper (rows [B, max_out, 6], flat candidate index anchor * C + class, count), written
    directly from the callbacks' definition: candidates (multi-label: every (anchor, class) in nonzero order; single-label:
    best class per anchor) passing the threshold (> or >=), top-k by (score desc, position asc), torchvision-style NMS."""
    import numpy as np

    B, Lc, C = scores.shape
    inclusive = (not multi_label) if thr_inclusive is None else bool(thr_inclusive)
    out = torch.zeros((B, max_out, 6))
    oidx = torch.full((B, max_out), -1, dtype=torch.int32)
    cnt = torch.zeros((B,), dtype=torch.int32)
    for b in range(B):
        bb, ss = boxes[b].float(), scores[b].float()
        if multi_label:
            i, j = ((ss >= score_thr) if inclusive else (ss > score_thr)).nonzero(as_tuple=False).T
            conf = ss[i, j]
        else:
            cmax, cidx = ss.max(1)
            m = (cmax >= score_thr) if inclusive else (cmax > score_thr)
            i, j, conf = m.nonzero().flatten(), cidx[m], cmax[m]
        if conf.shape[0] > top_k:
            order = torch.from_numpy(np.lexsort((np.arange(conf.shape[0]), -conf.numpy()))[:top_k].copy())
            i, j, conf = i[order], j[order], conf[order]
        bx = bb[i].numpy()
        keep = O.nms_numpy(bx, conf.numpy(), iou_thr) if class_agnostic else O.batched_nms_numpy(bx, conf.numpy(), j.numpy(), iou_thr)
        keep = keep[:max_out]
        n = len(keep)
        cnt[b] = n
        if n:
            out[b, :n] = torch.from_numpy(np.concatenate([bx[keep], conf.numpy()[keep, None], j.numpy()[keep, None].astype(np.float32)], 1))
            oidx[b, :n] = torch.from_numpy((i.numpy()[keep] * C + j.numpy()[keep]).astype(np.int32))
    return out, oidx, cnt


# ---------------------------------------------------------------------------------------------------- training subset
def conv_dgrad(dy, w_crsk, x_shape, R, S, stride, pad, out=None, accumulate=False):
    n, c, h, w = x_shape
    Kk = dy.shape[1]
    wt = w_crsk[..., :Kk].permute(3, 0, 1, 2).float()  # CRSK -> OIHW
    dx = torch.nn.grad.conv2d_input((n, wt.shape[1], h, w), wt, dy.float(), stride=stride, padding=pad)
    if out is None:
        out = K.empty_nhwc(n, c, h, w, dy.device)
        accumulate = False
    out[:, : wt.shape[1]].copy_(_bf16(out[:, : wt.shape[1]].float() + dx if accumulate else dx))
    return out


def conv_wgrad(x, dy, R, S, stride, pad, dw_krsc=None):
    n, c, h, w = x.shape
    Kk = dy.shape[1]
    if dw_krsc is None:
        dw_krsc = K.zeros((Kk, R, S, c), torch.float32, x.device)  # through the step arena, like the CUDA wrapper
    g = torch.nn.grad.conv2d_weight(x.float(), (Kk, c, R, S), dy.float(), stride=stride, padding=pad)
    dw_krsc += g.permute(0, 2, 3, 1)
    return dw_krsc


def wgrad_to_oihw(dw_krsc, C, out=None, accumulate=False):
    g = dw_krsc[..., :C].permute(0, 3, 1, 2)
    if out is None:
        return g.contiguous()
    g = g.reshape(out.shape)  # an nn.Linear slot is [K, C]: the same memory as OIHW [K, C, 1, 1] (the kernel writes through a pointer)
    out.copy_(out + g if accumulate else g)
    return out


def weight_prepare_batch(entries, device):
    return list(entries), len(entries), sum(e[2].numel() + (e[3].numel() if e[3] is not None else 0) for e in entries)


def run_weight_prepare_batch(table, n, total):
    assert len(table) == n
    for w, scale, krsc, crsk, c_pad, add_identity, *extra in table:
        k2, c2 = weight_prepare(w, c_pad=c_pad, want_crsk=crsk is not None, scale=scale, add_identity=add_identity)
        kp, koff, etaps, etap = extra[0] if extra else (0, 0, 0, 0)
        Kk = w.shape[0]
        if etaps:  # 1x1 source = tap `etap` of an etaps-tap destination (krsc: this filter's rows of the wider KRSC, crsk: the wider CRSK)
            r, s = divmod(etap, int(round(etaps**0.5)))
            krsc[:, r, s, :].copy_(k2[:, 0, 0, :])
            if crsk is not None:
                crsk[:, r, s, koff : koff + Kk].copy_(c2[:, 0, 0, :Kk])
        elif kp:
            krsc.copy_(k2)
            if crsk is not None:
                crsk[..., koff : koff + Kk].copy_(c2[..., :Kk])
        else:
            krsc.copy_(k2)
            if crsk is not None:
                crsk.copy_(c2)


def wgrad_to_oihw_batch_table(entries, device):
    return list(entries), len(entries), sum(e[2].numel() for e in entries)


def run_wgrad_to_oihw_batch(table, n, total):
    assert len(table) == n
    for dw, C, g, accumulate in table:
        wgrad_to_oihw(dw, C, out=g, accumulate=accumulate)


def qarep_alpha_finish_table(entries, device):
    return list(entries), len(entries)


def run_qarep_alpha_finish(table, n):
    assert len(table) == n
    for dw1, C, w1, alpha, dab, bias1, g_w1, g_bias, g_alpha in table:
        g = dw1[:, 0, 0, :C]
        acc = (g * w1[:, :, 0, 0]).sum()
        g_w1.add_((alpha * g).reshape(g_w1.shape))
        if dab is not None:
            if bias1 is not None:
                acc = acc + (dab * bias1).sum()
            if g_bias is not None:
                g_bias.add_(alpha * dab)
        if g_alpha is not None:
            g_alpha.add_(acc.reshape(g_alpha.shape))


def _cv(t):
    return t.float().view(1, -1, 1, 1)


def _update_running(rm, rv, mean, var_biased, M, momentum):
    if rm is not None:
        rm.mul_(1 - momentum).add_(momentum * mean)
        rv.mul_(1 - momentum).add_(momentum * var_biased * (M / max(M - 1, 1)))


def _span(t, c):
    """The kernels read / write C entries from a pointer: two layers that share one GEMM pass the FIRST layer's tensor and rely on the
    second's following it in memory (functional.dual_conv_bn_act_ready checked that).  Same view here."""
    if t is None or t.numel() == c:
        return t
    t = t.detach()
    return torch.as_strided(t, (c,), (1,), t.storage_offset())


def bn_act_fwd(x, stats, gamma, beta, running_mean, running_var, eps, momentum, act, residual=None, sample_scale=None):
    n, c, h, w = x.shape
    M = n * h * w
    gamma, beta, running_mean, running_var = _span(gamma, c), _span(beta, c), _span(running_mean, c), _span(running_var, c)
    if stats is None:  # wide layers: the BatchNorm launch computes the sums of the stored bf16 values itself
        xd = x.double()
        stats = torch.stack([xd.sum((0, 2, 3)), (xd * xd).sum((0, 2, 3))])[None]
    tot = stats.sum(0)  # [2, C] fp64
    mean = tot[0] / M
    var = (tot[1] / M - mean * mean).clamp_min(0)
    rstd = torch.rsqrt(var + eps)
    y = (x.float() - _cv(mean)) * _cv(rstd) * _cv(gamma) + _cv(beta)
    if sample_scale is not None:
        y = y * sample_scale.float().view(-1, 1, 1, 1)
    if residual is not None:
        y = y + residual.float()
    out = K.empty_nhwc(n, c, h, w, x.device)
    out.copy_(_bf16(_act(y, act)))
    _update_running(running_mean, running_var, mean.float(), var.float(), M, momentum)
    return out, mean.float(), rstd.float()


def _mask(dout, out, act):
    code = K.act_code(act)
    if code == K.ACT_RELU:
        return dout.float() * (out.float() > 0)
    if code == K.ACT_NONE:
        return dout.float()
    raise NotImplementedError("CPU stand-in: only relu / identity activations have a backward")


def bn_act_bwd(dy, x, y, gamma, mean, rstd, eps, act, want_residual_grad=False, dgamma=None, dbeta=None, beta=None, sample_scale=None, dy2=None):
    n, c, h, w = x.shape
    M = n * h * w
    gamma, beta, dgamma, dbeta = _span(gamma, c), _span(beta, c), _span(dgamma, c), _span(dbeta, c)
    if dy2 is not None:  # the gradient arrives as two channel ranges (two layers that shared one GEMM)
        dy = torch.cat([dy.float(), dy2.float()], 1)
    if y is None:  # the mask is recomputed from x with the forward pass's FMA
        scale = gamma.float() * rstd
        y = x.float() * _cv(scale) + _cv(beta.float() - mean * scale)
    dz_res = _mask(dy, y, act)
    dz = dz_res if sample_scale is None else dz_res * sample_scale.float().view(-1, 1, 1, 1)
    xh = (x.float() - _cv(mean)) * _cv(rstd)
    sb = dz.sum((0, 2, 3))
    sg = (dz * xh).sum((0, 2, 3))
    dx32 = _cv(gamma.float() * rstd) * (dz - _cv(sb / M) - xh * _cv(sg / M))
    dx = K.empty_nhwc(n, c, h, w, x.device)
    dx.copy_(_bf16(dx32))
    dres = None
    if want_residual_grad:
        dres = K.empty_nhwc(n, c, h, w, x.device)
        dres.copy_(_bf16(dz_res))
    dgamma = K.zeros((c,), torch.float32, x.device) if dgamma is None else dgamma
    dbeta = K.zeros((c,), torch.float32, x.device) if dbeta is None else dbeta
    dgamma += sg
    dbeta += sb
    return dx, dres, dgamma, dbeta


def channel_stats(x):
    st = K.zeros((1, 2, x.shape[1]), torch.float64, x.device)
    s = x.double()
    st[0, 0] += s.sum((0, 2, 3))
    st[0, 1] += (s * s).sum((0, 2, 3))
    return st


def channel_dot(a, b):
    out = K.zeros((a.shape[1],), torch.float64, a.device)
    out += (a.double() * b.double()).sum((0, 2, 3))
    return out


def _bn_batch(t, gamma, beta, eps):
    mean = t.mean((0, 2, 3), keepdim=True)
    var = t.var((0, 2, 3), unbiased=False, keepdim=True)
    out = (t - mean) * torch.rsqrt(var + eps)
    if gamma is not None:
        out = out * gamma.view(1, -1, 1, 1)
    if beta is not None:
        out = out + beta.view(1, -1, 1, 1)
    return out, mean.flatten(), var.flatten()


def _qarep_pre(y3, u, gamma3, beta3, bias1a, gamma_p, beta_p, eps3, eps_post, use_post_bn):
    """Pre-activation of the train-mode block: bn3(y3) + u + alpha*bias1 [-> post_bn]  (qarepvgg_block.py:184-204)."""
    b3, m3, v3 = _bn_batch(y3, gamma3, beta3, eps3)
    z = b3 + u
    if bias1a is not None:
        z = z + bias1a.view(1, -1, 1, 1)
    if not use_post_bn:
        return z, (m3, v3, None, None)
    zp, mz, vz = _bn_batch(z, gamma_p, beta_p, eps_post)
    return zp, (m3, v3, mz, vz)


def qarep_fwd(y3, u, gamma3, beta3, bias1a, gamma_p, beta_p, rm3, rv3, rmp, rvp, eps3, eps_post, momentum, act, use_post_bn=True, residual=None, res_alpha=None):
    n, c, h, w = y3.shape
    M = n * h * w
    f = lambda t: None if t is None else t.detach().float()  # noqa: E731
    pre, (m3, v3, mz, vz) = _qarep_pre(y3.float(), u.float(), f(gamma3), f(beta3), f(bias1a), f(gamma_p), f(beta_p), eps3, eps_post, use_post_bn)
    out = K.empty_nhwc(n, c, h, w, y3.device)
    res = _act(pre, act)
    if residual is not None:  # the bottleneck's learnable shortcut in the same pass; the block's own output is rounded first, like the kernel
        res = res_alpha.detach().float() * residual.float() + _bf16(res).float()
    out.copy_(_bf16(res))
    _update_running(rm3, rv3, m3, v3, M, momentum)
    if use_post_bn:
        _update_running(rmp, rvp, mz, vz, M, momentum)
    coef = torch.zeros((9, c), dtype=torch.float32)  # the coefficient table is private to the CUDA kernels
    if residual is not None:
        # the kernels recompute the activation mask from y3, u and the coefficient table (they never read `out`); with a shortcut added
        # `out` no longer carries the mask, so the stand-in keeps it next to its (otherwise empty) coefficient table
        _QAREP_MASK[coef.data_ptr()] = (coef, pre > 0)
    return out, coef


_QAREP_MASK = {}


def qarep_bwd(dout, out, y3, u, coef, gamma3, gamma_p, eps3, eps_post, act, use_post_bn=True, acc=None, out_grads=None):
    n, c, h, w = y3.shape
    kept = _QAREP_MASK.pop(coef.data_ptr(), None)
    dpre = dout.float() * kept[1] if kept is not None and K.act_code(act) == K.ACT_RELU else _mask(dout, out, act)
    leaf = lambda t: t.detach().float().clone().requires_grad_(True)  # noqa: E731
    y3l, ul, g3l = leaf(y3), leaf(u), leaf(gamma3)
    b3l, abl = torch.zeros(c, requires_grad=True), torch.zeros(c, requires_grad=True)  # additive constants: only their gradients matter
    gpl = leaf(gamma_p) if use_post_bn else None
    bpl = torch.zeros(c, requires_grad=True) if use_post_bn else None
    with torch.enable_grad():
        pre, _ = _qarep_pre(y3l, ul, g3l, b3l, abl, gpl, bpl, eps3, eps_post, use_post_bn)
        wrt = [y3l, ul, g3l, b3l, abl] + ([gpl, bpl] if use_post_bn else [])
        grads = torch.autograd.grad(pre, wrt, dpre, allow_unused=True)
    gy3, gu, gg3, gb3, gab = grads[:5]
    ggp, gbp = (grads[5], grads[6]) if use_post_bn else (None, None)
    dy3, du = out_grads if out_grads is not None else (K.empty_nhwc(n, c, h, w, y3.device), K.empty_nhwc(n, c, h, w, y3.device))
    dy3.copy_(_bf16(gy3))
    du.copy_(_bf16(gu))
    acc = acc or (None,) * 5
    outs = []
    for slot, g in zip(acc, (gg3, gb3, gab, ggp, gbp)):
        t = slot if slot is not None else K.zeros((c,), torch.float32, y3.device)
        if g is not None:
            t += g
        outs.append(t)
    return (dy3, du, *outs)


def maxpool_bwd(dy, idx, x_shape, k, stride, pad):
    n, c, h, w = x_shape
    dx = torch.zeros((n, c, h * w), dtype=torch.float32)
    dx.scatter_add_(2, idx.reshape(n, c, -1), dy.float().reshape(n, c, -1))
    return dx.view(n, c, h, w)


def head_grad_scatter(grad, n, hw, L_total, anchor_base, dy):
    _, gC, hf, wf = dy.shape
    dy.copy_(_bf16(grad[:, anchor_base : anchor_base + hw, :gC].reshape(n, hf, wf, gC).permute(0, 3, 1, 2)))


def tal_assign(d, cls_logits, reg_distri, anchor_points, stride_tensor, gt_boxes, gt_labels, gt_valid, sums):
    st = stride_tensor.view(-1, 1)
    pred = O.bbox_decode(anchor_points / st, reg_distri.detach()) * st
    lab, box, score = O.tal_assign(cls_logits.detach().sigmoid(), pred, anchor_points, gt_labels.long().unsqueeze(-1), gt_boxes.float(), gt_valid.float().unsqueeze(-1),
                                   bg_index=d.ncls, topk=d.topk, alpha=d.alpha, beta=d.beta)  # fmt: skip
    sums[3] += score.sum().double()
    return lab.int(), box, score.sum(-1)


def atss_assign(*args, **kwargs):
    """The product header's arithmetic behind a serial host driver (tests/host_atss.py)."""
    import host_atss

    return host_atss.atss_assign(*args, **kwargs)


def dfl_iou_loss(d, cls_logits, reg_distri, anchor_points, stride_tensor, al, ab, asc, sums, grad_scale=1.0, want_grad=True, focal_alpha=None):
    C, reg_max = d.ncls, d.reg_max
    st = stride_tensor.view(-1, 1)
    cl, rd = cls_logits.detach().clone().requires_grad_(True), reg_distri.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        lab = al.long()
        onehot = F.one_hot(lab, C + 1)[..., :C].float()
        cls_sum = O.varifocal_loss(cl, onehot * asc.unsqueeze(-1), onehot) if focal_alpha is None else O.focal_loss(cl, onehot * asc.unsqueeze(-1), alpha=focal_alpha)
        pts_s = anchor_points / st
        pred = O.bbox_decode(pts_s, rd)
        pos = lab != C
        if bool(pos.any()):
            wgt = asc[pos].unsqueeze(-1)
            pb, gb = pred[pos], (ab / st)[pos]
            iou_sum = ((O.giou_loss if d.iou_type == 0 else O.ciou_loss)(pb, gb) * wgt).sum()
            pts = pts_s.unsqueeze(0).expand(lab.shape[0], -1, 2)[pos]
            ltrb = torch.cat([pts - gb[:, :2], gb[:, 2:] - pts], -1).clip(0, reg_max - 0.01)
            dfl_sum = (O.df_loss(rd[pos].reshape(-1, 4, reg_max + 1), ltrb) * wgt).sum()
        else:
            iou_sum, dfl_sum = rd.sum() * 0.0, rd.sum() * 0.0
        sums[0] += cls_sum.detach().double()
        sums[1] += iou_sum.detach().double()
        sums[2] += dfl_sum.detach().double()
        nrm = float(sums[3].clamp_min(1.0))
        total = (d.w_cls * cls_sum + d.w_iou * iou_sum + d.w_dfl * dfl_sum) / nrm
        gc = gr = None
        if want_grad:
            gc, gr = torch.autograd.grad(total * grad_scale, [cl, rd])
    items = torch.stack([d.w_cls * sums[0] / nrm, d.w_iou * sums[1] / nrm, d.w_dfl * sums[2] / nrm, total.detach().double()]).float()
    return items, gc, gr


def sgd_step(p, g, mom, hp):
    lr, mu, wd, gs, nesterov = [float(v) for v in hp]
    gg = g * gs + wd * p
    mom.mul_(mu).add_(gg)
    p.sub_(lr * (gg + mu * mom if nesterov else mom))


def adamw_step(p, g, m, v, hp):
    lr, b1, b2, eps, wd, bc1, bc2, gs = [float(x) for x in hp]
    gg = g * gs
    p.mul_(1 - lr * wd)
    m.mul_(b1).add_((1 - b1) * gg)
    v.mul_(b2).add_((1 - b2) * gg * gg)
    p.sub_((lr / bc1) * m / ((v / bc2).sqrt() + eps))


def ema_update(ema, p, decay_dev):
    dcy = float(decay_dev.reshape(-1)[0])
    ema.mul_(dcy).add_((1 - dcy) * p)


def avgpool_fwd(x):
    y = torch.empty((x.shape[0], x.shape[1], 1, 1), dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y.copy_(_bf16(x.float().mean((2, 3), keepdim=True)))
    return y


def avgpool_bwd(dy, hw_shape):
    h, w = hw_shape
    dx = torch.empty((dy.shape[0], dy.shape[1], h, w), dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dx.copy_(_bf16((dy.float() / (h * w)).expand(-1, -1, h, w)))
    return dx


# ---- pose loss: the product's own kernel arithmetic (csrc/pose_loss_math.cuh) compiled for the host
_POSE_HOST = {}


def _pose_host():
    if "h" not in _POSE_HOST:
        import tempfile

        import host_pose_loss

        _POSE_HOST["dir"] = tempfile.mkdtemp(prefix="sgb_pose_host_")
        _POSE_HOST["h"] = host_pose_loss.build(_POSE_HOST["dir"])
    return _POSE_HOST["h"]


def pose_tal_assign(d, cls_logits, reg_distri, pose_coords, anchor_points, stride_tensor, gt_boxes, gt_poses, gt_crowd, gt_valid, sigmas, sums):
    import host_pose_loss

    dummy = torch.zeros((d.B, d.L, d.J))
    r = host_pose_loss.run(_pose_host(), d, cls_logits, reg_distri, pose_coords, dummy, anchor_points, stride_tensor, gt_boxes, gt_poses, gt_crowd, gt_valid, sigmas)
    sums[3] += r["sums"][3]
    sums[6] += r["sums"][6]
    _POSE_HOST["targets"] = (gt_crowd, gt_valid)
    return r["assigned_gt"], r["assigned_score"]


def pose_loss(d, cls_logits, reg_distri, pose_coords, pose_logits, anchor_points, stride_tensor, gt_boxes, gt_poses, sigmas, agt, asc, sums, grad_scale=1.0, want_grad=True):
    import host_pose_loss

    gt_crowd, gt_valid = _POSE_HOST["targets"]  # the host driver redoes the assignment (same inputs -> same result)
    r = host_pose_loss.run(_pose_host(), d, cls_logits, reg_distri, pose_coords, pose_logits, anchor_points, stride_tensor, gt_boxes, gt_poses, gt_crowd, gt_valid, sigmas, grad_scale)
    assert torch.equal(r["assigned_gt"], agt) and torch.equal(r["assigned_score"], asc)
    for k in (0, 1, 2, 4, 5):
        sums[k] += r["sums"][k]
    gc, gr, gp, gl = r["grads"]
    return r["items"], gc.reshape(cls_logits.shape), gr, gp, gl


_TRAINING = dict(stem_patches=stem_patches, conv_dgrad=conv_dgrad, conv_wgrad=conv_wgrad, wgrad_to_oihw=wgrad_to_oihw, weight_prepare_batch=weight_prepare_batch,
                 run_weight_prepare_batch=run_weight_prepare_batch, wgrad_to_oihw_batch_table=wgrad_to_oihw_batch_table, run_wgrad_to_oihw_batch=run_wgrad_to_oihw_batch, qarep_alpha_finish_table=qarep_alpha_finish_table, run_qarep_alpha_finish=run_qarep_alpha_finish,
                 bn_act_fwd=bn_act_fwd, bn_act_bwd=bn_act_bwd, channel_stats=channel_stats, channel_dot=channel_dot, qarep_fwd=qarep_fwd, qarep_bwd=qarep_bwd,
                 maxpool_bwd=maxpool_bwd, head_grad_scatter=head_grad_scatter, tal_assign=tal_assign, atss_assign=atss_assign, dfl_iou_loss=dfl_iou_loss, sgd_step=sgd_step,
                 adamw_step=adamw_step, ema_update=ema_update, pose_tal_assign=pose_tal_assign, pose_loss=pose_loss, avgpool_fwd=avgpool_fwd, avgpool_bwd=avgpool_bwd)  # fmt: skip


def preprocess_u8(*args, **kwargs):
    """The product's own pre-processing arithmetic (csrc/preprocess_math.cuh) compiled for the host."""
    import host_preprocess

    return host_preprocess.preprocess_u8(*args, **kwargs)


def detection_matching(*args, **kwargs):
    """The product header's arithmetic behind a serial host driver (tests/host_detection_match.py)."""
    import host_detection_match

    return host_detection_match.detection_matching(*args, **kwargs)


_SUBSET = dict(conv_fprop=conv_fprop, weight_prepare=weight_prepare, convt2x2_fprop=convt2x2_fprop, nchw_f32_to_nhwc_bf16=nchw_f32_to_nhwc_bf16,
               nhwc_bf16_to_nchw_f32=nhwc_bf16_to_nchw_f32, bn_act_infer=bn_act_infer, maxpool_fwd=maxpool_fwd, axpby=axpby, scale_add=scale_add, scale_add_dot=scale_add_dot,
               dfl_decode=dfl_decode, pose_keypoint_decode=pose_keypoint_decode, batched_nms=batched_nms, preprocess_u8=preprocess_u8, detection_matching=detection_matching)  # fmt: skip


def install(monkeypatch):
    """Routes the inference subset of kernels.py to the CPU stand-ins and makes every other kernel call fail loudly."""

    def refuse(name, *args):
        raise NotImplementedError(f"CPU stand-in: {name} is outside the inference subset")

    monkeypatch.setattr(K, "require_cuda", lambda t, name="tensor": None)
    monkeypatch.setattr(L, "call", refuse)
    for name, fn in _SUBSET.items():
        monkeypatch.setattr(K, name, fn)


def install_training(monkeypatch):
    """install() plus the training subset; host-pinned staging buffers become plain host tensors."""
    install(monkeypatch)
    for name, fn in _TRAINING.items():
        monkeypatch.setattr(K, name, fn)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
