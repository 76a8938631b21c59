"""Torch (CPU) restatements of the C-ABI operators the training path calls (usot_b200/train.py), with the SAME argument lists.
TEST INFRASTRUCTURE, used two ways:

  * tests/test_train_graph_cpu.py swaps them in for the CUDA library (``install``) and runs ``usot_b200.train.forward_train`` +
    ``.backward()`` on the CPU: the Python side of the training path (graph wiring, BatchNorm modes, the stride-2 dgrad decomposition,
    layouts, detach points) is then checked against the live reference's 234 parameter gradients without a GPU;
  * tests/test_gpu_train_ops.py compares every CUDA operator with its restatement here on random inputs.

Every backward formula below is written from the mathematics (or delegated to torch autograd), not from the CUDA source.
"""
import contextlib

import torch
import torch.nn.functional as F

import usot_oracle as O


# ---- forward ops (tensor-level stand-ins for usot_b200.ops functions) -----------------------------------------------------------
def conv2d_nhwc(x, weight_oihw, scale, shift, stride=1, padding=(0, 0), dilation=(1, 1), residual=None, relu=False, precision="fp32", in_scale=None):
    if in_scale is not None:
        x = x * in_scale
    y = F.conv2d(x.permute(0, 3, 1, 2), weight_oihw, None, stride, padding, dilation)
    y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    y = y.permute(0, 2, 3, 1)
    if residual is not None:
        y = y + residual
    return (F.relu(y) if relu else y).contiguous()


def pred_conv(x, weight_oihw, bias, mode=0, mul=0.1, adjust=None, bias4=None):
    assert mode == 0
    return (mul * F.conv2d(x.permute(0, 3, 1, 2), weight_oihw, bias, 1, 1)).contiguous()


def maxpool3x3s2p1_nhwc(x, split=False):
    return F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).contiguous()


def conf_fusion(conf, value, nq):
    b = conf.shape[0] // nq
    e = torch.exp(torch.clamp(conf, max=4, min=-6)).view((b, nq) + tuple(conf.shape[1:]))
    return ((e / e.sum(dim=1, keepdim=True)) * value.view_as(e)).sum(dim=1)


def weighted_bce(pred, label):
    return torch.as_tensor(O.weighted_bce(pred, label), dtype=torch.float32)


def iou_loss(bbox, target, weight):
    return O.iou_loss(bbox, target, weight)


def cycle_glue(off_cls, mem_cls, off_bbox, cls_ratio, search_size=255, search_feature_size=25):
    n, r = off_bbox.shape[0], off_bbox.shape[-1]
    res = cls_ratio * off_cls.reshape(n, -1) + (1 - cls_ratio) * mem_cls.reshape(n, -1)
    best = res.max(dim=1)
    to_img = O.pred_offset_to_image_bbox(off_bbox, r, search_size).view(n, 4, -1).transpose(1, 2)
    box = torch.gather(to_img, 1, best.indices.view(n, 1, 1).repeat(1, 1, 4)).view(n, 4)
    return O.image_bbox_to_prpool_bbox(box, search_feature_size, search_size).float(), best.values, best.indices.int()


def prroi_pool2d(features, rois, ph, pw, scale):
    return O._PrRoIPoolFn.apply(features, rois) if features.requires_grad else O.prroi_pool2d(features, rois, ph, pw, scale)


def xcorr_depthwise(x, kernel):
    return O.xcorr_depthwise(x, kernel)


# ---- pointer-level C-ABI calls (tensors stand in for the pointers) ---------------------------------------------------------------
def _conv_from_kn(w_kn, cin, cout, kh, kw):
    return w_kn.view(kh, kw, cin, cout).permute(3, 2, 0, 1).contiguous()


def usot_nhwc_to_nchw(x, n, h, w, c, out, stream):
    out.copy_(x.view(n, h, w, c).permute(0, 3, 1, 2))


def usot_nchw_to_nhwc(x, n, c, h, w, out, stream):
    out.copy_(x.view(n, c, h, w).permute(0, 2, 3, 1))


def usot_stem_conv_raw(x, n, size, w_kn, out, stream):
    w = w_kn.t().reshape(64, 3, 7, 7)
    out.copy_(F.conv2d(x, w, None, 2, 0).permute(0, 2, 3, 1))


@torch.enable_grad()
def usot_stem_conv_wgrad(x, gout, n, size, gw_kn, stream):
    wt = torch.zeros(64, 3, 7, 7, requires_grad=True)
    y = F.conv2d(x, wt, None, 2, 0)
    (g,) = torch.autograd.grad(y, wt, gout.view(n, y.shape[2], y.shape[3], 64).permute(0, 3, 1, 2))
    gw_kn.copy_(g.reshape(64, 147).t())


@torch.enable_grad()
def usot_conv2d_wgrad_nhwc(x, gout, n, h, w, cin, cout, kh, kw, stride, ph, pw, dh, dw, gw_kn, precision, stream):
    xx = x.view(n, h, w, cin).permute(0, 3, 1, 2).detach().clone()
    wt = torch.zeros(cout, cin, kh, kw, requires_grad=True)
    y = F.conv2d(xx, wt, None, stride, (ph, pw), (dh, dw))
    (g,) = torch.autograd.grad(y, wt, gout.view(n, y.shape[2], y.shape[3], cout).permute(0, 3, 1, 2))
    gw_kn.copy_(g.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout))


@torch.enable_grad()
def usot_conv2d_dgrad_nhwc(gout, w_kn, n, h, w, cin, cout, kh, kw, stride, ph, pw, dh, dw, gin, stream):
    xx = torch.zeros(n, cin, h, w, requires_grad=True)
    y = F.conv2d(xx, _conv_from_kn(w_kn, cin, cout, kh, kw), None, stride, (ph, pw), (dh, dw))
    (g,) = torch.autograd.grad(y, xx, gout.view(n, y.shape[2], y.shape[3], cout).permute(0, 3, 1, 2))
    gin.copy_(g.permute(0, 2, 3, 1))


def usot_bn_stats(x, bias, m, c, mean, var, stream):
    v = x.reshape(m, c).double() + (0 if bias is None else bias.double())
    mean.copy_(v.mean(0))
    var.copy_(v.var(0, unbiased=False))


def usot_bn_apply(x, bias, mean, var, eps, gamma, beta, residual, relu, m, c, y, invstd, stream):
    v = x.reshape(m, c) + (0 if bias is None else bias)
    is_ = 1.0 / torch.sqrt(var.double() + eps)
    o = (v - mean) * is_.float() * gamma + beta
    if residual is not None:
        o = o + residual.reshape(m, c)
    y.copy_((F.relu(o) if relu else o).view_as(y))
    if invstd is not None:
        invstd.copy_(is_.float())


def usot_bn_backward(gy, y, x, bias, mean, invstd, gamma, train, relu, m, c, gx, ggamma, gbeta, gres, stream):
    """Batch-norm backward from the textbook formulas (double precision):
       dz = gy * [y > 0];  dbeta = sum dz;  dgamma = sum dz * xhat;
       train: dx = gamma * invstd * (dz - dbeta / m - xhat * dgamma / m);   eval: dx = gamma * invstd * dz."""
    dz = gy.reshape(m, c).double()
    if relu:
        dz = dz * (y.reshape(m, c) > 0)
    xh = (x.reshape(m, c).double() + (0 if bias is None else bias.double()) - mean.double()) * invstd.double()
    db, dg = dz.sum(0), (dz * xh).sum(0)
    t = dz - (db + xh * dg) / m if train else dz
    gx.copy_((gamma.double() * invstd.double() * t).float().view_as(gx))
    ggamma.copy_(dg.float())
    gbeta.copy_(db.float())
    if gres is not None:
        gres.copy_(dz.float().view_as(gres))


def usot_pow2_scale(x, numel, target_log2, y, scale2, stream):
    import math
    m = float(x.abs().max())
    s = 1.0 if not (m > 0 and math.isfinite(m)) else 2.0 ** (target_log2 + 1 - math.frexp(m)[1])
    scale2.copy_(torch.tensor([s, 1.0 / s]))
    if y is not None:
        y.copy_(x * s)


def usot_channel_sum(x, m, c, out, stream):
    out.copy_(x.reshape(m, c).double().sum(0).float())


@torch.enable_grad()
def usot_maxpool3x3s2p1_backward_nhwc(x, gout, n, h, w, c, gin, stream):
    xx = x.view(n, h, w, c).permute(0, 3, 1, 2).detach().clone().requires_grad_(True)
    y = F.max_pool2d(xx, 3, 2, 1)
    (g,) = torch.autograd.grad(y, xx, gout.view(n, y.shape[2], y.shape[3], c).permute(0, 3, 1, 2))
    gin.copy_(g.permute(0, 2, 3, 1))


@torch.enable_grad()
def usot_conf_fusion_backward(conf, value, gout, b, nq, per_map, gconf, gvalue, stream):
    cf = conf.detach().clone().requires_grad_(True)
    va = value.detach().clone().requires_grad_(True)
    out = conf_fusion(cf, va, nq)
    g1, g2 = torch.autograd.grad(out, (cf, va), gout.view_as(out))
    gconf.copy_(g1)
    gvalue.copy_(g2)


def usot_weighted_sum3(x0, x1, x2, w3, numel, out, stream):
    out.copy_(w3[0] * x0 + w3[1] * x1 + w3[2] * x2)


def usot_weighted_sum3_backward(x0, x1, x2, w3, g, numel, g0, g1, g2, gw, stream):
    g0.copy_(w3[0] * g); g1.copy_(w3[1] * g); g2.copy_(w3[2] * g)
    gw.copy_(torch.stack([(g.double() * x.double()).sum() for x in (x0, x1, x2)]).float())


@torch.enable_grad()
def usot_weighted_bce_backward(pred, label, count, gloss, gpred, stream):
    p = pred.detach().clone().requires_grad_(True)
    loss = O.weighted_bce(p, label)
    if not torch.is_tensor(loss) or not loss.requires_grad:
        gpred.zero_()
        return
    (g,) = torch.autograd.grad(loss, p, gloss.reshape(()))
    gpred.copy_(g)


@torch.enable_grad()
def usot_iou_loss_backward(bbox, target, weight, n, cells, gloss, gbbox, stream):
    b = bbox.detach().clone().requires_grad_(True)
    (g,) = torch.autograd.grad(O.iou_loss(b, target, weight), b, gloss.reshape(()))
    gbbox.copy_(g)


@contextlib.contextmanager
def install():
    """Swap the CUDA library out of usot_b200.train / usot_b200.ops for the restatements above (CPU tensors)."""
    import sys
    from usot_b200 import _lib, ops, train
    me = sys.modules[__name__]
    saved = {}

    def patch(obj, name, val):
        saved[(obj, name)] = getattr(obj, name)
        setattr(obj, name, val)

    patch(train, "_lib_call", lambda name, dev, *args: getattr(me, name)(*args))
    patch(train, "_stream", lambda t: 0)
    patch(_lib, "ptr", lambda t: t)
    for name in ("conv2d_nhwc", "pred_conv", "maxpool3x3s2p1_nhwc", "conf_fusion", "weighted_bce", "iou_loss", "cycle_glue", "prroi_pool2d",
                 "xcorr_depthwise"):
        patch(ops, name, getattr(me, name))
    try:
        yield
    finally:
        for (obj, name), val in saved.items():
            setattr(obj, name, val)
