"""CPU oracle for the USOT per-frame forward path.  TEST INFRASTRUCTURE ONLY.

This file is a functional restatement (torch fp32 on the CPU + numpy for
PrRoIPool) of the reference algorithm behind ``lib/models`` ->
``USOT.template()/track()/extract_memory_feature()/forward()`` and of the tensor
path of ``lib/tracker/usot_tracker.py``.  It is the *checker* for the CUDA path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it.  Nothing under ``usot_b200/`` imports
it and the product never falls back to it.

Pinning: every function below is checked against the *live* reference modules
(imported from /root/reference with a harness-side ``.cuda`` shim) by
``oracle/gen_golden.py``, which also writes the fixtures under ``tests/golden``.
The reference itself ships no golden vectors / tests for this path
(SURVEY.md §4), and the conv/BN arithmetic lives in PyTorch (the reference pins
torch 1.7.1, ``README.md:67``; here torch 2.11 CPU) -- so the pin is "reference
modules run here on identical inputs".  PrRoIPool has no CPU implementation in
the reference (``lib/models/prroi_pool/functional.py:62-63``); its restatement
follows ``src/prroi_pooling_gpu_impl.cu:37-42,71-106,149-212`` and is checked on
the GPU box against the reference ``.cu`` compiled unchanged into
``oracle/_ref/libprroi_ref.so`` (see ``oracle/Makefile``).

All tensors are NCHW float32 like the reference.  ``sd`` is a reference-layout
``state_dict`` (444 tensors).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
BN_EPS = 1e-5  # nn.BatchNorm2d default, used everywhere in the reference

# ----------------------------------------------------------------------------
# Architecture tables (restated from the reference constructors)
# ----------------------------------------------------------------------------
# ResNet_plus2(Bottleneck, [3, 4, 6, 3], used_layers=[3])  lib/models/modules.py:61-135
#   layer1: planes 64,  3 blocks, stride 1, dil 1, 1x1 shortcut          (modules.py:110-115)
#   layer2: planes 128, 4 blocks, stride 2, dil 1, 3x3 s2 p0 shortcut    (modules.py:116-126)
#   layer3: planes 256, 6 blocks, stride 1, dil 2, 3x3 s1 p1 shortcut    (modules.py:116-126)
LAYERS = (("layer1", 64, 3, 1, 1), ("layer2", 128, 4, 2, 1), ("layer3", 256, 6, 1, 2))
BK = "features.features."


def _bottleneck_cfg(stride: int, dilation: int, has_down: bool) -> Tuple[int, int]:
    """(padding, dilation) of Bottleneck.conv2 -- lib/models/modules.py:18-29."""
    padding = 2 - stride
    if has_down and dilation > 1:
        dilation = dilation // 2
        padding = dilation
    if dilation > 1:
        padding = dilation
    return padding, dilation


def conv_specs() -> List[dict]:
    """Every conv of the model with its reference hyper-parameters, in state_dict order.

    Each entry: name (state_dict prefix of the conv), cin, cout, k, stride,
    pad (h, w), dil (h, w), bias (bool), bn (state_dict prefix or None).
    """
    specs: List[dict] = []

    def add(name, cin, cout, k, stride=1, pad=0, dil=1, bias=False, bn=None):
        pad = (pad, pad) if isinstance(pad, int) else tuple(pad)
        dil = (dil, dil) if isinstance(dil, int) else tuple(dil)
        specs.append(dict(name=name, cin=cin, cout=cout, k=k, stride=stride, pad=pad, dil=dil,
                          bias=bias, bn=bn))

    add(BK + "conv1", 3, 64, 7, stride=2, pad=0, bn=BK + "bn1")  # modules.py:70-72
    inplanes = 64
    for lname, planes, blocks, stride, dilation in LAYERS:
        for i in range(blocks):
            p = f"{BK}{lname}.{i}."
            has_down = i == 0
            s = stride if i == 0 else 1
            pad2, dil2 = _bottleneck_cfg(s, dilation, has_down)
            add(p + "conv1", inplanes, planes, 1, bn=p + "bn1")
            add(p + "conv2", planes, planes, 3, stride=s, pad=pad2, dil=dil2, bn=p + "bn2")
            add(p + "conv3", planes, planes * 4, 1, bn=p + "bn3")
            if has_down:
                if stride == 1 and dilation == 1:  # modules.py:110-115
                    add(p + "downsample.0", inplanes, planes * 4, 1, bn=p + "downsample.1")
                else:  # modules.py:116-126
                    dpad = dilation // 2 if dilation > 1 else 0
                    add(p + "downsample.0", inplanes, planes * 4, 3, stride=stride, pad=dpad,
                        bn=p + "downsample.1")
                inplanes = planes * 4
    add("neck.downsample.0", 1024, 256, 1, bn="neck.downsample.1")  # connect.py:287-290
    for enc in ("cls_encode", "reg_encode"):  # connect.py:20-53
        for m, dil in (("matrix11", (1, 1)), ("matrix12", (2, 1)), ("matrix21", (1, 2))):
            for br in ("k", "s"):
                p = f"connect_model.{enc}.{m}_{br}."
                add(p + "0", 256, 256, 3, dil=dil, bn=p + "1")
    for g in ("conf_gen", "value_gen"):  # connect.py:112-121
        p = f"connect_model.conf_fusion.{g}."
        add(p + "0", 256, 256, 3, pad=1, bias=True, bn=p + "1")
    for tw in ("bbox_tower", "cls_tower", "cls_memory_tower"):  # connect.py:178-209
        for i in range(4):
            p = f"connect_model.{tw}."
            add(p + str(3 * i), 256, 256, 3, pad=1, bias=True, bn=p + str(3 * i + 1))
    add("connect_model.bbox_pred", 256, 4, 3, pad=1, bias=True)  # connect.py:212-216
    add("connect_model.cls_pred", 256, 1, 3, pad=1, bias=True)
    add("connect_model.cls_memory_pred", 256, 1, 3, pad=1, bias=True)
    return specs


def make_state_dict(seed: int = 0, damp: Optional[float] = None) -> SD:
    """Seeded synthetic weights with the reference's state_dict keys and shapes.

    Conv weights ~ N(0, sqrt(2 / (k*k*cout))) as in modules.py:96-102 (applied to every conv
    here except the three Cout<=4 prediction convs, which get std 0.03 so that ``exp`` stays sane); conv biases ~ U(-0.1, 0.1); BN weight 1 (``bn3.weight`` = ``damp`` if given, which
    tames the chaotic error growth of a random residual net), BN bias 0,
    running stats (0, 1) until ``calibrate_bn`` fills them.
    """
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for s in conv_specs():
        n = s["k"] * s["k"] * s["cout"]
        std = 0.03 if s["name"].endswith("_pred") else math.sqrt(2.0 / n)  # FCOS-style small init for the 3 pred convs
        sd[s["name"] + ".weight"] = torch.randn(s["cout"], s["cin"], s["k"], s["k"], generator=g) * std
        if s["bias"]:
            sd[s["name"] + ".bias"] = (torch.rand(s["cout"], generator=g) - 0.5) * 0.2
        if s["bn"] is not None:
            c = s["cout"]
            w = torch.ones(c)
            if damp is not None and s["bn"].endswith("bn3"):
                w = w * damp
            sd[s["bn"] + ".weight"] = w
            sd[s["bn"] + ".bias"] = torch.zeros(c)
            sd[s["bn"] + ".running_mean"] = torch.zeros(c)
            sd[s["bn"] + ".running_var"] = torch.ones(c)
            sd[s["bn"] + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd["connect_model.cls_dw.weight"] = torch.tensor([1.0, 0.5, 1.5])  # connect.py:84 (ones in the reference)
    sd["connect_model.reg_dw.weight"] = torch.tensor([0.7, 1.2, 1.0])
    sd["connect_model.adjust"] = 0.1 * torch.ones(1)  # connect.py:218
    sd["connect_model.bias"] = torch.ones(1, 4, 1, 1)  # connect.py:219
    return sd


def bn_stat_keys(sd: SD) -> List[str]:
    return [k for k in sd if k.endswith("running_mean") or k.endswith("running_var")]


# ----------------------------------------------------------------------------
# Building blocks
# ----------------------------------------------------------------------------
class _Calib:
    """When active, BN layers use batch statistics and record them into ``sd`` (a one-shot
    stand-in for the reference's train()-mode running-stat update with momentum=None)."""

    def __init__(self):
        self.on = False


_CAL = _Calib()


def _bn(sd: SD, x: torch.Tensor, p: str) -> torch.Tensor:
    if _CAL.on:
        mean = x.mean(dim=(0, 2, 3))
        var_b = x.var(dim=(0, 2, 3), unbiased=False)
        n = x.numel() // x.shape[1]
        sd[p + ".running_mean"] = mean.clone()
        sd[p + ".running_var"] = (var_b * (n / max(n - 1, 1))).clone()  # unbiased, as torch stores it
        sd[p + ".num_batches_tracked"] = torch.tensor(1, dtype=torch.long)
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], True, 0.0, BN_EPS)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def _conv(sd: SD, x: torch.Tensor, p: str, stride=1, pad=0, dil=1) -> torch.Tensor:
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=pad, dilation=dil)


def bottleneck(sd: SD, x: torch.Tensor, p: str, stride: int, dilation: int, has_down: bool,
               down_k: int, down_stride: int, down_pad: int) -> torch.Tensor:
    """Bottleneck.forward -- lib/models/modules.py:37-58."""
    pad2, dil2 = _bottleneck_cfg(stride, dilation, has_down)
    out = F.relu(_bn(sd, _conv(sd, x, p + "conv1"), p + "bn1"))
    out = F.relu(_bn(sd, _conv(sd, out, p + "conv2", stride, pad2, dil2), p + "bn2"))
    out = _bn(sd, _conv(sd, out, p + "conv3"), p + "bn3")
    residual = x
    if has_down:
        residual = _bn(sd, _conv(sd, x, p + "downsample.0", down_stride, down_pad), p + "downsample.1")
    return F.relu(out + residual)


def backbone(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """ResNet_plus2.forward -> p3 -- lib/models/modules.py:137-151 (only p3 is used, models.py:174,181)."""
    x = F.relu(_bn(sd, _conv(sd, x, BK + "conv1", 2, 0), BK + "bn1"))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for lname, planes, blocks, stride, dilation in LAYERS:
        for i in range(blocks):
            p = f"{BK}{lname}.{i}."
            if i == 0:
                if stride == 1 and dilation == 1:
                    dk, dpad = 1, 0
                else:
                    dk, dpad = 3, (dilation // 2 if dilation > 1 else 0)
                x = bottleneck(sd, x, p, stride, dilation, True, dk, stride, dpad)
            else:
                x = bottleneck(sd, x, p, 1, dilation, False, 0, 1, 0)
    return x


def neck(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """AdjustLayer.downsample: 1x1 + BN, no ReLU -- lib/models/connect.py:287-296."""
    return _bn(sd, _conv(sd, x, "neck.downsample.0"), "neck.downsample.1")


def backbone_neck(sd: SD, x: torch.Tensor) -> torch.Tensor:
    return neck(sd, backbone(sd, x))


# ---- PrRoIPool forward (numpy float32) --------------------------------------
def _f32(v):
    return np.float32(v)


def prroi_pool2d(features: torch.Tensor, rois: torch.Tensor, pooled_h: int = 7, pooled_w: int = 7,
                 spatial_scale: float = 1.0) -> torch.Tensor:
    """PrRoIPoolingForward -- lib/models/prroi_pool/src/prroi_pooling_gpu_impl.cu:149-212.

    ``rois`` is (N, 5) = [batch_idx, x1, y1, x2, y2].  Loops follow the kernel: cells
    [floor(win_start), ceil(win_end)) in w then h, four corner terms per cell (:71-106), zero
    outside the map (:37-42), divide by the bin area, 0 when the area is 0 (:189-193).
    Vectorised over channels only; float32 arithmetic throughout.
    """
    feat = features.detach().cpu().numpy().astype(np.float32)
    r = rois.detach().cpu().numpy().astype(np.float32)
    n_rois = r.shape[0]
    _, C, H, W = feat.shape
    out = np.zeros((n_rois, C, pooled_h, pooled_w), np.float32)
    scale = _f32(spatial_scale)

    def get(data, h, w):  # PrRoIPoolingGetData :37-42
        if h < 0 or w < 0 or h >= H or w >= W:
            return np.zeros((C,), np.float32)
        return data[:, h, w]

    def g(lim, a):  # the 1-D antiderivative difference used in :80-81
        return lim - _f32(0.5) * lim * lim - a + _f32(0.5) * a * a

    for n in range(n_rois):
        b = int(r[n, 0])
        data = feat[b]
        sw, sh, ew, eh = (r[n, 1] * scale, r[n, 2] * scale, r[n, 3] * scale, r[n, 4] * scale)
        roi_w = max(ew - sw, _f32(0.0))
        roi_h = max(eh - sh, _f32(0.0))
        bin_h = _f32(roi_h / _f32(pooled_h))
        bin_w = _f32(roi_w / _f32(pooled_w))
        win_size = max(_f32(0.0), _f32(bin_w * bin_h))
        if win_size == 0:
            continue
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                ws_w = _f32(sw + _f32(bin_w * _f32(pw)))
                ws_h = _f32(sh + _f32(bin_h * _f32(ph)))
                we_w = _f32(ws_w + bin_w)
                we_h = _f32(ws_h + bin_h)
                s_w, e_w = int(math.floor(ws_w)), int(math.ceil(we_w))
                s_h, e_h = int(math.floor(ws_h)), int(math.ceil(we_h))
                acc = np.zeros((C,), np.float32)
                for wi in range(s_w, e_w):
                    for hi in range(s_h, e_h):
                        y0 = max(ws_h, _f32(hi))
                        x0 = max(ws_w, _f32(wi))
                        y1 = min(we_h, _f32(hi + 1))
                        x1 = min(we_w, _f32(wi + 1))
                        # PrRoIPoolingMatCalculation :71-106 with (s_h, s_w, e_h, e_w) = (hi, wi, hi+1, wi+1)
                        a, bt = _f32(x0 - _f32(wi)), _f32(y0 - _f32(hi))
                        la, lb = _f32(x1 - _f32(wi)), _f32(y1 - _f32(hi))
                        cell = get(data, hi, wi) * _f32(g(la, a) * g(lb, bt))
                        a2, la2 = _f32(_f32(wi + 1) - x1), _f32(_f32(wi + 1) - x0)
                        cell = cell + get(data, hi, wi + 1) * _f32(g(la2, a2) * g(lb, bt))
                        b2, lb2 = _f32(_f32(hi + 1) - y1), _f32(_f32(hi + 1) - y0)
                        cell = cell + get(data, hi + 1, wi) * _f32(g(la, a) * g(lb2, b2))
                        cell = cell + get(data, hi + 1, wi + 1) * _f32(g(la2, a2) * g(lb2, b2))
                        acc = acc + cell
                out[n, :, ph, pw] = acc / win_size
    return torch.from_numpy(out)


def prroi_pool2d_backward(grad_out: torch.Tensor, rois: torch.Tensor, feat_shape, spatial_scale: float = 1.0) -> torch.Tensor:
    """PrRoIPoolingBackward -- lib/models/prroi_pool/src/prroi_pooling_gpu_impl.cu:214-272 (gradient w.r.t. the features; the
    transpose of ``prroi_pool2d``, which is linear in the features: every corner of every cell of every bin receives
    coef * top_diff / bin_area, :108-147).  float32, vectorised over channels."""
    go = grad_out.detach().cpu().numpy().astype(np.float32)
    r = rois.detach().cpu().numpy().astype(np.float32)
    B, C, H, W = feat_shape
    n_rois, _, pooled_h, pooled_w = go.shape
    gf = np.zeros((B, C, H, W), np.float32)
    scale = _f32(spatial_scale)

    def g(lim, a):
        return lim - _f32(0.5) * lim * lim - a + _f32(0.5) * a * a

    def add(b, h, w, v):  # PrRoIPoolingDistributeDiff: nothing outside the map
        if 0 <= h < H and 0 <= w < W:
            gf[b, :, h, w] += v

    for n in range(n_rois):
        b = int(r[n, 0])
        sw, sh, ew, eh = (r[n, 1] * scale, r[n, 2] * scale, r[n, 3] * scale, r[n, 4] * scale)
        roi_w = max(ew - sw, _f32(0.0))
        roi_h = max(eh - sh, _f32(0.0))
        bin_h = _f32(roi_h / _f32(pooled_h))
        bin_w = _f32(roi_w / _f32(pooled_w))
        win_size = max(_f32(0.0), _f32(bin_w * bin_h))
        if win_size == 0:
            continue
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                ws_w = _f32(sw + _f32(bin_w * _f32(pw)))
                ws_h = _f32(sh + _f32(bin_h * _f32(ph)))
                we_w = _f32(ws_w + bin_w)
                we_h = _f32(ws_h + bin_h)
                s_w, e_w = int(math.floor(ws_w)), int(math.ceil(we_w))
                s_h, e_h = int(math.floor(ws_h)), int(math.ceil(we_h))
                top = go[n, :, ph, pw] / win_size
                for wi in range(s_w, e_w):
                    for hi in range(s_h, e_h):
                        y0 = max(ws_h, _f32(hi))
                        x0 = max(ws_w, _f32(wi))
                        y1 = min(we_h, _f32(hi + 1))
                        x1 = min(we_w, _f32(wi + 1))
                        a, bt = _f32(x0 - _f32(wi)), _f32(y0 - _f32(hi))
                        la, lb = _f32(x1 - _f32(wi)), _f32(y1 - _f32(hi))
                        a2, la2 = _f32(_f32(wi + 1) - x1), _f32(_f32(wi + 1) - x0)
                        b2, lb2 = _f32(_f32(hi + 1) - y1), _f32(_f32(hi + 1) - y0)
                        add(b, hi, wi, top * _f32(g(la, a) * g(lb, bt)))
                        add(b, hi, wi + 1, top * _f32(g(la2, a2) * g(lb, bt)))
                        add(b, hi + 1, wi, top * _f32(g(la, a) * g(lb2, b2)))
                        add(b, hi + 1, wi + 1, top * _f32(g(la2, a2) * g(lb2, b2)))
    return torch.from_numpy(gf)


class _PrRoIPoolFn(torch.autograd.Function):
    """Differentiable (w.r.t. the features) wrapper used by the gradient fixtures; the roi coordinates get no gradient, exactly
    like the detached boxes of the training forward (lib/models/models.py:271-272) and the label boxes."""

    @staticmethod
    def forward(ctx, features, rois):
        ctx.save_for_backward(rois)
        ctx.feat_shape = tuple(features.shape)
        return prroi_pool2d(features, rois, 7, 7, 1.0)

    @staticmethod
    def backward(ctx, grad_out):
        (rois,) = ctx.saved_tensors
        return prroi_pool2d_backward(grad_out, rois, ctx.feat_shape, 1.0), None


def prpool_feature(features: torch.Tensor, bboxs: torch.Tensor) -> torch.Tensor:
    """USOT_.prpool_feature -- lib/models/models.py:164-171."""
    idx = torch.arange(0, features.shape[0]).view(-1, 1).float()
    rois = torch.cat((idx, bboxs.detach().float().cpu()), dim=1)
    if features.requires_grad:
        return _PrRoIPoolFn.apply(features, rois)
    return prroi_pool2d(features, rois, 7, 7, 1.0)


# ---- head -------------------------------------------------------------------
def _cbr(sd: SD, x: torch.Tensor, p: str, pad=0, dil=1) -> torch.Tensor:
    """conv3x3 -> BN -> ReLU, (p + '0', p + '1') Sequential naming."""
    return F.relu(_bn(sd, _conv(sd, x, p + "0", 1, pad, dil), p + "1"))


def matrix_encode(sd: SD, enc: str, z: Optional[torch.Tensor] = None, x: Optional[torch.Tensor] = None):
    """matrix.forward -- lib/models/connect.py:55-74: three parallel 3x3 convs on the SAME input."""
    p = f"connect_model.{enc}."
    zs = xs = None
    if x is not None:
        xs = [_cbr(sd, x, p + "matrix11_s."), _cbr(sd, x, p + "matrix12_s.", 0, (2, 1)),
              _cbr(sd, x, p + "matrix21_s.", 0, (1, 2))]
    if z is not None:
        zs = [_cbr(sd, z, p + "matrix11_k."), _cbr(sd, z, p + "matrix12_k.", 0, (2, 1)),
              _cbr(sd, z, p + "matrix21_k.", 0, (1, 2))]
    return zs, xs


def xcorr_depthwise(x: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    """lib/models/connect.py:147-157 (same view trick, so kernel batch 1 broadcasts over x batch)."""
    batch, channel, h_k, w_k = kernel.shape
    _, _, h_x, w_x = x.shape
    x = x.reshape(-1, batch * channel, h_x, w_x)
    kernel = kernel.reshape(batch * channel, 1, h_k, w_k)
    out = F.conv2d(x, kernel, groups=batch * channel)
    return out.view(-1, channel, out.size(2), out.size(3))


def groupdw(weight: torch.Tensor, z: Sequence[torch.Tensor], x: Sequence[torch.Tensor]) -> torch.Tensor:
    """GroupDW.forward -- lib/models/connect.py:86-102."""
    w = F.softmax(weight, 0)
    s = 0
    for i in range(3):
        s = s + w[i] * xcorr_depthwise(x[i], z[i])
    return s


def conf_fusion(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """Conf_Fusion.forward -- lib/models/connect.py:123-144.  x: (B, Nq, C, H, W)."""
    batch, mem, ch, h, w = x.shape
    x = x.reshape(-1, ch, h, w)
    conf = _cbr(sd, x, "connect_model.conf_fusion.conf_gen.", 1)
    conf = torch.exp(torch.clamp(conf, max=4, min=-6)).view(batch, mem, ch, h, w)
    conf_norm = conf / conf.sum(dim=1, keepdim=True)
    value = _cbr(sd, x, "connect_model.conf_fusion.value_gen.", 1).view(batch, mem, ch, h, w)
    return (conf_norm * value).sum(dim=1)


def tower(sd: SD, x: torch.Tensor, name: str) -> torch.Tensor:
    """4 x [conv3x3 p1 + bias, BN, ReLU] -- lib/models/connect.py:178-209."""
    for i in range(4):
        p = f"connect_model.{name}."
        x = F.relu(_bn(sd, _conv(sd, x, p + str(3 * i), 1, 1), p + str(3 * i + 1)))
    return x


def connect(sd: SD, search: torch.Tensor, kernel: Optional[torch.Tensor] = None,
            memory_kernel: Optional[torch.Tensor] = None, memory_confidence: Optional[torch.Tensor] = None,
            cls_x_store: Optional[List[torch.Tensor]] = None):
    """box_tower_reg.forward -- lib/models/connect.py:221-281.  Returns the same 5-tuple."""
    x_bbox = cls = cls_x = reg_x = None
    if kernel is not None:
        cls_z, cls_x = matrix_encode(sd, "cls_encode", kernel, search)
        reg_z, reg_x = matrix_encode(sd, "reg_encode", kernel, search)
        cls_dw = groupdw(sd["connect_model.cls_dw.weight"], cls_z, cls_x)
        reg_dw = groupdw(sd["connect_model.reg_dw.weight"], reg_z, reg_x)
        x_reg = tower(sd, reg_dw, "bbox_tower")
        x_bbox = torch.exp(sd["connect_model.adjust"] * _conv(sd, x_reg, "connect_model.bbox_pred", 1, 1)
                           + sd["connect_model.bias"])
        cls = 0.1 * _conv(sd, tower(sd, cls_dw, "cls_tower"), "connect_model.cls_pred", 1, 1)
        if memory_kernel is None:
            return x_bbox, cls, cls_x, reg_x, None
    if memory_kernel is not None:
        if cls_x_store is None:
            cls_mem_zs, cls_x_store = matrix_encode(sd, "cls_encode", memory_kernel, search)
        else:
            cls_mem_zs, _ = matrix_encode(sd, "cls_encode", memory_kernel, None)
        batch, mem = memory_confidence.shape  # only the shape is used (connect.py:258)
        rep = []
        for cx in cls_x_store:
            _, c, h, w = cx.shape
            rep.append(cx.view(batch, 1, c, h, w).repeat(1, mem, 1, 1, 1).view(-1, c, h, w))
        dw = groupdw(sd["connect_model.cls_dw.weight"], cls_mem_zs, rep)
        _, c, h, w = dw.shape
        fused = conf_fusion(sd, dw.view(batch, mem, c, h, w))
        cls_mem = 0.1 * _conv(sd, tower(sd, fused, "cls_memory_tower"), "connect_model.cls_memory_pred", 1, 1)
        if kernel is not None:
            return x_bbox, cls, cls_x, reg_x, cls_mem
        return None, None, None, None, cls_mem
    return None


# ---- model façade -----------------------------------------------------------
def template(sd: SD, z: torch.Tensor, template_bbox: Optional[torch.Tensor] = None, pr_pool: bool = True):
    """USOT_.template -- lib/models/models.py:173-177 (+ AdjustLayer crop paths, connect.py:298-314)."""
    x_ori = backbone_neck(sd, z)
    if not pr_pool:
        return x_ori[:, :, 4:-4, 4:-4]
    return prpool_feature(x_ori, template_bbox)


def track(sd: SD, zf: torch.Tensor, x: torch.Tensor, template_mem: Optional[torch.Tensor] = None,
          score_mem: Optional[torch.Tensor] = None):
    """USOT_.track -- lib/models/models.py:179-198."""
    xf = backbone_neck(sd, x)
    if template_mem is not None:
        bbox, cls, _, _, cls_mem = connect(sd, xf, kernel=zf, memory_kernel=template_mem, memory_confidence=score_mem)
        return cls, bbox, cls_mem, xf
    bbox, cls, _, _, _ = connect(sd, xf, kernel=zf)
    return cls, bbox, None, None


def extract_memory_feature(sd: SD, ori_x=None, xf=None, search_bbox=None):
    """USOT_.extract_memory_feature -- lib/models/models.py:200-206."""
    if ori_x is not None:
        xf = backbone_neck(sd, ori_x)
    return prpool_feature(xf, search_bbox)


def calibrate_bn(sd: SD, z: torch.Tensor, x: torch.Tensor, bbox: torch.Tensor, n_mem: int = 3) -> None:
    """Populate every BN's running stats from one batch-statistics pass over the whole model
    (SURVEY.md §8c oracle recipe, step 2).  Raw default-init weights overflow ``exp`` otherwise."""
    assert z.shape[0] == 1
    try:
        with torch.no_grad():
            _CAL.on = True  # template/search share the backbone BNs: calibrate on the search branch
            xf = backbone_neck(sd, x)
            _CAL.on = False
            zf = prpool_feature(backbone_neck(sd, z), bbox)
            mem = prpool_feature(xf, torch.tensor([[6.0, 7.0, 17.0, 18.0]]).repeat(x.shape[0], 1))
            mem = mem.repeat_interleave(n_mem, 0)
            _CAL.on = True  # encoders (_s from x, _k from the memory kernels), towers, conf_fusion
            connect(sd, xf, kernel=zf, memory_kernel=mem, memory_confidence=torch.ones(x.shape[0], n_mem))
    finally:
        _CAL.on = False


# ---- training forward (cycle memory) ----------------------------------------
def _grids(score_size=25, search_size=255, sf_size=25, stride=8):
    """USOT_.grids -- lib/models/models.py:102-129."""
    sz = score_size
    x, y = np.meshgrid(np.arange(0, sz) - np.floor(float(sz // 2)), np.arange(0, sz) - np.floor(float(sz // 2)))
    gx = torch.Tensor(x * stride + search_size // 2)
    gy = torch.Tensor(y * stride + search_size // 2)
    axis = (np.arange(0, sf_size) - np.floor(float(sf_size // 2))) * stride + search_size // 2
    return gx, gy, axis


def pred_offset_to_image_bbox(bbox_pred: torch.Tensor, score_size=25, search_size=255) -> torch.Tensor:
    """lib/models/models.py:131-148."""
    gx, gy, _ = _grids(score_size, search_size)
    gx, gy = gx[None, None], gy[None, None]
    return torch.cat([gx - bbox_pred[:, 0:1], gy - bbox_pred[:, 1:2], gx + bbox_pred[:, 2:3], gy + bbox_pred[:, 3:4]], 1)


def image_bbox_to_prpool_bbox(image_bbox: torch.Tensor, sf_size=25, search_size=255) -> torch.Tensor:
    """lib/models/models.py:150-162."""
    _, _, axis = _grids(25, search_size, sf_size)
    reg_min, reg_max = axis[0], axis[-1]
    sz = 2 * (sf_size // 2)
    gap = (reg_max - reg_min) / sz
    image_bbox = torch.clamp(image_bbox, max=reg_max + 2 * gap, min=reg_min - 2 * gap)
    return (image_bbox - reg_min) * (1.0 / gap)


def _cls_loss(pred, label, select):
    """lib/models/models.py:42-47 (returns 0 for a 0-d selection, quirk E5)."""
    if len(select.size()) == 0:
        return 0
    return F.binary_cross_entropy_with_logits(torch.index_select(pred, 0, select), torch.index_select(label, 0, select))


def weighted_bce(pred, label):
    """lib/models/models.py:49-58."""
    pred, label = pred.reshape(-1), label.reshape(-1)
    pos = label.eq(1).nonzero().squeeze()
    neg = label.eq(0).nonzero().squeeze()
    return _cls_loss(pred, label, pos) * 0.5 + _cls_loss(pred, label, neg) * 0.5


def iou_loss(bbox_pred, reg_target, reg_weight):
    """add_iouloss + _IOULoss -- lib/models/models.py:60-100."""
    p = bbox_pred.permute(0, 2, 3, 1).reshape(-1, 4)
    t = reg_target.reshape(-1, 4)
    idx = torch.nonzero(reg_weight.reshape(-1) > 0).squeeze(1)
    p, t = p[idx], t[idx]
    t_area = (t[:, 0] + t[:, 2]) * (t[:, 1] + t[:, 3])
    p_area = (p[:, 0] + p[:, 2]) * (p[:, 1] + p[:, 3])
    w_i = torch.min(p[:, 0], t[:, 0]) + torch.min(p[:, 2], t[:, 2])
    h_i = torch.min(p[:, 3], t[:, 3]) + torch.min(p[:, 1], t[:, 1])
    a_i = w_i * h_i
    a_u = t_area + p_area - a_i
    return (-torch.log((a_i + 1.0) / (a_u + 1.0))).mean()


def forward_train(sd: SD, template_img, search, label, reg_target, reg_weight, template_bbox,
                  search_memory=None, search_bbox=None, cls_ratio=0.40, detail: bool = False):
    """USOT_.forward -- lib/models/models.py:208-295 (eval-mode BN)."""
    zf = prpool_feature(backbone_neck(sd, template_img), template_bbox)
    xf = backbone_neck(sd, search)
    if search_memory is None:
        bbox_pred, cls_pred, _, _, _ = connect(sd, xf, kernel=zf)
        return weighted_bce(cls_pred, label), None, iou_loss(bbox_pred, reg_target, reg_weight)
    bbox_pred, cls_pred, cls_x, _, _ = connect(sd, xf, kernel=zf)
    reg_loss = iou_loss(bbox_pred, reg_target, reg_weight)
    cls_loss_ori = weighted_bce(cls_pred, label)
    batch, mem, cx, hx, wx = search_memory.shape
    xf_mem = backbone_neck(sd, search_memory.reshape(-1, cx, hx, wx))
    spf = prpool_feature(xf, search_bbox)
    spf = spf.view(batch, 1, *spf.shape[1:]).repeat(1, mem, 1, 1, 1).view(-1, *spf.shape[1:])
    zf_mem = zf.view(batch, 1, *zf.shape[1:]).repeat(1, mem, 1, 1, 1).view(-1, *zf.shape[1:])
    off_bbox, off_cls, fwd_store, _, _ = connect(sd, xf_mem, kernel=zf_mem)
    _, _, _, _, mem_cls = connect(sd, xf_mem, memory_kernel=spf, memory_confidence=torch.ones(batch * mem, 1),
                                  cls_x_store=fwd_store)
    mem_cls = mem_cls.view(batch, mem, -1)
    off_cls = off_cls.view(batch, mem, -1)
    res = cls_ratio * off_cls + (1 - cls_ratio) * mem_cls
    best = res.max(dim=2)
    arg = best.indices.view(batch, mem, 1, 1).repeat(1, 1, 1, 4)
    to_img = pred_offset_to_image_bbox(off_bbox).view(batch, mem, 4, -1).transpose(2, 3)
    best_box = torch.gather(to_img, dim=2, index=arg).view(batch * mem, 4)
    best_score = best.values
    pool_box = image_bbox_to_prpool_bbox(best_box)
    pooled = prpool_feature(xf_mem, pool_box)
    _, _, _, _, back = connect(sd, xf, memory_kernel=pooled, memory_confidence=best_score, cls_x_store=cls_x)
    cls_memory_loss = weighted_bce(back, label)
    if detail:
        return dict(cls_loss=cls_loss_ori, cls_memory_loss=cls_memory_loss, reg_loss=reg_loss, cls_pred=cls_pred,
                    bbox_pred=bbox_pred, forward_argmax=best.indices, best_box=best_box, pool_box=pool_box,
                    backward_map=back)
    return cls_loss_ori, cls_memory_loss, reg_loss


# ---- tracker tensor path ------------------------------------------------------
def tracker_update(cls_score, bbox_pred, cls_memory, target_sz_scaled, window, instance_size=255, score_size=25,
                   ratio=0.3, penalty_k=0.021, window_influence=0.321, stride=8):
    """The tensor path of USOTTracker.update -- lib/tracker/usot_tracker.py:137-163 (numpy, float64
    where the reference's numpy promotes).  Returns (r_max, c_max, pscore, penalty, cls, box4)."""
    cls = torch.sigmoid(cls_score).squeeze().cpu().numpy()
    cmem = torch.sigmoid(cls_memory).squeeze().cpu().numpy()
    cls = ratio * cls + (1 - ratio) * cmem
    bp = bbox_pred.squeeze().cpu().numpy()
    sz = score_size
    gx, gy = np.meshgrid(np.arange(0, sz) - np.floor(float(sz // 2)), np.arange(0, sz) - np.floor(float(sz // 2)))
    gx = gx * stride + instance_size // 2
    gy = gy * stride + instance_size // 2
    x1, y1, x2, y2 = gx - bp[0], gy - bp[1], gx + bp[2], gy + bp[3]

    def change(r):
        return np.maximum(r, 1.0 / r)

    def szf(w, h):
        pad = (w + h) * 0.5
        return np.sqrt((w + pad) * (h + pad))

    s_c = change(szf(x2 - x1, y2 - y1) / szf(target_sz_scaled[0], target_sz_scaled[1]))
    r_c = change((target_sz_scaled[0] / target_sz_scaled[1]) / ((x2 - x1) / (y2 - y1)))
    penalty = np.exp(-(r_c * s_c - 1) * penalty_k)
    pscore = penalty * cls
    pscore = pscore * (1 - window_influence) + window * window_influence
    r_max, c_max = np.unravel_index(pscore.argmax(), pscore.shape)
    return int(r_max), int(c_max), pscore, penalty, cls, np.array([x1[r_max, c_max], y1[r_max, c_max],
                                                                    x2[r_max, c_max], y2[r_max, c_max]])


# ---- synthetic inputs (shared by tests, golden generation, bench) --------------
def synth_inputs(seed: int, batch: int, search_size: int = 255, n_templates: int = 1):
    """U[0,255) crops plus plausible PrPool boxes in feature coordinates."""
    g = torch.Generator().manual_seed(seed)
    z = torch.rand(n_templates, 3, 127, 127, generator=g) * 255.0
    x = torch.rand(batch, 3, search_size, search_size, generator=g) * 255.0
    tb = torch.tensor([[3.3, 4.1, 10.6, 11.2]]).repeat(n_templates, 1) + torch.rand(n_templates, 4, generator=g)
    sb = torch.tensor([[8.2, 7.4, 16.9, 15.3]]).repeat(batch, 1) + torch.rand(batch, 4, generator=g) * 2.0
    return z, x, tb, sb
