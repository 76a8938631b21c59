"""Training path of the USOT façade: ``USOT.forward`` with an autograd graph, so that the reference's training
step -- ``loss.backward()`` at scripts/train_usot.py:229-236 over the three losses of lib/models/models.py:208-295 -- runs on this
library's kernels, in the reference's train()-mode BatchNorm semantics (batch statistics per feature_extractor / head call, running
statistics updated with momentum 0.1) or with running statistics (eval()).

Every arithmetic op is one of this library's CUDA kernels behind the C ABI (include/usot_b200.h), wrapped in a
``torch.autograd.Function`` whose backward calls the matching gradient kernel:

    conv (raw)          forward: usot_conv2d_nhwc (tcgen05 fp16x3 / fp32 FMA) or usot_stem_conv_raw / usot_pred_conv
                        dgrad:   the SAME forward kernels with transposed, flipped filters (stride 1; stride 2 = four parity
                                 sub-problems), usot_conv2d_dgrad_nhwc for the thin prediction convs
                        wgrad:   usot_conv2d_wgrad_nhwc
    BatchNorm (+bias, +residual, +ReLU)   usot_bn_stats / usot_bn_apply / usot_bn_backward / usot_channel_sum
    MaxPool             usot_maxpool3x3s2p1_nhwc / _backward_nhwc
    PrRoIPool           usot_prroi_pool_forward / _backward           (lib/models/prroi_pool/functional.py:41-81)
    xcorr, GroupDW      usot_xcorr_depthwise / _backward, usot_weighted_sum3 / _backward   (lib/models/connect.py:86-102,147-157)
    Conf_Fusion         usot_conf_fusion / _backward                  (lib/models/connect.py:123-144)
    losses              usot_weighted_bce / usot_iou_loss + their _backward kernels      (lib/models/models.py:42-100)
    forward-track glue  usot_cycle_glue (detached in the reference: models.py:273-274)

torch supplies the autograd graph, device memory and a handful of scalar / 3-element glue ops (softmax of the three GroupDW weights,
``exp(adjust * y + bias)`` on the (n,4,25,25) box map); no torch convolution, normalisation, pooling or loss kernel runs.
Activations are NHWC fp32.  There is no CPU fallback.
"""
import torch
from torch.autograd import Function

from . import _lib, ops
from .ops import _stream

BN_EPS = 1e-5

# Arithmetic of the backward convs (dgrad on the forward kernels, wgrad): None = the same mode as the forward convs (tcgen05 fp16x3 by
# default).  Gradient magnitudes sit far below fp16's normal range, so in the tensor-core modes every gradient map is first multiplied by a
# power of two computed on the device from its absmax (usot_pow2_scale) and the factor is divided out in the conv epilogue -- exact.
# "fp32" forces the fp32 FMA kernels.
BWD_PRECISION = None
GRAD_TARGET_LOG2 = 10


class Mode:
    """How one forward_train call runs: the arithmetic of the forward convs (the façade's precision mode) and the BatchNorm flavour.
    ``train=None`` (default): every BatchNorm follows ITS OWN ``module.training`` flag, exactly like torch -- the reference freezes backbone
    layers by putting their BatchNorms in eval() (scripts/train_usot.py:74-102) and then calls ``model.train()`` (:153); both states behave
    here as they do there.  ``train=True / False`` forces one flavour for the whole call."""

    def __init__(self, train, precision):
        self.train, self.precision = (None if train is None else bool(train)), precision

    def bn_train(self, bn):
        return bn.training if self.train is None else self.train


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _lib_call(name, dev, *args):
    with torch.cuda.device(dev):
        _lib.check(getattr(_lib.load(), name)(*args))


# ---------------------------------------------------------------------------------------------------------------------------------
# layout changes (our transpose kernels; differentiable)
# ---------------------------------------------------------------------------------------------------------------------------------
class _ToNCHW(Function):
    @staticmethod
    def forward(ctx, x):  # (n,h,w,c) contiguous -> (n,c,h,w) contiguous
        x = _c(x)
        n, h, w, c = x.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        _lib_call("usot_nhwc_to_nchw", x.device, _lib.ptr(x), n, h, w, c, _lib.ptr(out), _stream(x))
        return out

    @staticmethod
    def backward(ctx, g):
        return _ToNHWC.apply(g)


class _ToNHWC(Function):
    @staticmethod
    def forward(ctx, x):  # (n,c,h,w) contiguous -> (n,h,w,c) contiguous
        x = _c(x)
        n, c, h, w = x.shape
        out = torch.empty((n, h, w, c), dtype=torch.float32, device=x.device)
        _lib_call("usot_nchw_to_nhwc", x.device, _lib.ptr(x), n, c, h, w, _lib.ptr(out), _stream(x))
        return out

    @staticmethod
    def backward(ctx, g):
        return _ToNCHW.apply(g)


def to_nchw(x):
    return _ToNCHW.apply(x)


def to_nhwc(x):
    return _ToNHWC.apply(x)


# ---------------------------------------------------------------------------------------------------------------------------------
# convolution
# ---------------------------------------------------------------------------------------------------------------------------------
def _w_kn(weight_oihw):
    cout, cin, kh, kw = weight_oihw.shape
    return weight_oihw.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout).contiguous()


def _conv_raw(x, weight_oihw, stride, pad, dil, precision):
    """x NHWC, raw convolution (no bias / affine / ReLU) on the forward kernels."""
    cout = weight_oihw.shape[0]
    one = torch.ones(cout, dtype=torch.float32, device=x.device)
    return ops.conv2d_nhwc(x, weight_oihw, one, torch.zeros_like(one), stride=stride, padding=pad, dilation=dil, precision=precision)


def _wide(cin, cout, precision):
    """Can the forward GEMM kernels run a conv with these channel counts?"""
    return (cin % 64 == 0 and cout % 64 == 0) if precision != "fp32" else (cin % 16 == 0 and cout % 64 == 0)


def _prescale(grad_out, n_scale_channels, materialise=True):
    """Power of two s chosen on the device from max|grad_out| (no host sync).  Returns (s * grad_out or None, per-channel vector 1/s, s (1,))."""
    g = torch.empty_like(grad_out) if materialise else None
    sc2 = torch.empty(2, dtype=torch.float32, device=grad_out.device)
    _lib_call("usot_pow2_scale", grad_out.device, _lib.ptr(grad_out), grad_out.numel(), GRAD_TARGET_LOG2, _lib.ptr(g), _lib.ptr(sc2), _stream(grad_out))
    return g, sc2[1:2].expand(n_scale_channels).contiguous(), sc2[0:1]


def _dgrad_gemm(grad_out, weight_oihw, pad, dil, precision):
    """Stride-1 input gradient on the forward conv kernel (transposed, flipped filter).  Tensor-core modes: the gradient is multiplied by a
    device-chosen power of two inside the fp32 -> split-fp16 conversion and the factor is divided out by the conv's per-channel scale."""
    if precision == "fp32":
        return ops.conv2d_nhwc_input_grad(grad_out, weight_oihw, pad, dil, precision=precision)
    w_t, pad_t = ops.dgrad_weights(weight_oihw, pad, dil)
    _, inv, s = _prescale(grad_out, w_t.shape[0], materialise=False)
    return ops.conv2d_nhwc(grad_out, w_t, inv, torch.zeros_like(inv), stride=1, padding=pad_t, dilation=dil, precision=precision, in_scale=s)


def conv_dgrad(grad_out, weight_oihw, in_hw, stride, pad, dil, precision="fp32"):
    """Gradient w.r.t. the input of a conv.  grad_out NHWC (n,ho,wo,cout) -> NHWC (n,h,w,cin)."""
    cout, cin, kh, kw = weight_oihw.shape
    n, ho, wo, _ = grad_out.shape
    h, w = in_hw
    (ph, pw), (dh, dw) = _pair(pad), _pair(dil)
    grad_out = _c(grad_out)
    if not _wide(cout, cin, precision):   # thin conv: generic gather kernel
        gi = torch.empty((n, h, w, cin), dtype=torch.float32, device=grad_out.device)
        _lib_call("usot_conv2d_dgrad_nhwc", grad_out.device, _lib.ptr(grad_out), _lib.ptr(_w_kn(weight_oihw)), n, h, w, cin, cout, kh, kw, stride,
                  ph, pw, dh, dw, _lib.ptr(gi), _stream(grad_out))
        return gi
    if stride == 1:
        gi = _dgrad_gemm(grad_out, weight_oihw, (ph, pw), (dh, dw), precision)
        assert tuple(gi.shape[1:3]) == (h, w), (gi.shape, h, w)
        return gi
    # stride 2 (layer2.0.conv2, layer2.0.downsample; dilation 1): input rows of parity class r only see the taps kh = r (mod 2), and for
    # those the gradient is a stride-1 "full" correlation of grad_out with the sub-filter W[r::2] -> the forward kernel again.
    assert stride == 2 and (dh, dw) == (1, 1), "dgrad supports stride 1 (any dilation) and stride 2 (dilation 1)"
    gi = torch.zeros((n, h, w, cin), dtype=torch.float32, device=grad_out.device)
    for rh in range(2):
        for rw in range(2):
            sub = weight_oihw[:, :, rh::2, rw::2]
            if sub.shape[2] == 0 or sub.shape[3] == 0:
                continue
            part = _dgrad_gemm(grad_out, sub.contiguous(), (0, 0), (1, 1), precision)   # (n, ho+Jh-1, wo+Jw-1, cin)
            # part[u, v] is the gradient of input pixel (y, x) = (2u + rh - ph, 2v + rw - pw)
            ys = [(2 * u + rh - ph, u) for u in range(part.shape[1]) if 0 <= 2 * u + rh - ph < h]
            xs = [(2 * v + rw - pw, v) for v in range(part.shape[2]) if 0 <= 2 * v + rw - pw < w]
            if not ys or not xs:
                continue
            gi[:, ys[0][0]:ys[-1][0] + 1:2, xs[0][0]:xs[-1][0] + 1:2, :] = part[:, ys[0][1]:ys[-1][1] + 1, xs[0][1]:xs[-1][1] + 1, :]
    return gi


def conv_wgrad(x, grad_out, weight_shape, stride, pad, dil, precision="fp32"):
    """Gradient w.r.t. the OIHW weight.  x NHWC (n,h,w,cin), grad_out NHWC (n,ho,wo,cout)."""
    cout, cin, kh, kw = weight_shape
    n, h, w, _ = x.shape
    (ph, pw), (dh, dw) = _pair(pad), _pair(dil)
    gw = torch.empty((kh * kw * cin, cout), dtype=torch.float32, device=x.device)
    _lib_call("usot_conv2d_wgrad_nhwc", x.device, _lib.ptr(_c(x)), _lib.ptr(_c(grad_out)), n, h, w, cin, cout, kh, kw, stride, ph, pw, dh, dw,
              _lib.ptr(gw), _lib.PRECISIONS[precision], _stream(x))
    return gw.view(kh, kw, cin, cout).permute(3, 2, 0, 1).contiguous()


class _Conv(Function):
    """Raw nn.Conv2d (no bias) on NHWC maps, Cin % 64 == 0 and Cout % 64 == 0."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad, dil, precision):
        x = _c(x)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, _pair(pad), _pair(dil), BWD_PRECISION or precision)
        return _conv_raw(x, weight, stride, _pair(pad), _pair(dil), precision)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        stride, pad, dil, prec = ctx.cfg
        g = _c(g)
        gx = conv_dgrad(g, weight, x.shape[1:3], stride, pad, dil, prec) if ctx.needs_input_grad[0] else None
        gw = conv_wgrad(x, g, weight.shape, stride, pad, dil, prec) if ctx.needs_input_grad[1] else None
        return gx, gw, None, None, None, None


class _StemConv(Function):
    """conv1: 7x7 / stride 2 / pad 0, 3 -> 64, NCHW image in, NHWC map out (fp32 FMA)."""

    @staticmethod
    def forward(ctx, x_nchw, weight):
        x_nchw = _c(x_nchw.float())
        n, c, s, _ = x_nchw.shape
        ho = (s - 7) // 2 + 1
        wk = weight.reshape(64, 147).t().contiguous()
        out = torch.empty((n, ho, ho, 64), dtype=torch.float32, device=x_nchw.device)
        _lib_call("usot_stem_conv_raw", x_nchw.device, _lib.ptr(x_nchw), n, s, _lib.ptr(wk), _lib.ptr(out), _stream(x_nchw))
        ctx.save_for_backward(x_nchw)
        return out

    @staticmethod
    def backward(ctx, g):
        (x_nchw,) = ctx.saved_tensors   # (the input image needs no gradient)
        n, _, s, _ = x_nchw.shape
        gw = torch.empty((147, 64), dtype=torch.float32, device=x_nchw.device)
        _lib_call("usot_stem_conv_wgrad", x_nchw.device, _lib.ptr(x_nchw), _lib.ptr(_c(g)), n, s, _lib.ptr(gw), _stream(x_nchw))
        return None, gw.t().reshape(64, 3, 7, 7).contiguous()


class _PredConv(Function):
    """bbox_pred / cls_pred / cls_memory_pred: 3x3 p1, 256 -> 1 | 4, with bias; NHWC in, raw NCHW (n,cout,r,r) out."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = _c(x)
        ctx.save_for_backward(x, weight)
        return ops.pred_conv(x, weight, bias, mode=0, mul=1.0)

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        g_nhwc = to_nhwc(_c(g))
        gx = conv_dgrad(g_nhwc, weight, x.shape[1:3], 1, (1, 1), (1, 1)) if ctx.needs_input_grad[0] else None
        gw = conv_wgrad(x, g_nhwc, weight.shape, 1, (1, 1), (1, 1)) if ctx.needs_input_grad[1] else None
        gb = channel_sum(g_nhwc) if ctx.needs_input_grad[2] else None
        return gx, gw, gb


def channel_sum(x_nhwc):
    c = x_nhwc.shape[-1]
    m = x_nhwc.numel() // c
    out = torch.empty(c, dtype=torch.float32, device=x_nhwc.device)
    _lib_call("usot_channel_sum", x_nhwc.device, _lib.ptr(_c(x_nhwc)), m, c, _lib.ptr(out), _stream(x_nhwc))
    return out


# ---------------------------------------------------------------------------------------------------------------------------------
# BatchNorm (+ conv bias) (+ residual) (+ ReLU)
# ---------------------------------------------------------------------------------------------------------------------------------
class _BNAct(Function):
    @staticmethod
    def forward(ctx, x, conv_bias, gamma, beta, mean, var, residual, relu, train):
        """mean / var: batch statistics (train) or running statistics (eval) of x + conv_bias; both are constants of the graph here --
        in train mode their dependence on x is folded into the backward kernel."""
        x = _c(x)
        c = x.shape[-1]
        m = x.numel() // c
        y = torch.empty_like(x)
        invstd = torch.empty(c, dtype=torch.float32, device=x.device)
        residual = None if residual is None else _c(residual)
        _lib_call("usot_bn_apply", x.device, _lib.ptr(x), _lib.ptr(conv_bias), _lib.ptr(mean), _lib.ptr(var), BN_EPS, _lib.ptr(gamma), _lib.ptr(beta),
                  _lib.ptr(residual), int(relu), m, c, _lib.ptr(y), _lib.ptr(invstd), _stream(x))
        ctx.save_for_backward(x, y if relu else None, conv_bias, gamma, mean, invstd)
        ctx.cfg = (bool(relu), bool(train), residual is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        x, y, conv_bias, gamma, mean, invstd = ctx.saved_tensors
        relu, train, has_res = ctx.cfg
        g = _c(g)
        c = x.shape[-1]
        m = x.numel() // c
        gx = torch.empty_like(x)
        ggamma = torch.empty(c, dtype=torch.float32, device=x.device)
        gbeta = torch.empty_like(ggamma)
        gres = torch.empty_like(x) if has_res else None
        _lib_call("usot_bn_backward", x.device, _lib.ptr(g), _lib.ptr(y), _lib.ptr(x), _lib.ptr(conv_bias), _lib.ptr(mean), _lib.ptr(invstd),
                  _lib.ptr(gamma), int(train), int(relu), m, c, _lib.ptr(gx), _lib.ptr(ggamma), _lib.ptr(gbeta), _lib.ptr(gres), _stream(x))
        gbias = channel_sum(gx) if (conv_bias is not None and ctx.needs_input_grad[1]) else None
        return gx, gbias, ggamma, gbeta, None, None, gres, None, None


def batchnorm(x, bn, conv_bias=None, residual=None, relu=False, train=False):
    """nn.BatchNorm2d ``bn`` applied to x + conv_bias (the bias of the preceding conv), + residual, + ReLU.  train=True: batch
    statistics, and the module's running statistics are updated like torch does (momentum 0.1 or cumulative average, unbiased variance)."""
    c = x.shape[-1]
    m = x.numel() // c
    if train:
        mean = torch.empty(c, dtype=torch.float32, device=x.device)
        var = torch.empty_like(mean)
        xd = _c(x.detach())
        _lib_call("usot_bn_stats", x.device, _lib.ptr(xd), _lib.ptr(None if conv_bias is None else conv_bias.detach()), m, c, _lib.ptr(mean),
                  _lib.ptr(var), _stream(x))
        if bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():
                bn.num_batches_tracked += 1
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)   # (None = cumulative average)
                bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
                bn.running_var.mul_(1 - mom).add_(var, alpha=mom * m / max(m - 1, 1))
    else:
        mean, var = bn.running_mean, bn.running_var
    return _BNAct.apply(x, conv_bias, bn.weight, bn.bias, mean, var, residual, relu, train)


# ---------------------------------------------------------------------------------------------------------------------------------
# pooling, correlation, fusion, losses
# ---------------------------------------------------------------------------------------------------------------------------------
class _MaxPool(Function):
    @staticmethod
    def forward(ctx, x):
        x = _c(x)
        ctx.save_for_backward(x)
        return ops.maxpool3x3s2p1_nhwc(x)

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        n, h, w, c = x.shape
        gi = torch.empty_like(x)
        _lib_call("usot_maxpool3x3s2p1_backward_nhwc", x.device, _lib.ptr(x), _lib.ptr(_c(g)), n, h, w, c, _lib.ptr(gi), _stream(x))
        return gi


class _WeightedSum3(Function):
    @staticmethod
    def forward(ctx, x0, x1, x2, w3):
        x0, x1, x2, w3 = _c(x0), _c(x1), _c(x2), _c(w3)
        out = torch.empty_like(x0)
        _lib_call("usot_weighted_sum3", x0.device, _lib.ptr(x0), _lib.ptr(x1), _lib.ptr(x2), _lib.ptr(w3), x0.numel(), _lib.ptr(out), _stream(x0))
        ctx.save_for_backward(x0, x1, x2, w3)
        return out

    @staticmethod
    def backward(ctx, g):
        x0, x1, x2, w3 = ctx.saved_tensors
        g = _c(g)
        g0, g1, g2 = torch.empty_like(x0), torch.empty_like(x1), torch.empty_like(x2)
        gw = torch.empty(3, dtype=torch.float32, device=x0.device)
        _lib_call("usot_weighted_sum3_backward", x0.device, _lib.ptr(x0), _lib.ptr(x1), _lib.ptr(x2), _lib.ptr(w3), _lib.ptr(g), x0.numel(),
                  _lib.ptr(g0), _lib.ptr(g1), _lib.ptr(g2), _lib.ptr(gw), _stream(x0))
        return g0, g1, g2, gw


class _ConfFusion(Function):
    @staticmethod
    def forward(ctx, conf, value, nq):
        conf, value = _c(conf), _c(value)
        ctx.save_for_backward(conf, value)
        ctx.nq = nq
        return ops.conf_fusion(conf, value, nq)

    @staticmethod
    def backward(ctx, g):
        conf, value = ctx.saved_tensors
        gc, gv = torch.empty_like(conf), torch.empty_like(value)
        b = conf.shape[0] // ctx.nq
        _lib_call("usot_conf_fusion_backward", conf.device, _lib.ptr(conf), _lib.ptr(value), _lib.ptr(_c(g)), b, ctx.nq, conf[0].numel(), _lib.ptr(gc),
                  _lib.ptr(gv), _stream(conf))
        return gc, gv, None


class _WeightedBCE(Function):
    @staticmethod
    def forward(ctx, pred, label):
        pred, label = _c(pred), _c(label)
        ctx.save_for_backward(pred, label)
        return ops.weighted_bce(pred, label)

    @staticmethod
    def backward(ctx, g):
        pred, label = ctx.saved_tensors
        gp = torch.empty_like(pred)
        _lib_call("usot_weighted_bce_backward", pred.device, _lib.ptr(pred), _lib.ptr(label), pred.numel(), _lib.ptr(_c(g.reshape(1))), _lib.ptr(gp),
                  _stream(pred))
        return gp, None


class _IoULoss(Function):
    @staticmethod
    def forward(ctx, bbox, target, weight):
        bbox, target, weight = _c(bbox), _c(target), _c(weight)
        ctx.save_for_backward(bbox, target, weight)
        return ops.iou_loss(bbox, target, weight)

    @staticmethod
    def backward(ctx, g):
        bbox, target, weight = ctx.saved_tensors
        gb = torch.empty_like(bbox)
        n, _, r, _ = bbox.shape
        _lib_call("usot_iou_loss_backward", bbox.device, _lib.ptr(bbox), _lib.ptr(target), _lib.ptr(weight), n, r * r, _lib.ptr(_c(g.reshape(1))),
                  _lib.ptr(gb), _stream(bbox))
        return gb, None, None


# ---------------------------------------------------------------------------------------------------------------------------------
# the network (mirrors lib/models/modules.py, connect.py, models.py; parameters come from the usot_b200.USOT module)
# ---------------------------------------------------------------------------------------------------------------------------------
_LAYERS = (("layer1", 3, 1, 1), ("layer2", 4, 2, 1), ("layer3", 6, 1, 2))   # name, blocks, stride, dilation (modules.py:76-95)


def _bottleneck(blk, x, stride, dilation, has_down, down_k, down_stride, down_pad, md):
    """Bottleneck.forward (lib/models/modules.py:37-58)."""
    pad2 = 2 - stride
    dil2 = dilation
    if has_down and dil2 > 1:
        dil2 = dil2 // 2
        pad2 = dil2
    if dil2 > 1:
        pad2 = dil2
    P = md.precision
    out = batchnorm(_Conv.apply(x, blk.conv1.weight, 1, 0, 1, P), blk.bn1, relu=True, train=md.bn_train(blk.bn1))
    out = batchnorm(_Conv.apply(out, blk.conv2.weight, stride, pad2, dil2, P), blk.bn2, relu=True, train=md.bn_train(blk.bn2))
    out = _Conv.apply(out, blk.conv3.weight, 1, 0, 1, P)
    residual = x
    if has_down:
        residual = batchnorm(_Conv.apply(x, blk.downsample[0].weight, down_stride, down_pad, 1, P), blk.downsample[1], train=md.bn_train(blk.downsample[1]))
    return batchnorm(out, blk.bn3, residual=residual, relu=True, train=md.bn_train(blk.bn3))


def backbone_neck(net, x_nchw, md):
    """feature_extractor + neck (lib/models/modules.py:137-151, connect.py:294-296): (n,3,S,S) image -> NHWC (n,F,F,256)."""
    f = net.features.features
    x = batchnorm(_StemConv.apply(x_nchw, f.conv1.weight), f.bn1, relu=True, train=md.bn_train(f.bn1))
    x = _MaxPool.apply(x)
    for lname, blocks, stride, dilation in _LAYERS:
        layer = getattr(f, lname)
        for i in range(blocks):
            if i == 0:
                if stride == 1 and dilation == 1:
                    dk, dpad = 1, 0
                else:
                    dk, dpad = 3, (dilation // 2 if dilation > 1 else 0)
                x = _bottleneck(layer[i], x, stride, dilation, True, dk, stride, dpad, md)
            else:
                x = _bottleneck(layer[i], x, 1, dilation, False, 0, 1, 0, md)
    return batchnorm(_Conv.apply(x, net.neck.downsample[0].weight, 1, 0, 1, md.precision), net.neck.downsample[1], train=md.bn_train(net.neck.downsample[1]))


def prpool_feature(feat_nhwc, boxes):
    """USOT_.prpool_feature (lib/models/models.py:164-171): PrRoIPool 7x7 of each map with its own box.  NHWC in / out."""
    n = feat_nhwc.shape[0]
    idx = torch.arange(0, n, device=feat_nhwc.device, dtype=torch.float32).view(-1, 1)
    rois = torch.cat((idx, boxes.to(feat_nhwc.device, torch.float32)), dim=1)
    return to_nhwc(ops.prroi_pool2d(to_nchw(feat_nhwc), rois, 7, 7, 1.0))


def _cbr(seq, x, dil, pad, md):
    conv, bn = seq[0], seq[1]
    return batchnorm(_Conv.apply(x, conv.weight, 1, pad, dil, md.precision), bn, conv_bias=conv.bias, relu=True, train=md.bn_train(bn))


def _matrix_encode(enc, z, x, md):
    """matrix.forward (lib/models/connect.py:55-74)."""
    zs = xs = None
    if x is not None:
        xs = [_cbr(enc.matrix11_s, x, 1, 0, md), _cbr(enc.matrix12_s, x, (2, 1), 0, md), _cbr(enc.matrix21_s, x, (1, 2), 0, md)]
    if z is not None:
        zs = [_cbr(enc.matrix11_k, z, 1, 0, md), _cbr(enc.matrix12_k, z, (2, 1), 0, md), _cbr(enc.matrix21_k, z, (1, 2), 0, md)]
    return zs, xs


def _groupdw(dw, zs, xs):
    """GroupDW.forward (lib/models/connect.py:86-102).  zs / xs NHWC lists; kernel batch 1 or == search batch."""
    w = torch.softmax(dw.weight, 0)
    maps = [ops.xcorr_depthwise(to_nchw(x), to_nchw(z)) for x, z in zip(xs, zs)]
    return to_nhwc(_WeightedSum3.apply(maps[0], maps[1], maps[2], w))


def _tower(seq, x, md):
    for i in range(4):
        conv, bn = seq[3 * i], seq[3 * i + 1]
        x = batchnorm(_Conv.apply(x, conv.weight, 1, 1, 1, md.precision), bn, conv_bias=conv.bias, relu=True, train=md.bn_train(bn))
    return x


def connect(head, md, search, kernel=None, memory_kernel=None, memory_confidence=None, cls_x_store=None):
    """box_tower_reg.forward (lib/models/connect.py:221-281) on NHWC maps.  Returns (bbox, cls, cls_x, reg_x, cls_mem) with the score /
    box maps in the reference's NCHW."""
    x_bbox = cls = cls_x = reg_x = None
    if kernel is not None:
        cls_z, cls_x = _matrix_encode(head.cls_encode, kernel, search, md)
        reg_z, reg_x = _matrix_encode(head.reg_encode, kernel, search, md)
        cls_dw = _groupdw(head.cls_dw, cls_z, cls_x)
        reg_dw = _groupdw(head.reg_dw, reg_z, reg_x)
        x_reg = _tower(head.bbox_tower, reg_dw, md)
        x_bbox = torch.exp(head.adjust * _PredConv.apply(x_reg, head.bbox_pred.weight, head.bbox_pred.bias) + head.bias)
        cls = 0.1 * _PredConv.apply(_tower(head.cls_tower, cls_dw, md), head.cls_pred.weight, head.cls_pred.bias)
        if memory_kernel is None:
            return x_bbox, cls, cls_x, reg_x, None
    if cls_x_store is None:
        cls_mem_zs, cls_x_store = _matrix_encode(head.cls_encode, memory_kernel, search, md)
    else:
        cls_mem_zs, _ = _matrix_encode(head.cls_encode, memory_kernel, None, md)
    batch, mem = memory_confidence.shape   # only the shape is used (connect.py:258)
    rep = [cx.unsqueeze(1).expand(-1, mem, -1, -1, -1).reshape((-1,) + tuple(cx.shape[1:])) for cx in cls_x_store]
    dw = _groupdw(head.cls_dw, cls_mem_zs, rep)
    cf = head.conf_fusion
    conf = _cbr(cf.conf_gen, dw, 1, 1, md)
    value = _cbr(cf.value_gen, dw, 1, 1, md)
    fused = _ConfFusion.apply(conf, value, mem)
    cls_mem = 0.1 * _PredConv.apply(_tower(head.cls_memory_tower, fused, md), head.cls_memory_pred.weight, head.cls_memory_pred.bias)
    if kernel is not None:
        return x_bbox, cls, cls_x, reg_x, cls_mem
    return None, None, None, None, cls_mem


def forward_train(net, template, search, label, reg_target, reg_weight, template_bbox, search_memory=None, search_bbox=None, cls_ratio=0.40,
                  train=None):
    """USOT_.forward (lib/models/models.py:208-295) with an autograd graph.  ``train=None`` (default): each BatchNorm follows its own
    ``.training`` flag like torch; ``True`` / ``False`` force batch / running statistics everywhere.  Returns (cls_loss, cls_memory_loss or None, reg_loss), 0-d tensors that support ``.backward()``."""
    md = Mode(train, getattr(net, "precision", "fp16x3"))
    head = net.connect_model
    dev = search.device
    f32 = lambda t: None if t is None else t.to(dev, torch.float32)
    label, reg_target, reg_weight = f32(label), f32(reg_target), f32(reg_weight)
    z_ori = backbone_neck(net, template, md)
    xf = backbone_neck(net, search, md)
    if net.pr_pool:
        zf = prpool_feature(z_ori, template_bbox)
    else:
        zf = z_ori[:, 4:-4, 4:-4, :].contiguous()
    if search_memory is None:
        bbox_pred, cls_pred, _, _, _ = connect(head, md, xf, kernel=zf)
        return _WeightedBCE.apply(cls_pred, label), None, _IoULoss.apply(bbox_pred, reg_target, reg_weight)
    bbox_pred, cls_pred, cls_x, _, _ = connect(head, md, xf, kernel=zf)
    reg_loss = _IoULoss.apply(bbox_pred, reg_target, reg_weight)
    cls_loss_ori = _WeightedBCE.apply(cls_pred, label)
    batch, mem, cx, hx, wx = search_memory.shape
    xf_mem = backbone_neck(net, search_memory.reshape(-1, cx, hx, wx), md)
    spf = prpool_feature(xf, search_bbox)
    spf = spf.unsqueeze(1).expand(-1, mem, -1, -1, -1).reshape((-1,) + tuple(spf.shape[1:]))
    zf_mem = zf.unsqueeze(1).expand(-1, mem, -1, -1, -1).reshape((-1,) + tuple(zf.shape[1:]))
    off_bbox, off_cls, fwd_store, _, _ = connect(head, md, xf_mem, kernel=zf_mem)
    _, _, _, _, mem_cls = connect(head, md, xf_mem, memory_kernel=spf, memory_confidence=torch.ones(batch * mem, 1), cls_x_store=fwd_store)
    r = off_cls.shape[-1]
    size = 255 + (xf.shape[1] - 31) * 8
    with torch.no_grad():   # best_forward_bbox_pool / best_forward_cls_score are detached in the reference (models.py:273-274)
        pool_box, best_score, _ = ops.cycle_glue(off_cls.detach(), mem_cls.detach(), off_bbox.detach(), cls_ratio, size, r)
    pooled = prpool_feature(xf_mem, pool_box)
    _, _, _, _, back = connect(head, md, xf, memory_kernel=pooled, memory_confidence=best_score.view(batch, mem), cls_x_store=cls_x)
    return cls_loss_ori, _WeightedBCE.apply(back, label), reg_loss
