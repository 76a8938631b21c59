"""Operator-level Python wrappers over the C ABI (same names / argument meaning as the reference ops).

    prroi_pool2d      lib/models/prroi_pool/functional.py:41-84  (forward only)
    xcorr_depthwise   lib/models/connect.py:147-157
    groupdw_xcorr     lib/models/connect.py:86-102 (fused, NHWC)
    conv2d_nhwc       nn.Conv2d + folded BatchNorm2d (+residual)(+ReLU)
    pred_conv         bbox_pred / cls_pred / cls_memory_pred + epilogue, lib/models/connect.py:235-241,274-275
    stem_conv, maxpool3x3s2p1_nhwc   conv1+bn1+relu / maxpool, lib/models/modules.py:70-75,138-141
    stem_maxpool                     the same four modules as one fused tensor-core kernel
    conf_fusion       reduction of Conf_Fusion.forward, lib/models/connect.py:123-144
    cycle_glue        forward-tracking argmax / box maps, lib/models/models.py:262-274
    weighted_bce, iou_loss           lib/models/models.py:42-100

torch is used for device memory and the current stream only.
"""
import torch

from . import _lib


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            # same contract as the reference op (functional.py:62-63), extended to every op of this package
            raise NotImplementedError("usot_b200 only supports GPU (cuda) tensors; there is no CPU fallback")


def _need_float(*ts):
    for t in ts:
        if t is not None and t.dtype != torch.float32:
            raise AssertionError("usot_b200 ops only take float32 input, got {}".format(t.dtype))


def as_nhwc(t):
    """Logical-NCHW tensor -> tensor whose memory is contiguous NHWC (no copy if it already is)."""
    return t.permute(0, 2, 3, 1).contiguous()


def nhwc_view(t_nhwc):
    """Contiguous NHWC tensor -> logical NCHW view (channels-last strides)."""
    return t_nhwc.permute(0, 3, 1, 2)


class PrRoIPool2DFunction(torch.autograd.Function):
    """Forward + both backward passes of Precise RoI Pooling, same contract as lib/models/prroi_pool/functional.py:41-81."""

    @staticmethod
    def forward(ctx, features, rois, pooled_height, pooled_width, spatial_scale):
        _need_float(features, rois)
        _need_cuda(features, rois)
        pooled_height, pooled_width, spatial_scale = int(pooled_height), int(pooled_width), float(spatial_scale)
        features, rois = features.contiguous(), rois.contiguous()
        n, c, h, w = features.shape
        out = torch.empty((rois.shape[0], c, pooled_height, pooled_width), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            _lib.check(_lib.load().usot_prroi_pool_forward(_lib.ptr(features), _lib.ptr(rois), _lib.ptr(out), n, rois.shape[0], c, h, w,
                                                           pooled_height, pooled_width, spatial_scale, _stream(features)))
        ctx.params = (pooled_height, pooled_width, spatial_scale)
        ctx.save_for_backward(features, rois, out)
        return out

    @staticmethod
    def backward(ctx, grad_output):
        features, rois, out = ctx.saved_tensors
        ph, pw, scale = ctx.params
        n, c, h, w = features.shape
        grad_input = grad_coor = None
        grad_output = grad_output.contiguous().float()
        lib = _lib.load()
        with torch.cuda.device(features.device):
            if ctx.needs_input_grad[0]:
                grad_input = torch.empty_like(features)
                _lib.check(lib.usot_prroi_pool_backward(_lib.ptr(rois), _lib.ptr(grad_output), _lib.ptr(grad_input), n, rois.shape[0], c, h, w,
                                                        ph, pw, scale, _stream(features)))
            if ctx.needs_input_grad[1]:
                grad_coor = torch.empty_like(rois)
                _lib.check(lib.usot_prroi_pool_coor_backward(_lib.ptr(features), _lib.ptr(rois), _lib.ptr(out), _lib.ptr(grad_output),
                                                             _lib.ptr(grad_coor), rois.shape[0], c, h, w, ph, pw, scale, _stream(features)))
        return grad_input, grad_coor, None, None, None


def prroi_pool2d(features, rois, pooled_height, pooled_width, spatial_scale):
    """Drop-in for lib.models.prroi_pool.functional.prroi_pool2d (differentiable w.r.t. features and roi coordinates)."""
    return PrRoIPool2DFunction.apply(features, rois, pooled_height, pooled_width, spatial_scale)


class XCorrDepthwiseFunction(torch.autograd.Function):
    """Depth-wise cross-correlation with both gradients (lib/models/connect.py:147-157 + the autograd of its F.conv2d)."""

    @staticmethod
    def forward(ctx, x, kernel):
        _need_float(x, kernel)
        _need_cuda(x, kernel)
        x, kernel = x.contiguous(), kernel.contiguous()
        bx, c, hx, wx = x.shape
        bk, ck, hk, wk = kernel.shape
        assert c == ck, "channel mismatch"
        out = torch.empty((bx, c, hx - hk + 1, wx - wk + 1), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().usot_xcorr_depthwise(_lib.ptr(x), _lib.ptr(kernel), _lib.ptr(out), bx, bk, c, hx, wx, hk, wk, _stream(x)))
        ctx.save_for_backward(x, kernel)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, kernel = ctx.saved_tensors
        bx, c, hx, wx = x.shape
        bk, _, hk, wk = kernel.shape
        grad_out = grad_out.contiguous().float()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gk = torch.empty_like(kernel) if ctx.needs_input_grad[1] else None
        if gx is not None or gk is not None:
            with torch.cuda.device(x.device):
                _lib.check(_lib.load().usot_xcorr_depthwise_backward(_lib.ptr(x), _lib.ptr(kernel), _lib.ptr(grad_out), _lib.ptr(gx), _lib.ptr(gk),
                                                                     bx, bk, c, hx, wx, hk, wk, _stream(x)))
        return gx, gk


def xcorr_depthwise(x, kernel):
    """Drop-in for lib.models.connect.xcorr_depthwise (NCHW; kernel batch 1 or equal to the search batch); differentiable."""
    return XCorrDepthwiseFunction.apply(x, kernel)


def groupdw_xcorr(x, z, weight, n_out=None):
    """x = [x11, x12, x21], z = [z11, z12, z21] contiguous NHWC tensors; weight (3,) raw.  Returns NHWC (n_out,R,R,C)."""
    _need_float(*x, *z, weight)
    _need_cuda(*x, *z, weight)
    x = [t.contiguous() for t in x]
    z = [t.contiguous() for t in z]
    nx, h11, w11, c = x[0].shape
    f = h11 + 2
    assert tuple(x[1].shape) == (nx, f - 4, f - 2, c) and tuple(x[2].shape) == (nx, f - 2, f - 4, c), "bad search map shapes"
    nz = z[0].shape[0]
    assert tuple(z[0].shape[1:]) == (5, 5, c) and tuple(z[1].shape[1:]) == (3, 5, c) and tuple(z[2].shape[1:]) == (5, 3, c)
    n_out = n_out or max(nx, nz)
    out = torch.empty((n_out, f - 6, f - 6, c), dtype=torch.float32, device=x[0].device)
    with torch.cuda.device(out.device):
        _lib.check(_lib.load().usot_groupdw_xcorr(*[_lib.ptr(t) for t in x], *[_lib.ptr(t) for t in z], _lib.ptr(weight.contiguous()),
                                                  _lib.ptr(out), nx, nz, n_out, c, f, _stream(out)))
    return out


def conv2d_nhwc(x, weight_oihw, scale, shift, stride=1, padding=(0, 0), dilation=(1, 1), residual=None, relu=False, precision="fp32", in_scale=None):
    """x NHWC (n,h,w,cin); weight in the reference's OIHW layout (repacked here); returns NHWC.  ``in_scale`` (1-element CUDA tensor,
    tensor-core precisions only): x is multiplied by it inside the fp32 -> split-fp16 conversion (usot_conv2d_nhwc_scaled)."""
    _need_float(x, weight_oihw, scale, shift, residual)
    _need_cuda(x, weight_oihw, scale, shift, residual)
    n, h, w, cin = x.shape
    cout, cin2, kh, kw = weight_oihw.shape
    assert cin == cin2
    ph, pw = (padding, padding) if isinstance(padding, int) else padding
    dh, dw = (dilation, dilation) if isinstance(dilation, int) else dilation
    ho = (h + 2 * ph - dh * (kh - 1) - 1) // stride + 1
    wo = (w + 2 * pw - dw * (kw - 1) - 1) // stride + 1
    w_kn = weight_oihw.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout).contiguous()
    out = torch.empty((n, ho, wo, cout), dtype=torch.float32, device=x.device)
    if in_scale is not None:
        assert residual is None and not relu
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().usot_conv2d_nhwc_scaled(_lib.ptr(x.contiguous()), _lib.ptr(in_scale), n, h, w, cin, _lib.ptr(w_kn), cout, kh, kw, stride,
                                                           ph, pw, dh, dw, _lib.ptr(scale.contiguous()), _lib.ptr(shift.contiguous()), _lib.ptr(out),
                                                           _lib.PRECISIONS[precision], _stream(x)))
        return out
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().usot_conv2d_nhwc(_lib.ptr(x.contiguous()), n, h, w, cin, _lib.ptr(w_kn), cout, kh, kw, stride, ph, pw, dh, dw,
                                                _lib.ptr(scale.contiguous()), _lib.ptr(shift.contiguous()),
                                                _lib.ptr(None if residual is None else residual.contiguous()), int(bool(relu)),
                                                _lib.ptr(out), _lib.PRECISIONS[precision], _stream(x)))
    return out


def pred_conv(x, weight_oihw, bias, mode=0, mul=0.1, adjust=None, bias4=None):
    """x NHWC (n,r,r,256); weight (cout,256,3,3) OIHW as in the reference, cout in {1,4}; returns NCHW (n,cout,r,r).
    mode 0: mul * (conv(x) + bias); mode 1: exp(adjust * (conv(x) + bias) + bias4)."""
    _need_float(x, weight_oihw, bias, adjust, bias4)
    _need_cuda(x, weight_oihw, bias, adjust, bias4)
    n, r, r2, c = x.shape
    cout = weight_oihw.shape[0]
    assert r == r2 and tuple(weight_oihw.shape) == (cout, c, 3, 3)
    w = weight_oihw.permute(2, 3, 0, 1).reshape(9, cout, c).contiguous()
    out = torch.empty((n, cout, r, r), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().usot_pred_conv(_lib.ptr(x.contiguous()), n, r, c, _lib.ptr(w), _lib.ptr(bias.contiguous()), cout, int(mode),
                                              float(mul), _lib.ptr(None if adjust is None else adjust.contiguous()),
                                              _lib.ptr(None if bias4 is None else bias4.reshape(-1).contiguous()), _lib.ptr(out), _stream(x)))
    return out


def dgrad_weights(weight_oihw, padding=(0, 0), dilation=(1, 1)):
    """Host side of conv dgrad for a STRIDE-1 convolution: the gradient w.r.t. the input is itself a stride-1 convolution of
    grad_out with the spatially flipped, channel-transposed filter and padding d*(k-1) - p.  Returns (weight' (Cin,Cout,kh,kw),
    padding').  Pure tensor bookkeeping (checked against torch autograd on the CPU, tests/test_host_logic.py)."""
    ph, pw = (padding, padding) if isinstance(padding, int) else padding
    dh, dw = (dilation, dilation) if isinstance(dilation, int) else dilation
    kh, kw = weight_oihw.shape[2], weight_oihw.shape[3]
    pad = (dh * (kh - 1) - ph, dw * (kw - 1) - pw)
    if pad[0] < 0 or pad[1] < 0:
        raise ValueError("dgrad as a forward conv needs padding <= dilation * (k - 1)")
    return weight_oihw.flip(2, 3).permute(1, 0, 2, 3).contiguous(), pad


def conv2d_nhwc_input_grad(grad_out, weight_oihw, padding=(0, 0), dilation=(1, 1), precision="fp32"):
    """dgrad of every stride-1 conv of the network (40 of the 43 backbone convs, the neck, encoders, towers, conf/value generators)
    on the SAME kernel as the forward pass: grad_out NHWC (n,ho,wo,Cout) -> grad_in NHWC (n,h,w,Cin).  Training path groundwork
    (SURVEY.md §8f-3); wgrad and the two stride-2 layers need their own kernels."""
    w_t, pad = dgrad_weights(weight_oihw, padding, dilation)
    cin = w_t.shape[0]
    one = torch.ones(cin, dtype=torch.float32, device=grad_out.device)
    return conv2d_nhwc(grad_out, w_t, one, torch.zeros_like(one), stride=1, padding=pad, dilation=dilation, precision=precision)


# ---- small stand-alone operators (each one reference op, exposed for op-level parity tests) ---------------------------------
def maxpool3x3s2p1_nhwc(x, split=False):
    """MaxPool2d(3, stride 2, padding 1) of the stem (lib/models/modules.py:75,141).  x NHWC (n,h,w,C) -> NHWC (n,ho,wo,C).
    ``split=True`` runs the engine's variant that writes split-fp16 planes and returns hi + lo."""
    _need_float(x)
    _need_cuda(x)
    n, h, w, c = x.shape
    out = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().usot_maxpool3x3s2p1_nhwc(_lib.ptr(x.contiguous()), n, h, w, c, None if split else _lib.ptr(out),
                                                        _lib.ptr(out) if split else None, _stream(x)))
    return out


def stem_conv(x, weight_oihw, scale, shift, precision="fp32"):
    """conv1 7x7/2 p0 + folded BN + ReLU (lib/models/modules.py:70-74,138-140).  x NCHW (n,3,S,S) on the GPU; weight (64,3,7,7),
    scale / shift (64) on the host or device (copied to the host: the entry packs them itself).  Returns NHWC (n,HO,HO,64)."""
    _need_float(x)
    _need_cuda(x)
    n, c, s, s2 = x.shape
    assert c == 3 and s == s2 and tuple(weight_oihw.shape) == (64, 3, 7, 7)
    ho = (s - 7) // 2 + 1
    host = [t.detach().to("cpu", torch.float32).contiguous() for t in (weight_oihw, scale, shift)]
    out = torch.empty((n, ho, ho, 64), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().usot_stem_conv(_lib.ptr(x.contiguous()), n, s, _lib.ptr(host[0]), _lib.ptr(host[1]), _lib.ptr(host[2]),
                                              _lib.ptr(out), _lib.PRECISIONS[precision], _stream(x)))
    return out


def stem_maxpool(x, weight_oihw, scale, shift, precision="fp16x3"):
    """conv1 7x7/2 p0 + folded BN + ReLU + maxpool 3x3/2 p1 (lib/models/modules.py:70-75,138-141) as ONE tensor-core kernel over the
    space-to-depth image (the engine's path in the tcgen05 modes; crops up to 261 pixels).  Returns NHWC (n,PO,PO,64) fp32 = hi + lo of
    the split-fp16 planes the kernel writes."""
    _need_float(x)
    _need_cuda(x)
    n, c, s, s2 = x.shape
    assert c == 3 and s == s2 and tuple(weight_oihw.shape) == (64, 3, 7, 7)
    ho = (s - 7) // 2 + 1
    po = (ho - 1) // 2 + 1
    host = [t.detach().to("cpu", torch.float32).contiguous() for t in (weight_oihw, scale, shift)]
    out = torch.empty((n, po, po, 64), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().usot_stem_maxpool(_lib.ptr(x.contiguous()), n, s, _lib.ptr(host[0]), _lib.ptr(host[1]), _lib.ptr(host[2]),
                                                 _lib.ptr(out), _lib.PRECISIONS[precision], _stream(x)))
    return out


def conf_fusion(conf, value, nq):
    """Reduction of Conf_Fusion.forward (lib/models/connect.py:123-144): conf / value (B*nq, ...) -> (B, ...)."""
    _need_float(conf, value)
    _need_cuda(conf, value)
    assert conf.shape == value.shape and conf.shape[0] % nq == 0
    b = conf.shape[0] // nq
    per_map = conf[0].numel()
    out = torch.empty((b,) + tuple(conf.shape[1:]), dtype=torch.float32, device=conf.device)
    with torch.cuda.device(conf.device):
        _lib.check(_lib.load().usot_conf_fusion(_lib.ptr(conf.contiguous()), _lib.ptr(value.contiguous()), b, nq, per_map, _lib.ptr(out),
                                                _stream(conf)))
    return out


def cycle_glue(off_cls, mem_cls, off_bbox, cls_ratio, search_size=255, search_feature_size=25):
    """lib/models/models.py:262-274: returns (pool_box (n,4), best_score (n), best_idx (n) int32)."""
    _need_float(off_cls, mem_cls, off_bbox)
    _need_cuda(off_cls, mem_cls, off_bbox)
    n, r = off_bbox.shape[0], off_bbox.shape[-1]
    assert tuple(off_bbox.shape) == (n, 4, r, r) and off_cls.numel() == n * r * r == mem_cls.numel()
    dev = off_cls.device
    box = torch.empty((n, 4), dtype=torch.float32, device=dev)
    score = torch.empty((n,), dtype=torch.float32, device=dev)
    idx = torch.empty((n,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().usot_cycle_glue(_lib.ptr(off_cls.contiguous()), _lib.ptr(mem_cls.contiguous()), _lib.ptr(off_bbox.contiguous()), n, r,
                                               int(search_size), int(search_feature_size), float(cls_ratio), _lib.ptr(box), _lib.ptr(score),
                                               _lib.ptr(idx), _stream(off_cls)))
    return box, score, idx


def weighted_bce(pred, label):
    """_weighted_BCE (lib/models/models.py:49-58) -> 0-d tensor."""
    _need_float(pred, label)
    _need_cuda(pred, label)
    out = torch.empty((1,), dtype=torch.float32, device=pred.device)
    with torch.cuda.device(pred.device):
        _lib.check(_lib.load().usot_weighted_bce(_lib.ptr(pred.contiguous()), _lib.ptr(label.contiguous()), pred.numel(), _lib.ptr(out), _stream(pred)))
    return out[0]


def iou_loss(bbox_pred, reg_target, reg_weight):
    """add_iouloss (lib/models/models.py:85-100): bbox_pred (n,4,R,R), reg_target (n,R,R,4), reg_weight (n,R,R) -> 0-d tensor."""
    _need_float(bbox_pred, reg_target, reg_weight)
    _need_cuda(bbox_pred, reg_target, reg_weight)
    n, _, r, _ = bbox_pred.shape
    out = torch.empty((1,), dtype=torch.float32, device=bbox_pred.device)
    with torch.cuda.device(bbox_pred.device):
        _lib.check(_lib.load().usot_iou_loss(_lib.ptr(bbox_pred.contiguous()), _lib.ptr(reg_target.contiguous()), _lib.ptr(reg_weight.contiguous()),
                                             n, r * r, _lib.ptr(out), _stream(bbox_pred)))
    return out[0]
