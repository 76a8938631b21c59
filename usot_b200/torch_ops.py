"""``torch.ops.usot_b200.*``: dispatcher registration of the stand-alone operators (SURVEY.md §8b sketched a TORCH_LIBRARY boundary).

The reference binds its one native op through ``PYBIND11_MODULE`` (lib/models/prroi_pool/src/prroi_pooling_gpu.c:109-112) and calls it
as ``_prroi_pooling.prroi_pooling_forward_cuda(features, rois, ph, pw, scale)`` (functional.py:57).  This module registers the same
three entry points -- and the other operators of this library -- with the PyTorch dispatcher under the ``usot_b200`` namespace, CUDA
key only (calling them with CPU tensors raises NotImplementedError from the dispatcher: there is no CPU fallback).  The registered
kernels are thin: they forward to the C ABI (include/usot_b200.h) through usot_b200.ops.

    import usot_b200.torch_ops
    out = torch.ops.usot_b200.prroi_pooling_forward(features, rois, 7, 7, 1.0)
"""
import torch

from . import _lib, ops
from .ops import _stream

_LIB = torch.library.Library("usot_b200", "DEF")
_LIB.define("prroi_pooling_forward(Tensor features, Tensor rois, int pooled_height, int pooled_width, float spatial_scale) -> Tensor")
_LIB.define("prroi_pooling_backward(Tensor features, Tensor rois, Tensor output, Tensor output_diff, int pooled_height, int pooled_width, "
            "float spatial_scale) -> Tensor")
_LIB.define("prroi_pooling_coor_backward(Tensor features, Tensor rois, Tensor output, Tensor output_diff, int pooled_height, int pooled_width, "
            "float spatial_scale) -> Tensor")
_LIB.define("xcorr_depthwise(Tensor x, Tensor kernel) -> Tensor")
_LIB.define("groupdw_xcorr(Tensor[] x, Tensor[] z, Tensor weight) -> Tensor")
_LIB.define("conv2d_nhwc(Tensor x, Tensor weight, Tensor scale, Tensor shift, int stride, int[] padding, int[] dilation, Tensor? residual, "
            "bool relu, str precision) -> Tensor")
_LIB.define("pred_conv(Tensor x, Tensor weight, Tensor bias, int mode, float mul, Tensor? adjust, Tensor? bias4) -> Tensor")


def _prroi_forward(features, rois, ph, pw, scale):
    features, rois = features.contiguous(), rois.contiguous()
    n, c, h, w = features.shape
    out = torch.empty((rois.shape[0], c, ph, pw), dtype=torch.float32, device=features.device)
    with torch.cuda.device(features.device):
        _lib.check(_lib.load().usot_prroi_pool_forward(_lib.ptr(features), _lib.ptr(rois), _lib.ptr(out), n, rois.shape[0], c, h, w, ph, pw, float(scale),
                                                       _stream(features)))
    return out


def _prroi_backward(features, rois, output, output_diff, ph, pw, scale):   # same argument list as prroi_pooling_backward_cuda (functional.py:69-72)
    n, c, h, w = features.shape
    grad = torch.empty_like(features)
    with torch.cuda.device(features.device):
        _lib.check(_lib.load().usot_prroi_pool_backward(_lib.ptr(rois.contiguous()), _lib.ptr(output_diff.contiguous()), _lib.ptr(grad), n, rois.shape[0], c,
                                                        h, w, ph, pw, float(scale), _stream(features)))
    return grad


def _prroi_coor_backward(features, rois, output, output_diff, ph, pw, scale):
    n, c, h, w = features.shape
    grad = torch.empty_like(rois)
    with torch.cuda.device(features.device):
        _lib.check(_lib.load().usot_prroi_pool_coor_backward(_lib.ptr(features.contiguous()), _lib.ptr(rois.contiguous()), _lib.ptr(output.contiguous()),
                                                             _lib.ptr(output_diff.contiguous()), _lib.ptr(grad), rois.shape[0], c, h, w, ph, pw,
                                                             float(scale), _stream(features)))
    return grad


_LIB.impl("prroi_pooling_forward", _prroi_forward, "CUDA")
_LIB.impl("prroi_pooling_backward", _prroi_backward, "CUDA")
_LIB.impl("prroi_pooling_coor_backward", _prroi_coor_backward, "CUDA")
_LIB.impl("xcorr_depthwise", lambda x, k: ops.XCorrDepthwiseFunction.apply(x, k), "CUDA")
_LIB.impl("groupdw_xcorr", lambda x, z, w: ops.groupdw_xcorr(list(x), list(z), w), "CUDA")
_LIB.impl("conv2d_nhwc", lambda x, w, sc, sh, stride, pad, dil, res, relu, prec: ops.conv2d_nhwc(x, w, sc, sh, stride, tuple(pad), tuple(dil), res, relu, prec),
          "CUDA")
_LIB.impl("pred_conv", lambda x, w, b, mode, mul, adjust, bias4: ops.pred_conv(x, w, b, mode, mul, adjust, bias4), "CUDA")

OPS = ("prroi_pooling_forward", "prroi_pooling_backward", "prroi_pooling_coor_backward", "xcorr_depthwise", "groupdw_xcorr", "conv2d_nhwc", "pred_conv")
