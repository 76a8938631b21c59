"""GPU parity of the stand-alone operators (through the C ABI) against the oracle / torch-fp32 CPU reference."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import usot_oracle as O
from helpers import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available(), "GPU tests need a B200"
    from usot_b200 import ops as _ops
    return _ops


def _ref_prroi(features, rois, ph, pw, scale):
    """The reference's own CUDA kernel (compiled unchanged into oracle/_ref) through ctypes."""
    path = os.path.join(ROOT, "oracle", "_ref", "libprroi_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libprroi_ref.so not built (make -C oracle)")
    lib = ctypes.CDLL(path)
    f, r = features.cuda().contiguous(), rois.cuda().contiguous()
    n, c, h, w = f.shape
    out = torch.zeros((r.shape[0], c, ph, pw), device="cuda")
    P = ctypes.c_void_p
    lib.PrRoIPoolingForwardGpu.argtypes = [P, P, P, P] + [ctypes.c_int] * 5 + [ctypes.c_float, ctypes.c_int]
    lib.PrRoIPoolingForwardGpu(P(torch.cuda.current_stream().cuda_stream), P(f.data_ptr()), P(r.data_ptr()), P(out.data_ptr()), c, h, w,
                               ph, pw, ctypes.c_float(scale), out.numel())
    torch.cuda.synchronize()
    return out.cpu()


@pytest.mark.parametrize("h,n", [(15, 3), (31, 4), (33, 2)])
def test_prroi_vs_reference_kernel_and_oracle(ops, h, n):
    g = torch.Generator().manual_seed(5 + h)
    feat = torch.randn(n, 256, h, h, generator=g)
    lo = torch.rand(n, 2, generator=g) * (h / 2) - 1.5  # some boxes start outside the map (usot_tracker.py:349)
    sz = torch.rand(n, 2, generator=g) * (h / 2) + 0.3
    rois = torch.cat([torch.arange(n).float().view(-1, 1).flip(0), lo, lo + sz], 1)
    rois[-1, 3] = rois[-1, 1]  # zero-width roi -> zeros (prroi_pooling_gpu_impl.cu:189-193)
    ours = ops.prroi_pool2d(feat.cuda(), rois.cuda(), 7, 7, 1.0).cpu()
    oracle = O.prroi_pool2d(feat, rois, 7, 7, 1.0)
    ref = _ref_prroi(feat, rois, 7, 7, 1.0)
    assert rel_err(oracle, ref) <= 5e-6, "oracle restatement drifted from the reference kernel"
    assert rel_err(ours, ref) <= 5e-6
    assert float(ours[-1].abs().max()) == 0.0


def test_prroi_empty_and_errors(ops):
    out = ops.prroi_pool2d(torch.zeros(1, 8, 15, 15, device="cuda"), torch.zeros(0, 5, device="cuda"), 7, 7, 1.0)
    assert tuple(out.shape) == (0, 8, 7, 7)
    with pytest.raises(NotImplementedError):
        ops.prroi_pool2d(torch.zeros(1, 8, 15, 15), torch.zeros(1, 5), 7, 7, 1.0)


@pytest.mark.parametrize("bx,bk,hx,wx,hk,wk", [(3, 3, 29, 29, 5, 5), (4, 1, 27, 29, 3, 5), (2, 2, 31, 29, 5, 3), (1, 1, 7, 7, 7, 7)])
def test_xcorr_depthwise(ops, bx, bk, hx, wx, hk, wk):
    g = torch.Generator().manual_seed(bx * 100 + hx)
    x, k = torch.randn(bx, 256, hx, wx, generator=g), torch.randn(bk, 256, hk, wk, generator=g)
    ref = O.xcorr_depthwise(x, k) if bk == bx else O.xcorr_depthwise(x, k)  # view trick broadcasts bk == 1
    ours = ops.xcorr_depthwise(x.cuda(), k.cuda()).cpu()
    assert ours.shape == ref.shape
    assert rel_err(ours, ref) <= 2e-6


def _groupdw_ref(xs, zs, w, rep):
    xs = [t.permute(0, 3, 1, 2).repeat_interleave(rep, 0) for t in xs]
    zs = [t.permute(0, 3, 1, 2) for t in zs]
    return O.groupdw(w, zs, xs).permute(0, 2, 3, 1)


@pytest.mark.parametrize("F_,nx,nz,n_out,strips", [(31, 3, 1, 3, 3), (31, 2, 2, 2, 2), (33, 2, 6, 6, 3), (33, 1, 1, 1, 2), (31, 2, 14, 14, 3),
                                                   (31, 6, 2, 6, 0), (31, 3, 1, 3, 0), (33, 2, 6, 6, 0), (31, 2, 14, 14, 0),
                                                   (29, 2, 1, 2, 0), (27, 3, 3, 3, 0),
                                                   (31, 6, 2, 6, -1), (33, 2, 6, 6, -1), (31, 2, 14, 14, -1)])
def test_groupdw_fused(ops, F_, nx, nz, n_out, strips):
    """strips = 0 selects the TMA-pipelined packed-FFMA2 kernel (default), -1 the scalar-FMA TMA kernel; 2 / 3 the
    register-staged variants.  F = 29 / 27 exercise the masked (ragged last strip) path of the default kernel."""
    from usot_b200 import _lib
    _lib.check(_lib.load().usot_set_tunable(b"groupdw_tma", 2 if strips == 0 else (1 if strips < 0 else 0)))
    if strips > 0:
        _lib.check(_lib.load().usot_set_tunable(b"groupdw_strips", strips))
    g = torch.Generator().manual_seed(F_ + nx)
    C = 256
    xs = [torch.randn(nx, F_ - 2, F_ - 2, C, generator=g), torch.randn(nx, F_ - 4, F_ - 2, C, generator=g),
          torch.randn(nx, F_ - 2, F_ - 4, C, generator=g)]
    zs = [torch.randn(nz, 5, 5, C, generator=g), torch.randn(nz, 3, 5, C, generator=g), torch.randn(nz, 5, 3, C, generator=g)]
    w = torch.tensor([0.3, -0.2, 0.9])
    zs_ref = [t.repeat_interleave(n_out // nz, 0) if 1 < nz < n_out else t for t in zs]
    ref = _groupdw_ref(xs, zs_ref, w, n_out // nx)
    ours = ops.groupdw_xcorr([t.cuda() for t in xs], [t.cuda() for t in zs], w.cuda(), n_out).cpu()
    assert ours.shape == ref.shape
    assert rel_err(ours, ref) <= 3e-6  # pure fp32 FMA, only the summation order differs
    _lib.check(_lib.load().usot_set_tunable(b"groupdw_strips", 3))
    _lib.check(_lib.load().usot_set_tunable(b"groupdw_tma", 2))


@pytest.mark.parametrize("F_", [31, 33, 29])
def test_groupdw_variants_bit_identical(ops, F_):
    """The three GroupDW kernels accumulate every output element in the same order with IEEE fma.rn (FFMA2 = two independent
    fma.rn), so they must agree bit for bit."""
    from usot_b200 import _lib
    g = torch.Generator().manual_seed(F_)
    C, nx = 256, 5
    xs = [torch.randn(nx, F_ - 2, F_ - 2, C, generator=g).cuda(), torch.randn(nx, F_ - 4, F_ - 2, C, generator=g).cuda(),
          torch.randn(nx, F_ - 2, F_ - 4, C, generator=g).cuda()]
    zs = [torch.randn(1, 5, 5, C, generator=g).cuda(), torch.randn(1, 3, 5, C, generator=g).cuda(), torch.randn(1, 5, 3, C, generator=g).cuda()]
    w = torch.tensor([0.1, 0.7, -0.4]).cuda()
    outs = []
    for mode in (2, 1, 0):
        _lib.check(_lib.load().usot_set_tunable(b"groupdw_tma", mode))
        outs.append(ops.groupdw_xcorr(xs, zs, w, nx).clone())
    _lib.check(_lib.load().usot_set_tunable(b"groupdw_tma", 2))
    _lib.check(_lib.load().usot_set_tunable(b"groupdw_warps4", 0))   # r02: the default FFMA2 kernel runs four consumer warps at F = 31 / 33
    outs.append(ops.groupdw_xcorr(xs, zs, w, nx).clone())
    _lib.check(_lib.load().usot_set_tunable(b"groupdw_warps4", 1))
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]) and torch.equal(outs[0], outs[3])


def test_groupdw_linearity_full_size(ops):
    """Size-independent property at the BASELINE batch (256 crops): xcorr is linear in the search maps."""
    torch.manual_seed(0)
    C, F_, B = 256, 31, 256
    mk = lambda *s: torch.randn(*s, device="cuda")
    xa = [mk(B, F_ - 2, F_ - 2, C), mk(B, F_ - 4, F_ - 2, C), mk(B, F_ - 2, F_ - 4, C)]
    xb = [mk(B, F_ - 2, F_ - 2, C), mk(B, F_ - 4, F_ - 2, C), mk(B, F_ - 2, F_ - 4, C)]
    zs = [mk(1, 5, 5, C), mk(1, 3, 5, C), mk(1, 5, 3, C)]
    w = torch.tensor([1.0, 0.5, 1.5], device="cuda")
    ya, yb = ops.groupdw_xcorr(xa, zs, w, B), ops.groupdw_xcorr(xb, zs, w, B)
    yab = ops.groupdw_xcorr([a + 2.0 * b for a, b in zip(xa, xb)], zs, w, B)
    assert rel_err(yab, ya + 2.0 * yb) <= 1e-5
    # and sample 17 of the batch equals the same sample run alone (no cross-sample leakage)
    y17 = ops.groupdw_xcorr([t[17:18].contiguous() for t in xa], zs, w, 1)
    assert torch.equal(y17[0], ya[17])


@pytest.mark.parametrize("cout,mode,r,n,min_batch", [(4, 1, 25, 3, 1), (1, 0, 25, 5, 1), (4, 1, 27, 2, 1), (1, 0, 27, 2, 1), (4, 1, 19, 2, 1),
                                                      (1, 0, 9, 3, 1), (4, 1, 25, 3, 0), (1, 0, 25, 3, 0), (4, 1, 25, 70, 48)])
def test_pred_conv(ops, cout, mode, r, n, min_batch):
    """bbox_pred / cls_pred heads (connect.py:235-241): TMA-streamed per-image kernel (min_batch >= 1 forces it at these small
    batches; 48 is the default dispatch) and the warp-per-pixel kernel (min_batch = 0) against torch fp32."""
    from usot_b200 import _lib
    _lib.check(_lib.load().usot_set_tunable(b"pred_tma_min_batch", min_batch))
    g = torch.Generator().manual_seed(100 * cout + r)
    x = torch.randn(n, r, r, 256, generator=g)
    w = torch.randn(cout, 256, 3, 3, generator=g) * 0.02
    b = torch.randn(cout, generator=g) * 0.1
    adjust, bias4 = torch.tensor([0.7]), torch.randn(1, 4, 1, 1, generator=g) * 0.3
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1)
    ref = (0.1 * y if mode == 0 else torch.exp(adjust.double() * y + bias4.double())).float()
    ours = ops.pred_conv(x.cuda(), w.cuda(), b.cuda(), mode=mode, mul=0.1, adjust=adjust.cuda(), bias4=bias4.cuda()).cpu()
    _lib.check(_lib.load().usot_set_tunable(b"pred_tma_min_batch", 48))
    assert ours.shape == ref.shape
    assert rel_err(ours, ref) <= 5e-6


@pytest.mark.parametrize("cout,mode,r", [(4, 1, 25), (1, 0, 25), (4, 1, 27), (1, 0, 19)])
def test_pred_conv_variants_bit_identical(ops, cout, mode, r):
    """The per-image TMA kernel and the small-batch kernel build the same sequential fma chains: results must not depend on
    which one the batch size selects."""
    from usot_b200 import _lib
    g = torch.Generator().manual_seed(cout + r)
    x = torch.randn(3, r, r, 256, generator=g).cuda()
    w = (torch.randn(cout, 256, 3, 3, generator=g) * 0.02).cuda()
    b = (torch.randn(cout, generator=g) * 0.1).cuda()
    adjust, bias4 = torch.tensor([0.7]).cuda(), (torch.randn(4, generator=g) * 0.3).cuda()
    outs = []
    for min_batch in (1, 0):
        _lib.check(_lib.load().usot_set_tunable(b"pred_tma_min_batch", min_batch))
        outs.append(ops.pred_conv(x, w, b, mode=mode, mul=0.1, adjust=adjust, bias4=bias4).clone())
    _lib.check(_lib.load().usot_set_tunable(b"pred_tma_min_batch", 48))
    assert torch.equal(outs[0], outs[1])


def test_empty_inputs_are_noops(ops):
    """Zero rois / crops / maps: every stand-alone op returns an empty tensor of the right shape and leaves no CUDA error behind
    (the reference's PrRoIPool launches a zero-size grid and exit()s on the resulting launch error, prroi_pooling_gpu_impl.cu:20-27)."""
    from usot_b200 import tracker_ops
    feat = torch.randn(2, 8, 9, 9).cuda()
    out = ops.prroi_pool2d(feat, torch.zeros(0, 5).cuda(), 7, 7, 1.0)
    assert tuple(out.shape) == (0, 8, 7, 7)
    y = ops.pred_conv(torch.zeros(0, 25, 25, 256).cuda(), torch.zeros(1, 256, 3, 3).cuda(), torch.zeros(1).cuda())
    assert tuple(y.shape) == (0, 1, 25, 25)
    x = ops.xcorr_depthwise(torch.zeros(0, 256, 29, 29).cuda(), torch.zeros(1, 256, 5, 5).cuda())
    assert tuple(x.shape) == (0, 256, 25, 25)
    frames = torch.zeros(1, 32, 32, 3, dtype=torch.uint8).cuda()
    c = tracker_ops.crop_resize(frames, torch.zeros(0, 4, dtype=torch.int32).cuda(), torch.zeros(0, 3, dtype=torch.uint8).cuda(), 127)
    assert tuple(c.shape) == (0, 3, 127, 127)
    torch.cuda.synchronize()
    assert torch.isfinite(ops.prroi_pool2d(feat, torch.tensor([[0.0, 1.0, 1.0, 5.0, 5.0]]).cuda(), 7, 7, 1.0)).all()  # still healthy


@pytest.mark.parametrize("bx,bk,hk,wk", [(3, 3, 5, 5), (4, 1, 5, 5), (2, 1, 3, 5), (2, 2, 5, 3)])
def test_xcorr_depthwise_backward_vs_autograd(ops, bx, bk, hk, wk):
    """Both gradients of the depth-wise xcorr against torch autograd through the oracle's F.conv2d formulation (float64)."""
    g = torch.Generator().manual_seed(bx * 10 + bk)
    c, hx, wx = 64, 29 if hk == 5 else 27, 29 if wk == 5 else 27
    x = torch.randn(bx, c, hx, wx, generator=g)
    k = torch.randn(bk, c, hk, wk, generator=g)
    go = torch.randn(bx, c, hx - hk + 1, wx - wk + 1, generator=g)
    xr, kr = x.double().requires_grad_(True), k.double().requires_grad_(True)
    O.xcorr_depthwise(xr, kr).backward(go.double())
    xc, kc = x.cuda().requires_grad_(True), k.cuda().requires_grad_(True)
    ops.xcorr_depthwise(xc, kc).backward(go.cuda())
    assert rel_err(xc.grad.cpu(), xr.grad.float()) <= 5e-6
    assert rel_err(kc.grad.cpu(), kr.grad.float()) <= 5e-6


@pytest.mark.parametrize("cin,cout,k,pad,dil,h", [(256, 256, 3, (2, 2), (2, 2), 13), (256, 1024, 1, (0, 0), (1, 1), 9), (256, 256, 3, (0, 0), (2, 1), 15),
                                                   (512, 1024, 3, (1, 1), (1, 1), 9)])
@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_conv_dgrad_on_the_forward_kernel(ops, cin, cout, k, pad, dil, h, precision):
    """Input gradient of the stride-1 convs computed by the forward conv kernel with transposed, flipped filters."""
    g = torch.Generator().manual_seed(cin + k)
    x = torch.randn(2, cin, h, h, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, k, k, generator=g, dtype=torch.float64) / (cin * k * k) ** 0.5
    y = torch.nn.functional.conv2d(x, w, None, 1, pad, dil)
    go = torch.randn(y.shape, generator=g, dtype=torch.float64)
    y.backward(go)
    ours = ops.conv2d_nhwc_input_grad(go.float().permute(0, 2, 3, 1).contiguous().cuda(), w.float().cuda(), pad, dil, precision)
    assert tuple(ours.shape) == (2, h, h, cin)
    assert rel_err(ours.permute(0, 3, 1, 2).cpu(), x.grad.float()) <= (5e-6 if precision == "fp32" else 3e-5)


CONV_CASES = [
    # cin, cout, k, stride, pad, dil, h, w, residual, relu
    (64, 64, 1, 1, (0, 0), (1, 1), 17, 17, False, True),
    (64, 256, 1, 1, (0, 0), (1, 1), 9, 11, True, True),
    (128, 128, 3, 2, (0, 0), (1, 1), 21, 21, False, True),     # layer2.0.conv2
    (256, 512, 3, 2, (0, 0), (1, 1), 15, 15, False, False),    # layer2.0.downsample
    (256, 256, 3, 1, (2, 2), (2, 2), 13, 13, False, True),     # layer3.x.conv2 (dilated)
    (512, 1024, 3, 1, (1, 1), (1, 1), 9, 9, False, False),     # layer3.0.downsample
    (256, 256, 3, 1, (0, 0), (2, 1), 15, 15, False, True),     # matrix12
    (256, 256, 3, 1, (0, 0), (1, 2), 7, 7, False, True),       # matrix21 on the template
    (256, 256, 3, 1, (1, 1), (1, 1), 25, 25, False, True),     # tower
    (1024, 256, 1, 1, (0, 0), (1, 1), 31, 31, False, False),   # neck
]


CONV_TOL = {"fp32": 5e-6, "fp16x3": 3e-5, "fp16": 3e-3}  # fp16x3: tensor-core accumulation truncates, error grows ~3e-9*K


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "fp16"])
def test_conv2d_nhwc(ops, case, precision):
    cin, cout, k, stride, pad, dil, h, w, use_res, relu = case
    g = torch.Generator().manual_seed(cin + cout + k)
    n = 3
    x = torch.randn(n, cin, h, w, generator=g)
    wgt = torch.randn(cout, cin, k, k, generator=g) / np.sqrt(cin * k * k)
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    ref = F.conv2d(x, wgt, None, stride, pad, dil) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    res = torch.randn_like(ref) if use_res else None
    if use_res:
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    ours = ops.conv2d_nhwc(x.permute(0, 2, 3, 1).contiguous().cuda(), wgt.cuda(), scale.cuda(), shift.cuda(), stride, pad, dil,
                           None if res is None else res.permute(0, 2, 3, 1).contiguous().cuda(), relu, precision)
    ours = ours.permute(0, 3, 1, 2).cpu()
    assert ours.shape == ref.shape
    err = rel_err(ours, ref)
    print(precision, case[:6], f"{err:.2e}")
    assert err <= CONV_TOL[precision]


def _ref_prroi_backward(features, rois, out, grad_out, ph, pw, scale, coor):
    """The reference's own backward kernels (oracle/_ref, prroi_pooling_gpu_impl.cu:404-443) through ctypes."""
    path = os.path.join(ROOT, "oracle", "_ref", "libprroi_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libprroi_ref.so not built (make -C oracle)")
    lib = ctypes.CDLL(path)
    f, r, o, g = [t.cuda().contiguous() for t in (features, rois, out, grad_out)]
    n, c, h, w = f.shape
    res = torch.zeros_like(r) if coor else torch.zeros_like(f)
    fn = lib.PrRoIPoolingCoorBackwardGpu if coor else lib.PrRoIPoolingBackwardGpu
    P = ctypes.c_void_p
    fn.argtypes = [P] * 6 + [ctypes.c_int] * 5 + [ctypes.c_float, ctypes.c_int, ctypes.c_int]
    fn(P(torch.cuda.current_stream().cuda_stream), P(f.data_ptr()), P(r.data_ptr()), P(o.data_ptr()), P(g.data_ptr()), P(res.data_ptr()),
       c, h, w, ph, pw, ctypes.c_float(scale), o.numel(), res.numel())
    torch.cuda.synchronize()
    return res.cpu()


@pytest.mark.parametrize("h,n", [(15, 3), (31, 4)])
def test_prroi_backward_vs_reference_kernels(ops, h, n):
    g = torch.Generator().manual_seed(17 + h)
    feat = torch.randn(n, 64, h, h, generator=g)
    lo = torch.rand(n, 2, generator=g) * (h / 2) - 1.0
    sz = torch.rand(n, 2, generator=g) * (h / 2) + 0.5
    rois = torch.cat([torch.arange(n).float().view(-1, 1), lo, lo + sz], 1)
    grad_out = torch.randn(n, 64, 7, 7, generator=g)
    f = feat.cuda().requires_grad_(True)
    r = rois.cuda().requires_grad_(True)
    out = ops.prroi_pool2d(f, r, 7, 7, 1.0)
    out.backward(grad_out.cuda())
    ref_f = _ref_prroi_backward(feat, rois, out.detach().cpu(), grad_out, 7, 7, 1.0, coor=False)
    ref_r = _ref_prroi_backward(feat, rois, out.detach().cpu(), grad_out, 7, 7, 1.0, coor=True)
    assert rel_err(f.grad, ref_f) <= 1e-5      # atomics: summation order differs
    assert rel_err(r.grad[:, 1:], ref_r[:, 1:]) <= 1e-4
    assert float(r.grad[:, 0].abs().max()) == 0.0


def test_prroi_backward_is_the_adjoint_of_forward(ops):
    """Size-independent property: forward is linear in the features, so <P f, g> == <f, P^T g>."""
    torch.manual_seed(3)
    f = torch.randn(8, 256, 31, 31, device="cuda", requires_grad=True)
    rois = torch.cat([torch.arange(8.0).view(-1, 1), torch.rand(8, 2) * 10 + 2, torch.rand(8, 2) * 10 + 15], 1).cuda()
    g = torch.randn(8, 256, 7, 7, device="cuda")
    out = ops.prroi_pool2d(f, rois, 7, 7, 1.0)
    out.backward(g)
    lhs, rhs = float((out.detach().double() * g.double()).sum()), float((f.detach().double() * f.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
