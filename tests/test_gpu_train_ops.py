"""GPU tests of the training-path operators (usot_b200/csrc/train_kernels.cu and the dgrad routes of usot_b200/train.py) against the
torch-CPU restatements in tests/train_ref.py (autograd of the same modules the reference trains, scripts/train_usot.py:229-236)."""
import pytest
import torch
import torch.nn.functional as F

import train_ref as R
from helpers import rel_err

pytestmark = pytest.mark.gpu


def _run(name, *args):
    """Call the C-ABI op `name` with GPU tensors (stream = torch's current stream) and return nothing (outputs are arguments)."""
    from usot_b200 import _lib, train
    dev = next(a for a in args if torch.is_tensor(a)).device
    conv = [(_lib.ptr(a) if (torch.is_tensor(a) or a is None) else a) for a in args]
    train._lib_call(name, dev, *conv, torch.cuda.current_stream(dev).cuda_stream)


CONVS = [  # n, h, w, cin, cout, k, stride, pad, dil
    (2, 63, 63, 3, 64, 7, 2, (0, 0), (1, 1)),       # stem
    (2, 15, 15, 64, 256, 1, 1, (0, 0), (1, 1)),     # bottleneck 1x1
    (3, 31, 31, 128, 128, 3, 2, (0, 0), (1, 1)),    # layer2.0.conv2 (stride 2)
    (2, 33, 31, 256, 512, 3, 2, (0, 0), (1, 1)),    # layer2.0.downsample (stride 2), ragged size
    (2, 15, 15, 64, 64, 3, 1, (2, 2), (2, 2)),      # layer3 conv2 (dilation 2)
    (2, 15, 17, 64, 64, 3, 1, (0, 0), (2, 1)),      # encoder matrix12
    (2, 15, 17, 64, 128, 3, 1, (0, 0), (1, 2)),     # encoder matrix21
    (2, 25, 25, 256, 4, 3, 1, (1, 1), (1, 1)),      # bbox_pred
    (5, 25, 25, 256, 1, 3, 1, (1, 1), (1, 1)),      # cls_pred
    (1, 7, 7, 256, 256, 3, 1, (0, 0), (1, 1)),      # template-side encoder
]


PREC_TOL = {"fp32": 2e-5, "fp16x3": 5e-5, "fp16": 6e-3}


@pytest.mark.parametrize("precision", ["fp32", "fp16x3", "fp16"])
@pytest.mark.parametrize("cfg", CONVS, ids=[f"{c[3]}to{c[4]}k{c[5]}s{c[6]}" + (f"d{c[8][0]}{c[8][1]}" if c[8] != (1, 1) else "") + f"_{c[1]}x{c[2]}" for c in CONVS])
def test_conv_wgrad_vs_autograd(cfg, precision):
    """fp32: the FMA kernel.  fp16x3 / fp16: the tcgen05 kernel (MN-major operands) for the wide layers -- with gradients of magnitude 1e-6,
    far below fp16's normal range, so the device-side power-of-two scaling is exercised -- and the FMA kernel for the thin ones."""
    from usot_b200 import train
    n, h, w, cin, cout, k, stride, pad, dil = cfg
    g = torch.Generator().manual_seed(h * 100 + cin)
    x = torch.randn(n, h, w, cin, generator=g) * 3.0
    ho = (h + 2 * pad[0] - dil[0] * (k - 1) - 1) // stride + 1
    wo = (w + 2 * pad[1] - dil[1] * (k - 1) - 1) // stride + 1
    gy = torch.randn(n, ho, wo, cout, generator=g) * 1e-6
    ref = torch.empty(k * k * cin, cout)
    R.usot_conv2d_wgrad_nhwc(x, gy, n, h, w, cin, cout, k, k, stride, pad[0], pad[1], dil[0], dil[1], ref, 0, 0)
    out = train.conv_wgrad(x.cuda(), gy.cuda(), (cout, cin, k, k), stride, pad, dil, precision)      # OIHW
    ref_oihw = ref.view(k, k, cin, cout).permute(3, 2, 0, 1)
    err = rel_err(out, ref_oihw)
    print("wgrad", cfg, precision, f"{err:.2e}")
    assert err <= PREC_TOL[precision]


@pytest.mark.parametrize("cfg", [c for c in CONVS if c[3] != 3], ids=lambda c: f"{c[3]}to{c[4]}k{c[5]}s{c[6]}_{c[1]}x{c[2]}")
@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_conv_dgrad_vs_autograd(cfg, precision):
    """stride 1 (any dilation / padding) and the stride-2 parity decomposition on the forward kernels (fp32 FMA, and tcgen05 fp16x3 with the
    gradient pre-scaled on the device); thin convs on the gather kernel."""
    from usot_b200 import train
    n, h, w, cin, cout, k, stride, pad, dil = cfg
    g = torch.Generator().manual_seed(h * 10 + cout)
    wt = torch.randn(cout, cin, k, k, generator=g) * 0.1
    x = torch.zeros(n, cin, h, w, requires_grad=True)
    y = F.conv2d(x, wt, None, stride, pad, dil)
    gy = torch.randn(y.shape, generator=g) * 1e-7
    (ref,) = torch.autograd.grad(y, x, gy)
    out = train.conv_dgrad(gy.permute(0, 2, 3, 1).contiguous().cuda(), wt.cuda(), (h, w), stride, pad, dil, precision=precision)
    assert tuple(out.shape) == (n, h, w, cin)
    assert rel_err(out.permute(0, 3, 1, 2), ref) <= PREC_TOL[precision]


def test_pow2_scale():
    from usot_b200 import train
    for mag in (3e-9, 1.0, 700.0):
        x = torch.randn(4096, generator=torch.Generator().manual_seed(1)) * mag
        g, inv, _ = train._prescale(x.cuda(), 8)
        m = float(g.abs().max())
        assert 1024.0 <= m < 2048.0
        s = 1.0 / float(inv[0])
        assert torch.equal(g.cpu(), x * s) and abs(torch.log2(torch.tensor(s)).item() - round(torch.log2(torch.tensor(s)).item())) == 0.0
    g, inv, _ = train._prescale(torch.zeros(64, device="cuda"), 4)
    assert float(inv[0]) == 1.0 and float(g.abs().max()) == 0.0


def test_stem_raw_and_device_weight_pack():
    """usot_stem_conv_raw = bare conv1; usot_conv2d_nhwc's tensor-core path now packs / splits the weights on the device: same result as fp32."""
    from usot_b200 import ops, train
    g = torch.Generator().manual_seed(1)
    x = torch.rand(2, 3, 127, 127, generator=g) * 255
    w = torch.randn(64, 3, 7, 7, generator=g) * 0.05
    wc = w.cuda().requires_grad_(True)
    out = train._StemConv.apply(x.cuda(), wc)
    wr = w.clone().requires_grad_(True)
    ref = F.conv2d(x, wr, None, 2, 0)
    assert rel_err(out.permute(0, 3, 1, 2), ref) <= 2e-6
    gy = torch.randn(ref.shape, generator=g) * 1e-3
    ref.backward(gy)
    out.backward(gy.permute(0, 2, 3, 1).contiguous().cuda())
    assert rel_err(wc.grad, wr.grad) <= 2e-5      # usot_stem_conv_wgrad (im2col on the fly, (channel, tap) as the GEMM's M dimension)
    xa = torch.randn(2, 15, 15, 128, generator=g)
    wa = torch.randn(256, 128, 3, 3, generator=g) * 0.03
    one, zero = torch.ones(256), torch.zeros(256)
    ref = F.conv2d(xa.permute(0, 3, 1, 2), wa, None, 1, 1).permute(0, 2, 3, 1)
    for prec, tol in (("fp32", 2e-6), ("fp16x3", 2e-5), ("fp16", 4e-3)):
        o = ops.conv2d_nhwc(xa.cuda(), wa.cuda(), one.cuda(), zero.cuda(), padding=(1, 1), precision=prec)
        assert rel_err(o, ref) <= tol, prec


@pytest.mark.parametrize("m,c", [(2 * 31 * 31, 256), (7, 64), (5000, 1024), (2 * 25 * 25, 4)])
def test_bn_stats_and_channel_sum(m, c):
    g = torch.Generator().manual_seed(m + c)
    x = torch.randn(m, c, generator=g) * 3 + 5
    bias = torch.randn(c, generator=g)
    mean, var = torch.empty(c), torch.empty(c)
    R.usot_bn_stats(x, bias, m, c, mean, var, 0)
    gm, gv = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    _run("usot_bn_stats", x.cuda(), bias.cuda(), m, c, gm, gv)
    assert torch.allclose(gm.cpu(), mean, rtol=1e-6, atol=1e-6) and torch.allclose(gv.cpu(), var, rtol=2e-6, atol=1e-7)
    s = torch.empty(c, device="cuda")
    _run("usot_channel_sum", x.cuda(), m, c, s)
    assert torch.allclose(s.cpu(), x.double().sum(0).float(), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("train_mode", [True, False])
@pytest.mark.parametrize("bias,residual,relu", [(False, False, True), (True, False, True), (False, True, True), (False, False, False), (True, True, False)])
def test_bn_forward_backward_vs_torch(train_mode, bias, residual, relu):
    """nn.BatchNorm2d (+ conv bias)(+ shortcut)(+ ReLU) forward and backward against torch autograd, batch and running statistics."""
    from usot_b200 import train
    g = torch.Generator().manual_seed(int(train_mode) * 8 + int(bias) * 4 + int(residual) * 2 + int(relu))
    n, h, w, c = 3, 9, 11, 64
    x = (torch.randn(n, c, h, w, generator=g) * 2 + 1).requires_grad_(True)
    b = torch.randn(c, generator=g).requires_grad_(True) if bias else None
    res = torch.randn(n, c, h, w, generator=g).requires_grad_(True) if residual else None
    bn = torch.nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.normal_(generator=g); bn.running_mean.normal_(generator=g); bn.running_var.uniform_(0.5, 2.0, generator=g)
    import copy
    bn_ref = copy.deepcopy(bn).train(train_mode)
    v = x + (b.view(1, -1, 1, 1) if bias else 0)
    y = bn_ref(v)
    if residual:
        y = y + res
    if relu:
        y = F.relu(y)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    # ours (NHWC, CUDA)
    bn_c = copy.deepcopy(bn).cuda().train(train_mode)
    xc = x.detach().permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
    bc = b.detach().cuda().requires_grad_(True) if bias else None
    rc = res.detach().permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True) if residual else None
    yo = train.batchnorm(xc, bn_c, conv_bias=bc, residual=rc, relu=relu, train=train_mode)
    yo.backward(gy.permute(0, 2, 3, 1).contiguous().cuda())
    assert rel_err(yo.permute(0, 3, 1, 2), y) <= 5e-6
    assert rel_err(xc.grad.permute(0, 3, 1, 2), x.grad) <= 2e-5
    assert rel_err(bn_c.weight.grad, bn_ref.weight.grad) <= 2e-5 and rel_err(bn_c.bias.grad, bn_ref.bias.grad) <= 2e-5
    if bias:
        if train_mode:   # exactly zero in exact arithmetic; both sides hold rounding noise
            assert float(bc.grad.abs().max()) <= 1e-4 * float(gy.abs().sum() / c)
        else:
            assert rel_err(bc.grad, b.grad) <= 2e-5
    if residual:
        assert rel_err(rc.grad.permute(0, 3, 1, 2), res.grad) <= 2e-6
    if train_mode:   # running statistics updated like torch (momentum 0.1, unbiased variance)
        assert torch.allclose(bn_c.running_mean.cpu(), bn_ref.running_mean, rtol=1e-5, atol=1e-6)
        assert torch.allclose(bn_c.running_var.cpu(), bn_ref.running_var, rtol=1e-5, atol=1e-6)
        assert int(bn_c.num_batches_tracked) == 1
    else:
        assert torch.equal(bn_c.running_mean.cpu(), bn.running_mean)


@pytest.mark.parametrize("shape", [(2, 61, 61, 64), (1, 8, 7, 16), (2, 5, 5, 4), (1, 1, 1, 8)])
def test_maxpool_backward_vs_torch_incl_ties(shape):
    from usot_b200 import train
    g = torch.Generator().manual_seed(sum(shape))
    x = F.relu(torch.randn(shape, generator=g))          # post-ReLU map: many exact zeros -> tied maxima (first one wins in torch)
    x[0, :2, :2, 0] = 1.5                                  # tied positive maxima as well
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    y = F.max_pool2d(xr, 3, 2, 1)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy)
    xc = x.cuda().requires_grad_(True)
    yo = train._MaxPool.apply(xc)
    yo.backward(gy.permute(0, 2, 3, 1).contiguous().cuda())
    assert torch.equal(yo.detach().cpu().permute(0, 3, 1, 2), y.detach())
    assert rel_err(xc.grad.permute(0, 3, 1, 2), xr.grad) <= 1e-6


@pytest.mark.parametrize("b,nq", [(2, 7), (3, 1), (1, 3)])
def test_conf_fusion_backward_vs_autograd(b, nq):
    from usot_b200 import train
    g = torch.Generator().manual_seed(b + nq)
    conf = (F.relu(torch.randn(b * nq, 5, 5, 64, generator=g) * 4.0)).requires_grad_(True)   # post-ReLU; exceeds the upper clamp (4) in places
    value = torch.randn(b * nq, 5, 5, 64, generator=g).requires_grad_(True)
    out = R.conf_fusion(conf, value, nq)
    gy = torch.randn(out.shape, generator=g)
    out.backward(gy)
    cc, vc = conf.detach().cuda().requires_grad_(True), value.detach().cuda().requires_grad_(True)
    oo = train._ConfFusion.apply(cc, vc, nq)
    oo.backward(gy.cuda())
    assert rel_err(oo, out) <= 2e-6 and rel_err(vc.grad, value.grad) <= 1e-5
    if nq == 1:   # a single slot: the normalised weight is identically 1, so the confidence gets NO gradient (0 in exact arithmetic)
        assert float(cc.grad.abs().max()) <= 1e-5 * float(gy.abs().max() * value.abs().max())
    else:
        assert rel_err(cc.grad, conf.grad) <= 1e-5


def test_weighted_sum3_and_groupdw_vs_autograd():
    """GroupDW = softmax-weighted sum of three depth-wise correlations (connect.py:86-102): forward and all gradients (maps, kernels, weight)."""
    import types
    import usot_oracle as O
    from usot_b200 import train
    g = torch.Generator().manual_seed(9)
    xs = [torch.randn(2, 64, 13, 13, generator=g).requires_grad_(True), torch.randn(2, 64, 11, 13, generator=g).requires_grad_(True),
          torch.randn(2, 64, 13, 11, generator=g).requires_grad_(True)]
    zs = [torch.randn(2, 64, 5, 5, generator=g).requires_grad_(True), torch.randn(2, 64, 3, 5, generator=g).requires_grad_(True),
          torch.randn(2, 64, 5, 3, generator=g).requires_grad_(True)]
    wt = torch.tensor([1.0, 0.5, 1.5], requires_grad=True)
    out = O.groupdw(wt, zs, xs)
    gy = torch.randn(out.shape, generator=g)
    out.backward(gy)
    to = lambda t: t.detach().permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
    xc, zc = [to(t) for t in xs], [to(t) for t in zs]
    wc = wt.detach().cuda().requires_grad_(True)
    oo = train._groupdw(types.SimpleNamespace(weight=wc), zc, xc)
    oo.backward(gy.permute(0, 2, 3, 1).contiguous().cuda())
    assert rel_err(oo.permute(0, 3, 1, 2), out) <= 5e-6
    for a, b in zip(xc + zc, xs + zs):
        assert rel_err(a.grad.permute(0, 3, 1, 2), b.grad) <= 2e-5
    assert rel_err(wc.grad, wt.grad) <= 2e-5


def test_loss_backward_vs_autograd():
    import usot_oracle as O
    from usot_b200 import train
    g = torch.Generator().manual_seed(17)
    pred = (torch.randn(3, 1, 25, 25, generator=g) * 2).requires_grad_(True)
    label = torch.zeros(3, 25, 25)
    label[:, 10:15, 10:15] = 1.0
    label[0, 0, :5] = 0.5                                   # ignored cells
    loss = O.weighted_bce(pred, label) * 1.7
    loss.backward()
    pc = pred.detach().cuda().requires_grad_(True)
    lo = train._WeightedBCE.apply(pc, label.cuda()) * 1.7
    lo.backward()
    assert abs(float(lo) - float(loss)) <= 2e-6 * abs(float(loss)) and rel_err(pc.grad, pred.grad) <= 1e-5
    one = torch.zeros(1, 25, 25)
    one[0, 3, 3] = 1.0                                      # exactly one positive: that class contributes neither loss nor gradient
    p1 = pred.detach()[:1].clone().requires_grad_(True)
    O.weighted_bce(p1, one).backward()
    p1c = pred.detach()[:1].cuda().requires_grad_(True)
    train._WeightedBCE.apply(p1c, one.cuda()).backward()
    assert rel_err(p1c.grad, p1.grad) <= 1e-5 and float(p1c.grad[0, 0, 3, 3]) == 0.0
    bbox = (torch.rand(4, 4, 25, 25, generator=g) * 60 + 0.5).requires_grad_(True)
    target = torch.rand(4, 25, 25, 4, generator=g) * 40 + 5
    weight = (torch.rand(4, 25, 25, generator=g) > 0.9).float()
    l2 = O.iou_loss(bbox, target, weight)
    l2.backward()
    bc = bbox.detach().cuda().requires_grad_(True)
    l2o = train._IoULoss.apply(bc, target.cuda(), weight.cuda())
    l2o.backward()
    assert abs(float(l2o) - float(l2)) <= 3e-6 * abs(float(l2)) and rel_err(bc.grad, bbox.grad) <= 1e-5
