"""Stand-alone GPU tests of the small kernels that were only covered end to end in round 1: the stem conv (CUDA-core and
tensor-core), the stem max-pool (fp32 and split-fp16 outputs), the Conf_Fusion reduction, the cycle-memory glue (arg-max, box
maps) and the two losses -- each against the oracle function that restates the reference op."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import usot_oracle as O
from helpers import load_weights, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-6), ("fp16x3", 2e-5), ("fp16", 4e-3)])
@pytest.mark.parametrize("size,n", [(127, 2), (255, 3), (271, 1)])
def test_stem_conv_vs_oracle(precision, tol, size, n):
    """conv1 + bn1 + relu (lib/models/modules.py:138-140) on raw 0..255 crops, calibrated BN statistics."""
    from usot_b200 import ops
    sd = load_weights("damp025")
    p = "features.features."
    x = O.synth_inputs(90 + size, batch=n, search_size=size)[1]
    with torch.no_grad():
        ref = F.relu(O._bn(sd, O._conv(sd, x, p + "conv1", 2, 0), p + "bn1"))
    scale = sd[p + "bn1.weight"].double() / torch.sqrt(sd[p + "bn1.running_var"].double() + 1e-5)
    shift = sd[p + "bn1.bias"].double() - sd[p + "bn1.running_mean"].double() * scale
    out = ops.stem_conv(x.cuda(), sd[p + "conv1.weight"], scale.float(), shift.float(), precision)
    assert tuple(out.shape) == (n, (size - 7) // 2 + 1, (size - 7) // 2 + 1, 64)
    err = rel_err(out.permute(0, 3, 1, 2), ref)
    print("stem", precision, size, f"{err:.2e}")
    assert err <= tol


@pytest.mark.parametrize("precision,tol", [("fp16x3", 2e-5), ("fp16", 4e-3)])
@pytest.mark.parametrize("size,n", [(127, 1), (127, 9), (255, 1), (255, 3), (255, 40), (261, 2), (191, 5), (63, 2)])
def test_fused_stem_maxpool_vs_oracle(precision, tol, size, n):
    """conv1 + bn1 + relu + maxpool (lib/models/modules.py:138-141) as ONE tcgen05 kernel over the space-to-depth image (overlapping-row
    TMA view = im2col, pooling in the epilogue), for every band plan these batch sizes select, against the fp32 oracle."""
    from usot_b200 import ops
    sd = load_weights("damp025")
    p = "features.features."
    g = torch.Generator().manual_seed(700 + size + n)
    x = torch.rand(n, 3, size, size, generator=g) * 255.0
    with torch.no_grad():
        ref = F.max_pool2d(F.relu(O._bn(sd, O._conv(sd, x, p + "conv1", 2, 0), p + "bn1")), 3, 2, 1)
    scale = sd[p + "bn1.weight"].double() / torch.sqrt(sd[p + "bn1.running_var"].double() + 1e-5)
    shift = sd[p + "bn1.bias"].double() - sd[p + "bn1.running_mean"].double() * scale
    out = ops.stem_maxpool(x.cuda(), sd[p + "conv1.weight"], scale.float(), shift.float(), precision)
    assert tuple(out.shape) == tuple(ref.permute(0, 2, 3, 1).shape)
    err = rel_err(out.permute(0, 3, 1, 2), ref)
    print("stem+pool", precision, size, n, f"{err:.2e}")
    assert err <= tol


@pytest.mark.parametrize("shape", [(2, 125, 125, 64), (1, 61, 61, 64), (3, 8, 7, 16), (1, 1, 1, 4), (2, 2, 5, 8)])
def test_maxpool_fp32_is_exact_and_split_rounds_like_fp16x2(shape):
    from usot_b200 import ops
    g = torch.Generator().manual_seed(sum(shape))
    x = (torch.rand(shape, generator=g) - 0.3) * 7.0   # includes negatives: padding must act as -inf, not 0 (modules.py:75)
    ref = F.max_pool2d(x.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    out = ops.maxpool3x3s2p1_nhwc(x.cuda())
    assert out.shape == ref.shape and torch.equal(out.cpu(), ref)
    sp = ops.maxpool3x3s2p1_nhwc(x.cuda(), split=True).cpu()
    hi = ref.half().float()
    expect = hi + (ref - hi).half().float()          # value = rn16(v) + rn16(v - rn16(v))
    assert torch.equal(sp, expect)
    assert float((sp - ref).abs().max()) <= 2.0 ** -21 * float(ref.abs().max())


@pytest.mark.parametrize("b,nq", [(1, 7), (3, 2), (2, 1), (5, 4)])
def test_conf_fusion_reduction_vs_oracle(b, nq):
    """connect.py:130-144 on generator outputs that exercise both clamp limits (-6, 4)."""
    from usot_b200 import ops
    g = torch.Generator().manual_seed(b * 10 + nq)
    conf = F.relu(torch.randn(b * nq, 25, 25, 256, generator=g) * 4.0)   # post-ReLU like the reference's generator
    conf[0, 0, 0, :8] = torch.tensor([-9.0, -6.0, -5.99, 0.0, 3.99, 4.0, 4.01, 30.0])
    value = torch.randn(b * nq, 25, 25, 256, generator=g)
    e = torch.exp(torch.clamp(conf, max=4, min=-6)).view(b, nq, 25, 25, 256)
    ref = ((e / e.sum(dim=1, keepdim=True)) * value.view(b, nq, 25, 25, 256)).sum(dim=1)
    out = ops.conf_fusion(conf.cuda(), value.cuda(), nq)
    assert rel_err(out, ref) <= 2e-6


@pytest.mark.parametrize("n,r,size,sf", [(6, 25, 255, 25), (2, 27, 271, 27), (1, 25, 255, 25)])
def test_cycle_glue_vs_oracle(n, r, size, sf):
    """argmax (first index on ties), box gather and image->PrPool box map of models.py:131-162,262-274, incl. clamped boxes."""
    from usot_b200 import ops
    g = torch.Generator().manual_seed(n * 100 + r)
    off_cls, mem_cls = torch.randn(n, 1, r, r, generator=g), torch.randn(n, 1, r, r, generator=g)
    off_bbox = torch.rand(n, 4, r, r, generator=g) * 150.0    # large offsets: some boxes leave the search area and get clamped
    off_cls[0].fill_(0.5); mem_cls[0].fill_(0.5)              # all-ties map -> index 0
    ratio = 0.4
    res = ratio * off_cls.view(n, -1) + (1 - ratio) * mem_cls.view(n, -1)
    best = res.max(dim=1)
    to_img = O.pred_offset_to_image_bbox(off_bbox, r, size).view(n, 4, -1).transpose(1, 2)
    box = torch.gather(to_img, 1, best.indices.view(n, 1, 1).repeat(1, 1, 4)).view(n, 4)
    ref_box = O.image_bbox_to_prpool_bbox(box, sf, size)
    pb, score, idx = ops.cycle_glue(off_cls.cuda(), mem_cls.cuda(), off_bbox.cuda(), ratio, size, sf)
    assert torch.equal(idx.cpu().long(), best.indices)
    assert torch.allclose(score.cpu(), best.values, rtol=0, atol=1e-6)
    assert torch.allclose(pb.cpu(), ref_box.float(), rtol=1e-6, atol=1e-5)


def test_weighted_bce_vs_oracle_incl_quirks():
    """models.py:42-58: class means, the 0-d selection quirk (a class with exactly ONE member contributes 0) and soft labels ignored."""
    from usot_b200 import ops
    g = torch.Generator().manual_seed(3)
    pred = torch.randn(4, 25, 25, generator=g) * 3.0
    label = torch.zeros(4, 25, 25)
    label[:, 10:15, 10:15] = 1.0
    for lab in (label, torch.cat([label[:1] * 0, label[1:]]), ):
        ref = O.weighted_bce(pred, lab)
        out = ops.weighted_bce(pred.cuda(), lab.cuda())
        assert abs(float(out) - float(ref)) <= 2e-6 * abs(float(ref))
    one = torch.zeros(1, 25, 25)
    one[0, 3, 3] = 1.0                                        # exactly one positive: loss_pos == 0 in the reference
    ref = O.weighted_bce(pred[:1], one)
    out = ops.weighted_bce(pred[:1].cuda(), one.cuda())
    assert abs(float(out) - float(ref)) <= 2e-6 * abs(float(ref))
    big = torch.tensor([[-80.0, 80.0, 0.0, 30.0]])            # saturating logits stay finite
    lab = torch.tensor([[1.0, 0.0, 1.0, 0.0]])
    assert abs(float(ops.weighted_bce(big.cuda(), lab.cuda())) - float(O.weighted_bce(big, lab))) <= 1e-5 * float(O.weighted_bce(big, lab))


@pytest.mark.parametrize("n", [1, 4, 16])
def test_iou_loss_vs_oracle(n):
    from usot_b200 import ops
    g = torch.Generator().manual_seed(40 + n)
    bbox = torch.rand(n, 4, 25, 25, generator=g) * 60.0 + 0.5
    target = torch.rand(n, 25, 25, 4, generator=g) * 40.0 + 5.0
    weight = (torch.rand(n, 25, 25, generator=g) > 0.9).float()
    weight[0, 0, 0] = 1.0
    ref = O.iou_loss(bbox, target, weight)
    out = ops.iou_loss(bbox.cuda(), target.cuda(), weight.cuda())
    assert abs(float(out) - float(ref)) <= 3e-6 * abs(float(ref))


def test_torch_ops_dispatch_to_the_c_abi():
    """torch.ops.usot_b200.* (usot_b200/torch_ops.py): same results as the ctypes wrappers / the oracle."""
    import usot_b200.torch_ops  # noqa: F401
    from usot_b200 import ops
    g = torch.Generator().manual_seed(4)
    f = torch.randn(2, 8, 9, 9, generator=g)
    rois = torch.tensor([[0, 1.3, 0.7, 6.9, 7.2], [1, -1.0, 2.0, 4.5, 10.5]])
    out = torch.ops.usot_b200.prroi_pooling_forward(f.cuda(), rois.cuda(), 7, 7, 1.0)
    assert rel_err(out, O.prroi_pool2d(f, rois, 7, 7, 1.0)) <= 5e-6
    gout = torch.randn(2, 8, 7, 7, generator=g)
    gin = torch.ops.usot_b200.prroi_pooling_backward(f.cuda(), rois.cuda(), out, gout.cuda(), 7, 7, 1.0)
    assert rel_err(gin, O.prroi_pool2d_backward(gout, rois, f.shape)) <= 5e-6
    x, k = torch.randn(2, 8, 11, 11, generator=g), torch.randn(2, 8, 3, 3, generator=g)
    assert rel_err(torch.ops.usot_b200.xcorr_depthwise(x.cuda(), k.cuda()), O.xcorr_depthwise(x, k)) <= 2e-6
