"""GPU parity of the model façade (USOT.template / track / extract_memory_feature) against the oracle and the committed
golden fixtures of the live reference.  Bar (BASELINE.json north_star): <= 1e-3 relative (max-abs / max-abs(ref)), exact
argmax of the response maps."""
import numpy as np
import pytest
import torch

import usot_oracle as O
from helpers import golden, load_weights, rel_err, subsample

pytestmark = pytest.mark.gpu
TOL = 1e-3
PRECISIONS = ["fp32", "fp16x3", "fp16"]
# fp32 / fp16x3 are the parity-graded modes (north-star bar: 1e-3, exact argmax).  "fp16" is the single-pass fast mode
# (BASELINE config 3): validated against the same oracle with its own, looser, stated tolerance and no argmax claim.
MODE_TOL = {"fp32": TOL, "fp16x3": TOL, "fp16": {"damp025": 3e-2, "raw": 5e-1}}


def tol_for(precision, wname):
    t = MODE_TOL[precision]
    return t[wname] if isinstance(t, dict) else t


def _net(wname, precision, settings=None):
    from usot_b200 import USOT
    net = USOT(settings, precision=precision)
    net.load_state_dict(load_weights(wname), strict=True)
    return net.eval().cuda()


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("wname", ["damp025", "raw"])
def test_config1_pair_forward_vs_golden(wname, precision):
    net, g = _net(wname, precision), golden(wname)
    z, x, tb, sb = O.synth_inputs(7, batch=1)
    net.pr_pool = False
    net.template(z.cuda())
    cls, bbox, cmem, xf = net.track(x.cuda())
    assert cmem is None and xf is None
    assert tuple(net.zf.shape) == (1, 256, 7, 7) and tuple(cls.shape) == (1, 1, 25, 25) and tuple(bbox.shape) == (1, 4, 25, 25)
    errs = dict(zf=rel_err(net.zf, g["c1_zf"]), cls=rel_err(cls, g["c1_cls"]), bbox=rel_err(bbox, g["c1_bbox"]))
    print(wname, "c1", precision, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) <= tol_for(precision, wname), errs
    if precision != "fp16":
        assert int(cls.flatten().argmax()) == int(np.argmax(g["c1_cls"]))


@pytest.mark.parametrize("precision", PRECISIONS)
@pytest.mark.parametrize("wname,tag,S,B,seed", [("damp025", "m255", 255, 2, 21), ("raw", "m255", 255, 2, 21),
                                                ("damp025", "m271", 271, 1, 22), ("raw", "m271", 271, 1, 22)])
def test_track_with_memory_vs_golden(wname, tag, S, B, seed, precision):
    net, g = _net(wname, precision), golden(wname)
    z, x, tb, sb = O.synth_inputs(seed, batch=B, search_size=S)
    nq = 7
    mem_src = O.synth_inputs(31, batch=B * nq, search_size=S)[1]
    mem_box = torch.from_numpy(g[f"{tag}_mem_box"])
    # tracker call pattern (lib/tracker/usot_tracker.py:71,105-106,258-261,196-199)
    net.pr_pool = True
    net.template(z.cuda(), template_bbox=tb.cuda())
    mem = net.extract_memory_feature(ori_x=mem_src.cuda(), search_bbox=mem_box.cuda())
    mem_host = torch.cat([m.unsqueeze(0) for m in mem.cpu().detach()], dim=0)
    cls, bbox, cmem, xf = net.track(x.cuda(), template_mem=mem_host.cuda(), score_mem=torch.full((B, nq), 0.9).cuda())
    feat = net.extract_memory_feature(xf=xf, search_bbox=sb.cuda())
    F_ = 31 if S == 255 else 33
    assert tuple(xf.shape) == (B, 256, F_, F_) and tuple(feat.shape) == (B, 256, 7, 7)
    errs = dict(mem=rel_err(mem[:, ::8], g[f"{tag}_mem_sub"]), zf=rel_err(net.zf, g[f"{tag}_zf"]), cls=rel_err(cls, g[f"{tag}_cls"]),
                bbox=rel_err(bbox, g[f"{tag}_bbox"]), cls_mem=rel_err(cmem, g[f"{tag}_cls_mem"]),
                xf=rel_err(subsample(xf), g[f"{tag}_xf_sub"]), feat=rel_err(feat, g[f"{tag}_feat"]))
    print(wname, tag, precision, {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) <= tol_for(precision, wname), errs
    if precision == "fp16":
        return
    ratio = 0.3  # experiments/test/USOT.yaml:7
    mix_ref = ratio * torch.sigmoid(torch.from_numpy(g[f"{tag}_cls"])) + (1 - ratio) * torch.sigmoid(torch.from_numpy(g[f"{tag}_cls_mem"]))
    mix = ratio * torch.sigmoid(cls.cpu()) + (1 - ratio) * torch.sigmoid(cmem.cpu())
    for b in range(B):
        assert int(cls[b].flatten().argmax()) == int(np.argmax(g[f"{tag}_cls"][b]))
        assert int(cmem[b].flatten().argmax()) == int(np.argmax(g[f"{tag}_cls_mem"][b]))
        assert int(mix[b].flatten().argmax()) == int(mix_ref[b].flatten().argmax())


@pytest.mark.parametrize("precision", PRECISIONS)
def test_backbone_vs_oracle_fresh_inputs(precision):
    """Not only the fixtures: fresh seeded inputs compared with the oracle run on this box's CPU."""
    wname = "damp025"
    sd = load_weights(wname)
    net = _net(wname, precision)
    x = O.synth_inputs(1234, batch=3)[1]
    with torch.no_grad():
        ref = O.backbone_neck(sd, x)
    ours = net.backbone_neck(x.cuda())
    assert tuple(ours.shape) == (3, 256, 31, 31)
    err = rel_err(ours, ref)
    print("backbone fresh", precision, f"{err:.2e}")
    assert err <= tol_for(precision, wname)


@pytest.mark.parametrize("precision", PRECISIONS)
def test_batch_256_independence(precision):
    """BASELINE config 2 size: 256 crops through backbone + xcorr heads; every sample must equal the same sample run
    alone (crops are independent units -- the property that makes the multi-GPU sharding exact)."""
    net = _net("damp025", precision)
    z, x1, tb, sb = O.synth_inputs(77, batch=4)
    x = x1.repeat(64, 1, 1, 1)  # 256 crops, 4 distinct
    net.pr_pool = True
    net.template(z.cuda(), template_bbox=tb.cuda())
    cls, bbox, _, _ = net.track(x.cuda())
    assert tuple(cls.shape) == (256, 1, 25, 25)
    cls4, bbox4, _, _ = net.track(x1.cuda())
    for i in (0, 1, 2, 3, 100, 255):
        assert torch.equal(cls[i], cls4[i % 4]), "sample result depends on its position in the batch"
        assert torch.equal(bbox[i], bbox4[i % 4])
    assert torch.isfinite(cls).all() and torch.isfinite(bbox).all()


def test_track_before_template_and_bad_shapes():
    net = _net("damp025", "fp32")
    with pytest.raises(RuntimeError):
        net.track(torch.zeros(1, 3, 255, 255, device="cuda"))
    net.pr_pool = True
    with pytest.raises(ValueError):
        net.template(torch.zeros(1, 3, 127, 127, device="cuda"))


def test_weight_reload_is_picked_up():
    net = _net("damp025", "fp32")
    x = O.synth_inputs(5, batch=1)[1].cuda()
    a = net.backbone_neck(x).clone()
    with torch.no_grad():
        net.neck.downsample[1].bias.add_(1.0)
    b = net.backbone_neck(x)
    assert torch.allclose(b, a + 1.0, atol=1e-5)
