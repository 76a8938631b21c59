"""The oracle must reproduce the fixtures produced from the LIVE reference (oracle/gen_golden.py)."""
import numpy as np
import pytest
import torch

import usot_oracle as O
from helpers import golden, load_weights, rel_err, subsample

TOL = 2e-5  # oracle vs reference fixtures: same arithmetic, tolerance only covers thread-count dependent summation order


@pytest.mark.parametrize("wname", ["damp025", "raw"])
def test_config1_pair_forward(wname):
    sd, g = load_weights(wname), golden(wname)
    z, x, tb, sb = O.synth_inputs(7, batch=1)
    with torch.no_grad():
        zf = O.template(sd, z, pr_pool=False)
        cls, bbox, _, _ = O.track(sd, zf, x)
    assert rel_err(zf, g["c1_zf"]) <= TOL
    assert rel_err(cls, g["c1_cls"]) <= TOL
    assert rel_err(bbox, g["c1_bbox"]) <= TOL
    assert int(cls.argmax()) == int(np.argmax(g["c1_cls"]))


@pytest.mark.parametrize("wname,tag,S,B,seed", [("damp025", "m255", 255, 2, 21), ("raw", "m271", 271, 1, 22)])
def test_track_with_memory(wname, tag, S, B, seed):
    sd, g = load_weights(wname), golden(wname)
    z, x, tb, sb = O.synth_inputs(seed, batch=B, search_size=S)
    nq = 7
    mem_src = O.synth_inputs(31, batch=B * nq, search_size=S)[1]
    mem_box = torch.from_numpy(g[f"{tag}_mem_box"])
    with torch.no_grad():
        mem = O.extract_memory_feature(sd, ori_x=mem_src, search_bbox=mem_box)
        zf = O.template(sd, z, tb)
        cls, bbox, cls_mem, xf = O.track(sd, zf, x, mem, torch.full((B, nq), 0.9))
        feat = O.extract_memory_feature(sd, xf=xf, search_bbox=sb)
    assert rel_err(mem[:, ::8], g[f"{tag}_mem_sub"]) <= TOL
    assert rel_err(zf, g[f"{tag}_zf"]) <= TOL
    assert rel_err(cls, g[f"{tag}_cls"]) <= TOL
    assert rel_err(bbox, g[f"{tag}_bbox"]) <= TOL
    assert rel_err(cls_mem, g[f"{tag}_cls_mem"]) <= TOL
    assert rel_err(subsample(xf), g[f"{tag}_xf_sub"]) <= TOL
    assert rel_err(feat, g[f"{tag}_feat"]) <= TOL
    for b in range(B):
        assert int(cls[b].argmax()) == int(np.argmax(g[f"{tag}_cls"][b]))
        assert int(cls_mem[b].argmax()) == int(np.argmax(g[f"{tag}_cls_mem"][b]))


def test_prroi_properties():
    """Known-answer checks of the PrRoIPool restatement: a constant map pools to the constant wherever the bin lies inside
    the map, a bin fully outside pools to 0, and a zero-area roi gives 0 (prroi_pooling_gpu_impl.cu:189-193)."""
    feat = torch.full((1, 3, 15, 15), 2.5)
    rois = torch.tensor([[0, 2.0, 3.0, 9.0, 10.0], [0, 30.0, 30.0, 40.0, 40.0], [0, 4.0, 4.0, 4.0, 9.0]])
    out = O.prroi_pool2d(feat, rois, 7, 7, 1.0)
    assert torch.allclose(out[0], torch.full((3, 7, 7), 2.5), atol=1e-5)
    assert float(out[1].abs().max()) == 0.0
    assert float(out[2].abs().max()) == 0.0
    # linear ramp f(h,w) = w: the exact bin average is the bin centre
    ramp = torch.arange(15.0).view(1, 1, 1, 15).repeat(1, 1, 15, 1)
    out = O.prroi_pool2d(ramp, torch.tensor([[0, 1.0, 1.0, 8.0, 8.0]]), 7, 7, 1.0)
    centres = 1.0 + (torch.arange(7.0) + 0.5)
    assert torch.allclose(out[0, 0, 3], centres, atol=1e-5)


def test_tracker_update_matches_numpy_reference_formulae():
    g = torch.Generator().manual_seed(3)
    cls, cmem = torch.randn(1, 1, 25, 25, generator=g), torch.randn(1, 1, 25, 25, generator=g)
    bbox = torch.rand(1, 4, 25, 25, generator=g) * 30 + 5
    window = np.outer(np.hanning(25), np.hanning(25))
    r, c, pscore, penalty, mixed, box = O.tracker_update(cls, bbox, cmem, (60.0, 40.0), window)
    assert pscore.shape == (25, 25) and 0 <= r < 25 and 0 <= c < 25
    assert pscore[r, c] == pscore.max()
    assert box[2] > box[0] and box[3] > box[1]
