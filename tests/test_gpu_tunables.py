"""Performance knobs must never change results: every alternative code path (TMA-store epilogue, TMA residual, tensor-core
stem, TMA GroupDW, N-tile cap) is compared bit-for-bit / to rounding against the default path on the same inputs."""
import pytest
import torch

import usot_oracle as O
from helpers import load_weights

pytestmark = pytest.mark.gpu


def _set(name, v):
    from usot_b200 import _lib
    _lib.check(_lib.load().usot_set_tunable(name.encode(), v))


DEFAULTS = {"tc_pdl": 1, "groupdw_row_split": 1, "tc_latency_split": 1, "tc_l2_prefetch": 0, "tc_tma_f32": 1, "tc_fuse_cross": 1, "tc_tma_store": 1, "tc_tma_res": 1, "stem_tc": 1, "groupdw_tma": 2, "pred_tma_min_batch": 48, "tc_bn_max": 256, "tc_split_bn_max": 128, "groupdw_warps4": 1, "conf_fusion_fused": 1, "tc_multi_image_tiles": 1, "tc_skip_pad_rows": 1, "tc_res_ahead": 1, "stem_pool_fused": 1, "graph_max_batch": 8}


@pytest.fixture()
def net():
    from usot_b200 import USOT
    n = USOT(precision="fp16x3")
    n.load_state_dict(load_weights("damp025"))
    yield n.eval().cuda()
    for k, v in DEFAULTS.items():
        _set(k, v)


@pytest.mark.parametrize("knob,value,exact", [("tc_fuse_cross", 0, False), ("tc_tma_f32", 0, True), ("tc_l2_prefetch", 1, True), ("tc_latency_split", 0, True), ("groupdw_row_split", 0, True), ("tc_pdl", 0, True), ("tc_tma_store", 0, True), ("tc_tma_res", 0, True), ("groupdw_tma", 0, True), ("groupdw_tma", 1, True), ("groupdw_warps4", 0, True), ("stem_pool_fused", 0, False), ("tc_multi_image_tiles", 0, True), ("tc_skip_pad_rows", 0, True), ("tc_res_ahead", 2, True), ("conf_fusion_fused", 0, True),
                                              ("stem_tc", 0, False), ("pred_tma_min_batch", 1, True), ("tc_bn_max", 64, False)])
def test_knob_keeps_results(net, knob, value, exact):
    z, x, tb, sb = O.synth_inputs(91, batch=3)
    net.template(z.cuda(), tb.cuda())
    ref = net.track(x.cuda())
    _set(knob, value)
    net.template(z.cuda(), tb.cuda())
    alt = net.track(x.cuda())
    for a, b in zip(ref[:2], alt[:2]):
        if exact:  # same arithmetic, only the data movement differs
            assert torch.equal(a, b)
        else:      # different summation order / fp32 CUDA-core stem: rounding-level differences only
            assert float((a - b).abs().max() / b.abs().max()) <= 2e-4
            assert torch.equal(a.flatten(1).argmax(1), b.flatten(1).argmax(1))


def test_cuda_graph_replay_matches_eager(net):
    """track() at small batch replays a captured CUDA graph from the 2nd call on; results must equal the eager path bit for bit,
    also when inputs / template / memory change between replays."""
    z, x, tb, sb = O.synth_inputs(92, batch=2)
    nq = 3
    net.template(z.cuda(), tb.cuda())
    mem = net.extract_memory_feature(ori_x=x[:1].repeat(2 * nq, 1, 1, 1).cuda(), search_bbox=sb[:1].repeat(2 * nq, 1).cuda())
    score = torch.full((2, nq), 0.9).cuda()
    _set("graph_max_batch", 0)
    eager = [t.clone() for t in net.track(x.cuda(), mem, score)]
    x2 = (x * 0.5 + 17.0).cuda()
    eager2 = [t.clone() for t in net.track(x2, mem, score)]
    _set("graph_max_batch", 8)
    for rep in range(4):  # call 1 eager, call 2 captures, calls 3-4 replay
        out = net.track(x.cuda(), mem, score)
        for a, b in zip(out, eager):
            assert torch.equal(a, b), f"replay {rep} differs"
    out2 = net.track(x2, mem, score)
    for a, b in zip(out2, eager2):
        assert torch.equal(a, b)


@pytest.mark.parametrize("precision", ["fp16x3", "fp16"])
@pytest.mark.parametrize("batch,nq", [(3, 3), (2, 7), (12, 1)])
def test_fused_conf_fusion_equals_two_convs_plus_reduction(precision, batch, nq):
    """conf_gen || value_gen as ONE conv with the Conf_Fusion reduction in its epilogue (conv_tc.cu EPI = 1; connect.py:123-144) builds
    the same sums in the same order as the two convs + conf_fusion_kernel: cls_mem must agree bit for bit."""
    from usot_b200 import USOT
    n = USOT(precision=precision)
    n.load_state_dict(load_weights("damp025"))
    n = n.eval().cuda()
    try:
        _set("graph_max_batch", 0)
        z, x, tb, sb = O.synth_inputs(93, batch=batch)
        n.template(z.cuda(), tb.cuda())
        src = O.synth_inputs(94, batch=4)
        feats = n.extract_memory_feature(ori_x=src[1].cuda(), search_bbox=src[3].cuda())
        pick = torch.tensor([(b * nq + q) * 3 % 4 for b in range(batch) for q in range(nq)]).cuda()
        mem = feats[pick].contiguous(memory_format=torch.channels_last)
        score = torch.full((batch, nq), 0.9).cuda()
        _set("conf_fusion_fused", 0)
        ref = [t.clone() for t in n.track(x.cuda(), mem, score)]
        _set("conf_fusion_fused", 1)
        out = n.track(x.cuda(), mem, score)
        assert torch.isfinite(out[2]).all()
        assert torch.equal(out[2], ref[2]), float((out[2] - ref[2]).abs().max() / ref[2].abs().max())
        assert torch.equal(out[0], ref[0]) and torch.equal(out[1], ref[1])
    finally:
        for k, v in DEFAULTS.items():
            _set(k, v)


@pytest.mark.parametrize("precision,tol", [("fp16x3", 2e-4), ("fp16", 3e-2)])
@pytest.mark.parametrize("batch,size", [(3, 255), (13, 255), (40, 255), (9, 127), (7, 271)])
def test_fused_stem_maxpool_feature_map_agrees_with_two_kernels(precision, tol, batch, size):
    """stem + maxpool as one TMA-fed tcgen05 kernel over the space-to-depth image (conv_tc.cu EPI = 2; modules.py:70-74,138-141) vs
    stem_tc_kernel + maxpool_kernel: same products, another summation order -> the neck feature map agrees to rounding (271-pixel crops,
    whose 133-wide map keeps the two kernels, agree exactly)."""
    from usot_b200.engine import Engine
    eng = Engine(torch.device("cuda", 0), precision)
    eng.load_state_dict(load_weights("damp025"))
    try:
        g = torch.Generator().manual_seed(500 + batch + size)
        x = (torch.rand(batch, 3, size, size, generator=g) * 255.0).cuda()
        _set("stem_pool_fused", 0)
        ref = eng.backbone_neck(x).clone()
        _set("stem_pool_fused", 1)
        out = eng.backbone_neck(x)
        assert torch.isfinite(out).all()
        err = float((out - ref).abs().max() / ref.abs().max())
        print("stem+pool fused vs two kernels", precision, batch, size, f"{err:.2e}")
        assert err <= tol
        if size > 261:
            assert torch.equal(out, ref)
    finally:
        for k, v in DEFAULTS.items():
            _set(k, v)
        eng.close()


@pytest.mark.parametrize("precision", ["fp16x3", "fp16"])
@pytest.mark.parametrize("batch", [4, 5, 13])
def test_multi_image_tiles_keep_results_bit_identical(precision, batch):
    """M tiles that span several images (one TMA box {64 ch, bw, bh, bimg}: a 31-wide map fills 124 of 128 MMA rows with ONE row of FOUR
    images instead of four rows of one) change only which pixels share a tile: every output element keeps its accumulation order, so
    track() with a memory queue must agree bit for bit with the one-image-per-tile plan, also when the batch is not a multiple of bimg."""
    from usot_b200 import USOT
    n = USOT(precision=precision)
    n.load_state_dict(load_weights("damp025"))
    n = n.eval().cuda()
    try:
        _set("graph_max_batch", 0)
        z, x, tb, sb = O.synth_inputs(95, batch=batch)
        n.template(z.cuda(), tb.cuda())
        nq = 2
        feats = n.extract_memory_feature(ori_x=x[:4].cuda(), search_bbox=sb[:4].cuda())
        pick = torch.tensor([(b * nq + q) % 4 for b in range(batch) for q in range(nq)]).cuda()
        mem = feats[pick].contiguous(memory_format=torch.channels_last)
        score = torch.full((batch, nq), 0.9).cuda()
        _set("tc_multi_image_tiles", 0)
        n.template(z.cuda(), tb.cuda())
        ref = [t.clone() for t in n.track(x.cuda(), mem, score)]
        _set("tc_multi_image_tiles", 1)
        n.template(z.cuda(), tb.cuda())
        out = n.track(x.cuda(), mem, score)
        for a, b in zip(out, ref):
            assert torch.isfinite(a).all() and torch.equal(a, b), float((a - b).abs().max())
    finally:
        for k, v in DEFAULTS.items():
            _set(k, v)
