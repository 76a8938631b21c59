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


DEFAULTS = {"tc_pdl": 1, "groupdw_row_split": 1, "tc_latency_split": 1, "tc_l2_prefetch": 0, "tc_tma_f32": 1, "tc_fuse_cross": 1, "tc_tma_store": 1, "tc_tma_res": 1, "stem_tc": 1, "groupdw_tma": 2, "pred_tma_min_batch": 48, "tc_bn_max": 256, "tc_split_bn_max": 128, "groupdw_warps4": 1, "conf_fusion_fused": 1, "tc_multi_image_tiles": 1, "tc_skip_pad_rows": 1, "tc_res_ahead": 1, "stem_pool_fused": 1, "graph_max_batch": 8, "tc_cta_pair": 3}


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


PAIR_OPS = {  # name: (cin, cout, k, stride, pad, dil, h, residual, relu, n) -- n chosen so that every SM pair gets a pair tile
    "l3_down": (512, 1024, 3, 1, 1, 1, 31, False, False, 16),
    "l3_conv3_res": (256, 1024, 1, 1, 0, 1, 31, True, True, 16),
    "l3_conv2_dil2": (256, 256, 3, 1, 2, 2, 31, False, True, 24),
    "tower_odd_groups": (256, 256, 3, 1, 1, 1, 25, False, True, 25),
    "l2_conv2_s2": (128, 128, 3, 2, 1, 1, 63, False, True, 32),
    "ragged_last_group": (256, 256, 3, 1, 1, 1, 31, False, True, 30),
}


@pytest.mark.parametrize("precision", ["fp16x3", "fp16"])
@pytest.mark.parametrize("case", sorted(PAIR_OPS))
def test_cta_pair_conv_launches_are_bit_identical(precision, case):
    """tcgen05.mma.cta_group::2 (clusters of two CTAs, M = 256, each CTA staging half of the weight rows; conv_tc.cu PAIR = true) vs the
    one-CTA kernel: same K order and accumulators per output element, so the stand-alone conv op must agree bit for bit -- 3x3 / 1x1,
    stride 2, dilation, residual, an odd number of image groups (phantom partner tile) and a ragged last image group."""
    from usot_b200 import ops
    cin, cout, k, s, p, d, h, res, relu, n = PAIR_OPS[case]
    try:
        g = torch.Generator(device="cuda").manual_seed(len(case) * 131 + cin)
        x = torch.randn(n, h, h, cin, device="cuda", generator=g).relu_()
        w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
        sc = torch.rand(cout, device="cuda", generator=g) + 0.5
        sh = torch.randn(cout, device="cuda", generator=g) * 0.1
        ho = (h + 2 * p - d * (k - 1) - 1) // s + 1
        r = torch.randn(n, ho, ho, cout, device="cuda", generator=g) if res else None
        _set("tc_cta_pair", 0)
        y0 = ops.conv2d_nhwc(x, w, sc, sh, s, p, d, r, relu, precision)
        _set("tc_cta_pair", 7)
        y1 = ops.conv2d_nhwc(x, w, sc, sh, s, p, d, r, relu, precision)
        assert torch.isfinite(y1).all() and torch.equal(y0, y1), float((y0 - y1).abs().max())
    finally:
        for kk, v in DEFAULTS.items():
            _set(kk, v)


@pytest.mark.parametrize("precision", ["fp16x3", "fp16"])
@pytest.mark.parametrize("batch,nq", [(64, 0), (37, 0), (48, 3)])
def test_cta_pair_track_is_bit_identical(precision, batch, nq):
    """The whole track() call (split-plane TMA-store epilogue, TMA-prefetched residual, fp32 outputs) with every eligible conv launch
    running as CTA pairs vs none: identical outputs, with and without a memory queue, also when image-group counts are odd."""
    from usot_b200 import USOT
    net = USOT(precision=precision)
    net.load_state_dict(load_weights("damp025"))
    net = net.eval().cuda()
    try:
        _set("graph_max_batch", 0)
        z, x, tb, sb = O.synth_inputs(95, batch=2)
        xb = torch.cat([x * (1.0 + 0.01 * i) + i for i in range((batch + 1) // 2)])[:batch].cuda()
        outs = []
        for knob in (0, 7):
            _set("tc_cta_pair", knob)
            net.template(z[:1].cuda(), tb[:1].cuda())
            if nq:
                mem = net.extract_memory_feature(ori_x=xb[:1].repeat(batch * nq, 1, 1, 1), search_bbox=sb[:1].repeat(batch * nq, 1).cuda())
                out = net.track(xb, mem, torch.full((batch, nq), 0.9).cuda())
            else:
                out = net.track(xb)
            outs.append([t.clone() for t in out if torch.is_tensor(t)])
        for a, b in zip(*outs):
            assert torch.isfinite(a).all() and torch.equal(a, b), float((a - b).abs().max())
    finally:
        for kk, v in DEFAULTS.items():
            _set(kk, v)
