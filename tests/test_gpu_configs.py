"""Direct GPU-vs-oracle parity AT the sizes BASELINE.json states (VERDICT r01 Weak #1: the round-1 tests reached those sizes only
through a batch-independence argument).  The CPU oracle (pinned 0.0 against the live reference, tests/golden/PIN_REPORT.txt) runs
on sampled rows of the same batch on this box's host cores:

    config 2   track(x), batch 256, offline head                     8 sampled rows
    config 3   track(x, memory), batch 64, N_q = 7                   8 sampled rows
    config 4   cycle-memory training forward, batch 16, 3 memory frames   (whole batch: the losses are batch means)

Bar: fp32 / fp16x3 <= 1e-3 relative (max-abs / max-abs(ref)) with exact arg-max; the single-pass fp16 fast mode is run with
its own, looser, printed tolerance and no arg-max claim."""
import numpy as np
import pytest
import torch

import usot_oracle as O
from helpers import load_weights, rel_err

pytestmark = pytest.mark.gpu
TOL = {"fp32": 1e-3, "fp16x3": 1e-3, "fp16": 3e-2}   # weight set damp025 (the bench's weights)
ROWS_256 = [0, 1, 37, 100, 127, 128, 200, 255]
ROWS_64 = [0, 9, 17, 31, 32, 40, 55, 63]


def _net(precision, settings=None):
    from usot_b200 import USOT
    net = USOT(settings, precision=precision)
    net.load_state_dict(load_weights("damp025"), strict=True)
    return net.eval().cuda()


@pytest.mark.parametrize("precision", ["fp16x3", "fp16", "fp32"])
def test_config2_batch256_sampled_rows_vs_oracle(precision):
    """BASELINE configs[1] (the bench workload): 256 DISTINCT crops, one template, offline head."""
    sd = load_weights("damp025")
    net = _net(precision)
    z, x, tb, _ = O.synth_inputs(2024, batch=256)
    net.template(z.cuda(), tb.cuda())
    cls, bbox, _, _ = net.track(x.cuda())
    assert tuple(cls.shape) == (256, 1, 25, 25) and tuple(bbox.shape) == (256, 4, 25, 25)
    with torch.no_grad():
        zf = O.template(sd, z, tb)
        o_cls, o_bbox, _, _ = O.track(sd, zf, x[ROWS_256])
    errs = dict(cls=rel_err(cls[ROWS_256], o_cls), bbox=rel_err(bbox[ROWS_256], o_bbox))
    print(f"config2 B=256 {precision}: tolerance {TOL[precision]:.0e}", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) <= TOL[precision], errs
    if precision != "fp16":
        for j, r in enumerate(ROWS_256):
            assert int(cls[r].flatten().argmax()) == int(o_cls[j].flatten().argmax()), f"argmax differs at row {r}"


@pytest.mark.parametrize("precision", ["fp16x3", "fp16"])
def test_config3_batch64_nq7_sampled_rows_vs_oracle(precision):
    """BASELINE configs[2]: full USOT* head at batch 64 with a 7-entry memory queue per crop (448 memory templates)."""
    sd = load_weights("damp025")
    net = _net(precision)
    B, NQ = 64, 7
    z, x, tb, sb = O.synth_inputs(3033, batch=B)
    src = O.synth_inputs(3034, batch=16)[1]
    box = O.synth_inputs(3035, batch=16)[3]
    with torch.no_grad():
        feats = O.extract_memory_feature(sd, ori_x=src, search_bbox=box)          # 16 distinct (256,7,7) memory features
        zf = O.template(sd, z, tb)
    pick = torch.tensor([(b * NQ + q) * 5 % 16 for b in range(B) for q in range(NQ)])
    mem = feats[pick].contiguous()                                                # (B*NQ, 256, 7, 7), row b*NQ+q belongs to crop b
    score = torch.full((B, NQ), 0.9)
    net.template(z.cuda(), tb.cuda())
    cls, bbox, cmem, xf = net.track(x.cuda(), template_mem=mem.cuda(), score_mem=score.cuda())
    assert tuple(cmem.shape) == (B, 1, 25, 25) and tuple(xf.shape) == (B, 256, 31, 31)
    rows = torch.tensor(ROWS_64)
    mrows = (rows[:, None] * NQ + torch.arange(NQ)[None]).reshape(-1)
    with torch.no_grad():
        o_cls, o_bbox, o_mem, o_xf = O.track(sd, zf, x[rows], mem[mrows], score[rows])
    errs = dict(cls=rel_err(cls[ROWS_64], o_cls), bbox=rel_err(bbox[ROWS_64], o_bbox), cls_mem=rel_err(cmem[ROWS_64], o_mem),
                xf=rel_err(xf[ROWS_64], o_xf))
    print(f"config3 B=64 Nq=7 {precision}: tolerance {TOL[precision]:.0e}", {k: f"{v:.2e}" for k, v in errs.items()})
    assert max(errs.values()) <= TOL[precision], errs
    if precision != "fp16":
        for j, r in enumerate(ROWS_64):
            assert int(cls[r].flatten().argmax()) == int(o_cls[j].flatten().argmax())
            assert int(cmem[r].flatten().argmax()) == int(o_mem[j].flatten().argmax())


def _train_batch(B, M, seed):
    z, x, tb, sb = O.synth_inputs(seed, batch=B, n_templates=B)
    g = torch.Generator().manual_seed(seed + 1)
    smem = torch.rand(B, M, 3, 255, 255, generator=g) * 255.0
    label = torch.zeros(B, 25, 25)
    label[:, 10:15, 10:15] = 1.0
    reg_weight = torch.zeros(B, 25, 25)
    reg_weight[:, 11:14, 11:14] = 1.0
    reg_target = torch.rand(B, 25, 25, 4, generator=g) * 40.0 + 5.0
    return dict(template=z, search=x, search_memory=smem, label=label, reg_target=reg_target, reg_weight=reg_weight, template_bbox=tb,
                search_bbox=sb)


@pytest.mark.parametrize("precision", ["fp16x3"])
def test_config4_cycle_forward_batch16_m3_vs_oracle(precision):
    """BASELINE configs[3] per-GPU shard: cycle-memory forward at batch 16 with 3 memory frames (running-statistics BN)."""
    sd = load_weights("damp025")
    net = _net(precision, {"mem_size": 3, "pr_pool": True})
    b = _train_batch(16, 3, 4040)
    with torch.no_grad():
        d = O.forward_train(sd, b["template"], b["search"], b["label"], b["reg_target"], b["reg_weight"], b["template_bbox"],
                            b["search_memory"], b["search_bbox"], 0.4, detail=True)
    cu = {k: v.cuda() for k, v in b.items()}
    with torch.no_grad():
        eng = net._engine()
        zf, _ = eng.template(cu["template"], cu["template_bbox"])
        xf = eng.backbone_neck(cu["search"])
        xf_mem = eng.backbone_neck(cu["search_memory"].reshape(-1, 3, 255, 255))
        losses, back, pbox = eng.forward_train_heads(zf, xf, xf_mem, 3, cu["label"], cu["reg_target"], cu["reg_weight"], cu["search_bbox"], 0.4,
                                                     want_aux=True)
    ours = losses.cpu().numpy()
    ref = np.array([float(d["cls_loss"]), float(d["cls_memory_loss"]), float(d["reg_loss"])])
    print("config4 B=16 M=3", precision, "losses", ours, "oracle", ref)
    assert np.all(np.abs(ours - ref) <= 1e-3 * np.abs(ref))
    assert rel_err(pbox, d["pool_box"]) <= 1e-3
    assert rel_err(back, d["backward_map"]) <= 1e-3
    for i in range(16):
        assert int(back[i].flatten().argmax()) == int(d["backward_map"][i].flatten().argmax())


def test_graph_replay_sees_reloaded_weights():
    """a captured track() graph must not survive a re-pack of the weights.  track twice at batch 1 (eager, then
    capture + replay), load DIFFERENT weights into the same model (same engine), track again: the result must equal a fresh
    engine's and differ from the old one."""
    from usot_b200 import USOT
    net = _net("fp16x3")
    z, x, tb, _ = O.synth_inputs(5150, batch=1)
    net.template(z.cuda(), tb.cuda())
    outs = [net.track(x.cuda())[0].clone() for _ in range(3)]       # eager, capture, replay
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2])
    sd2 = load_weights("raw")
    junk = [torch.empty(64 << 20, device="cuda") for _ in range(4)]   # perturb the allocator so freed weight addresses are not simply reused
    net.load_state_dict(sd2, strict=True)
    net.template(z.cuda(), tb.cuda())
    a = net.track(x.cuda())[0].clone()
    b = net.track(x.cuda())[0].clone()
    del junk
    fresh = USOT(precision="fp16x3")
    fresh.load_state_dict(sd2, strict=True)
    fresh = fresh.eval().cuda()
    fresh.template(z.cuda(), tb.cuda())
    ref = fresh.track(x.cuda())[0]
    assert torch.equal(a, ref) and torch.equal(b, ref)
    assert not torch.equal(a, outs[0])


def test_data_edits_need_invalidate_and_get_it():
    """writes through .data do not bump version counters; invalidate() forces the re-pack."""
    net = _net("fp32")
    x = O.synth_inputs(5, batch=1)[1].cuda()
    a = net.backbone_neck(x).clone()
    net.neck.downsample[1].bias.data.add_(1.0)
    net.invalidate()
    b = net.backbone_neck(x)
    assert torch.allclose(b, a + 1.0, atol=1e-5)
    # re-assigning a Parameter object is caught by invalidate() as well
    net.neck.downsample[1].bias = torch.nn.Parameter(net.neck.downsample[1].bias.detach() + 1.0)
    net.invalidate()
    c = net.backbone_neck(x)
    assert torch.allclose(c, a + 2.0, atol=1e-5)


def test_calls_from_two_streams_are_ordered_on_the_device():
    """the engine's arena is shared by consecutive calls; calls issued on different streams are ordered by an
    event, so interleaving two streams gives the same results as one stream."""
    net = _net("fp16x3")
    z, x, tb, _ = O.synth_inputs(616, batch=12)
    net.template(z.cuda(), tb.cuda())
    xa, xb = x[:12].cuda(), torch.flip(x[:12], dims=[3]).cuda()
    ref_a, ref_b = net.track(xa)[0].clone(), net.track(xb)[0].clone()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    outs = []
    for i in range(6):
        with torch.cuda.stream(s1 if i % 2 == 0 else s2):
            outs.append(net.track(xa if i % 2 == 0 else xb)[0])
    torch.cuda.synchronize()
    for i, o in enumerate(outs):
        assert torch.equal(o, ref_a if i % 2 == 0 else ref_b)
