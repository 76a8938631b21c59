"""GPU parity of the cycle-memory training forward (USOT.forward, lib/models/models.py:208-295) and the tracker tensor path."""
import numpy as np
import pytest
import torch

import usot_oracle as O
from helpers import golden, load_weights, rel_err

pytestmark = pytest.mark.gpu


def _net(wname, precision, settings=None):
    from usot_b200 import USOT
    net = USOT(settings, precision=precision)
    net.load_state_dict(load_weights(wname), strict=True)
    return net.eval().cuda()


def _train_batch(B=2, M=3):
    z, x, tb, sb = O.synth_inputs(41, batch=B, n_templates=B)
    g = torch.Generator().manual_seed(42)
    smem = torch.rand(B, M, 3, 255, 255, generator=g) * 255.0
    label = torch.zeros(B, 25, 25)
    label[:, 10:15, 10:15] = 1.0
    reg_weight = torch.zeros(B, 25, 25)
    reg_weight[:, 11:14, 11:14] = 1.0
    reg_target = torch.rand(B, 25, 25, 4, generator=g) * 40.0 + 5.0
    return dict(template=z, search=x, search_memory=smem, label=label, reg_target=reg_target, reg_weight=reg_weight, template_bbox=tb,
                search_bbox=sb)


@pytest.mark.parametrize("precision", ["fp32", "fp16x3"])
def test_cycle_forward_vs_golden_and_oracle(precision):
    net = _net("damp025", precision, {"mem_size": 3, "pr_pool": True})
    g = golden("damp025")
    b = _train_batch()
    cu = {k: v.cuda() for k, v in b.items()}
    cls_loss, mem_loss, reg_loss = net(cu["template"], cu["search"], label=cu["label"], reg_target=cu["reg_target"], reg_weight=cu["reg_weight"],
                                       template_bbox=cu["template_bbox"], search_memory=cu["search_memory"], search_bbox=cu["search_bbox"],
                                       cls_ratio=0.4)
    ours = np.array([float(cls_loss), float(mem_loss), float(reg_loss)])
    ref = g["train_losses"]  # produced by the live reference module (oracle/gen_golden.py)
    print(precision, "losses", ours, "ref", ref)
    assert np.all(np.abs(ours - ref) / np.abs(ref) <= 1e-3)
    # auxiliary outputs: forward-tracked PrPool boxes and the backward response map
    eng = net._engine()
    zf, _ = eng.template(cu["template"], cu["template_bbox"])
    xf = eng.backbone_neck(cu["search"])
    xf_mem = eng.backbone_neck(cu["search_memory"].reshape(-1, 3, 255, 255))
    losses, back, pbox = eng.forward_train_heads(zf, xf, xf_mem, 3, cu["label"], cu["reg_target"], cu["reg_weight"], cu["search_bbox"], 0.4,
                                                 want_aux=True)
    assert rel_err(pbox, g["train_pool_box"]) <= 1e-3
    assert rel_err(back, g["train_backward_map"]) <= 1e-3
    assert int(back[0].flatten().argmax()) == int(np.argmax(g["train_backward_map"][0]))


def test_naive_siamese_branch_and_loss_edge_cases():
    """search_memory=None -> (cls_loss, None, reg_loss) (models.py:288-295); a class with exactly one member contributes 0."""
    sd = load_weights("damp025")
    net = _net("damp025", "fp32")
    b = _train_batch(B=2)
    b["label"][:] = 0.5          # neither positive nor negative ...
    b["label"][0, 3, 3] = 1.0    # ... except ONE positive (-> loss_pos = 0, quirk E5) and two negatives
    b["label"][1, 5, 5] = 0.0
    b["label"][1, 6, 6] = 0.0
    cu = {k: v.cuda() for k, v in b.items()}
    cls_loss, none, reg_loss = net(cu["template"], cu["search"], label=cu["label"], reg_target=cu["reg_target"], reg_weight=cu["reg_weight"],
                                   template_bbox=cu["template_bbox"])
    assert none is None
    with torch.no_grad():
        o = O.forward_train(sd, b["template"], b["search"], b["label"], b["reg_target"], b["reg_weight"], b["template_bbox"])
    assert abs(float(cls_loss) - float(o[0])) <= 1e-4 * max(1.0, abs(float(o[0])))
    assert abs(float(reg_loss) - float(o[2])) <= 1e-4 * abs(float(o[2]))


@pytest.mark.parametrize("R,S", [(25, 255), (27, 271)])
def test_tracker_postprocess_vs_numpy_reference(R, S):
    from usot_b200 import tracker_ops
    g = torch.Generator().manual_seed(R)
    for trial in range(5):
        cls, cmem = torch.randn(1, 1, R, R, generator=g), torch.randn(1, 1, R, R, generator=g)
        bbox = torch.rand(1, 4, R, R, generator=g) * 30 + 5
        tsz = (40.0 + 10 * trial, 55.0 - 5 * trial)
        window = np.outer(np.hanning(R), np.hanning(R))
        r, c, pscore, penalty, mixed, box = O.tracker_update(cls, bbox, cmem, tsz, window, instance_size=S, score_size=R)
        out = tracker_ops.postprocess(cls.cuda(), bbox.cuda(), cmem.cuda(), tracker_ops.cosine_window(R, "cuda"), tsz, instance_size=S).cpu().numpy()
        assert (int(out[0]), int(out[1])) == (r, c)
        assert np.allclose(out[2:6], box, rtol=1e-6, atol=1e-6)
        assert abs(out[6] - penalty[r, c]) <= 1e-6 and abs(out[7] - mixed[r, c]) <= 1e-6


def test_device_tracker_update_matches_reference_formulae():
    """usot_b200.tracker_ops.update_device (device queue + fused post-process) against a host restatement of
    USOTTracker.update (lib/tracker/usot_tracker.py:133-200) driven by the oracle's model outputs."""
    import types
    from usot_b200 import tracker_ops
    sd = load_weights("damp025")
    net = _net("damp025", "fp16x3")
    z, x, tb, sb = O.synth_inputs(55, batch=1)
    net.template(z.cuda(), tb.cuda())
    f0 = net.extract_memory_feature(ori_x=x.cuda(), search_bbox=sb.cuda())
    f1 = net.extract_memory_feature(ori_x=torch.flip(x, dims=[3]).cuda(), search_bbox=sb.cuda())
    q = tracker_ops.MemoryQueue([f0, f1], mem_queue_size=7)
    p = types.SimpleNamespace(instance_size=255, score_size=25, total_stride=8, ratio=0.3, penalty_k=0.021, window_influence=0.321, lr=0.730)
    target_pos, target_sz, scale_z = np.array([320.0, 240.0]), np.array([80.0, 60.0]), 0.9
    window = tracker_ops.cosine_window(25, "cuda")
    pos, sz, conf, feat = tracker_ops.update_device(net, x.cuda(), target_pos, target_sz * scale_z, window, scale_z, p, q)
    # reference formulae on the host with the oracle's maps
    with torch.no_grad():
        zf = O.template(sd, z, tb)
        mem, _ = q.select()
        cls, bbox, cmem, xf = O.track(sd, zf, x, mem.cpu().contiguous(), torch.full((1, 7), 0.9))
    r, c, pscore, penalty, mixed, box = O.tracker_update(cls, bbox, cmem, target_sz * scale_z, np.outer(np.hanning(25), np.hanning(25)))
    pw, ph = (box[2] - box[0]) / scale_z, (box[3] - box[1]) / scale_z
    lr = penalty[r, c] * mixed[r, c] * p.lr
    ref_pos = target_pos + np.array([(box[0] + box[2]) / 2 - 127, (box[1] + box[3]) / 2 - 127]) / scale_z
    tsz = target_sz * scale_z / scale_z
    ref_sz = tsz * (1 - lr) + lr * np.array([pw * lr + (1 - lr) * tsz[0], ph * lr + (1 - lr) * tsz[1]])
    assert np.allclose(pos, ref_pos, rtol=1e-4, atol=1e-3) and np.allclose(sz, ref_sz, rtol=1e-4, atol=1e-3)
    assert abs(conf - mixed[r, c]) <= 1e-4 and tuple(feat.shape) == (1, 256, 7, 7)
