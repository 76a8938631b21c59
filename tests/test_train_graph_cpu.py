"""The Python side of the training path (usot_b200/train.py: graph wiring, train / eval BatchNorm, running-statistics updates, the
stride-2 dgrad decomposition, layouts, detach points) checked WITHOUT a GPU: tests/train_ref.py stands in for the CUDA library and
``USOT.forward(...)`` + ``.backward()`` must reproduce the 234 parameter gradients of the LIVE reference
(tests/golden/grads_damp025.npz, written by oracle/gen_grad_golden.py) in both BatchNorm regimes.  The CUDA operators themselves
are compared with the same restatements in tests/test_gpu_train_ops.py, and the whole path end to end on the B200 in
tests/test_gpu_train_backward.py."""
import os

import numpy as np
import pytest
import torch

import train_ref
import usot_oracle as O
from helpers import GOLD, load_weights


def _inputs(B, M):
    z, x, tb, sb = O.synth_inputs(51, batch=B, n_templates=B)
    g = torch.Generator().manual_seed(52)
    smem = torch.rand(B, M, 3, 255, 255, generator=g) * 255.0
    label = torch.zeros(B, 25, 25)
    label[:, 10:15, 10:15] = 1.0
    reg_weight = torch.zeros(B, 25, 25)
    reg_weight[:, 11:14, 11:14] = 1.0
    reg_target = torch.rand(B, 25, 25, 4, generator=g) * 40.0 + 5.0
    return z, x, tb, sb, smem, label, reg_target, reg_weight


def check_against_golden(net, mode, losses, loss_rtol=2e-4, slack=1.5):
    """Shared with the GPU test.  Two fixtures written by oracle/gen_grad_golden.py from the LIVE reference:

    grads_damp025.npz    float32 reference gradients: L2 norm + 8 samples per parameter (234 parameters);
    grads64_damp025.npz  64 samples per parameter of the float32 reference gradients AND of the same algorithm in float64.

    The random-weight network is ill-conditioned on purpose, and the float32 reference's own gradients sit 6.2e-3 (eval BN) / 3.4e-3
    (train BN) away from exact arithmetic (global relative L2; stored in the fixture).  That rounding noise is the floor for any float32-class
    implementation, so the bar is: OUR distance to the exact gradients may not exceed ``slack`` x the REFERENCE's distance to them
    (+1e-3), measured on the same samples -- globally and per parameter tensor (per tensor: with a wider factor, single tensors are noisier);
    and losses / gradient norms agree with the float32 reference to a few 1e-3."""
    gold = np.load(os.path.join(GOLD, "grads_damp025.npz"))
    g64 = np.load(os.path.join(GOLD, "grads64_damp025.npz"))
    assert np.allclose([float(v.detach()) for v in losses], gold[f"{mode}/losses"], rtol=loss_rtol), (losses, gold[f"{mode}/losses"])
    names = [k[len(mode) + 3:] for k in gold.files if k.startswith(f"{mode}/n:")]
    assert len(names) == 234
    params = dict(net.named_parameters())
    num_o = num_r = den = 0.0
    worst_norm, worst_t, n_checked = 0.0, (0.0, ""), 0
    for k in names:
        assert params[k].grad is not None, f"no gradient for {k}"
        g = params[k].grad.detach().double().cpu().flatten()
        ex, r32 = g64[f"{mode}/exact/s:{k}"], g64[f"{mode}/ref32/s:{k}"]
        ex_norm, ref_norm = float(g64[f"{mode}/exact/n:{k}"][0]), float(gold[f"{mode}/n:{k}"][0])
        step = max(1, g.numel() // 64)
        ours = g[::step][:64].numpy()
        if ex_norm <= 1e-6:   # constants in front of a train-mode BatchNorm (conv biases, the neck's beta): 0 in exact arithmetic, rounding noise in float32
            assert float(g.norm()) <= 1e-5, (k, float(g.norm()), ref_norm)
            continue
        n_checked += 1
        e_o, e_r, d = float(((ours - ex) ** 2).sum()), float(((r32 - ex) ** 2).sum()), float((ex ** 2).sum())
        num_o, num_r, den = num_o + e_o, num_r + e_r, den + d
        norm_err = abs(float(g.norm()) - ex_norm) / ex_norm
        worst_norm = max(worst_norm, norm_err)
        rel_t, rel_r = np.sqrt(e_o / max(d, 1e-300)), np.sqrt(e_r / max(d, 1e-300))
        if rel_t > worst_t[0]:
            worst_t = (rel_t, k)
        assert norm_err <= 2e-2, (k, float(g.norm()), ex_norm)
        assert rel_t <= max(4.0 * rel_r, 3e-2), (k, rel_t, rel_r)
    assert n_checked >= 200
    ours_g, ref_g = np.sqrt(num_o / den), np.sqrt(num_r / den)
    print(f"{mode}: gradient vs float64 arithmetic, global relative L2 on the fixture samples: ours {ours_g:.3e}, float32 reference {ref_g:.3e}; "
          f"worst per-parameter norm error {worst_norm:.2e}; worst per-parameter relative L2 {worst_t[0]:.2e} ({worst_t[1]})")
    assert ours_g <= slack * ref_g + 1e-3, (ours_g, ref_g)
    return ours_g, ref_g


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_training_graph_reproduces_reference_gradients_on_cpu(mode):
    from usot_b200 import USOT
    gold = np.load(os.path.join(GOLD, "grads_damp025.npz"))
    B, M = int(gold["B"]), int(gold["M"])
    net = USOT({"mem_size": M, "pr_pool": True}, precision="fp32")
    net.load_state_dict(load_weights("damp025"), strict=True)
    net.train(mode == "train")
    stats_before = {k: v.clone() for k, v in net.state_dict().items() if k.endswith("running_mean")}
    z, x, tb, sb, smem, label, reg_target, reg_weight = _inputs(B, M)
    with train_ref.install():
        from usot_b200 import train
        losses = train.forward_train(net, z, x, label, reg_target, reg_weight, tb, smem, sb, 0.4)
        (losses[0] + losses[1] + losses[2]).backward()
    check_against_golden(net, mode, losses)
    moved = sum(int(not torch.equal(v, net.state_dict()[k])) for k, v in stats_before.items())
    assert moved == (len(stats_before) if mode == "train" else 0)   # train(): every BatchNorm updated its running statistics; eval(): none


def test_stride2_dgrad_decomposition_matches_autograd():
    """conv_dgrad for stride 2 = four parity sub-problems on the forward kernel (here: its CPU stand-in)."""
    from usot_b200 import train
    g = torch.Generator().manual_seed(5)
    for (h, w, k, pad) in ((63, 63, 3, 0), (15, 14, 3, 1), (9, 9, 1, 0), (12, 11, 5, 2)):
        x = torch.randn(2, 64, h, w, generator=g, requires_grad=True)
        wt = torch.randn(64, 64, k, k, generator=g)
        y = torch.nn.functional.conv2d(x, wt, None, 2, pad)
        gy = torch.randn(y.shape, generator=g)
        (ref,) = torch.autograd.grad(y, x, gy)
        with train_ref.install():
            out = train.conv_dgrad(gy.permute(0, 2, 3, 1).contiguous(), wt, (h, w), 2, pad, 1, precision="fp32")
        assert torch.allclose(out.permute(0, 3, 1, 2), ref, atol=1e-4, rtol=1e-4), (h, w, k, pad)


def test_tensor_core_backward_route_prescales_and_unscales_exactly():
    """In the tcgen05 modes conv_dgrad multiplies the gradient by a device-chosen power of two and divides it out in the conv epilogue:
    the route (here on the CPU stand-ins) must give the same result as the unscaled fp32 route, to rounding."""
    from usot_b200 import train
    g = torch.Generator().manual_seed(6)
    wt = torch.randn(64, 128, 3, 3, generator=g) * 0.1
    gy = torch.randn(2, 9, 9, 64, generator=g) * 1e-7
    with train_ref.install():
        a = train.conv_dgrad(gy, wt, (9, 9), 1, 1, 1, precision="fp32")
        b = train.conv_dgrad(gy, wt, (9, 9), 1, 1, 1, precision="fp16x3")
        c = train.conv_dgrad(gy[:, :4, :4].contiguous(), wt, (9, 9), 2, 0, 1, precision="fp16x3")
        d = train.conv_dgrad(gy[:, :4, :4].contiguous(), wt, (9, 9), 2, 0, 1, precision="fp32")
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-14) and torch.allclose(c, d, rtol=1e-5, atol=1e-14)


def test_naive_siamese_branch_and_partly_frozen_backbone_follow_torch_semantics():
    """search_memory=None (models.py:288-295) with the reference's freezing recipe (scripts/train_usot.py:74-102): layer1 / layer2 / stem frozen
    (requires_grad False, their BatchNorms in eval()), everything else in train().  Gradients must match the oracle run the same way, frozen
    parameters must get none, and only the train()-mode BatchNorms may update their running statistics."""
    from usot_b200 import USOT
    sd = load_weights("damp025")
    net = USOT({"mem_size": 2, "pr_pool": True}, precision="fp32")
    net.load_state_dict(sd, strict=True)
    net.train()
    frozen = ("features.features.conv1", "features.features.bn1", "features.features.layer1", "features.features.layer2")
    for k, p in net.named_parameters():
        if k.startswith(frozen):
            p.requires_grad_(False)
    for k, m in net.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d) and k.startswith(frozen):
            m.eval()
    z, x, tb, sb, smem, label, reg_target, reg_weight = _inputs(2, 2)
    before = {k: v.clone() for k, v in net.state_dict().items() if k.endswith("running_mean")}
    with train_ref.install():
        from usot_b200 import train
        cls_loss, none, reg_loss = train.forward_train(net, z, x, label, reg_target, reg_weight, tb)   # (USOT.forward itself refuses CPU tensors)
        assert none is None
        (cls_loss + reg_loss).backward()
    # oracle with the same mixed BatchNorm modes: batch statistics only where the module is in train()
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")) and not k.startswith(frozen))
              for k, v in sd.items()}
    orig_bn = O._bn

    def mixed_bn(sd_, t, p):
        O._CAL.on = not p.startswith(frozen)
        try:
            return orig_bn(sd_, t, p)
        finally:
            O._CAL.on = False

    O._bn = mixed_bn
    try:
        lo = O.forward_train(params, z, x, label, reg_target, reg_weight, tb)
    finally:
        O._bn = orig_bn
    (lo[0] + lo[2]).backward()
    assert abs(float(cls_loss) - float(lo[0])) <= 1e-5 * abs(float(lo[0])) and abs(float(reg_loss) - float(lo[2])) <= 1e-5 * abs(float(lo[2]))
    checked = 0
    for k, p in net.named_parameters():
        if k.startswith(frozen):
            assert p.grad is None
            continue
        ref = params[k].grad
        if ref is None or float(ref.abs().max()) < 1e-7:
            continue
        # (the three GroupDW weights sit behind a softmax: their gradient is a difference of large, nearly equal sums -- rounding-noise dominated)
        tol = 0.15 if k.endswith("_dw.weight") else 1e-2
        assert abs(float(p.grad.norm()) - float(ref.norm())) <= tol * float(ref.norm()), k
        checked += 1
    assert checked > 100
    moved = {k for k, v in before.items() if not torch.equal(v, net.state_dict()[k])}
    assert moved and all(not k.startswith(frozen) for k in moved)
    assert all((k in moved) for k in before if k.startswith("features.features.layer3"))
