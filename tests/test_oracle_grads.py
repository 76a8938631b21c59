"""Backward of the training forward (SURVEY.md §8f-3 groundwork): the oracle differentiated by torch autograd must reproduce the
parameter gradients of the LIVE reference (oracle/gen_grad_golden.py -> tests/golden/grads_damp025.npz), in the running-statistics
BN regime usot_b200's forward implements and in the batch-statistics regime the reference trains in.  The CUDA training path of a
later round is checked against the same fixture."""
import os

import numpy as np
import pytest
import torch

import usot_oracle as O
from helpers import GOLD, load_weights, rel_err


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


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_oracle_gradients_match_reference_fixture(mode):
    gold = np.load(os.path.join(GOLD, "grads_damp025.npz"))
    B, M = int(gold["B"]), int(gold["M"])
    sd = load_weights("damp025")
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    z, x, tb, sb, smem, label, reg_target, reg_weight = _inputs(B, M)
    O._CAL.on = mode == "train"
    try:
        losses = O.forward_train(params, z, x, label, reg_target, reg_weight, tb, smem, sb, 0.4)
    finally:
        O._CAL.on = False
    assert np.allclose([float(v.detach()) for v in losses], gold[f"{mode}/losses"], rtol=2e-5)
    (losses[0] + losses[1] + losses[2]).backward()
    names = [k[len(mode) + 3:] for k in gold.files if k.startswith(f"{mode}/n:")]
    assert len(names) == 234
    worst = 0.0
    for k in names:
        g = params[k].grad.detach().double().flatten()
        step = max(1, g.numel() // 8)
        ref_norm, ref_s = float(gold[f"{mode}/n:{k}"][0]), gold[f"{mode}/s:{k}"]
        if ref_norm > 1e-10:
            worst = max(worst, abs(float(g.norm()) - ref_norm) / ref_norm)
            assert np.allclose(g[::step][:8].numpy(), ref_s, rtol=5e-3, atol=1e-4 * ref_norm), k
    assert worst <= 1e-3  # same autograd over the same ops; the slack covers thread-count dependent summation order


def test_prroi_backward_is_the_adjoint_of_the_forward():
    torch.manual_seed(0)
    f = torch.randn(2, 6, 9, 9)
    rois = torch.tensor([[0, 1.3, 0.7, 6.9, 7.2], [1, -1.0, 2.0, 4.5, 10.5], [0, 3.0, 3.0, 3.0, 5.0], [1, 0.0, 0.0, 8.0, 8.0]])
    g = torch.randn(4, 6, 7, 7)
    out = O.prroi_pool2d(f, rois)
    gf = O.prroi_pool2d_backward(g, rois, f.shape)
    lhs, rhs = float((out.double() * g.double()).sum()), float((f.double() * gf.double()).sum())
    assert abs(lhs - rhs) <= 1e-6 * abs(lhs)
    assert float(gf[:, :, :, :].abs().sum()) > 0 and float(out[2].abs().max()) == 0.0  # zero-area roi pools to 0 and gets no gradient
