"""End-to-end GPU test of the training path: ``USOT.forward`` with autograd on the B200 -- tcgen05 forward convs, dgrad
on the forward kernels, wgrad / BatchNorm / pooling / correlation / loss gradient kernels -- must reproduce the gradients of all 234
parameters of the LIVE reference (oracle/gen_grad_golden.py: scripts/train_usot.py:229-236's ``loss.backward()`` on the CPU at B=2, M=2),
with running-statistics BatchNorm (eval) and with batch statistics (train, what the reference trains in).  The bar and its rationale
(the reference's own float32 rounding noise on this ill-conditioned random-weight network) are in test_train_graph_cpu.check_against_golden."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, load_weights
from test_train_graph_cpu import _inputs, check_against_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp16x3", "fp32"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_all_234_parameter_gradients_vs_live_reference_fixture(mode, precision):
    from usot_b200 import USOT
    gold = np.load(os.path.join(GOLD, "grads_damp025.npz"))
    B, M = int(gold["B"]), int(gold["M"])
    net = USOT({"mem_size": M, "pr_pool": True}, precision=precision)
    net.load_state_dict(load_weights("damp025"), strict=True)
    net = net.cuda().train(mode == "train")
    z, x, tb, sb, smem, label, reg_target, reg_weight = [t.cuda() for t in _inputs(B, M)]
    # the reference's call (scripts/train_usot.py:196-199) and backward (:229-233)
    cls_loss, mem_loss, reg_loss = net(z, x, label, reg_target=reg_target, reg_weight=reg_weight, template_bbox=tb, search_memory=smem,
                                       search_bbox=sb, cls_ratio=0.4)
    loss = cls_loss + mem_loss + reg_loss
    assert loss.requires_grad
    loss.backward()
    torch.cuda.synchronize()
    check_against_golden(net, mode, (cls_loss, mem_loss, reg_loss), loss_rtol=1e-3, slack=2.0)


def test_optimizer_step_changes_the_forward_only_engine_too():
    """One SGD step on the training path, then the inference engine (track) must see the new weights (version counters bump -> re-pack)."""
    from usot_b200 import USOT
    import usot_oracle as O
    net = USOT({"mem_size": 2, "pr_pool": True}, precision="fp16x3")
    net.load_state_dict(load_weights("damp025"), strict=True)
    net = net.cuda().eval()
    z, x, tb, sb, smem, label, reg_target, reg_weight = [t.cuda() for t in _inputs(2, 2)]
    with torch.no_grad():
        net.template(z[:1], tb[:1])
        before = net.track(x)[0].clone()
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)
    losses = net(z, x, label, reg_target=reg_target, reg_weight=reg_weight, template_bbox=tb, search_memory=smem, search_bbox=sb)
    opt.zero_grad()
    (losses[0] + losses[1] + losses[2]).backward()
    opt.step()
    with torch.no_grad():
        net.template(z[:1], tb[:1])
        after = net.track(x)[0]
        l2 = net(z, x, label, reg_target=reg_target, reg_weight=reg_weight, template_bbox=tb, search_memory=smem, search_bbox=sb)   # engine path (no_grad + eval)
    assert not torch.equal(before, after)
    assert float(l2[0] + l2[1] + l2[2]) < float(losses[0] + losses[1] + losses[2])   # a small step along -grad lowers the loss


def test_cuda_graphed_training_step_matches_eager_steps():
    """usot_b200.dist.GraphedTrainStep: the whole step (forward + backward + SGD) captured once and replayed must follow the eager steps
    (same losses step by step; parameters agree to the rounding noise of the split-K atomics)."""
    from usot_b200 import USOT
    from usot_b200.dist import GradientReducer, GraphedTrainStep, train_step_sharded
    z, x, tb, sb, smem, label, reg_target, reg_weight = [t.cuda() for t in _inputs(2, 2)]
    batch = dict(template=z, search=x, search_memory=smem, label=label, reg_target=reg_target, reg_weight=reg_weight, template_bbox=tb, search_bbox=sb)

    def make():
        net = USOT({"mem_size": 2, "pr_pool": True}, precision="fp16x3")
        net.load_state_dict(load_weights("damp025"), strict=True)
        net = net.cuda().train()
        red = GradientReducer(net.parameters())
        opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9)
        return net, red, opt

    net_e, red_e, opt_e = make()
    eager = [torch.stack(train_step_sharded(net_e, batch, red_e, opt_e)).cpu() for _ in range(5)]
    net_g, red_g, opt_g = make()
    gstep = GraphedTrainStep(net_g, red_g, opt_g, batch, warmup=3)      # 3 eager warm-up steps + 1 captured (not executed) step
    assert sum(gstep.launches.values()) > 1000 and gstep.launches["conv_wgrad"] > 100    # launches of this library recorded into the graph
    graphed = [torch.stack(gstep(batch)).cpu().clone() for _ in range(2)]   # = steps 4 and 5
    for a, b in zip(eager[3:], graphed):
        assert torch.allclose(a, b, rtol=2e-4, atol=1e-6), (a, b)
    assert float(eager[4].sum()) < float(eager[0].sum())                  # the loss goes down over the five SGD steps
    pe, pg = dict(net_e.named_parameters()), dict(net_g.named_parameters())
    for k in ("connect_model.cls_pred.weight", "features.features.layer3.5.conv3.weight", "features.features.conv1.weight", "neck.downsample.1.weight"):
        # (the split-K weight-gradient atomics land in a run-dependent order: five SGD steps later the stem filter -- the end of the backward
        #  chain of this ill-conditioned random network -- differs by 0.7e-4 ... 1.2e-4 from run to run, the other tensors by < 1e-6)
        print("graphed vs eager parameter", k, f"{rel_err_t(pg[k], pe[k]):.3e}")
        assert rel_err_t(pg[k], pe[k]) <= (4e-4 if k.endswith("features.conv1.weight") else 1e-4), k


def rel_err_t(a, b):
    return float((a.detach() - b.detach()).abs().max() / b.detach().abs().max().clamp_min(1e-30))
