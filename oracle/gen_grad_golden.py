"""Pin the BACKWARD of the training forward against the LIVE reference and write tests/golden/grads_damp025.npz.
TEST INFRASTRUCTURE for SURVEY.md §8f-3 (conv dgrad / wgrad, BN backward, xcorr / PrRoIPool / loss gradients): the CUDA training
path of a later round is checked against these parameter gradients.  Run in the build container only (needs /root/reference):

    python oracle/gen_grad_golden.py

The unmodified reference ``USOT.forward`` (lib/models/models.py:208-295) runs the cycle-memory forward at B=2, M=2 on the CPU and
``(cls_loss + cls_memory_loss + reg_loss).backward()`` produces the gradient of every parameter, in two BN regimes:
  eval   running statistics (what usot_b200's forward implements), and
  train  batch statistics per feature_extractor / head call (what scripts/train_usot.py runs).
Harness-side stand-ins as in gen_golden.py: ``.cuda()`` no-ops and the reference's GPU-only PrRoIPool symbol replaced by the
oracle's autograd-capable restatement (forward :149-212, backward :214-272 of prroi_pooling_gpu_impl.cu; the backward is
adjoint-tested against the forward here and its CUDA counterpart is pinned on the GPU box against the reference .cu).  The oracle's
functional forward differentiated by torch autograd must reproduce the reference gradients; per-parameter L2 norms and eight
strided samples per tensor are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("USOT_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(1, REF)

import usot_oracle as O  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self

import lib.models.models as ref_models  # noqa: E402
import lib.models.prroi_pool.prroi_pool as ref_prroi_mod  # noqa: E402


def _ref_prroi(f, r, ph, pw, s):
    assert (ph, pw, s) == (7, 7, 1.0)
    return O._PrRoIPoolFn.apply(f, r) if f.requires_grad else O.prroi_pool2d(f, r, ph, pw, s)


ref_models.prroi_pool2d = _ref_prroi
ref_prroi_mod.prroi_pool2d = _ref_prroi

GOLD = os.path.join(ROOT, "tests", "golden")
B, M = 2, 2


def load_weights(name="damp025", seed=11, damp=0.25):
    sd = O.make_state_dict(seed, damp)
    st = np.load(os.path.join(GOLD, f"bnstats_{name}.npz"))
    for k in O.bn_stat_keys(sd):
        sd[k] = torch.from_numpy(st[k].copy())
    return sd


def inputs():
    z, x, tb, sb = O.synth_inputs(51, batch=B, n_templates=B)
    g = torch.Generator().manual_seed(52)
    smem = torch.rand(B, M, 3, 255, 255, generator=g) * 255.0
    label = torch.zeros(B, 25, 25)
    label[:, 10:15, 10:15] = 1.0
    reg_weight = torch.zeros(B, 25, 25)
    reg_weight[:, 11:14, 11:14] = 1.0
    reg_target = torch.rand(B, 25, 25, 4, generator=g) * 40.0 + 5.0
    return z, x, tb, sb, smem, label, reg_target, reg_weight


def grad_summary(named_grads):
    out = {}
    for k, g in named_grads.items():
        flat = g.detach().double().flatten()
        step = max(1, flat.numel() // 8)
        out["n:" + k] = np.array([float(flat.norm())])
        out["s:" + k] = flat[::step][:8].numpy()
    return out


def grad_samples64(named_grads):
    """64 strided samples per tensor (all of it when smaller) + its L2 norm: enough for a sample-based relative-L2 error."""
    out = {}
    for k, g in named_grads.items():
        flat = g.detach().double().flatten()
        step = max(1, flat.numel() // 64)
        out["n:" + k] = np.array([float(flat.norm())])
        out["s:" + k] = flat[::step][:64].numpy()
    return out


def oracle_grads_fp64(sd, train_bn):
    """The same algorithm in float64 (PrRoIPool stays float32 numpy, cast around): the 'exact arithmetic' gradients.  The distance between
    these and the float32 reference gradients is the reference's OWN rounding noise -- the floor any float32-class implementation can be
    held to on this (deliberately ill-conditioned, random-weight) network."""
    pp = O.prpool_feature
    O.prpool_feature = lambda f, b: pp(f.float(), b).to(f.dtype)
    try:
        params = {k: (v.double() if v.dtype.is_floating_point else v).clone().requires_grad_(
            v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))) for k, v in sd.items()}
        ins = [t.double() for t in inputs()]
        z, x, tb, sb, smem, label, reg_target, reg_weight = ins
        O._CAL.on = train_bn
        try:
            losses = O.forward_train(params, z, x, label, reg_target, reg_weight, tb, smem, sb, 0.4)
        finally:
            O._CAL.on = False
        (losses[0] + losses[1] + losses[2]).backward()
        return [float(v.detach()) for v in losses], {k: v.grad for k, v in params.items() if v.requires_grad and v.grad is not None}
    finally:
        O.prpool_feature = pp


def oracle_grads(sd, train_bn):
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var")))
              for k, v in sd.items()}
    z, x, tb, sb, smem, label, reg_target, reg_weight = inputs()
    O._CAL.on = train_bn  # batch statistics (also overwrites the running stats of this private copy)
    try:
        losses = O.forward_train(params, z, x, label, reg_target, reg_weight, tb, smem, sb, 0.4)
    finally:
        O._CAL.on = False
    (losses[0] + losses[1] + losses[2]).backward()
    return [float(v.detach()) for v in losses], {k: v.grad for k, v in params.items() if v.requires_grad and v.grad is not None}


def reference_grads(sd, train_bn):
    net = ref_models.USOT({"mem_size": M, "pr_pool": True})
    net.load_state_dict(sd, strict=True)
    net.train(train_bn)
    if train_bn:
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.momentum = None
    z, x, tb, sb, smem, label, reg_target, reg_weight = inputs()
    losses = net(z, x, label=label, reg_target=reg_target, reg_weight=reg_weight, template_bbox=tb, search_memory=smem, search_bbox=sb,
                 cls_ratio=0.4)
    (losses[0] + losses[1] + losses[2]).backward()
    return [float(v.detach()) for v in losses], {k: p.grad for k, p in net.named_parameters() if p.grad is not None}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    sd = load_weights()
    out = {"B": B, "M": M}
    for mode, train_bn in (("eval", False), ("train", True)):
        rl, rg = reference_grads(sd, train_bn)
        ol, og = oracle_grads(sd, train_bn)
        assert set(rg.keys()) == set(og.keys()), (sorted(set(rg) ^ set(og))[:5])
        worst = 0.0
        for k in rg:
            denom = float(rg[k].abs().max())
            err = float((rg[k] - og[k]).abs().max()) / max(denom, 1e-30)
            worst = max(worst, err if denom > 1e-12 else 0.0)
        loss_err = max(abs(a - b) / abs(b) for a, b in zip(ol, rl))
        print(f"{mode}: {len(rg)} parameter gradients, losses {np.round(rl, 6)}, worst rel-maxabs(oracle - reference) = {worst:.3e}, loss err {loss_err:.1e}")
        assert worst <= 2e-4 and loss_err <= 1e-5, "oracle autograd does not reproduce the reference gradients"
        summ = grad_summary(rg)
        out.update({f"{mode}/{k}": v for k, v in summ.items()})
        out[f"{mode}/losses"] = np.array(rl, np.float64)
    np.savez_compressed(os.path.join(GOLD, "grads_damp025.npz"), **out)
    print("wrote tests/golden/grads_damp025.npz")
    # second fixture: 64 samples per tensor of (a) the float32 reference gradients and (b) the float64 'exact' gradients
    out64 = {"B": B, "M": M}
    for mode, train_bn in (("eval", False), ("train", True)):
        _, rg = reference_grads(sd, train_bn)
        l64, g64 = oracle_grads_fp64(sd, train_bn)
        assert set(rg) == set(g64)
        out64.update({f"{mode}/ref32/{k}": v for k, v in grad_samples64(rg).items()})
        out64.update({f"{mode}/exact/{k}": v for k, v in grad_samples64(g64).items()})
        out64[f"{mode}/exact_losses"] = np.array(l64)
        num = sum(float(((rg[k].double() - g64[k]) ** 2).sum()) for k in rg)
        den = sum(float((g64[k] ** 2).sum()) for k in rg)
        out64[f"{mode}/ref32_vs_exact_global_rel_l2"] = np.array([np.sqrt(num / den)])
        print(f"{mode}: float32 reference vs float64 arithmetic: global relative L2 of the gradient = {np.sqrt(num / den):.3e}")
    np.savez_compressed(os.path.join(GOLD, "grads64_damp025.npz"), **out64)
    print("wrote tests/golden/grads64_damp025.npz")


if __name__ == "__main__":
    main()
