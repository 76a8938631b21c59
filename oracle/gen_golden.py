"""Pin the oracle against the LIVE reference and write the golden fixtures.  TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py            # writes tests/golden/*.npz, prints the pin report

What it does
  1. imports the unmodified reference ``lib.models.models.USOT`` from /root/reference with a
     harness-side ``.cuda()`` no-op shim (the reference calls ``.cuda()`` in its constructors,
     ``lib/models/models.py:119-120``, ``lib/models/connect.py:219``);
  2. builds seeded synthetic weights (``usot_oracle.make_state_dict``), calibrates BN running stats,
     and loads them with ``load_state_dict(strict=True)`` -- which also pins the 444-key contract;
  3. runs the reference modules and the oracle restatement on identical inputs and asserts they
     agree (max-abs / max-abs(ref) <= 2e-6, argmax identical);
  4. stores BN stats + reference outputs as small fixtures so the GPU box (which has no
     /root/reference) can rebuild the exact state_dict from the seed and compare.

PrRoIPool: the reference has no CPU path (``prroi_pool/functional.py:62-63``), so for the
pr_pool=True pipelines the reference's ``prroi_pool2d`` symbol is substituted by the oracle's numpy
restatement *inside this script only*; PrRoIPool itself is pinned on the GPU box against the
reference .cu compiled unchanged (``oracle/Makefile`` -> ``oracle/_ref/libprroi_ref.so``,
``tests/test_gpu_prroi.py``).
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

torch.Tensor.cuda = lambda self, *a, **k: self  # harness-side shim, see docstring
torch.nn.Module.cuda = lambda self, *a, **k: self

import lib.models.models as ref_models  # noqa: E402
import lib.models.connect as ref_connect  # noqa: E402
import lib.models.prroi_pool.prroi_pool as ref_prroi_mod  # noqa: E402

# PrRoIPool substitution (see docstring): both call sites bind the name at import time.
ref_models.prroi_pool2d = lambda f, r, ph, pw, s: O.prroi_pool2d(f, r, ph, pw, s)
ref_prroi_mod.prroi_pool2d = lambda f, r, ph, pw, s: O.prroi_pool2d(f, r, ph, pw, s)

GOLD = os.path.join(ROOT, "tests", "golden")
WEIGHT_SETS = {"damp025": dict(seed=11, damp=0.25), "raw": dict(seed=12, damp=None)}


def rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def build_weights(name):
    cfg = WEIGHT_SETS[name]
    sd = O.make_state_dict(cfg["seed"], cfg["damp"])
    z, x, tb, sb = O.synth_inputs(1000 + cfg["seed"], batch=4)
    O.calibrate_bn(sd, z, x, tb, n_mem=3)
    return sd


def load_weights(name):
    """Rebuild the state_dict on any machine: seeded weights + stored BN stats."""
    cfg = WEIGHT_SETS[name]
    sd = O.make_state_dict(cfg["seed"], cfg["damp"])
    st = np.load(os.path.join(GOLD, f"bnstats_{name}.npz"))
    for k in O.bn_stat_keys(sd):
        sd[k] = torch.from_numpy(st[k].copy())
    return sd


def subsample(xf):
    return xf[:, ::16, ::3, ::3].contiguous()


def main():
    os.makedirs(GOLD, exist_ok=True)
    report = []
    torch.set_num_threads(os.cpu_count() or 1)
    for wname in WEIGHT_SETS:
        sd = build_weights(wname)
        np.savez_compressed(os.path.join(GOLD, f"bnstats_{wname}.npz"),
                            **{k: sd[k].numpy() for k in O.bn_stat_keys(sd)})
        net = ref_models.USOT({"mem_size": 3, "pr_pool": True})
        missing = net.load_state_dict(sd, strict=True)
        assert len(net.state_dict()) == 444, len(net.state_dict())
        net.eval()
        out = {}
        with torch.no_grad():
            # ---- config 1: single pair, plumbing (pr_pool=False centre crop + offline track) ----
            z, x, tb, sb = O.synth_inputs(7, batch=1)
            net.pr_pool = False
            net.template(z)
            r_cls, r_bbox, _, _ = net.track(x)
            zf_o = O.template(sd, z, pr_pool=False)
            o_cls, o_bbox, _, _ = O.track(sd, zf_o, x)
            report.append((wname, "c1.zf_crop", rel(zf_o, net.zf)))
            report.append((wname, "c1.cls", rel(o_cls, r_cls)))
            report.append((wname, "c1.bbox", rel(o_bbox, r_bbox)))
            assert int(o_cls.argmax()) == int(r_cls.argmax())
            out.update(c1_zf=net.zf.numpy(), c1_cls=r_cls.numpy(), c1_bbox=r_bbox.numpy())

            # ---- inference with memory: pr_pool template + Nq=7 memory queue (tracker call pattern) ----
            for tag, S, B in (("m255", 255, 2), ("m271", 271, 1)):
                z, x, tb, sb = O.synth_inputs(21 if S == 255 else 22, batch=B, search_size=S)
                net.pr_pool = True
                net.template(z, template_bbox=tb)
                nq = 7
                mem_src = O.synth_inputs(31, batch=B * nq, search_size=S)[1]
                mem_box = sb.repeat_interleave(nq, 0) + 0.25 * torch.arange(B * nq).view(-1, 1) / nq
                r_mem = net.extract_memory_feature(ori_x=mem_src, search_bbox=mem_box)
                o_mem = O.extract_memory_feature(sd, ori_x=mem_src, search_bbox=mem_box)
                report.append((wname, f"{tag}.memfeat", rel(o_mem, r_mem)))
                score = torch.full((B, nq), 0.9)
                r_cls, r_bbox, r_cmem, r_xf = net.track(x, template_mem=r_mem, score_mem=score)
                zf_o = O.template(sd, z, tb)
                o_cls, o_bbox, o_cmem, o_xf = O.track(sd, zf_o, x, r_mem, score)
                for nm, a, b in (("zf", zf_o, net.zf), ("cls", o_cls, r_cls), ("bbox", o_bbox, r_bbox),
                                 ("cls_mem", o_cmem, r_cmem), ("xf", o_xf, r_xf)):
                    report.append((wname, f"{tag}.{nm}", rel(a, b)))
                for b in range(B):
                    assert int(o_cls[b].argmax()) == int(r_cls[b].argmax())
                    assert int(o_cmem[b].argmax()) == int(r_cmem[b].argmax())
                r_feat = net.extract_memory_feature(xf=r_xf, search_bbox=sb)
                out.update({f"{tag}_zf": net.zf.numpy(), f"{tag}_cls": r_cls.numpy(), f"{tag}_bbox": r_bbox.numpy(),
                            f"{tag}_cls_mem": r_cmem.numpy(), f"{tag}_xf_sub": subsample(r_xf).numpy(),
                            f"{tag}_xf_sum": r_xf.sum(dim=(1, 2, 3)).numpy(), f"{tag}_mem_sub": r_mem[:, ::8].numpy(),
                            f"{tag}_feat": r_feat.numpy(), f"{tag}_mem_box": mem_box.numpy()})

            # ---- training forward (cycle memory), config-4 shape at B=2, M=3, eval-mode BN ----
            if wname == "damp025":
                B, M = 2, 3
                z, x, tb, sb = O.synth_inputs(41, batch=B, n_templates=B)
                g = torch.Generator().manual_seed(42)
                smem = torch.rand(B, M, 3, 255, 255, generator=g) * 255.0
                label = torch.zeros(B, 25, 25)
                label[:, 10:15, 10:15] = 1.0
                reg_weight = torch.zeros(B, 25, 25)
                reg_weight[:, 11:14, 11:14] = 1.0
                reg_target = torch.rand(B, 25, 25, 4, generator=g) * 40.0 + 5.0
                r_losses = net(z, x, label=label, reg_target=reg_target, reg_weight=reg_weight, template_bbox=tb,
                               search_memory=smem, search_bbox=sb, cls_ratio=0.4)
                # the reference mutates its grid attributes in forward(); rebuild for a clean module state
                o = O.forward_train(sd, z, x, label, reg_target, reg_weight, tb, smem, sb, 0.4, detail=True)
                for nm, a, b in (("cls_loss", o["cls_loss"], r_losses[0]), ("cls_memory_loss", o["cls_memory_loss"], r_losses[1]),
                                 ("reg_loss", o["reg_loss"], r_losses[2])):
                    report.append((wname, f"train.{nm}", abs(float(a) - float(b)) / abs(float(b))))
                out.update(train_losses=np.array([float(v) for v in r_losses], np.float64),
                           train_forward_argmax=o["forward_argmax"].numpy(), train_backward_map=o["backward_map"].numpy(),
                           train_pool_box=o["pool_box"].numpy())
        np.savez_compressed(os.path.join(GOLD, f"golden_{wname}.npz"), **out)

    worst = 0.0
    for w, n, v in report:
        print(f"{w:8s} {n:22s} rel-maxabs = {v:.3e}")
        worst = max(worst, v)
    print("worst:", worst)
    assert worst <= 2e-6, "oracle does not match the live reference"
    with open(os.path.join(GOLD, "PIN_REPORT.txt"), "w") as f:
        f.write("oracle (oracle/usot_oracle.py) vs live reference modules (/root/reference, torch %s CPU)\n" % torch.__version__)
        for w, n, v in report:
            f.write(f"{w:8s} {n:22s} rel-maxabs = {v:.3e}\n")
    print("golden fixtures written to", GOLD)


if __name__ == "__main__":
    main()
