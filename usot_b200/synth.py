"""Seeded synthetic weights / inputs for benchmarking and smoke runs (no checkpoints or datasets are reachable offline).

The reference ships no weights (README.md:88-95 points at Google Drive) and its raw default init overflows ``exp`` in the
regression head, so benchmarks use N(0, sqrt(2/(k*k*cout))) conv weights (the law of lib/models/modules.py:96-102) with
BN running statistics calibrated once offline and stored in tests/golden/bnstats_<name>.npz.  tests/test_synth.py checks
that this generator and the oracle's produce identical tensors.
"""
import math
import os

import numpy as np
import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHT_SETS = {"damp025": dict(seed=11, damp=0.25), "raw": dict(seed=12, damp=None)}


def _conv_table():
    """(name, bn, cin, cout, k, bias) of every conv in state_dict order."""
    t = [("features.features.conv1", "features.features.bn1", 3, 64, 7, False)]
    inplanes = 64
    for lname, planes, blocks, shortcut_k in (("layer1", 64, 3, 1), ("layer2", 128, 4, 3), ("layer3", 256, 6, 3)):
        for i in range(blocks):
            p = f"features.features.{lname}.{i}."
            t.append((p + "conv1", p + "bn1", inplanes, planes, 1, False))
            t.append((p + "conv2", p + "bn2", planes, planes, 3, False))
            t.append((p + "conv3", p + "bn3", planes, planes * 4, 1, False))
            if i == 0:
                t.append((p + "downsample.0", p + "downsample.1", inplanes, planes * 4, shortcut_k, False))
                inplanes = planes * 4
    t.append(("neck.downsample.0", "neck.downsample.1", 1024, 256, 1, False))
    for enc in ("cls_encode", "reg_encode"):
        for m in ("matrix11", "matrix12", "matrix21"):
            for br in ("k", "s"):
                p = f"connect_model.{enc}.{m}_{br}."
                t.append((p + "0", p + "1", 256, 256, 3, False))
    for g in ("conf_gen", "value_gen"):
        p = f"connect_model.conf_fusion.{g}."
        t.append((p + "0", p + "1", 256, 256, 3, True))
    for tw in ("bbox_tower", "cls_tower", "cls_memory_tower"):
        for i in range(4):
            p = f"connect_model.{tw}."
            t.append((p + str(3 * i), p + str(3 * i + 1), 256, 256, 3, True))
    t.append(("connect_model.bbox_pred", None, 256, 4, 3, True))
    t.append(("connect_model.cls_pred", None, 256, 1, 3, True))
    t.append(("connect_model.cls_memory_pred", None, 256, 1, 3, True))
    return t


def synthetic_state_dict(name="damp025", calibrated=True):
    cfg = WEIGHT_SETS[name]
    g = torch.Generator().manual_seed(cfg["seed"])
    sd = {}
    for conv, bn, cin, cout, k, bias in _conv_table():
        std = 0.03 if conv.endswith("_pred") else math.sqrt(2.0 / (k * k * cout))
        sd[conv + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * std
        if bias:
            sd[conv + ".bias"] = (torch.rand(cout, generator=g) - 0.5) * 0.2
        if bn is not None:
            w = torch.ones(cout)
            if cfg["damp"] is not None and bn.endswith("bn3"):
                w = w * cfg["damp"]
            sd[bn + ".weight"] = w
            sd[bn + ".bias"] = torch.zeros(cout)
            sd[bn + ".running_mean"] = torch.zeros(cout)
            sd[bn + ".running_var"] = torch.ones(cout)
            sd[bn + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
    sd["connect_model.cls_dw.weight"] = torch.tensor([1.0, 0.5, 1.5])
    sd["connect_model.reg_dw.weight"] = torch.tensor([0.7, 1.2, 1.0])
    sd["connect_model.adjust"] = 0.1 * torch.ones(1)
    sd["connect_model.bias"] = torch.ones(1, 4, 1, 1)
    if calibrated:
        st = np.load(os.path.join(_ROOT, "tests", "golden", f"bnstats_{name}.npz"))
        for k in st.files:
            sd[k] = torch.from_numpy(st[k].copy())
    return sd


def synthetic_inputs(seed, batch, search_size=255, n_templates=1):
    """U[0,255) crops (raw BGR range, lib/utils/track_utils.py:24-27) and PrPool boxes in feature coordinates."""
    g = torch.Generator().manual_seed(seed)
    z = torch.rand(n_templates, 3, 127, 127, generator=g) * 255.0
    x = torch.rand(batch, 3, search_size, search_size, generator=g) * 255.0
    tb = torch.tensor([[3.3, 4.1, 10.6, 11.2]]).repeat(n_templates, 1) + torch.rand(n_templates, 4, generator=g)
    sb = torch.tensor([[8.2, 7.4, 16.9, 15.3]]).repeat(batch, 1) + torch.rand(batch, 4, generator=g) * 2.0
    return z, x, tb, sb
