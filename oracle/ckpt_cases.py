"""Synthetic checkpoint files shared by oracle/gen_ckpt_pin.py (pin against the live reference) and tests/test_checkpoint.py.
TEST INFRASTRUCTURE ONLY."""
import os

import torch


def synthetic_checkpoints(dirname, seed=5):
    """The three files (also used by tests/test_checkpoint.py).  Returns {tag: path}."""
    from usot_b200 import USOT
    g = torch.Generator().manual_seed(seed)
    base = {k: torch.randn(v.shape, generator=g) * 0.05 if v.dtype.is_floating_point else v.clone()
            for k, v in USOT().state_dict().items()}
    paths = {}
    paths["dp"] = os.path.join(dirname, "checkpoint_e30.pth")
    torch.save({"epoch": 30, "arch": "USOT", "state_dict": {"module." + k: v for k, v in base.items()}}, paths["dp"])
    paths["online"] = os.path.join(dirname, "online.pth")
    torch.save({("feature_extractor." + k if i % 2 else k): v for i, (k, v) in enumerate(base.items())}, paths["online"])
    moco = {}
    for k, v in base.items():
        if not k.startswith("features.features."):
            continue
        name = k.replace("features.features.", "module.encoder_q.")
        if name in ("module.encoder_q.layer2.0.downsample.0.weight", "module.encoder_q.layer3.0.downsample.0.weight"):
            v = v[:, :, 1:2, 1:2].clone()  # MoCo's ResNet-50 has 1x1 shortcut kernels
        moco[name] = v
        moco[name.replace("encoder_q", "encoder_k")] = v + 1.0
    moco["module.encoder_q.fc.0.weight"] = torch.randn(8, 8, generator=g)
    moco["module.queue"] = torch.randn(4, 4, generator=g)
    paths["moco"] = os.path.join(dirname, "moco_v2_800ep.pth")
    torch.save({"epoch": 800, "state_dict": moco}, paths["moco"])
    return paths
