"""Shared test helpers: golden weights, comparison metric."""
import os

import numpy as np
import torch

import usot_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WEIGHT_SETS = {"damp025": dict(seed=11, damp=0.25), "raw": dict(seed=12, damp=None)}  # must match oracle/gen_golden.py


def load_weights(name):
    """Seeded synthetic weights + the stored calibrated BN statistics (tests/golden/bnstats_*.npz)."""
    cfg = WEIGHT_SETS[name]
    sd = O.make_state_dict(cfg["seed"], cfg["damp"])
    st = np.load(os.path.join(GOLD, f"bnstats_{name}.npz"))
    for k in O.bn_stat_keys(sd):
        sd[k] = torch.from_numpy(st[k].copy())
    return sd


def golden(name):
    return np.load(os.path.join(GOLD, f"golden_{name}.npz"))


def rel_err(a, ref):
    """max-abs(a - ref) / max-abs(ref): the parity metric of SURVEY.md §8c."""
    a = torch.as_tensor(a).detach().float().cpu()
    ref = torch.as_tensor(ref).detach().float().cpu()
    return float((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30))


def subsample(xf):
    return xf[:, ::16, ::3, ::3].contiguous()
