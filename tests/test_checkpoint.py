"""Checkpoint pipeline (lib/utils/train_utils.py:92-156): usot_b200.checkpoint against the hashes produced by the LIVE reference
load_pretrain (oracle/gen_ckpt_pin.py), and the documented prefix / MoCo / key-report behaviour."""
import os

import numpy as np
import pytest
import torch

from ckpt_cases import synthetic_checkpoints
from helpers import GOLD
from usot_b200 import USOT
from usot_b200.checkpoint import check_keys, convert_moco, load_pretrain, prepare_state_dict, remove_prefix, state_dict_hash


@pytest.fixture(scope="module")
def ckpts(tmp_path_factory):
    return synthetic_checkpoints(str(tmp_path_factory.mktemp("ckpt")))


@pytest.mark.parametrize("tag", ["dp", "online", "moco"])
def test_load_pretrain_matches_reference_hash(ckpts, tag):
    pin = np.load(os.path.join(GOLD, "ckpt_pin.npz"))
    torch.manual_seed(1)
    net = load_pretrain(USOT(), ckpts[tag], print_unuse=False, verbose=False)
    sd = net.state_dict()
    sub = {k: v for k, v in sd.items() if tag != "moco" or k.startswith("features.features.")}
    assert state_dict_hash(sub) == bytes(pin[tag]).hex()


def test_moco_embedding_and_key_report(ckpts, capsys):
    raw = torch.load(ckpts["moco"], map_location="cpu")
    sd = prepare_state_dict(raw, is_moco=True, verbose=False)
    assert all(k.startswith("features.features.") for k in sd)  # encoder_k / queue dropped, encoder_q renamed
    for layer in ("layer2", "layer3"):
        w = sd[f"features.features.{layer}.0.downsample.0.weight"]
        src = raw["state_dict"][f"module.encoder_q.{layer}.0.downsample.0.weight"]
        assert w.shape[2:] == (3, 3) and torch.equal(w[:, :, 1, 1], src[:, :, 0, 0])
        w2 = w.clone()
        w2[:, :, 1, 1] = 0
        assert float(w2.abs().max()) == 0.0
    net = USOT()
    missing, unused = check_keys(net, sd, verbose=True)
    out = capsys.readouterr().out
    assert "missing keys:" in out and "unused checkpoint keys:" in out
    assert any(k.startswith("connect_model.") for k in missing) and any(k.startswith("neck.") for k in missing)
    assert unused == ["features.features.fc.0.weight"]
    assert not any("num_batches_tracked" in k for k in missing)


def test_prefix_rules_and_empty_checkpoint():
    assert remove_prefix({"module.a.module.b": 1, "c": 2}, "module.", verbose=False) == {"a.module.b": 1, "c": 2}
    with pytest.raises(AssertionError, match="load NONE"):
        check_keys(USOT(), {"nothing.weight": torch.zeros(1)}, verbose=False)
    assert convert_moco({"module.encoder_k.x": torch.zeros(1)}) == {}


def test_hash_is_content_sensitive():
    a = USOT().state_dict()
    h0 = state_dict_hash(a)
    assert h0 == state_dict_hash({k: v.clone() for k, v in a.items()})
    b = {k: v.clone() for k, v in a.items()}
    b["connect_model.adjust"] += 1e-3
    assert state_dict_hash(b) != h0
