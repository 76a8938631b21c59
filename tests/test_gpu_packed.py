"""Packed-weight image (usot_engine_export_packed / usot_engine_import_packed) and the state_dict-hash keyed cache:
an engine restored from an image must compute exactly what the engine that packed the state_dict computes."""
import os

import pytest
import torch

import usot_oracle as O
from helpers import load_weights

pytestmark = pytest.mark.gpu


def _track(eng, z, tb, x):
    zf, _ = eng.template(z.cuda(), tb.cuda())
    return eng.track(x.cuda(), zf)


@pytest.mark.parametrize("precision", ["fp16x3", "fp32"])
def test_export_import_roundtrip(tmp_path, precision):
    from usot_b200 import Engine
    sd = load_weights("damp025")
    z, x, tb, sb = O.synth_inputs(5, batch=2)
    a = Engine("cuda:0", precision)
    a.load_state_dict(sd)
    ref = _track(a, z, tb, x)
    path = str(tmp_path / "w.usotw")
    n = a.export_packed(path)
    assert os.path.getsize(path) == n > 50e6
    b = Engine("cuda:0", precision)
    b.import_packed(path)  # no state_dict involved
    out = _track(b, z, tb, x)
    for u, v in zip(ref[:2], out[:2]):
        assert torch.equal(u, v)
    other = Engine("cuda:0", "fp16" if precision != "fp16" else "fp32")
    with pytest.raises(RuntimeError, match="precision"):
        other.import_packed(path)
    import numpy as np
    raw = np.fromfile(path, dtype=np.uint8)
    flipped = raw.copy()
    flipped[n // 3] ^= 0x40  # one flipped bit in the payload (the image carries a checksum now)
    flipped.tofile(path)
    with pytest.raises(RuntimeError, match="corrupt"):
        Engine("cuda:0", precision).import_packed(path)
    raw[: n // 2].tofile(path)  # truncated image
    with pytest.raises(RuntimeError, match="corrupt|truncated"):
        Engine("cuda:0", precision).import_packed(path)


def test_cache_keyed_by_state_dict_hash(tmp_path, monkeypatch):
    from usot_b200 import USOT
    monkeypatch.setenv("USOT_B200_WEIGHT_CACHE", str(tmp_path))
    sd = load_weights("damp025")
    z, x, tb, sb = O.synth_inputs(6, batch=1)

    def run(state):
        net = USOT(precision="fp16x3")
        net.load_state_dict(state)
        net = net.eval().cuda()
        net.template(z.cuda(), tb.cuda())
        return net.track(x.cuda())[:2]

    first = run(sd)                      # miss: packs and writes the image
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].endswith("_fp16x3_abi3.usotw")
    second = run(sd)                     # hit: restored from the image
    assert sorted(os.listdir(tmp_path)) == files
    for u, v in zip(first, second):
        assert torch.equal(u, v)
    sd2 = {k: v.clone() for k, v in sd.items()}
    sd2["connect_model.adjust"] = sd2["connect_model.adjust"] * 1.5
    third = run(sd2)                     # different weights -> different key, different result
    assert len(os.listdir(tmp_path)) == 2
    assert not torch.equal(first[1], third[1])
