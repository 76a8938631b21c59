"""usot_b200.synth (product-side benchmark weights) and the oracle generate identical tensors."""
import torch

import usot_oracle as O
from helpers import load_weights
from usot_b200.synth import synthetic_inputs, synthetic_state_dict


def test_synth_matches_oracle_generator():
    for name in ("damp025", "raw"):
        a, b = synthetic_state_dict(name), load_weights(name)
        assert set(a.keys()) == set(b.keys())
        for k in a:
            assert torch.equal(a[k], b[k]), k
    for got, ref in zip(synthetic_inputs(9, 2, 271), O.synth_inputs(9, 2, 271)):
        assert torch.equal(got, ref)
