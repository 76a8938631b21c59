"""Pin usot_b200/checkpoint.py against the LIVE reference ``load_pretrain`` (lib/utils/train_utils.py:92-128).  TEST INFRASTRUCTURE.
Run in the build container only (needs /root/reference):   python oracle/gen_ckpt_pin.py

Three synthetic checkpoint files are written to a temp dir -- a DataParallel-style one ('module.' prefix inside a
{'state_dict': ...} wrapper), a bare online-train one ('feature_extractor.' prefix) and a MoCo-v2 style one (path contains
"moco", 'module.encoder_q.*' keys with 1x1 layer2.0/layer3.0 shortcut kernels plus unrelated 'encoder_k' / fc keys).  Each is
loaded (a) by the unmodified reference function into the unmodified reference USOT and (b) by usot_b200.checkpoint.load_pretrain
into usot_b200.USOT; the two resulting state_dicts must be identical.  The reference needs CUDA only to place tensors
(``storage.cuda(device)``, ``.cuda()``), so harness-side no-op shims stand in for those calls.  The content hashes are stored
in tests/golden/ckpt_pin.npz; tests/test_checkpoint.py rebuilds the same files from the seed and must reproduce them.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("USOT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(1, HERE)
sys.path.insert(2, REF)

torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
torch.UntypedStorage.cuda = lambda self, *a, **k: self
torch.cuda.current_device = lambda: 0
torch.cuda.set_device = lambda *a, **k: None

import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_models", os.path.join(REF, "lib", "models", "models.py"))
from lib.utils.train_utils import load_pretrain as ref_load  # noqa: E402  (lib.utils resolves to the reference)

from usot_b200.checkpoint import load_pretrain, state_dict_hash  # noqa: E402
from ckpt_cases import synthetic_checkpoints  # noqa: E402


def ref_usot():
    # lib.models.models is shadowed by this repo when ROOT precedes the reference on sys.path: load the reference file itself
    ref_models = importlib.util.module_from_spec(spec)
    sys.modules["ref_models"] = ref_models
    spec.loader.exec_module(ref_models)
    return ref_models.USOT()


def main():
    from usot_b200 import USOT
    out = {}
    with tempfile.TemporaryDirectory() as d:
        paths = synthetic_checkpoints(d)
        for tag, path in paths.items():
            torch.manual_seed(1)
            ref = ref_load(ref_usot(), path, print_unuse=False)
            torch.manual_seed(1)
            ours = load_pretrain(USOT(), path, print_unuse=False, verbose=False)
            a, b = ref.state_dict(), ours.state_dict()
            assert list(a.keys()) == list(b.keys()), tag
            if tag != "moco":  # every tensor comes from the file; for moco the head keeps each model's own random init
                for k in a:
                    assert torch.equal(a[k], b[k]), (tag, k)
            else:
                for k in a:
                    if k.startswith("features.features."):
                        assert torch.equal(a[k], b[k]), (tag, k)
            sub = {k: v for k, v in b.items() if tag != "moco" or k.startswith("features.features.")}
            assert state_dict_hash(sub) == state_dict_hash({k: a[k] for k in sub})
            out[tag] = np.frombuffer(bytes.fromhex(state_dict_hash(sub)), np.uint8)
            print(tag, "reference == usot_b200.checkpoint on", len(sub), "tensors;", state_dict_hash(sub)[:16])
    np.savez(os.path.join(ROOT, "tests", "golden", "ckpt_pin.npz"), **out)


if __name__ == "__main__":
    main()
