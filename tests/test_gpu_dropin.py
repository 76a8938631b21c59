"""Drop-in acceptance (BASELINE.json north_star: "scripts/test_usot.py runs unchanged"): the reference's UNMODIFIED
scripts/test_usot.py (staged byte-for-byte in baseline/_ref by baseline/stage_reference.py; :134-161 builds the model through
``lib.models.models.__dict__['USOT']()``, ``load_pretrain``, ``.eval().cuda()`` and drives ``USOTTracker`` over a dataset) is run
as a subprocess over this repository's shadow of ``lib.models.models`` on the B200, twice:

  * with the reference's OWN host-side tracker loop, lib/tracker/usot_tracker.py:22-131,202-276 unmodified (host queue of memory
    features, ``.cpu()`` / ``torch.cat`` / ``.cuda()`` round trips of channels-last features, numpy post-processing), and
  * with the device-side tracker shadow (usot_b200/tracker.py).

The result files it writes must match tests/golden/dropin_results.npz, written by the same unmodified script running the
reference's own model on the CPU (oracle/gen_dropin_fixture.py)."""
import os

import numpy as np
import pytest

import dropin_utils as D
from helpers import GOLD

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def staged(tmp_path_factory):
    if not os.path.exists(os.path.join(D.REF, "MANIFEST.json")):
        pytest.skip("baseline/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
    import sys
    sys.path.insert(0, os.path.join(D.ROOT, "baseline"))
    import stage_reference
    assert stage_reference.verify(), "baseline/_ref differs from the reference it was staged from"
    names = D.build_dataset()
    tmp = tmp_path_factory.mktemp("dropin")
    ckpt = str(tmp / "synthetic_usot.pth")
    D.write_checkpoint(ckpt)
    return names, tmp, ckpt


@pytest.mark.parametrize("host_tracker", [True, False], ids=["reference_host_tracker", "device_tracker"])
@pytest.mark.parametrize("precision", ["fp16x3", "fp32"])
def test_unmodified_test_usot_script_over_the_shadow(staged, host_tracker, precision):
    names, tmp, ckpt = staged
    cwd = tmp / f"run_{precision}_{int(host_tracker)}"
    cwd.mkdir()
    res, log = D.run_test_usot(str(cwd), ckpt, host_tracker=host_tracker, shadow=True, extra_env={"USOT_B200_PRECISION": precision})
    print(log)
    gold = np.load(os.path.join(GOLD, "dropin_results.npz"))
    assert sorted(res) == sorted(names) == sorted(gold.files)
    for k in names:
        assert res[k].shape == gold[k].shape
        assert np.array_equal(res[k][0], gold[k][0])                      # frame 0 is the ground-truth box, written verbatim
        err = np.abs(res[k] - gold[k]).max()
        print(k, "host tracker" if host_tracker else "device tracker", precision, "max |box - reference| =", err)
        assert err <= 0.1, (k, err)                                         # pixels; same bar as the tracker trace test
