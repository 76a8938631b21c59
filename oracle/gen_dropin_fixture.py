"""Write tests/golden/dropin_results.npz: the result files the reference's UNMODIFIED scripts/test_usot.py produces with the
reference's own model and tracker on the CPU (baseline/run_reference_cpu.py: `.cuda()` no-ops + the numpy PrRoIPool stand-in) over
the two-video synthetic OTB-style dataset.  TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference staged by
baseline/stage_reference.py):   python oracle/gen_dropin_fixture.py
tests/test_gpu_dropin.py runs the same unmodified script over the shadowed lib.models.models on the B200 and compares."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [HERE, os.path.join(ROOT, "tests"), ROOT]

import dropin_utils as D  # noqa: E402


def main():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import stage_reference
    stage_reference.stage()
    names = D.build_dataset()
    with tempfile.TemporaryDirectory() as tmp:
        ckpt = os.path.join(tmp, "synthetic_usot.pth")
        D.write_checkpoint(ckpt)
        res, log = D.run_test_usot(tmp, ckpt, shadow=False)
    print(log)
    assert sorted(res) == sorted(names), res.keys()
    for k, v in res.items():
        print(k, v.shape, "\n", np.round(v, 3))
    np.savez(os.path.join(ROOT, "tests", "golden", "dropin_results.npz"), **res)
    print("wrote tests/golden/dropin_results.npz")


if __name__ == "__main__":
    main()
