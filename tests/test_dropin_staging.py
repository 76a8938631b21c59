"""CPU checks of the drop-in plumbing: the staged reference tree is byte-identical to the reference it was copied from, the
PEP-420 shadowing resolves as INTEGRATION.md says (model -> this repository, everything else -> the reference), and
USOT_B200_HOST_TRACKER=1 hands lib.tracker.usot_tracker back to the reference's own file."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
STUBS = os.path.join(ROOT, "baseline", "stubs")

needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "MANIFEST.json")), reason="baseline/_ref not staged")


@needs_ref
def test_staged_tree_is_unmodified():
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import stage_reference
    assert stage_reference.verify()
    if os.path.isdir("/root/reference"):   # build container: compare against the live tree as well
        import json
        with open(os.path.join(REF, "MANIFEST.json")) as f:
            files = json.load(f)["files"]
        assert len(files) >= 50
        for rel, h in files.items():
            assert stage_reference.sha256(os.path.join("/root/reference", rel)) == h, rel


def _probe(host_tracker):
    code = ("import lib.models.models as m, lib.tracker.usot_tracker as t, lib.utils.train_utils as u, lib.dataset_loader.benchmark as b\n"
            "import inspect\n"
            "print(m.__file__); print(inspect.getsourcefile(t.USOTTracker)); print(u.__file__); print(b.__file__)\n"
            "print(sorted(k for k in m.__dict__ if k.startswith('USOT')))\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, STUBS, REF]), USOT_B200_HOST_TRACKER="1" if host_tracker else "0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd="/tmp", timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout.strip().splitlines()


@needs_ref
@pytest.mark.parametrize("host_tracker", [False, True])
def test_namespace_shadowing_resolution(host_tracker):
    model_file, tracker_file, utils_file, bench_file, names = _probe(host_tracker)
    assert os.path.samefile(model_file, os.path.join(ROOT, "lib", "models", "models.py"))
    assert os.path.samefile(utils_file, os.path.join(REF, "lib", "utils", "train_utils.py"))
    assert os.path.samefile(bench_file, os.path.join(REF, "lib", "dataset_loader", "benchmark.py"))
    want = os.path.join(REF, "lib", "tracker", "usot_tracker.py") if host_tracker else os.path.join(ROOT, "usot_b200", "tracker.py")
    assert os.path.samefile(tracker_file, want)
    assert "USOT" in names and "USOT_" in names


def test_stand_in_modules_cover_what_the_reference_scripts_import():
    sys.path.insert(0, STUBS)
    try:
        from easydict import EasyDict
        e = EasyDict()
        e.arch = "USOT"
        e["dataset"] = "OTB"
        assert e.arch == "USOT" and e.dataset == "OTB" and e["arch"] == "USOT"
        from shapely.geometry import Polygon, box
        a, b = box(0, 0, 10, 10), Polygon([(5, 5), (15, 5), (15, 15), (5, 15)])
        assert a.intersection(b).area == 25 and a.union(b).area == 175
        import numpy as np
        import imgaug.augmenters as iaa
        from imgaug.augmentables.bbs import BoundingBox, BoundingBoxesOnImage
        img = np.arange(24).reshape(2, 4, 3)
        out, bbs = iaa.Sequential([iaa.Fliplr(1)])(image=img, bounding_boxes=BoundingBoxesOnImage([BoundingBox(x1=0, y1=0, x2=1, y2=2)], shape=img.shape))
        assert np.array_equal(out, img[:, ::-1]) and (bbs[0].x1, bbs[0].x2) == (3, 4)
    finally:
        sys.path.remove(STUBS)
        for k in [k for k in sys.modules if k.split(".")[0] in ("easydict", "shapely", "imgaug")]:
            del sys.modules[k]
