"""Pin oracle/crop_oracle.py against the LIVE cv2 and the LIVE reference crop, and write tests/golden/crop_golden.npz.
TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference and opencv-python):

    python oracle/gen_crop_golden.py

  1. ``resize_linear_u8`` vs ``cv2.resize`` (default INTER_LINEAR, uint8 HxWx3) on seeded random squares, with IPP enabled and
     disabled: must be bit-identical (up-scaling, down-scaling, the exact-2x INTER_AREA route, tiny and huge sources);
  2. ``get_subwindow_tracking`` vs the unmodified reference function (lib/utils/track_utils.py:30-119) on seeded frames with
     windows inside the frame, crossing every border and larger than the frame: patches bit-identical, crop_info equal;
  3. the case table + sha256 of every reference patch + a strided sub-sample are stored as the fixture.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("USOT_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(1, REF)

import cv2  # noqa: E402

import crop_oracle as C  # noqa: E402
from lib.utils.track_utils import get_subwindow_tracking as ref_crop  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")

# (frame seed, H, W, pos_x, pos_y, model_sz, original_sz, target_w, target_h, need_bbox)
CASES = [
    (1, 240, 320, 160.0, 120.0, 127, 90, 40.0, 30.0, 1),      # inside, up-scale (template with box)
    (1, 240, 320, 160.4, 119.6, 255, 181, 40.0, 30.0, 0),     # inside, up-scale (search)
    (2, 240, 320, 12.3, 20.7, 255, 300, 60.0, 50.0, 0),       # crosses left/top, down-scale
    (2, 240, 320, 310.0, 230.5, 255, 333, 60.0, 50.0, 0),     # crosses right/bottom
    (3, 120, 160, 80.0, 60.0, 255, 510, 30.0, 30.0, 0),       # window larger than the frame, exact 2x route
    (3, 120, 160, 81.5, 60.5, 127, 254, 30.0, 30.0, 1),       # exact 2x route for the template
    (4, 300, 300, 150.0, 150.0, 255, 255, 80.0, 80.0, 0),     # original_sz == model_sz: no resize
    (4, 300, 300, 2.5, 297.5, 127, 127, 80.0, 80.0, 1),       # no resize, heavy padding
    (5, 480, 640, 600.7, 50.2, 271, 415, 100.0, 80.0, 0),     # search size 271
    (5, 480, 640, 320.0, 240.0, 255, 37, 10.0, 8.0, 0),       # strong up-scale
    (6, 97, 131, -20.0, 300.0, 255, 200, 10.0, 8.0, 0),       # window entirely outside the frame: pure padding
    (6, 97, 131, 65.5, 48.5, 127, 1001, 10.0, 8.0, 1),        # huge down-scale
]


def frame(seed, h, w):
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def main():
    rng = np.random.default_rng(0)
    n = 0
    for ipp in (True, False):
        cv2.ipp.setUseIPP(ipp)
        for trial in range(150):
            ds = (127, 255, 271)[trial % 3]
            ss = int(rng.integers(2, 1200))
            if trial % 10 == 0:
                ss = 2 * ds
            if trial % 10 == 1:
                ss = ds + (1 if trial % 20 == 1 else -1)
            src = rng.integers(0, 256, (ss, ss, 3), dtype=np.uint8)
            assert np.array_equal(cv2.resize(src, (ds, ds)), C.resize_linear_u8(src, ds)), (ipp, ss, ds)
            n += 1
    cv2.ipp.setUseIPP(True)
    print(f"resize_linear_u8 == cv2.resize ({cv2.__version__}) bit for bit on {n} random squares (IPP on and off)")

    out = {"cases": np.array(CASES, np.float64)}
    for i, (seed, h, w, px, py, msz, osz, tw, th, nb) in enumerate(CASES):
        im = frame(seed, h, w)
        avg = np.mean(im, axis=(0, 1))
        pos, tsz = np.array([px, py]), np.array([tw, th])
        ref_t, ref_info = ref_crop(im, pos, msz, osz, avg, tsz, need_bbox=bool(nb))
        ref = ref_t.numpy()
        ours, info = C.get_subwindow_tracking(im, pos, msz, osz, avg, tsz, need_bbox=bool(nb))
        assert ref.dtype == np.float32 and ours.dtype == np.float32 and np.array_equal(ref, ours), i
        assert list(ref_info["crop_cords"]) == list(info["crop_cords"]) and list(ref_info["pad_info"]) == list(info["pad_info"]), i
        assert list(ref_info["original_image_bbox"]) == list(info["original_image_bbox"]), i
        if nb:
            assert np.allclose(ref_info["template_bbox"], info["template_bbox"], rtol=0, atol=0), i
            out[f"tbox_{i}"] = np.array(ref_info["template_bbox"], np.float64)
        out[f"sha_{i}"] = np.frombuffer(hashlib.sha256(ref.tobytes()).digest(), np.uint8)
        out[f"sub_{i}"] = ref[:, ::7, ::5].astype(np.uint8)
        out[f"cords_{i}"] = np.array(list(ref_info["crop_cords"]) + list(ref_info["pad_info"]), np.int64)
    np.savez_compressed(os.path.join(GOLD, "crop_golden.npz"), **out)
    print(f"get_subwindow_tracking == reference on {len(CASES)} cases; wrote tests/golden/crop_golden.npz")


if __name__ == "__main__":
    main()
