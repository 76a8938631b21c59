"""Pin oracle/tracker_oracle.py against the LIVE reference tracker and write tests/golden/tracker_trace.npz.  TEST INFRASTRUCTURE.
Run in the build container only (needs /root/reference):   python oracle/gen_tracker_pin.py

The unmodified ``lib.tracker.usot_tracker.USOTTracker`` (init + track) is driven over a seeded synthetic video with the
unmodified reference ``USOT`` model on the CPU.  Harness-side stand-ins, none of which touches the arithmetic under test:
``.cuda()`` no-ops; the reference's GPU-only PrRoIPool symbol replaced by the numpy restatement (as in gen_golden.py; PrRoIPool
itself is pinned on the GPU box against the reference .cu); a minimal ``imgaug`` with the 0.4.0 semantics of ``Fliplr(1)`` on an
image and one bounding box (imgaug is not installed here and is numpy-2 incompatible).  The oracle tracker must reproduce the
reference trace (target position / size / confidence per frame) to float rounding, with the same arg-max cell every frame.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("USOT_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
sys.path.insert(1, REF)

import tracker_oracle as T  # noqa: E402
import usot_oracle as O  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self

# ---- stand-in imgaug (Fliplr(1) on an HWC image + one BoundingBox, imgaug 0.4.0 semantics) ----
ia = types.ModuleType("imgaug")
iaa = types.ModuleType("imgaug.augmenters")
iab = types.ModuleType("imgaug.augmentables")
iabb = types.ModuleType("imgaug.augmentables.bbs")


class BoundingBox:
    def __init__(self, x1, y1, x2, y2):
        self.x1, self.y1, self.x2, self.y2 = x1, y1, x2, y2


class BoundingBoxesOnImage:
    def __init__(self, bounding_boxes, shape):
        self.bounding_boxes, self.shape = bounding_boxes, shape

    def __getitem__(self, i):
        return self.bounding_boxes[i]


class Fliplr:
    def __init__(self, p):
        assert p == 1


class Sequential:
    def __init__(self, children):
        assert len(children) == 1 and isinstance(children[0], Fliplr)

    def __call__(self, image, bounding_boxes):
        w = bounding_boxes.shape[1]
        out = [BoundingBox(w - b.x2, b.y1, w - b.x1, b.y2) for b in bounding_boxes.bounding_boxes]
        return image[:, ::-1], BoundingBoxesOnImage(out, bounding_boxes.shape)


iaa.Sequential, iaa.Fliplr = Sequential, Fliplr
iabb.BoundingBox, iabb.BoundingBoxesOnImage = BoundingBox, BoundingBoxesOnImage
ia.augmenters, ia.augmentables, iab.bbs = iaa, iab, iabb
sys.modules.update({"imgaug": ia, "imgaug.augmenters": iaa, "imgaug.augmentables": iab, "imgaug.augmentables.bbs": iabb})

import lib.models.models as ref_models  # noqa: E402
import lib.models.prroi_pool.prroi_pool as ref_prroi_mod  # noqa: E402
from lib.tracker.usot_tracker import USOTTracker  # noqa: E402

ref_models.prroi_pool2d = lambda f, r, ph, pw, s: O.prroi_pool2d(f, r, ph, pw, s)
ref_prroi_mod.prroi_pool2d = lambda f, r, ph, pw, s: O.prroi_pool2d(f, r, ph, pw, s)

GOLD = os.path.join(ROOT, "tests", "golden")
N_FRAMES = 6


def load_weights(name="damp025", seed=11, damp=0.25):
    sd = O.make_state_dict(seed, damp)
    st = np.load(os.path.join(GOLD, f"bnstats_{name}.npz"))
    for k in O.bn_stat_keys(sd):
        sd[k] = torch.from_numpy(st[k].copy())
    return sd


def oracle_trace(sd, seed, box):
    frames, pos0, sz0 = T.synthetic_video(seed=seed, n_frames=N_FRAMES, box=box)
    ostate = T.tracker_init(frames[0], pos0.copy(), sz0.copy(), T.OracleNet(sd))
    otrace, gaps, margins = [], [], []
    for im in frames[1:]:
        ostate = T.tracker_track(ostate, im)
        otrace.append(np.concatenate([ostate['target_pos'], ostate['target_sz'], [ostate['cls_score']]]))
        gaps.append(ostate['top2_gap'])
        margins.append(ostate['round_margin'])
    return np.array(otrace, np.float64), gaps, margins, ostate['p'].instance_size


def pin_one(sd, tag, box, want_size):
    """Choose a video whose trace is robust to 1e-3-class arithmetic differences (arg-max margin and the rounding margins of the
    crop geometry comfortably away from their decision boundaries on every frame), then require oracle == live reference on it."""
    for seed in range(3, 60):
        _, gaps, margins, size = oracle_trace(sd, seed, box)
        assert size == want_size, size
        print(tag, "video seed", seed, "min arg-max margin %.4f" % min(gaps), "min rounding margin %.3f" % min(margins))
        if min(gaps) >= 4e-3 and min(margins) >= 0.08:
            break
    else:
        raise SystemExit("no robust seed found")
    frames, pos0, sz0 = T.synthetic_video(seed=seed, n_frames=N_FRAMES, box=box)
    net = ref_models.USOT()
    net.load_state_dict(sd, strict=True)
    net.eval()
    tracker = USOTTracker(types.SimpleNamespace(arch="USOT"))
    ref_trace = []
    with torch.no_grad():
        state = tracker.init(frames[0], pos0.copy(), sz0.copy(), net)
        assert state['p'].instance_size == want_size
        for im in frames[1:]:
            state = tracker.track(state, im)
            ref_trace.append(np.concatenate([state['target_pos'], state['target_sz'], [state['cls_score']]]))
    ref_trace = np.array(ref_trace, np.float64)
    otrace, gaps, margins, _ = oracle_trace(sd, seed, box)
    err = np.abs(otrace - ref_trace).max()
    print(tag, "reference trace (x, y, w, h, conf):\n", np.round(ref_trace, 4))
    print(tag, "max |oracle - reference| over the trace:", err, " min arg-max margin:", min(gaps))
    assert err <= 1e-9, err
    return {f"{tag}trace": ref_trace, f"{tag}gaps": np.array(gaps), f"{tag}margins": np.array(margins), f"{tag}pos0": pos0, f"{tag}sz0": sz0,
            f"{tag}video_seed": seed, f"{tag}box": np.array(box)}


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    sd = load_weights()
    out = {"n_frames": N_FRAMES}
    out.update(pin_one(sd, "", (64, 48), 255))        # ordinary target: 255-pixel search window, 25x25 response
    out.update(pin_one(sd, "small_", (14, 12), 271))  # target < 0.4 % of the frame: 271-pixel window, 27x27 response
    np.savez(os.path.join(GOLD, "tracker_trace.npz"), **out)
    print("wrote tests/golden/tracker_trace.npz")


if __name__ == "__main__":
    main()
