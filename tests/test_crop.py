"""Crop path in front of the model (lib/utils/track_utils.py:30-119): CPU tests of the oracle against the fixtures made from the
LIVE reference + cv2 (oracle/gen_crop_golden.py), and of the host bookkeeping; GPU tests of the CUDA kernel against the oracle."""
import hashlib
import os

import numpy as np
import pytest
import torch

import crop_oracle as C
from helpers import GOLD


def _gold():
    return np.load(os.path.join(GOLD, "crop_golden.npz"))


def _frame(seed, h, w):
    return np.random.default_rng(int(seed)).integers(0, 256, (int(h), int(w), 3), dtype=np.uint8)


def _cases():
    g = _gold()
    for i, row in enumerate(g["cases"]):
        seed, h, w, px, py, msz, osz, tw, th, nb = row
        yield i, g, _frame(seed, h, w), np.array([px, py]), int(msz), int(osz), np.array([tw, th]), bool(nb)


def test_oracle_matches_reference_fixtures():
    n = 0
    for i, g, im, pos, msz, osz, tsz, nb in _cases():
        patch, info = C.get_subwindow_tracking(im, pos, msz, osz, np.mean(im, axis=(0, 1)), tsz, need_bbox=nb)
        assert patch.dtype == np.float32 and patch.shape == (3, msz, msz)
        assert np.array_equal(np.frombuffer(hashlib.sha256(patch.tobytes()).digest(), np.uint8), g[f"sha_{i}"]), i
        assert np.array_equal(patch[:, ::7, ::5].astype(np.uint8), g[f"sub_{i}"])
        assert list(info["crop_cords"]) + list(info["pad_info"]) == list(g[f"cords_{i}"])
        if nb:
            assert np.array_equal(np.array(info["template_bbox"], np.float64), g[f"tbox_{i}"])
        n += 1
    assert n == 12


def test_host_bookkeeping_matches_oracle():
    from usot_b200.tracker_ops import crop_geometry
    for i, g, im, pos, msz, osz, tsz, nb in _cases():
        xmin, ymin, info = crop_geometry(im.shape, pos, msz, osz, tsz, nb)
        assert (xmin, ymin) == C.context_window(pos, osz)
        assert list(info["crop_cords"]) + list(info["pad_info"]) == list(g[f"cords_{i}"])
        if nb:
            assert np.array_equal(np.array(info["template_bbox"], np.float64), g[f"tbox_{i}"])


def test_resize_properties():
    rng = np.random.default_rng(3)
    const = np.full((77, 77, 3), 93, np.uint8)
    assert (C.resize_linear_u8(const, 255) == 93).all()       # weights sum to 2048: constants are preserved
    src = rng.integers(0, 256, (127, 127, 3), dtype=np.uint8)
    assert np.array_equal(C.resize_linear_u8(src, 127), src)  # identity
    up = C.resize_linear_u8(src, 255)
    assert up.min() >= src.min() and up.max() <= src.max()    # convex combination


def test_crop_rejects_cpu_tensors():
    from usot_b200 import tracker_ops
    with pytest.raises(NotImplementedError):
        tracker_ops.crop_resize(torch.zeros(1, 8, 8, 3, dtype=torch.uint8), torch.zeros(1, 4, dtype=torch.int32),
                                torch.zeros(1, 3, dtype=torch.uint8), 127)


@pytest.mark.gpu
def test_gpu_crop_matches_oracle_bit_exact():
    from usot_b200 import tracker_ops
    for i, g, im, pos, msz, osz, tsz, nb in _cases():
        avg = np.mean(im, axis=(0, 1))
        ref, ref_info = C.get_subwindow_tracking(im, pos, msz, osz, avg, tsz, need_bbox=nb)
        ours, info = tracker_ops.get_subwindow_tracking(im, pos, msz, osz, avg, tsz, need_bbox=nb)
        assert ours.is_cuda and ours.dtype == torch.float32
        assert np.array_equal(ours.cpu().numpy(), ref), f"case {i}"
        assert np.array_equal(np.frombuffer(hashlib.sha256(ours.cpu().numpy().tobytes()).digest(), np.uint8), g[f"sha_{i}"])
        assert info["crop_cords"] == ref_info["crop_cords"] and info["pad_info"] == ref_info["pad_info"]


@pytest.mark.gpu
def test_gpu_crop_random_sizes_and_batches():
    """Seeded sweep over source sizes (every rounding class of the fixed-point weights), positions and frames, batched in one
    launch: (frame, window) pairs must each equal the oracle."""
    from usot_b200 import tracker_ops
    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (3, 96, 128, 3), dtype=np.uint8)
    fr = tracker_ops.upload_frame(frames)
    for msz in (127, 255):
        rows, fills, refs = [], [], []
        for k in range(24):
            f = int(rng.integers(0, 3))
            osz = int(rng.integers(3, 420)) if k % 6 else (2 * msz if k % 12 else msz)
            pos = rng.uniform(-30, 150, 2)
            avg = np.mean(frames[f], axis=(0, 1))
            xmin, ymin = C.context_window(pos, osz)
            rows.append([f, xmin, ymin, osz])
            fills.append(avg.astype(np.uint8))
            refs.append(C.get_subwindow_tracking(frames[f], pos, msz, osz, avg)[0])
        out = tracker_ops.crop_resize(fr, torch.tensor(rows, dtype=torch.int32).cuda(), torch.from_numpy(np.stack(fills)).cuda(), msz)
        assert np.array_equal(out.cpu().numpy(), np.stack(refs))


@pytest.mark.gpu
def test_gpu_crop_full_batch_properties():
    """256 crops of 255x255 (the BASELINE batch) in one launch: identical windows give identical patches, a constant frame gives
    constant patches, and the no-resize window reproduces the frame bytes."""
    from usot_b200 import tracker_ops
    rng = np.random.default_rng(5)
    frame = rng.integers(0, 256, (1, 480, 640, 3), dtype=np.uint8)
    fr = tracker_ops.upload_frame(frame)
    rows = [[0, 100 + (i % 4), 80, 300 + (i % 4)] for i in range(256)]
    fill = torch.full((256, 3), 7, dtype=torch.uint8).cuda()
    out = tracker_ops.crop_resize(fr, torch.tensor(rows, dtype=torch.int32).cuda(), fill, 255)
    assert tuple(out.shape) == (256, 3, 255, 255)
    assert torch.equal(out[0], out[4]) and torch.equal(out[3], out[255])
    same = tracker_ops.crop_resize(fr, torch.tensor([[0, 10, 20, 255]], dtype=torch.int32).cuda(), fill[:1], 255)
    assert np.array_equal(same[0].cpu().numpy(), frame[0, 20:275, 10:265].transpose(2, 0, 1).astype(np.float32))
    const = tracker_ops.upload_frame(np.full((120, 160, 3), 201, np.uint8))
    c = tracker_ops.crop_resize(const, torch.tensor([[0, 5, 5, 77]], dtype=torch.int32).cuda(), fill[:1], 255)
    assert float(c.min()) == 201.0 and float(c.max()) == 201.0


def test_crop_geometry_property_sweep():
    """Host bookkeeping of the device crop against the oracle's restatement of track_utils.py:41-55,81-115 over a seeded sweep
    that includes exact .5 ties (python's round-half-even), windows outside the frame and both need_bbox modes."""
    from usot_b200.tracker_ops import crop_geometry
    rng = np.random.default_rng(123)
    for k in range(400):
        h, w = int(rng.integers(20, 700)), int(rng.integers(20, 900))
        osz = int(rng.integers(2, 900))
        msz = (127, 255, 271)[k % 3]
        pos = rng.uniform(-100, 1000, 2)
        if k % 5 == 0:
            pos = np.floor(pos) + 0.5      # ties
        if k % 7 == 0:
            pos = np.floor(pos)
        tsz = rng.uniform(2, 300, 2)
        nb = bool(k % 2)
        im = np.zeros((h, w, 3), np.uint8)
        _, ref = C.get_subwindow_tracking(im, pos, msz, osz, np.zeros(3), tsz, need_bbox=nb) if osz * osz * 3 < 3e6 else (None, None)
        xmin, ymin, info = crop_geometry((h, w), pos, msz, osz, tsz, nb)
        assert (xmin, ymin) == C.context_window(pos, osz)
        if ref is not None:
            assert info["crop_cords"] == ref["crop_cords"] and info["pad_info"] == ref["pad_info"]
            assert info["original_image_bbox"] == ref["original_image_bbox"]
            if nb:
                assert info["template_bbox"] == ref["template_bbox"]
