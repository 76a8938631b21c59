"""CPU oracle for the SiamFC-style crop that feeds the forward path.  TEST INFRASTRUCTURE ONLY.

Restates, in numpy integer arithmetic,

  * ``get_subwindow_tracking``  /root/reference/lib/utils/track_utils.py:30-119  (context window, average-colour padding,
    crop, resize to the model size, CHW float32 conversion ``im_to_torch`` :24-27, template-box bookkeeping :81-111) and
  * the third-party arithmetic it calls: ``cv2.resize(uint8 HxWx3, (model_sz, model_sz))`` with the default INTER_LINEAR.
    OpenCV is not vendored in the reference (``requirements``: opencv-python, unpinned; this container has opencv-python
    4.13.0).  The published algorithm (modules/imgproc/src/resize.cpp: ``resizeGeneric_`` with ``HResizeLinear`` /
    ``VResizeLinear<uchar,int,short>``, ``INTER_RESIZE_COEF_BITS = 11``) is restated in ``resize_linear_u8``:
      - ``scale = 1 / (dsize / ssize)`` in double; per destination index ``f = (float)((d + 0.5) * scale - 0.5)``,
        ``s = floor(f)``, ``f -= s``; coefficients ``round_half_even((1 - f) * 2048)``, ``round_half_even(f * 2048)`` (int16);
      - horizontally, taps left of the image / at the last column are folded (``s < 0 -> s = 0, f = 0``;
        ``s >= w - 1 -> s = w - 1, f = 0``); vertically the coefficients are kept and the two ROW indices are clamped;
      - ``H[y][x] = S[y][s] * a0 + S[y][s + 1] * a1``;  ``dst = (((b0 * (H0 >> 4)) >> 16) + ((b1 * (H1 >> 4)) >> 16) + 2) >> 2``;
      - an exact 2x down-scale is routed to INTER_AREA: ``(p00 + p01 + p10 + p11 + 2) >> 2`` (resize.cpp, "is_area_fast").
Only ``tests/`` may import this file; nothing under ``usot_b200/`` does.

Pinning: ``oracle/gen_crop_golden.py`` checks ``resize_linear_u8`` bit for bit against the LIVE ``cv2.resize`` (IPP on and off)
and ``get_subwindow_tracking`` against the LIVE reference function imported from /root/reference, on seeded random frames,
and writes small fixtures to ``tests/golden/crop_golden.npz``.
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def linear_coeffs(ssize: int, dsize: int, vertical: bool):
    """Source index and the two int16 fixed-point weights per destination index (resize.cpp, linear branch)."""
    scale = np.float64(1.0) / (np.float64(dsize) / np.float64(ssize))
    ofs = np.zeros(dsize, np.int64)
    c0 = np.zeros(dsize, np.int64)
    c1 = np.zeros(dsize, np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(np.floor(f))
        f = np.float32(f - np.float32(s))
        if not vertical:
            if s < 0:
                s, f = 0, np.float32(0)
            if s >= ssize - 1:
                s, f = ssize - 1, np.float32(0)
        ofs[d] = s
        c0[d] = int(np.rint(np.float32((np.float32(1.0) - f) * np.float32(COEF_SCALE))))
        c1[d] = int(np.rint(np.float32(f * np.float32(COEF_SCALE))))
    return ofs, c0, c1


def resize_linear_u8(src: np.ndarray, dsize: int) -> np.ndarray:
    """cv2.resize(src, (dsize, dsize)) for a square uint8 HxWxC image, default interpolation."""
    h, w, _ = src.shape
    s = src.astype(np.int64)
    if h == dsize and w == dsize:
        return src.copy()
    if h == 2 * dsize and w == 2 * dsize:
        return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx, a0, a1 = linear_coeffs(w, dsize, vertical=False)
    sy, b0, b1 = linear_coeffs(h, dsize, vertical=True)
    sx1 = np.minimum(sx + 1, w - 1)
    hrow = s[:, sx, :] * a0[None, :, None] + s[:, sx1, :] * a1[None, :, None]
    h0 = hrow[np.clip(sy, 0, h - 1)]
    h1 = hrow[np.clip(sy + 1, 0, h - 1)]
    out = (((b0[:, None, None] * (h0 >> 4)) >> 16) + ((b1[:, None, None] * (h1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def context_window(pos, original_sz):
    """track_utils.py:41-47 (python's round = round-half-even)."""
    c = (original_sz + 1) / 2
    xmin = round(pos[0] - c)
    ymin = round(pos[1] - c)
    return int(xmin), int(ymin)


def get_subwindow_tracking(im: np.ndarray, pos, model_sz: int, original_sz: int, avg_chans, target_sz=None, need_bbox=False):
    """Restatement of track_utils.py:30-119.  Returns (CHW float32 patch, crop_info)."""
    crop_info = {}
    sz = original_sz
    r, c, k = im.shape
    xmin, ymin = context_window(pos, original_sz)
    xmax, ymax = xmin + sz - 1, ymin + sz - 1
    left_pad = int(max(0.0, -xmin))
    top_pad = int(max(0.0, -ymin))
    right_pad = int(max(0.0, xmax - c + 1))
    bottom_pad = int(max(0.0, ymax - r + 1))
    xmin_p, xmax_p, ymin_p, ymax_p = xmin + left_pad, xmax + left_pad, ymin + top_pad, ymax + top_pad
    # the padded canvas is uint8: the float channel means are truncated on assignment (track_utils.py:58-70)
    fill = np.asarray(avg_chans, np.float64).astype(np.uint8)
    yy = np.arange(ymin, ymin + sz)
    xx = np.arange(xmin, xmin + sz)
    inside = ((yy >= 0) & (yy < r))[:, None] & ((xx >= 0) & (xx < c))[None, :]
    patch = np.where(inside[:, :, None], im[np.clip(yy, 0, r - 1)[:, None], np.clip(xx, 0, c - 1)[None, :], :], fill[None, None, :])
    patch = patch.astype(np.uint8)
    out = resize_linear_u8(patch, model_sz) if model_sz != original_sz else patch
    if target_sz is not None:
        t_xmin = round(pos[0] - target_sz[0] / 2)
        t_xmax = round(pos[0] + target_sz[0] / 2)
        t_ymin = round(pos[1] - target_sz[1] / 2)
        t_ymax = round(pos[1] + target_sz[1] / 2)
        crop_info["original_image_bbox"] = [t_xmin, t_ymin, t_xmax, t_ymax]
        if need_bbox:
            patch_sz = patch.shape[0]
            x_slope = patch_sz / (xmax_p - xmin_p)
            y_slope = patch_sz / (ymax_p - ymin_p)
            scale_resize = out.shape[0] / patch_sz
            crop_info["template_bbox"] = [scale_resize * (left_pad - 1 + x_slope * (t_xmin - xmin_p)),
                                          scale_resize * (top_pad - 1 + y_slope * (t_ymin - ymin_p)),
                                          scale_resize * (left_pad - 1 + x_slope * (t_xmax - xmin_p)),
                                          scale_resize * (top_pad - 1 + y_slope * (t_ymax - ymin_p))]
    crop_info["crop_cords"] = [xmin_p, xmax_p, ymin_p, ymax_p]
    crop_info["pad_info"] = [top_pad, left_pad, r, c]
    return np.ascontiguousarray(out.transpose(2, 0, 1)).astype(np.float32), crop_info
