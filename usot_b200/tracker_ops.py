"""Device-side tensor path of the per-frame tracker update (lib/tracker/usot_tracker.py:137-163).

``postprocess`` replaces the reference's 3 blocking D2H copies + numpy post-processing by one kernel; the caller reads the
8 result doubles back with a single ``.cpu()`` (or keeps them on the device)."""
import numpy as np
import torch

from . import _lib
from .ops import _need_cuda, _stream


def cosine_window(score_size, device):
    """np.outer(np.hanning, np.hanning) as the reference builds it (usot_tracker.py:74-75), float64 on the device."""
    return torch.from_numpy(np.outer(np.hanning(score_size), np.hanning(score_size))).to(device)


def postprocess(cls_score, bbox_pred, cls_memory, window, target_sz_scaled, instance_size=255, ratio=0.3, penalty_k=0.021,
                window_influence=0.321):
    """cls_score / cls_memory (1,1,R,R), bbox_pred (1,4,R,R) float32 CUDA; window (R,R) float64 CUDA; target_sz_scaled = target_sz*scale_z.
    Returns a float64 CUDA tensor [r_max, c_max, x1, y1, x2, y2, penalty, mixed_score]."""
    _need_cuda(cls_score, bbox_pred, cls_memory, window)
    assert cls_score.shape[0] == 1 and cls_score.dtype == torch.float32 and window.dtype == torch.float64
    r = cls_score.shape[-1]
    out = torch.empty(8, dtype=torch.float64, device=cls_score.device)
    with torch.cuda.device(cls_score.device):
        _lib.check(_lib.load().usot_tracker_postprocess(_lib.ptr(cls_score.contiguous()), _lib.ptr(cls_memory.contiguous()),
                                                        _lib.ptr(bbox_pred.contiguous()), _lib.ptr(window.contiguous()), r, int(instance_size),
                                                        float(target_sz_scaled[0]), float(target_sz_scaled[1]), float(ratio), float(penalty_k),
                                                        float(window_influence), _lib.ptr(out), _stream(cls_score)))
    return out
