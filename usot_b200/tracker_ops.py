"""Device-side tensor path of the per-frame tracker update (lib/tracker/usot_tracker.py:137-163).

``postprocess`` replaces the reference's 3 blocking D2H copies + numpy post-processing by one kernel; the caller reads the
8 result doubles back with a single ``.cpu()`` (or keeps them on the device).

``get_subwindow_tracking`` / ``crop_resize`` replace the host crop in front of the model (lib/utils/track_utils.py:30-119:
context window, average-colour padding, ``cv2.resize``, ``im_to_torch``) by one kernel whose output is bit-identical to the
reference patch; the frame is uploaded once as uint8 (0.9 MB for 480x640) and every crop of that frame -- template, search,
several scales or several videos -- is produced on the device as the fp32 NCHW tensor ``USOT.template()/track()`` take."""
import numpy as np
import torch

from . import _lib
from .ops import _need_cuda, _stream, as_nhwc


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


class MemoryQueue:
    """Device-resident memory queue with the reference's sampling rule (lib/tracker/usot_tracker.py:222-256).

    The reference keeps every pooled feature on the HOST (``.cpu()`` each frame, usot_tracker.py:106,123,199) and re-uploads
    the N_q selected ones every frame (351 KB H2D).  Here the features never leave the GPU: they live in one growing NHWC
    buffer, the confidences (host scalars the selection rule needs) stay on the host, and ``select()`` gathers the N_q rows
    with a single index_select on the device.  The selected indices are exactly the reference's, including its documented
    start/end-index quirk (usot_tracker.py:239-242).
    """

    def __init__(self, init_features, mem_queue_size=7, capacity=512):
        """init_features: [feature, flipped_feature], each (1,256,7,7) (any strides) on the device."""
        f0 = init_features[0]
        self.device = f0.device
        self.nq = int(mem_queue_size)
        self._buf = torch.empty((capacity, 7, 7, f0.shape[1]), dtype=torch.float32, device=self.device)
        self._n = 0
        self.confidences = []
        self._init = [self._append_raw(f) for f in init_features]  # rows 0, 1
        self._first = self._init[0]  # state['memory_features'] = [memory_feature]
        self._mem_rows = [self._first]
        self.confidences = [0.9]

    def __len__(self):
        """len(state['memory_features']) of the reference."""
        return len(self._mem_rows)

    def _append_raw(self, feat_nchw):
        if self._n == self._buf.shape[0]:
            new = torch.empty((2 * self._n,) + tuple(self._buf.shape[1:]), dtype=torch.float32, device=self.device)
            new[: self._n].copy_(self._buf)
            self._buf = new
        self._buf[self._n].copy_(feat_nchw[0].permute(1, 2, 0))
        self._n += 1
        return self._n - 1

    def next_row(self):
        """(7,7,C) view of the next free row of the device buffer (grown if full): usot_engine_track_frame writes the new memory
        feature straight into it; ``commit_row`` then makes it part of the queue."""
        if self._n == self._buf.shape[0]:
            new = torch.empty((2 * self._n,) + tuple(self._buf.shape[1:]), dtype=torch.float32, device=self.device)
            new[: self._n].copy_(self._buf)
            self._buf = new
        return self._buf[self._n]

    def commit_row(self, confidence):
        self._mem_rows.append(self._n)
        self._n += 1
        self.confidences.append(float(confidence))

    def append(self, feature, confidence):
        """state['memory_features'].append(feat_mem); state['memory_confidences'].append(confidence)."""
        self._mem_rows.append(self._append_raw(feature))
        self.confidences.append(float(confidence))

    def selected_rows(self):
        """Buffer rows of the N_q memory templates and their confidences, in the reference's order."""
        rows, scores = list(self._init), [0.9, 0.9]
        mem_rows, conf = self._mem_rows, self.confidences
        n = len(conf)
        upd = self.nq - 3
        if n <= 1:
            rows += [mem_rows[0]] * (upd + 1)
            scores += [conf[0]] * (upd + 1)
        else:
            gap = (n - 1) / upd
            for i in range(upd):
                start = min(int(int(i * gap) * n), n - 1)
                end = min(int(int((i + 1) * gap) * n), n - 1)
                if start >= end:
                    rows.append(mem_rows[start])
                    scores.append(conf[start])
                else:
                    k = int(np.argmax(np.array(conf[start:end]))) + start
                    rows.append(mem_rows[k])
                    scores.append(conf[k])
            rows.append(mem_rows[-1])
            scores.append(conf[-1])
        return rows, scores

    def select(self):
        """(template_mem, score_mem) as USOT.track expects them: (N_q,256,7,7) channels-last view and (1,N_q), both on the device."""
        rows, scores = self.selected_rows()
        idx = torch.tensor(rows, dtype=torch.long, device=self.device)
        mem = self._buf.index_select(0, idx)  # (N_q,7,7,C) contiguous NHWC
        return mem.permute(0, 3, 1, 2), torch.tensor([scores], dtype=torch.float32, device=self.device)


def smooth_update(res, target_pos, target_sz, scale_z, p):
    """Host part of USOTTracker.update after the arg-max (usot_tracker.py:165-193): res = the 8 doubles of ``postprocess``;
    target_sz is the size already multiplied by scale_z.  Returns (new target_pos, new target_sz, confidence)."""
    x1, y1, x2, y2, penalty, score = res[2], res[3], res[4], res[5], res[6], res[7]
    half = p.instance_size // 2
    dx, dy = ((x1 + x2) / 2 - half) / scale_z, ((y1 + y2) / 2 - half) / scale_z
    pw, ph = (x2 - x1) / scale_z, (y2 - y1) / scale_z
    tsz = np.asarray(target_sz, dtype=np.float64) / scale_z
    lr = penalty * score * p.lr
    new_pos = np.array([target_pos[0] + dx, target_pos[1] + dy])
    blended = np.array([pw * lr + (1 - lr) * tsz[0], ph * lr + (1 - lr) * tsz[1]])
    new_sz = tsz * (1 - lr) + lr * blended
    return new_pos, new_sz, float(score)


def track_frame(net, frame, context_xmin, context_ymin, original_sz, fill, queue, window, target_sz, p):
    """One tracker frame in ONE library call (usot_engine_track_frame): crop -> memory gather -> track() -> post-process ->
    PrPool of the new memory feature into the queue's next row; nothing is copied to or from the host.  ``frame`` (1,H,W,3)
    uint8 CUDA, ``fill`` (1,3) uint8 CUDA, ``window`` (R,R) float64 CUDA, ``target_sz`` already multiplied by scale_z.
    Returns the (8,) float64 CUDA result of ``postprocess``; the caller reads it (one 64-byte D2H copy) and then calls
    ``queue.commit_row(confidence)``."""
    import ctypes
    eng = net._engine(frame.device)
    if net.zf is None or net.zf.shape[0] != 1:
        raise RuntimeError("track_frame needs a single template (call template() first)")
    zf = as_nhwc(net.zf)  # net.zf is a channels-last view of a contiguous NHWC tensor: no copy
    rows, _ = queue.selected_rows()
    rows_c = (ctypes.c_int32 * len(rows))(*rows)
    out_row = queue.next_row()
    result = torch.empty(8, dtype=torch.float64, device=frame.device)
    with torch.cuda.device(frame.device):
        _lib.check(_lib.load().usot_engine_track_frame(
            eng._h, _lib.ptr(frame), frame.shape[1], frame.shape[2], int(context_xmin), int(context_ymin), int(original_sz), _lib.ptr(fill),
            int(p.instance_size), _lib.ptr(zf), _lib.ptr(queue._buf), rows_c, len(rows), _lib.ptr(window), float(target_sz[0]),
            float(target_sz[1]), float(p.ratio), float(p.penalty_k), float(p.window_influence), int(p.total_stride), _lib.ptr(result),
            _lib.ptr(out_row), _stream(frame)))
    return result


def update_device(net, x_crops, target_pos, target_sz, window, scale_z, p, queue):
    """Device-side version of USOTTracker.update (lib/tracker/usot_tracker.py:133-200) for one frame.

    Same inputs / outputs as the reference method, except that the memory templates come from a ``MemoryQueue`` (features stay on
    the GPU) and ``window`` is the float64 CUDA tensor of ``cosine_window``.  One blocking D2H copy per frame (8 doubles)
    instead of the reference's four; the pooled memory feature is returned on the device.
    ``target_sz`` is the target size already multiplied by ``scale_z`` (as the reference passes it, usot_tracker.py:258-259).
    """
    template_mem, score_mem = queue.select()
    cls_score, bbox_pred, cls_memory, xf = net.track(x_crops, template_mem=template_mem, score_mem=score_mem)
    res = postprocess(cls_score, bbox_pred, cls_memory, window, target_sz, instance_size=p.instance_size, ratio=p.ratio,
                      penalty_k=p.penalty_k, window_influence=p.window_influence).cpu().numpy()
    new_pos, new_sz, score = smooth_update(res, target_pos, target_sz, scale_z, p)
    x1, y1, x2, y2 = res[2], res[3], res[4], res[5]
    # memory feature of the predicted box, PrPool'ed from xf on the device (usot_tracker.py:196-199, pool_label_search :329-350)
    sf = p.score_size
    axis0 = (0 - sf // 2) * p.total_stride + p.instance_size // 2
    axis1 = (sf - 1 - sf // 2) * p.total_stride + p.instance_size // 2
    slope = (2 * (sf // 2)) / (axis1 - axis0)
    gap = 1.0 / slope
    box = np.clip(np.array([x1, y1, x2, y2], np.float32), a_min=axis0 - gap, a_max=axis1 + gap)
    pool_box = torch.from_numpy(((box - axis0) * slope).astype(np.float32)[None]).to(x_crops.device)
    feat_mem = net.extract_memory_feature(xf=xf, search_bbox=pool_box)
    return new_pos, new_sz, float(score), feat_mem


def upload_frame(im, device="cuda"):
    """cv2 frame (H,W,3) -- or a stack (F,H,W,3) -- uint8 numpy -> (F,H,W,3) uint8 CUDA tensor (accepts an already-uploaded tensor)."""
    if isinstance(im, np.ndarray):
        if im.dtype != np.uint8 or im.ndim not in (3, 4) or im.shape[-1] != 3:
            raise AssertionError("frames must be (H, W, 3) or (F, H, W, 3) uint8 arrays as cv2.imread returns them")
        im = torch.from_numpy(np.ascontiguousarray(im)).to(device, non_blocking=True)
    _need_cuda(im)
    if im.dtype != torch.uint8:
        raise AssertionError("frames must be uint8, got {}".format(im.dtype))
    return im.reshape((-1,) + tuple(im.shape[-3:])).contiguous()


def crop_resize(frames, crops, fill, model_sz):
    """Batched crops.  frames (F,H,W,3) uint8 CUDA; crops (n,4) int32 [frame index, context_xmin, context_ymin, original_sz]
    (window in frame coordinates before padding); fill (n,3) uint8 = truncated channel means.  Returns (n,3,model_sz,model_sz)
    float32 CUDA, bit-identical to get_subwindow_tracking(...)[0] of the reference for each window."""
    _need_cuda(frames, crops, fill)
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.shape[3] == 3
    assert crops.dtype == torch.int32 and crops.dim() == 2 and crops.shape[1] == 4
    assert fill.dtype == torch.uint8 and tuple(fill.shape) == (crops.shape[0], 3)
    frames, crops, fill = frames.contiguous(), crops.contiguous(), fill.contiguous()
    n = crops.shape[0]
    out = torch.empty((n, 3, int(model_sz), int(model_sz)), dtype=torch.float32, device=frames.device)
    with torch.cuda.device(frames.device):
        _lib.check(_lib.load().usot_crop_resize(_lib.ptr(frames), frames.shape[0], frames.shape[1], frames.shape[2], _lib.ptr(crops),
                                                _lib.ptr(fill), n, int(model_sz), _lib.ptr(out), _stream(frames)))
    return out


def crop_geometry(im_shape, pos, model_sz, original_sz, target_sz=None, need_bbox=False):
    """Host bookkeeping of get_subwindow_tracking (track_utils.py:41-55,81-115): returns (context_xmin, context_ymin) in frame
    coordinates before padding, and the reference's crop_info dict (without the never-filled 'empty_mask' canvas)."""
    if isinstance(pos, float):
        pos = [pos, pos]
    sz = original_sz
    r, c = int(im_shape[0]), int(im_shape[1])
    half = (original_sz + 1) / 2
    xmin = round(pos[0] - half)
    xmax = xmin + sz - 1
    ymin = round(pos[1] - half)
    ymax = ymin + sz - 1
    left_pad = int(max(0., -xmin))
    top_pad = int(max(0., -ymin))
    right_pad = int(max(0., xmax - c + 1))
    bottom_pad = int(max(0., ymax - r + 1))
    info = {}
    cxmin, cxmax, cymin, cymax = xmin + left_pad, xmax + left_pad, ymin + top_pad, ymax + top_pad
    if target_sz is not None:
        t_xmin = round(pos[0] - target_sz[0] / 2)
        t_xmax = round(pos[0] + target_sz[0] / 2)
        t_ymin = round(pos[1] - target_sz[1] / 2)
        t_ymax = round(pos[1] + target_sz[1] / 2)
        info['original_image_bbox'] = [t_xmin, t_ymin, t_xmax, t_ymax]
        if need_bbox:
            patch_sz = int(cymax + 1) - int(cymin)
            x_slope = patch_sz / (cxmax - cxmin)
            y_slope = patch_sz / (cymax - cymin)
            scale_resize = (model_sz if model_sz != original_sz else patch_sz) / patch_sz
            info['template_bbox'] = [scale_resize * (left_pad - 1 + x_slope * (t_xmin - cxmin)),
                                     scale_resize * (top_pad - 1 + y_slope * (t_ymin - cymin)),
                                     scale_resize * (left_pad - 1 + x_slope * (t_xmax - cxmin)),
                                     scale_resize * (top_pad - 1 + y_slope * (t_ymax - cymin))]
    info['crop_cords'] = [cxmin, cxmax, cymin, cymax]
    info['pad_info'] = [top_pad, left_pad, r, c]
    return int(xmin), int(ymin), info


class FrameStager:
    """Per-video staging of cv2 frames: one pinned host buffer + one device buffer of the frame size, reused every frame, so the
    upload is a host memcpy into pinned memory followed by an asynchronous H2D copy on the current stream."""

    def __init__(self, shape, device):
        self.pinned = torch.empty(tuple(shape), dtype=torch.uint8).pin_memory()
        self.dev = torch.empty((1,) + tuple(shape), dtype=torch.uint8, device=device)
        self.done = torch.cuda.Event()
        self._first = True

    def upload(self, im):
        if tuple(im.shape) != tuple(self.pinned.shape) or im.dtype != np.uint8:
            raise AssertionError("frame shape / dtype changed inside one video")
        if not self._first:
            self.done.synchronize()  # the previous H2D copy has left the pinned buffer
        self._first = False
        self.pinned.numpy()[...] = im
        self.dev[0].copy_(self.pinned, non_blocking=True)
        self.done.record()
        return self.dev


def get_subwindow_tracking(im, pos, model_sz, original_sz, avg_chans, target_sz=None, out_mode='torch', need_bbox=False, vis=False,
                           fill=None):
    """Drop-in for lib.utils.track_utils.get_subwindow_tracking (same arguments, same crop_info) producing the patch on the GPU.

    ``im`` is the cv2 frame (numpy uint8, uploaded here) or an ``upload_frame`` tensor, so a caller cropping several windows of
    one frame uploads it once.  Returns ((3,model_sz,model_sz) float32 CUDA tensor, crop_info); the reference returns the same
    values on the CPU and its caller moves them with ``.cuda()`` (usot_tracker.py:66-71,99-105,219-220)."""
    if out_mode != 'torch':
        raise NotImplementedError("usot_b200.tracker_ops.get_subwindow_tracking only produces torch CUDA tensors")
    frames = upload_frame(im)
    h, w = frames.shape[1], frames.shape[2]
    xmin, ymin, info = crop_geometry((h, w), pos, model_sz, original_sz, target_sz, need_bbox)
    dev = frames.device
    crops = torch.tensor([[0, xmin, ymin, int(original_sz)]], dtype=torch.int32).to(dev, non_blocking=True)
    if fill is None:  # (callers cropping many windows of one video pass the cached device copy: the means are per-video constants)
        fill = fill_color(avg_chans, dev)
    return crop_resize(frames[:1], crops, fill, model_sz)[0], info


def fill_color(avg_chans, device):
    """(1,3) uint8 device tensor of the channel means truncated as the reference's uint8 canvas stores them (track_utils.py:58-70)."""
    return torch.from_numpy(np.asarray(avg_chans, np.float64).astype(np.uint8).reshape(1, 3)).to(device, non_blocking=True)
