"""Device-side mirror of the reference tracker class (lib/tracker/usot_tracker.py:12-420): same constructor, ``init(im,
target_pos, target_sz, model)`` / ``track(state, im)`` calls and ``state`` keys, so ``scripts/test_usot.py:60-75`` runs
unchanged when this module shadows ``lib.tracker.usot_tracker`` (lib/tracker/usot_tracker.py in this repository).

What moved to the GPU (the "tensor path of lib/tracker"):
  * the frame is uploaded ONCE as uint8 and every crop of it (template, search, flipped first-frame search) is produced by the
    bit-exact crop/pad/resize kernel (tracker_ops.get_subwindow_tracking <- track_utils.py:30-119);
  * the memory queue lives on the device (tracker_ops.MemoryQueue, reference sampling rule usot_tracker.py:222-256) -- the
    reference round-trips every pooled feature through the host (.cpu() at :106,123,199 and .cuda() at :259);
  * sigmoid / ratio mixing / box decoding / penalties / window / arg-max run in one kernel (tracker_ops.postprocess <- :137-163);
    one 64-byte D2H copy per frame instead of three score/box maps.
The first-frame left/right flip is ``imgaug.augmenters.Fliplr(1)`` in the reference (:18-20,109-116); it is a column reversal of
the crop and ``x1' = W - x2, x2' = W - x1`` on the box, done here with ``torch.flip`` (no imgaug dependency).
Host arithmetic that stays on the host (target window geometry, size/position smoothing, clamps) is restated 1:1.
"""
import os
import sys

import numpy as np
import torch

from . import tracker_ops
from .tracker_ops import get_subwindow_tracking


def python2round(f):
    """lib/utils/track_utils.py:121-127."""
    if round(f + 1) - round(f) != 1:
        return f + abs(f) / f * 0.5
    return round(f)


TEST_DEFAULTS = {"penalty_k": 0.021, "lr": 0.730, "window_influence": 0.321, "small_sz": 255, "big_sz": 271, "ratio": 0.3,
                 "mem_queue_size": 7}  # experiments/test/USOT.yaml (TEST section)


def load_test_config(arch):
    """The TEST section of experiments/test/<arch>.yaml (usot_tracker.py:34-41).  The file is looked up next to every
    ``sys.path`` entry (the reference tree is on the path when its scripts run); without it the shipped values are used."""
    for base in list(sys.path) + [os.environ.get("USOT_REFERENCE", "")]:
        path = os.path.join(base or ".", "experiments", "test", "{}.yaml".format(arch))
        if os.path.isfile(path):
            import yaml
            with open(path, "r") as f:
                return dict(yaml.load(f.read(), Loader=yaml.FullLoader)["TEST"])
    return dict(TEST_DEFAULTS)


class USOTConfig(object):
    """usot_tracker.py:386-420."""
    penalty_k = 0.021
    window_influence = 0.321
    lr = 0.730
    windowing = 'cosine'
    exemplar_size = 127
    instance_size = 255
    total_stride = 8
    score_size = (instance_size - exemplar_size) // total_stride + 1 + 8
    context_amount = 0.5
    tf_size = 15
    sf_size = 25
    ratio = 0.3
    mem_queue_size = 7

    def update(self, newparam=None):
        if newparam:
            for key, value in newparam.items():
                setattr(self, key, value)
            self.renew()

    def renew(self):
        self.score_size = (self.instance_size - self.exemplar_size) // self.total_stride + 1 + 8


class USOTTracker(object):
    def __init__(self, info):
        super(USOTTracker, self).__init__()
        self.info = info
        # per-frame path: one usot_engine_track_frame call (default) or the op-by-op path through the Python façade
        self.fused_frame = os.environ.get("USOT_B200_TRACK_FRAME", "1") != "0"

    # ---- per-video initialisation (usot_tracker.py:22-131) ----
    def init(self, im, target_pos, target_sz, model):
        model.pr_pool = True
        state = dict()
        p = USOTConfig()
        state['im_h'] = im.shape[0]
        state['im_w'] = im.shape[1]
        cfg_benchmark = load_test_config(getattr(self.info, "arch", "USOT"))
        p.update(cfg_benchmark)
        p.renew()
        if ((target_sz[0] * target_sz[1]) / float(state['im_h'] * state['im_w'])) < 0.004:
            p.instance_size = cfg_benchmark['big_sz']
        else:
            p.instance_size = cfg_benchmark['small_sz']
        p.renew()
        p.sf_size = p.score_size
        self.grids(p)
        net = model
        dev = next(net.parameters()).device
        stager = tracker_ops.FrameStager(im.shape, dev)
        frame = stager.upload(im)  # one H2D copy per frame; all crops below are cut on the device

        wc_z = target_sz[0] + p.context_amount * sum(target_sz)
        hc_z = target_sz[1] + p.context_amount * sum(target_sz)
        s_z = round(np.sqrt(wc_z * hc_z))
        avg_chans = np.mean(im, axis=(0, 1))
        fill = tracker_ops.fill_color(avg_chans, dev)
        z_crop, crop_info = get_subwindow_tracking(frame, target_pos, p.exemplar_size, s_z, avg_chans, target_sz, need_bbox=True, fill=fill)
        template_bbox = self.pool_label_template(p, crop_info['template_bbox'])
        template_bbox = torch.from_numpy(np.asarray([template_bbox], np.float32)).to(dev)
        net.template(z_crop.unsqueeze(0), template_bbox=template_bbox)

        if p.windowing == 'cosine':
            window = np.outer(np.hanning(p.score_size), np.hanning(p.score_size))
        else:
            window = np.ones((int(p.score_size), int(p.score_size)))
        state['p'] = p
        state['net'] = net
        state['avg_chans'] = avg_chans
        state['window'] = window
        state['window_dev'] = torch.from_numpy(window).to(dev)
        state['frame_stager'] = stager
        state['fill_dev'] = fill
        state['target_pos'] = target_pos
        state['target_sz'] = target_sz

        _, _, s_x = self._search_window(p, target_sz)
        x_crop, crop_info = get_subwindow_tracking(frame, target_pos, p.instance_size, python2round(s_x), avg_chans, target_sz, need_bbox=True,
                                                   fill=fill)
        search_bbox = crop_info['template_bbox']
        pool = torch.from_numpy(np.asarray([self.pool_label_search(p, search_bbox)], np.float32)).to(dev)
        memory_feature = net.extract_memory_feature(ori_x=x_crop.unsqueeze(0), search_bbox=pool)
        # left/right flipped first-frame crop and box (Fliplr(1))
        width = x_crop.shape[2]
        x_crop_aug = torch.flip(x_crop, dims=[2])
        bbox_aug = [self.clip_number(width - search_bbox[2], _max=x_crop.shape[1]), self.clip_number(search_bbox[1], _max=x_crop.shape[2]),
                    self.clip_number(width - search_bbox[0], _max=x_crop.shape[1]), self.clip_number(search_bbox[3], _max=x_crop.shape[2])]
        pool_aug = torch.from_numpy(np.asarray([self.pool_label_search(p, bbox_aug)], np.float32)).to(dev)
        memory_feature_aug = net.extract_memory_feature(ori_x=x_crop_aug.unsqueeze(0), search_bbox=pool_aug)

        queue = tracker_ops.MemoryQueue([memory_feature, memory_feature_aug], mem_queue_size=p.mem_queue_size)
        state['init_features'] = [memory_feature, memory_feature_aug]
        state['memory_queue'] = queue                    # device-resident state['memory_features']
        state['memory_features'] = queue
        state['memory_confidences'] = queue.confidences  # same list object the queue appends to
        return state

    # ---- per frame (usot_tracker.py:202-276; update :133-200 runs in tracker_ops.update_device) ----
    def track(self, state, im):
        p = state['p']
        net = state['net']
        target_pos = state['target_pos']
        target_sz = state['target_sz']
        _, scale_z, s_x = self._search_window(p, target_sz)
        frame = state['frame_stager'].upload(im)
        queue = state['memory_queue']
        if self.fused_frame:
            # one library call per frame: crop -> memory gather -> track() -> post-process -> PrPool into the queue's next row
            xmin, ymin, _ = tracker_ops.crop_geometry(frame.shape[1:3], target_pos, p.instance_size, python2round(s_x))
            res = tracker_ops.track_frame(net, frame, xmin, ymin, python2round(s_x), state['fill_dev'], queue, state['window_dev'],
                                          target_sz * scale_z, p).cpu().numpy()
            target_pos, target_sz, confidence = tracker_ops.smooth_update(res, target_pos, target_sz * scale_z, scale_z, p)
            queue.commit_row(confidence)
        else:
            x_crop, _ = get_subwindow_tracking(frame, target_pos, p.instance_size, python2round(s_x), state['avg_chans'], fill=state['fill_dev'])
            target_pos, target_sz, confidence, feat_mem = tracker_ops.update_device(net, x_crop.unsqueeze(0), target_pos, target_sz * scale_z,
                                                                                    state['window_dev'], scale_z, p, queue)
            queue.append(feat_mem, confidence)
        target_pos[0] = max(0, min(state['im_w'], target_pos[0]))
        target_pos[1] = max(0, min(state['im_h'], target_pos[1]))
        target_sz[0] = max(10, min(state['im_w'], target_sz[0]))
        target_sz[1] = max(10, min(state['im_h'], target_sz[1]))
        state['target_pos'] = target_pos
        state['target_sz'] = target_sz
        state['cls_score'] = confidence
        state['p'] = p
        return state

    @staticmethod
    def _search_window(p, target_sz):
        """usot_tracker.py:86-93,209-216: (s_z, scale_z, s_x)."""
        hc_z = target_sz[1] + p.context_amount * sum(target_sz)
        wc_z = target_sz[0] + p.context_amount * sum(target_sz)
        s_z = np.sqrt(wc_z * hc_z)
        scale_z = p.exemplar_size / s_z
        d_search = (p.instance_size - p.exemplar_size) / 2
        pad = d_search / scale_z
        return s_z, scale_z, s_z + 2 * pad

    def clip_number(self, num, _max=127.0, _min=0.0):
        if num >= _max:
            return _max
        elif num <= _min:
            return _min
        return num

    def grids(self, p):
        """usot_tracker.py:288-327."""
        sz = p.score_size
        x, y = np.meshgrid(np.arange(0, sz) - np.floor(float(sz // 2)), np.arange(0, sz) - np.floor(float(sz // 2)))
        self.grid_to_search_x = x * p.total_stride + p.instance_size // 2
        self.grid_to_search_y = y * p.total_stride + p.instance_size // 2
        tf_sz = p.tf_size
        x, y = np.meshgrid(np.arange(0, tf_sz) - np.floor(float(tf_sz // 2)), np.arange(0, tf_sz) - np.floor(float(tf_sz // 2)))
        self.grid_to_template = {}
        self.grid_to_template_x = x * p.total_stride + p.exemplar_size // 2
        self.grid_to_template_y = y * p.total_stride + p.exemplar_size // 2
        sf_sz = p.sf_size
        self.search_area_x_axis = (np.arange(0, sf_sz) - np.floor(float(sf_sz // 2))) * p.total_stride + p.instance_size // 2

    def pool_label_template(self, p, bbox):
        """usot_tracker.py:329-337."""
        reg_min = self.grid_to_template_x[0][0]
        reg_max = self.grid_to_template_x[-1][-1]
        bbox = np.clip(np.array(bbox, np.float32), a_max=reg_max, a_min=reg_min)
        slope = 2 * (p.tf_size // 2) / (reg_max - reg_min)
        return (bbox - reg_min) * slope

    def pool_label_search(self, p, bbox):
        """usot_tracker.py:339-362 (the documented 25-cell axis on the 31-cell map is kept)."""
        reg_min = self.search_area_x_axis[0]
        reg_max = self.search_area_x_axis[-1]
        slope = 2 * (p.sf_size // 2) / (reg_max - reg_min)
        gap = 1.0 / slope
        bbox = np.clip(np.array(bbox, np.float32), a_max=reg_max + gap, a_min=reg_min - gap)
        return (bbox - reg_min) * slope
