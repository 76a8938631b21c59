"""CPU oracle of the per-video tracker loop.  TEST INFRASTRUCTURE ONLY.

Restates ``USOTTracker`` / ``USOTConfig`` (/root/reference/lib/tracker/usot_tracker.py:12-420) on top of the other oracles:
the crops come from ``crop_oracle.get_subwindow_tracking`` (track_utils.py:30-119), the model calls from ``usot_oracle``
(``OracleNet`` exposes the reference façade ``template / track / extract_memory_feature`` over a state_dict) and the
tensor path of ``update`` from ``usot_oracle.tracker_update``.  The horizontal-flip augmentation of the first frame is
``imgaug.augmenters.Fliplr(1)`` in the reference (usot_tracker.py:18-20,109-116; imgaug is not vendored, 0.4.0 semantics:
image columns reversed, box ``x1' = W - x2``, ``x2' = W - x1``); it is restated directly.

Pinning: ``oracle/gen_tracker_pin.py`` drives the LIVE reference ``USOTTracker`` (reference model on the CPU behind harness
shims, a stand-in ``imgaug`` with the semantics above) and this file over the same synthetic video and requires identical
traces; the reference trace is stored in tests/golden/tracker_trace.npz.
"""
from __future__ import annotations

import numpy as np
import torch

import crop_oracle as C
import usot_oracle as O


class USOTConfig:
    """usot_tracker.py:386-420 with experiments/test/USOT.yaml applied (same values; adds small_sz / big_sz)."""
    penalty_k = 0.021
    window_influence = 0.321
    lr = 0.730
    windowing = 'cosine'
    exemplar_size = 127
    instance_size = 255
    total_stride = 8
    context_amount = 0.5
    tf_size = 15
    sf_size = 25
    ratio = 0.3
    mem_queue_size = 7
    small_sz = 255
    big_sz = 271

    def __init__(self):
        self.renew()

    def renew(self):
        self.score_size = (self.instance_size - self.exemplar_size) // self.total_stride + 1 + 8


class OracleNet:
    """The reference model façade (lib/models/models.py:173-206) over the CPU oracle."""

    def __init__(self, sd):
        self.sd = sd
        self.pr_pool = True
        self.zf = None

    def template(self, z, template_bbox=None):
        with torch.no_grad():
            self.zf = O.template(self.sd, z, template_bbox if self.pr_pool else None, pr_pool=self.pr_pool)

    def track(self, x, template_mem=None, score_mem=None):
        with torch.no_grad():
            return O.track(self.sd, self.zf, x, template_mem, score_mem)

    def extract_memory_feature(self, ori_x=None, xf=None, search_bbox=None):
        with torch.no_grad():
            return O.extract_memory_feature(self.sd, ori_x=ori_x, xf=xf, search_bbox=search_bbox)


def python2round(f):
    """track_utils.py:121-127."""
    if round(f + 1) - round(f) != 1:
        return f + abs(f) / f * 0.5
    return round(f)


def clip_number(num, _max=127.0, _min=0.0):
    return _max if num >= _max else (_min if num <= _min else num)


class Grids:
    """usot_tracker.py:299-327."""

    def __init__(self, p):
        sz = p.score_size
        x, y = np.meshgrid(np.arange(0, sz) - np.floor(float(sz // 2)), np.arange(0, sz) - np.floor(float(sz // 2)))
        self.grid_to_search_x = x * p.total_stride + p.instance_size // 2
        self.grid_to_search_y = y * p.total_stride + p.instance_size // 2
        tf = p.tf_size
        x, y = np.meshgrid(np.arange(0, tf) - np.floor(float(tf // 2)), np.arange(0, tf) - np.floor(float(tf // 2)))
        self.grid_to_template_x = x * p.total_stride + p.exemplar_size // 2
        self.search_area_x_axis = (np.arange(0, p.sf_size) - np.floor(float(p.sf_size // 2))) * p.total_stride + p.instance_size // 2


def pool_label_template(p, g, bbox):
    """usot_tracker.py:329-337."""
    reg_min, reg_max = g.grid_to_template_x[0][0], g.grid_to_template_x[-1][-1]
    bbox = np.clip(np.array(bbox, np.float32), a_max=reg_max, a_min=reg_min)
    return (bbox - reg_min) * (2 * (p.tf_size // 2) / (reg_max - reg_min))


def pool_label_search(p, g, bbox):
    """usot_tracker.py:339-362."""
    reg_min, reg_max = g.search_area_x_axis[0], g.search_area_x_axis[-1]
    slope = 2 * (p.sf_size // 2) / (reg_max - reg_min)
    gap = 1.0 / slope
    bbox = np.clip(np.array(bbox, np.float32), a_max=reg_max + gap, a_min=reg_min - gap)
    return (bbox - reg_min) * slope


def search_window(p, target_sz):
    """usot_tracker.py:86-93 / 209-216: (s_z, scale_z, s_x)."""
    hc_z = target_sz[1] + p.context_amount * sum(target_sz)
    wc_z = target_sz[0] + p.context_amount * sum(target_sz)
    s_z = np.sqrt(wc_z * hc_z)
    scale_z = p.exemplar_size / s_z
    pad = ((p.instance_size - p.exemplar_size) / 2) / scale_z
    return s_z, scale_z, s_z + 2 * pad


def select_memory(state, p):
    """usot_tracker.py:222-256: (list of memory features, list of scores) in the reference's order."""
    feats, conf = state['memory_features'], state['memory_confidences']
    template_mem = list(state['init_features'])
    score_mem = [0.9, 0.9]
    n = len(conf)
    upd = p.mem_queue_size - 3
    if n <= 1:
        template_mem += [feats[0]] * (upd + 1)
        score_mem += [conf[0]] * (upd + 1)
    else:
        gap = (n - 1) / upd
        for i in range(upd):
            start = min(int(int(i * gap) * n), n - 1)
            end = min(int(int((i + 1) * gap) * n), n - 1)
            if start >= end:
                template_mem.append(feats[start]); score_mem.append(conf[start])
            else:
                k = int(np.argmax(np.array(conf[start:end]))) + start
                template_mem.append(feats[k]); score_mem.append(conf[k])
        template_mem.append(feats[-1]); score_mem.append(conf[-1])
    return template_mem, score_mem


def tracker_init(im, target_pos, target_sz, net):
    """USOTTracker.init, usot_tracker.py:22-131."""
    net.pr_pool = True
    state = {'im_h': im.shape[0], 'im_w': im.shape[1]}
    p = USOTConfig()
    p.instance_size = p.big_sz if (target_sz[0] * target_sz[1]) / float(state['im_h'] * state['im_w']) < 0.004 else p.small_sz
    p.renew()
    p.sf_size = p.score_size
    g = Grids(p)
    wc_z = target_sz[0] + p.context_amount * sum(target_sz)
    hc_z = target_sz[1] + p.context_amount * sum(target_sz)
    s_z = round(np.sqrt(wc_z * hc_z))
    avg_chans = np.mean(im, axis=(0, 1))
    z_crop, info = C.get_subwindow_tracking(im, target_pos, p.exemplar_size, s_z, avg_chans, target_sz, need_bbox=True)
    template_bbox = torch.tensor(np.array([pool_label_template(p, g, info['template_bbox'])])).float()
    net.template(torch.from_numpy(z_crop).unsqueeze(0), template_bbox=template_bbox)
    window = np.outer(np.hanning(p.score_size), np.hanning(p.score_size))
    state.update(p=p, g=g, net=net, avg_chans=avg_chans, window=window, target_pos=target_pos, target_sz=target_sz)
    _, _, s_x = search_window(p, target_sz)
    x_crop, info = C.get_subwindow_tracking(im, target_pos, p.instance_size, python2round(s_x), avg_chans, target_sz, need_bbox=True)
    search_bbox = info['template_bbox']
    box = torch.tensor(np.array([pool_label_search(p, g, search_bbox)])).float()
    mem = net.extract_memory_feature(ori_x=torch.from_numpy(x_crop).unsqueeze(0), search_bbox=box)
    # left/right flipped first-frame crop (imgaug Fliplr(1) on image and box)
    w_img = x_crop.shape[2]
    x_aug = np.ascontiguousarray(x_crop[:, :, ::-1])
    bx1, bx2 = w_img - search_bbox[2], w_img - search_bbox[0]
    box_aug = [clip_number(bx1, _max=p.instance_size), clip_number(search_bbox[1], _max=p.instance_size),
               clip_number(bx2, _max=p.instance_size), clip_number(search_bbox[3], _max=p.instance_size)]
    box_aug = torch.tensor(np.array([pool_label_search(p, g, box_aug)])).float()
    mem_aug = net.extract_memory_feature(ori_x=torch.from_numpy(x_aug).unsqueeze(0), search_bbox=box_aug)
    state['init_features'] = [mem, mem_aug]
    state['memory_features'] = [mem]
    state['memory_confidences'] = [0.9]
    return state


def tracker_update(net, x_crops, target_pos, target_sz, window, scale_z, p, g, template_mem, score_mem):
    """USOTTracker.update, usot_tracker.py:133-200 (target_sz arrives multiplied by scale_z)."""
    cls_score, bbox_pred, cls_memory, xf = net.track(x_crops, template_mem=template_mem, score_mem=score_mem)
    r, c, pscore, penalty, cls, box = O.tracker_update(cls_score, bbox_pred, cls_memory, target_sz, window, instance_size=p.instance_size,
                                                       score_size=p.score_size, ratio=p.ratio, penalty_k=p.penalty_k,
                                                       window_influence=p.window_influence, stride=p.total_stride)
    x1, y1, x2, y2 = box
    diff_xs = ((x1 + x2) / 2 - p.instance_size // 2) / scale_z
    diff_ys = ((y1 + y2) / 2 - p.instance_size // 2) / scale_z
    pred_w, pred_h = (x2 - x1) / scale_z, (y2 - y1) / scale_z
    target_sz = target_sz / scale_z
    lr = penalty[r, c] * cls[r, c] * p.lr
    res_w = pred_w * lr + (1 - lr) * target_sz[0]
    res_h = pred_h * lr + (1 - lr) * target_sz[1]
    new_pos = np.array([target_pos[0] + diff_xs, target_pos[1] + diff_ys])
    new_sz = target_sz * (1 - lr) + lr * np.array([res_w, res_h])
    pool_box = torch.tensor(np.array([pool_label_search(p, g, [x1, y1, x2, y2])])).float()
    feat = net.extract_memory_feature(xf=xf, search_bbox=pool_box)
    top2 = np.sort(pscore.ravel())[-2:]
    return new_pos, new_sz, cls[r, c], feat, float(top2[1] - top2[0])


def tracker_track(state, im):
    """USOTTracker.track, usot_tracker.py:202-276.  Adds state['top2_gap'] (margin of the argmax, for choosing robust fixtures)."""
    p, g, net = state['p'], state['g'], state['net']
    target_pos, target_sz = state['target_pos'], state['target_sz']
    _, scale_z, s_x = search_window(p, target_sz)
    osz = python2round(s_x)
    x_crop, _ = C.get_subwindow_tracking(im, target_pos, p.instance_size, osz, state['avg_chans'])
    # distance of every rounded quantity of the crop geometry from its rounding boundary (fixture robustness, not reference logic)
    half = (osz + 1) / 2
    if len(state['memory_confidences']) <= 1:
        margin = 1.0  # first tracked frame: position and size are the caller's exact inputs
    else:
        vals = [s_x] + [v - half for v, hi in ((target_pos[0], state['im_w']), (target_pos[1], state['im_h'])) if 0 < v < hi]
        margin = min(abs(abs((v % 1.0) - 0.5)) for v in vals)
    feats, scores = select_memory(state, p)
    template_mem = torch.cat(feats, dim=0)
    score_mem = torch.tensor(scores).unsqueeze(0)
    target_pos, target_sz, conf, feat, gap = tracker_update(net, torch.from_numpy(x_crop).unsqueeze(0), target_pos, target_sz * scale_z,
                                                            state['window'], scale_z, p, g, template_mem, score_mem)
    state['memory_features'].append(feat)
    state['memory_confidences'].append(conf)
    target_pos[0] = max(0, min(state['im_w'], target_pos[0]))
    target_pos[1] = max(0, min(state['im_h'], target_pos[1]))
    target_sz[0] = max(10, min(state['im_w'], target_sz[0]))
    target_sz[1] = max(10, min(state['im_h'], target_sz[1]))
    state.update(target_pos=target_pos, target_sz=target_sz, cls_score=conf, top2_gap=gap, round_margin=margin)
    return state


def synthetic_video(seed=3, n_frames=6, h=240, w=320, box=(64, 48)):
    """Seeded uint8 frames: smooth-ish random background, a textured bright box (w, h) = ``box`` drifting by (+4, +3) px per
    frame.  Returns (frames, initial target_pos (cx, cy), target_sz (w, h)).  A box below 0.4 % of the frame area makes the
    tracker choose the 271-pixel search window (usot_tracker.py:43-48)."""
    rng = np.random.default_rng(seed)
    bw, bh = box
    bg = rng.integers(0, 120, (h // 8 + 1, w // 8 + 1, 3)).astype(np.float64)
    bg = np.kron(bg, np.ones((8, 8, 1)))[:h, :w] + rng.integers(0, 30, (h, w, 3))
    tex = rng.integers(150, 256, (bh, bw, 3)).astype(np.float64)
    frames = []
    for t in range(n_frames):
        f = bg.copy()
        y0, x0 = 90 + 3 * t, 120 + 4 * t
        f[y0:y0 + bh, x0:x0 + bw] = tex
        frames.append(np.clip(f, 0, 255).astype(np.uint8))
    return frames, np.array([120 + bw / 2.0, 90 + bh / 2.0]), np.array([float(bw), float(bh)])
