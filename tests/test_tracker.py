"""Per-video tracker loop (lib/tracker/usot_tracker.py): the oracle tracker against the trace of the LIVE reference tracker
(oracle/gen_tracker_pin.py -> tests/golden/tracker_trace.npz), and on the GPU the device-side ``usot_b200.tracker.USOTTracker``
(GPU crops, device memory queue, fused post-process) against the same trace."""
import os
import types

import numpy as np
import pytest
import torch

import tracker_oracle as T
from helpers import GOLD, load_weights


def _fixture(tag=""):
    """tag "" = ordinary target (255-pixel search window), "small_" = target below 0.4 % of the frame (271-pixel window, R = 27)."""
    z = np.load(os.path.join(GOLD, "tracker_trace.npz"))
    g = {k[len(tag):]: z[k] for k in z.files if k.startswith(tag) and (tag or not k.startswith("small_"))}
    g["n_frames"] = z["n_frames"]
    frames, pos0, sz0 = T.synthetic_video(seed=int(g["video_seed"]), n_frames=int(g["n_frames"]), box=tuple(int(v) for v in g["box"]))
    assert np.array_equal(pos0, g["pos0"]) and np.array_equal(sz0, g["sz0"])
    return g, frames, pos0, sz0


@pytest.mark.parametrize("tag,size", [("", 255), ("small_", 271)])
def test_oracle_tracker_reproduces_reference_trace(tag, size):
    g, frames, pos0, sz0 = _fixture(tag)
    state = T.tracker_init(frames[0], pos0.copy(), sz0.copy(), T.OracleNet(load_weights("damp025")))
    assert state['p'].instance_size == size and state['p'].score_size == (size - 127) // 8 + 9 and len(state['init_features']) == 2
    rows = []
    for im in frames[1:]:
        state = T.tracker_track(state, im)
        rows.append(np.concatenate([state['target_pos'], state['target_sz'], [state['cls_score']]]))
    # same arithmetic as the reference; the tolerance only covers thread-count dependent summation order inside torch
    assert np.abs(np.array(rows) - g["trace"]).max() <= 1e-3
    assert len(state['memory_confidences']) == len(frames)


def test_tracker_mirror_host_logic():
    """Pieces of the device tracker that need no GPU: config, grids and the pooling-box conversions equal the oracle's."""
    from usot_b200.tracker import USOTConfig, USOTTracker, load_test_config, python2round
    cfg = load_test_config("USOT")
    assert cfg["small_sz"] == 255 and cfg["big_sz"] == 271 and cfg["mem_queue_size"] == 7
    p, po = USOTConfig(), T.USOTConfig()
    p.update(cfg)
    for size in (255, 271):
        p.instance_size = po.instance_size = size
        p.renew(); po.renew()
        p.sf_size = po.sf_size = p.score_size
        assert p.score_size == po.score_size == (25 if size == 255 else 27)
        tr = USOTTracker(types.SimpleNamespace(arch="USOT"))
        tr.grids(p)
        g = T.Grids(po)
        box = [20.5, 31.25, 190.0, 260.75]
        assert np.array_equal(tr.pool_label_search(p, box), T.pool_label_search(po, g, box))
        assert np.array_equal(tr.pool_label_template(p, box), T.pool_label_template(po, g, box))
    assert python2round(2.5) == 3.0 and python2round(3.5) == 4.0 and python2round(-2.5) == -3.0 and python2round(2.4) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("precision,fused,tag", [("fp16x3", True, ""), ("fp32", True, ""), ("fp16x3", False, ""), ("fp16x3", True, "small_"),
                                                 ("fp32", False, "small_")])
def test_gpu_tracker_follows_reference_trace(precision, fused, tag):
    from usot_b200 import USOT
    from usot_b200.tracker import USOTTracker
    g, frames, pos0, sz0 = _fixture(tag)
    net = USOT(precision=precision)
    net.load_state_dict(load_weights("damp025"))
    net = net.eval().cuda()
    tracker = USOTTracker(types.SimpleNamespace(arch="USOT"))
    tracker.fused_frame = fused  # one usot_engine_track_frame call per frame vs the op-by-op path
    state = tracker.init(frames[0], pos0.copy(), sz0.copy(), net)
    assert tuple(net.zf.shape) == (1, 256, 7, 7) and state['p'].instance_size == (271 if tag else 255)
    rows = []
    for im in frames[1:]:
        state = tracker.track(state, im)
        rows.append(np.concatenate([state['target_pos'], state['target_sz'], [state['cls_score']]]))
    rows, ref = np.array(rows), g["trace"]
    # the arg-max cell and every rounded crop coordinate must coincide with the reference (fixture chosen with margins, see
    # gen_tracker_pin.py); what remains is the 1e-3-class arithmetic difference of the maps: a few hundredths of a pixel
    assert np.abs(rows[:, :4] - ref[:, :4]).max() <= 0.1, np.abs(rows - ref).max(axis=0)
    assert np.abs(rows[:, 4] - ref[:, 4]).max() <= 2e-3
    assert len(state['memory_confidences']) == len(frames) and len(state['memory_queue'].selected_rows()[0]) == 7


@pytest.mark.gpu
def test_fused_frame_equals_op_by_op_path():
    """usot_engine_track_frame must produce exactly what the separate crop / track / postprocess / PrPool calls produce."""
    from usot_b200 import USOT
    from usot_b200.tracker import USOTTracker
    g, frames, pos0, sz0 = _fixture()
    traces = []
    for fused in (True, False):
        net = USOT(precision="fp16x3")
        net.load_state_dict(load_weights("damp025"))
        net = net.eval().cuda()
        tracker = USOTTracker(types.SimpleNamespace(arch="USOT"))
        tracker.fused_frame = fused
        state = tracker.init(frames[0], pos0.copy(), sz0.copy(), net)
        rows = []
        for im in frames[1:4]:
            state = tracker.track(state, im)
            rows.append(np.concatenate([state['target_pos'], state['target_sz'], [state['cls_score']]]))
        q = state['memory_queue']
        traces.append((np.array(rows), q._buf[: q._n].clone()))
    assert np.array_equal(traces[0][0], traces[1][0])
    assert torch.allclose(traces[0][1], traces[1][1], rtol=0, atol=1e-6)  # pool box: float32 here vs numpy's float64 intermediate
