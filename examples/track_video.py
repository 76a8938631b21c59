"""Track one (synthetic) video with the drop-in classes, exactly as scripts/test_usot.py:60-97 drives the reference:

    PYTHONPATH=<this repo>:<USOT checkout> python examples/track_video.py [--weights USOT_star.pth] [--precision fp16x3]

``lib.models.models.USOT`` and ``lib.tracker.usot_tracker.USOTTracker`` resolve to usot_b200 (namespace shadows); without a
checkpoint the seeded synthetic weights of the test-suite are used, so the boxes are meaningless but the whole device path runs:
uint8 frame upload -> GPU crop -> track() with the memory queue -> fused post-process -> PrPool of the new memory feature.
Needs a B200 (there is no CPU fallback)."""
import argparse
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import lib.models.models as models  # noqa: E402  (the shadow in this repo)
from lib.tracker.usot_tracker import USOTTracker  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--weights", default=None, help="reference checkpoint (USOT.pth / USOT_star.pth / checkpoint_e*.pth)")
ap.add_argument("--precision", default="fp16x3", choices=["fp32", "fp16x3", "fp16"])
ap.add_argument("--frames", type=int, default=100)
args = ap.parse_args()

net = models.__dict__["USOT"](precision=args.precision)
if args.weights:
    from usot_b200.checkpoint import load_pretrain
    load_pretrain(net, args.weights, print_unuse=False)
else:
    from usot_b200.synth import synthetic_state_dict
    net.load_state_dict(synthetic_state_dict("damp025"))
net = net.eval().cuda()

rng = np.random.default_rng(15)  # a textured bright box drifting over a noisy background
bg = np.kron(rng.integers(0, 120, (31, 41, 3)).astype(np.float64), np.ones((8, 8, 1)))[:240, :320] + rng.integers(0, 30, (240, 320, 3))
tex = rng.integers(150, 256, (48, 64, 3)).astype(np.float64)
frames = []
for t in range(8):
    f = bg.copy()
    f[90 + 3 * t:138 + 3 * t, 120 + 4 * t:184 + 4 * t] = tex
    frames.append(np.clip(f, 0, 255).astype(np.uint8))
target_pos, target_sz = np.array([152.0, 114.0]), np.array([64.0, 48.0])
tracker = USOTTracker(types.SimpleNamespace(arch="USOT"))
state = tracker.init(frames[0], target_pos.copy(), target_sz.copy(), net)
torch.cuda.synchronize()
t0 = time.perf_counter()
for f in range(1, args.frames):
    state = tracker.track(state, frames[f % len(frames)])
    if f <= 5:
        print("frame %d: centre (%.1f, %.1f) size (%.1f, %.1f) confidence %.3f" % (f, *state["target_pos"], *state["target_sz"], state["cls_score"]))
torch.cuda.synchronize()
print("%.2f ms per frame" % ((time.perf_counter() - t0) / (args.frames - 1) * 1e3))
