"""Per-frame latency of the device-side tracker loop (usot_b200.tracker.USOTTracker: uint8 frame upload -> GPU crop -> track()
with the 7-entry memory queue -> fused post-process -> PrPool of the new memory feature) on one B200, synthetic 480x640 video.
Prints one JSON line.  Not a bench.py replacement: BASELINE's metric is batch-256 crops/s; this is the batch-1 serial path."""
import json
import sys
import time
import types

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
from usot_b200 import USOT  # noqa: E402
from usot_b200.synth import synthetic_state_dict  # noqa: E402
from usot_b200.tracker import USOTTracker  # noqa: E402

precision = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
n_frames = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rng = np.random.default_rng(0)
frames = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(8)]
net = USOT(precision=precision)
net.load_state_dict(synthetic_state_dict("damp025"))
net = net.eval().cuda()
tr = USOTTracker(types.SimpleNamespace(arch="USOT"))
state = tr.init(frames[0], np.array([320.0, 240.0]), np.array([90.0, 70.0]), net)
for i in range(20):
    state = tr.track(state, frames[i % 8])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(n_frames):
    state = tr.track(state, frames[i % 8])
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n_frames
print(json.dumps({"what": "device tracker loop, batch 1, 480x640 frames, N_q=7", "precision": precision, "frames": n_frames,
                  "ms_per_frame": dt * 1e3, "fps": 1.0 / dt, "queue_len": len(state["memory_features"])}))
