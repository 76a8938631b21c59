"""Where does a training step (BASELINE config 4 shard: 16 samples + 3 memory frames) spend its time?  Prints the host enqueue time per step
(no synchronisation), the synchronised step time, and the device-time breakdown by kernel from torch.profiler.
    python tools/train_step_profile.py [--graph]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from usot_b200 import USOT  # noqa: E402
from usot_b200.dist import GradientReducer, train_step_sharded  # noqa: E402
from usot_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402


def main():
    B, M = 16, 3
    dev = torch.device("cuda", 0)
    net = USOT({"mem_size": M, "pr_pool": True}, precision="fp16x3")
    net.load_state_dict(synthetic_state_dict("damp025"))
    net = net.cuda().train()
    z, x, tb, sb = synthetic_inputs(7, B, n_templates=B)
    g = torch.Generator().manual_seed(200)
    label = torch.zeros(B, 25, 25); label[:, 10:15, 10:15] = 1.0
    rw = torch.zeros(B, 25, 25); rw[:, 11:14, 11:14] = 1.0
    batch = dict(template=z, search=x, search_memory=torch.rand(B, M, 3, 255, 255, generator=g) * 255.0, label=label,
                 reg_target=torch.rand(B, 25, 25, 4, generator=g) * 40 + 5, reg_weight=rw, template_bbox=tb, search_bbox=sb)
    batch = {k: v.to(dev) for k, v in batch.items()}
    red = GradientReducer(net.parameters())
    opt = torch.optim.SGD(net.parameters(), lr=1e-6, momentum=0.9)
    step = lambda: train_step_sharded(net, batch, red, opt)
    if "--graph" in sys.argv:
        from usot_b200.dist import GraphedTrainStep
        gstep = GraphedTrainStep(net, red, opt, batch)
        step = lambda: gstep(batch)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        step()
    host = (time.perf_counter() - t0) / 5
    torch.cuda.synchronize()
    total = (time.perf_counter() - t0) / 5
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    rows = sorted(((e.self_device_time_total, e.count, e.key) for e in prof.key_averages() if e.self_device_time_total > 0), reverse=True)
    dev_ms = sum(r[0] for r in rows) / 1e3
    out = {"host_enqueue_ms_per_step": host * 1e3, "step_ms": total * 1e3, "device_busy_ms": dev_ms,
           "top_kernels": [{"ms": round(r[0] / 1e3, 3), "count": r[1], "name": r[2][:90]} for r in rows[:28]]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
