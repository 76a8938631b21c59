"""BASELINE config 4 (script, not a test): cycle-memory training forward, 3 memory frames, 16 samples per GPU, crops sharded
across the ranks, ONE NCCL all-gather of z_f overlapped with the search / memory backbones (usot_b200.dist).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/bench_cycle.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from usot_b200 import USOT  # noqa: E402
from usot_b200.dist import cycle_forward_sharded  # noqa: E402
from usot_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, M = 16, 3
    net = USOT({"mem_size": M, "pr_pool": True}, precision=os.environ.get("USOT_B200_PRECISION", "fp16x3"))
    net.load_state_dict(synthetic_state_dict("damp025"))
    net = net.eval().cuda()
    z, x, tb, sb = synthetic_inputs(100 + rank, B, n_templates=B)
    g = torch.Generator().manual_seed(200 + rank)
    label = torch.zeros(B, 25, 25)
    label[:, 10:15, 10:15] = 1.0
    rw = torch.zeros(B, 25, 25)
    rw[:, 11:14, 11:14] = 1.0
    batch = dict(template=z, search=x, search_memory=torch.rand(B, M, 3, 255, 255, generator=g) * 255.0, label=label,
                 reg_target=torch.rand(B, 25, 25, 4, generator=g) * 40 + 5, reg_weight=rw, template_bbox=tb, search_bbox=sb)
    batch = {k: v.to(dev) for k, v in batch.items()}
    for _ in range(3):
        losses = cycle_forward_sharded(net, batch)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 10
    e0.record()
    for _ in range(steps):
        losses = cycle_forward_sharded(net, batch)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": "cycle-memory training forward (BASELINE config 4)", "n_gpus": world, "global_batch": B * world,
                          "memory_frames": M, "ms_per_step": float(ms), "samples_per_s": B * world / float(ms) * 1e3,
                          "search_crops_per_s": B * world * (1 + M) / float(ms) * 1e3, "losses": [float(v) for v in losses]}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
