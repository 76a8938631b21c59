"""One data-parallel training step of the cycle-memory model on the sm_100a engine (the reference's scripts/train_usot.py:196-236 loop body):

    python examples/train_step.py                                        # one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/train_step.py

``USOT.forward`` in train() mode returns losses with an autograd graph; ``loss.backward()`` runs this library's dgrad / wgrad / BatchNorm /
pooling / correlation / loss-gradient kernels; ``GradientReducer`` all-reduces the gradients bucket by bucket while backward is still running
(it replaces nn.DataParallel's reduce_add_coalesced, scripts/train_usot.py:318).  Synthetic weights and data (no checkpoints / datasets offline).
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usot_b200 import USOT  # noqa: E402
from usot_b200.dist import GradientReducer, train_step_sharded  # noqa: E402
from usot_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, M = 4, 3
    net = USOT({"mem_size": M, "pr_pool": True})
    net.load_state_dict(synthetic_state_dict("damp025"))
    net = net.cuda().train()
    z, x, tb, sb = synthetic_inputs(100 + rank, B, n_templates=B)          # every rank: its own shard
    g = torch.Generator().manual_seed(200 + rank)
    label = torch.zeros(B, 25, 25); label[:, 10:15, 10:15] = 1.0
    rw = torch.zeros(B, 25, 25); rw[:, 11:14, 11:14] = 1.0
    batch = dict(template=z, search=x, search_memory=torch.rand(B, M, 3, 255, 255, generator=g) * 255.0, label=label,
                 reg_target=torch.rand(B, 25, 25, 4, generator=g) * 40 + 5, reg_weight=rw, template_bbox=tb, search_bbox=sb)
    batch = {k: v.cuda() for k, v in batch.items()}
    reducer = GradientReducer(net.parameters())
    opt = torch.optim.SGD(net.parameters(), lr=1e-4, momentum=0.9, weight_decay=1e-4)
    for step in range(5):
        cls_loss, mem_loss, reg_loss = train_step_sharded(net, batch, reducer, opt, loss_weights=(0.5, 0.5, 1.0))
        if rank == 0:
            print(f"step {step}: cls {float(cls_loss):.4f}  cls_memory {float(mem_loss):.4f}  reg {float(reg_loss):.4f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
