"""2-GPU NCCL run of the sharded cycle-memory forward (BASELINE config 4 at small scale) and of sharded inference:
per-rank results must equal the single-device results on the full batch."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from helpers import load_weights
    from test_gpu_train import _train_batch
    from usot_b200 import USOT
    from usot_b200.dist import cycle_forward_sharded, shard, track_sharded
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        net = USOT({"mem_size": 3, "pr_pool": True}, precision="fp16x3")
        net.load_state_dict(load_weights("damp025"))
        net = net.eval().cuda()
        full = {k: v.cuda() for k, v in _train_batch(B=4, M=3).items()}
        local = {k: shard(v, rank, world) for k, v in full.items()}
        losses = cycle_forward_sharded(net, local)
        ref = net(full["template"], full["search"], label=full["label"], reg_target=full["reg_target"], reg_weight=full["reg_weight"],
                  template_bbox=full["template_bbox"], search_memory=full["search_memory"], search_bbox=full["search_bbox"], cls_ratio=0.4)
        # reg/cls losses are means over per-rank subsets of equal size here (same #positives per sample) -> rank-mean == full-batch value
        for a, b in zip(losses, ref):
            assert abs(float(a) - float(b)) <= 2e-5 * max(1.0, abs(float(b))), (float(a), float(b))
        # sharded inference + gather == single-device inference
        net.template(full["template"][:1], full["template_bbox"][:1])
        cls_l, bbox_l, _, _ = track_sharded(net, local["search"], gather=True)
        cls_f, bbox_f, _, _ = net.track(full["search"])
        assert torch.equal(cls_l, cls_f) and torch.equal(bbox_l, bbox_f)
        # data-parallel TRAINING step (eval-mode BN so that shard and full batch see the same statistics): the all-reduced (mean) gradients
        # of the two shards must equal the single-device gradients of the full batch
        from usot_b200.dist import GradientReducer, train_step_sharded
        red = GradientReducer(net.parameters(), bucket_mb=8.0)
        l_local = train_step_sharded(net, local, red, None)
        assert len(red.buckets) >= 4 and sorted(red.launch_order) == list(range(len(red.buckets)))
        shard_grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        red.close()
        for p in net.parameters():
            p.grad = None
        lf = net(full["template"], full["search"], label=full["label"], reg_target=full["reg_target"], reg_weight=full["reg_weight"],
                 template_bbox=full["template_bbox"], search_memory=full["search_memory"], search_bbox=full["search_bbox"], cls_ratio=0.4)
        (lf[0] + lf[1] + lf[2]).backward()
        worst = 0.0
        for k, p in net.named_parameters():
            den = float(p.grad.abs().max())
            if den > 1e-8:
                worst = max(worst, float((shard_grads[k] - p.grad).abs().max()) / den)
        assert worst <= 2e-3, worst
        if rank == 0:
            out.put("ok")
    finally:
        dist.destroy_process_group()


def test_two_gpu_nccl_cycle_forward_and_sharded_track():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    mp.spawn(_worker, args=(2, port, q), nprocs=2, join=True)
    assert q.get() == "ok"
