"""Host-side multi-GPU logic on CPU: world_size-2 gloo run of the sharding rule, the z_f all-gather and the loss averaging."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from usot_b200.dist import ZfExchange, cycle_forward_sharded, shard_slice


def test_shard_slice_matches_torch_chunk():
    for total in (1, 7, 16, 128, 255, 256):
        for world in (1, 2, 3, 4, 8):
            chunks = torch.arange(total).chunk(world)
            for rank in range(world):
                lo, hi = shard_slice(total, rank, world)
                want = chunks[rank] if rank < len(chunks) else torch.arange(0)
                assert hi - lo == len(want)
                if len(want):
                    assert lo == int(want[0])


class _StubNet:
    """Stands in for usot_b200.USOT on the CPU: same forward() contract, deterministic per-sample 'losses'."""

    def forward(self, template, search, label=None, reg_target=None, reg_weight=None, template_bbox=None, search_memory=None,
                search_bbox=None, cls_ratio=0.4, zf_exchange=None):
        zf = template.mean(dim=(1, 2, 3)).view(-1, 1, 1, 1).repeat(1, 7, 7, 4)  # "template features" of the local shard
        wait = zf_exchange(zf)
        work = search.mean(dim=(1, 2, 3))                                         # stands in for the backbones that overlap
        zf_mine = wait()
        assert torch.equal(zf_mine, zf), "a rank must read back exactly its own rows of the gathered z_f"
        per_sample = zf_mine.mean(dim=(1, 2, 3)) + work
        return per_sample.mean(), (per_sample * 2).mean(), (per_sample * 3).mean()


def _worker(rank, world, port, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        template, search = torch.rand(total, 3, 8, 8, generator=g), torch.rand(total, 3, 8, 8, generator=g)
        lo, hi = shard_slice(total, rank, world)
        batch = dict(template=template[lo:hi], search=search[lo:hi], label=None, reg_target=None, reg_weight=None, template_bbox=None,
                     search_memory=None, search_bbox=None)
        l0, l1, l2 = cycle_forward_sharded(_StubNet(), batch)
        full = template.mean(dim=(1, 2, 3)) + search.mean(dim=(1, 2, 3))  # single-process result over the whole batch
        assert abs(float(l0) - float(full.mean())) < 1e-6 and abs(float(l1) - 2 * float(full.mean())) < 1e-6
        assert abs(float(l2) - 3 * float(full.mean())) < 1e-6
        ex = ZfExchange()
        zf = torch.full((hi - lo, 7, 7, 4), float(rank))
        mine = ex(zf)()
        assert ex.gathered.shape[0] == total and torch.equal(mine, zf)
        for r in range(world):
            assert float(ex.gathered[r * (hi - lo)].mean()) == float(r)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_allgather_and_loss_average():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 16), nprocs=2, join=True)


# ---- gradient all-reduce (replaces DataParallel's reduce_add_coalesced, scripts/train_usot.py:318) ------------------------------
def _grad_worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from usot_b200.dist import GradientReducer
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 3))
        net[2].bias.requires_grad_(False)                                   # a frozen parameter must simply be skipped
        red = GradientReducer(net.parameters(), bucket_mb=0.01)             # ~2.6k floats per bucket -> several buckets
        assert len(red.buckets) >= 3
        g = torch.Generator().manual_seed(1)
        x_all, y_all = torch.randn(8, 16, generator=g), torch.randn(8, 3, generator=g)
        lo, hi = rank * 4, rank * 4 + 4
        for step in range(2):                                               # second step: buckets are re-zeroed, views survive
            red.zero_grad()
            loss = ((net(x_all[lo:hi]) - y_all[lo:hi]) ** 2).mean()
            loss.backward()
            assert red.launch_order[0] == 0 and sorted(red.launch_order) == list(range(len(red.buckets)))   # heads-first, every bucket once
            red.finish()
            # reference: mean over ranks of the per-rank gradients == gradient of the mean of the per-rank losses
            ref_net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 3))
            ref_net.load_state_dict(net.state_dict())
            full = sum(((ref_net(x_all[r * 4:r * 4 + 4]) - y_all[r * 4:r * 4 + 4]) ** 2).mean() for r in range(world)) / world
            full.backward()
            for (k, p), (_, q) in zip(net.named_parameters(), ref_net.named_parameters()):
                if not p.requires_grad:
                    assert p.grad is None
                    continue
                assert torch.allclose(p.grad, q.grad, atol=1e-6), k
                assert p.grad.data_ptr() >= red.buckets[red._of[p]]["flat"].data_ptr()   # still a view into its bucket
        red.close()
    finally:
        dist.destroy_process_group()


def test_world2_gloo_bucketed_gradient_allreduce():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_grad_worker, args=(2, port), nprocs=2, join=True)
