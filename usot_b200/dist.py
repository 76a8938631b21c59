"""Multi-GPU plumbing: one process per GPU, crops sharded on dim 0, ONE collective.

* Inference (`track`) shards embarrassingly: every search crop is an independent unit given its template / memory features
  (eval-mode BN), so ranks exchange nothing on the data path.  ``shard_slice`` is the contiguous torch.chunk
  rule nn.DataParallel.scatter uses in the reference (scripts/train_usot.py:318).
* Cycle-memory training forward (BASELINE config 4): the template features z_f (n_local,7,7,256) are all-gathered over NCCL
  (NVLink/NVSwitch) on a side stream while the search / memory backbones run; each rank then consumes its own rows of the
  gathered tensor, so per-sample results equal the single-device reference, and the three losses are averaged over ranks
  exactly like the reference averages over DataParallel replicas (scripts/train_usot.py:201-206).
"""
import torch
import torch.distributed as dist


def shard_slice(total, rank, world):
    """Rows [lo, hi) of a dim-0 batch owned by `rank` (torch.chunk semantics: ceil-sized leading chunks)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    chunk = -(-total // world)
    lo = min(rank * chunk, total)
    return lo, min(lo + chunk, total)


def shard(t, rank, world):
    lo, hi = shard_slice(t.shape[0], rank, world)
    return t[lo:hi]


class ZfExchange:
    """all_gather_into_tensor of the local z_f on a side stream; calling the returned closure waits and returns this rank's
    rows of the gathered tensor (a view, no copy)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.gathered = None
        self._side = None

    def __call__(self, zf_local):
        n = zf_local.shape[0]
        self.gathered = torch.empty((self.world * n,) + tuple(zf_local.shape[1:]), dtype=zf_local.dtype, device=zf_local.device)
        if zf_local.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=zf_local.device)
            self._side.wait_stream(torch.cuda.current_stream(zf_local.device))  # z_f must be complete before it is sent
            with torch.cuda.stream(self._side):
                work = dist.all_gather_into_tensor(self.gathered, zf_local.contiguous(), group=self.group, async_op=True)
            zf_local.record_stream(self._side)
        else:  # gloo / CPU tensors (host-logic tests)
            work = dist.all_gather_into_tensor(self.gathered, zf_local.contiguous(), group=self.group, async_op=True)

        def wait():
            work.wait()
            if zf_local.is_cuda:
                torch.cuda.current_stream(zf_local.device).wait_stream(self._side)
            return self.gathered[self.rank * n:(self.rank + 1) * n]

        return wait


def cycle_forward_sharded(net, batch, group=None, cls_ratio=0.40):
    """BASELINE config 4: `batch` holds THIS rank's shard (template, search, search_memory, label, reg_target, reg_weight,
    template_bbox, search_bbox).  Returns the three losses averaged over ranks (0-d tensors)."""
    ex = ZfExchange(group)
    losses = net.forward(batch["template"], batch["search"], label=batch["label"], reg_target=batch["reg_target"],
                         reg_weight=batch["reg_weight"], template_bbox=batch["template_bbox"], search_memory=batch["search_memory"],
                         search_bbox=batch["search_bbox"], cls_ratio=cls_ratio, zf_exchange=ex)
    out = torch.stack([losses[0], losses[1] if losses[1] is not None else torch.zeros_like(losses[0]), losses[2]]).float()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    out = out / ex.world
    return out[0], out[1], out[2]


def track_sharded(net, x_local, template_mem_local=None, score_mem_local=None, gather=False, group=None):
    """Inference over a dim-0 shard; with gather=True the small score / box maps (12.5 KB per crop) are all-gathered."""
    cls, bbox, cls_mem, xf = net.track(x_local, template_mem_local, score_mem_local)
    if not gather:
        return cls, bbox, cls_mem, xf
    world = dist.get_world_size(group)

    def ag(t):
        if t is None:
            return None
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=group)
        return out

    return ag(cls), ag(bbox), ag(cls_mem), xf
