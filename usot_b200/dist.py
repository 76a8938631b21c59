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
    """BASELINE config 4, forward only: `batch` holds THIS rank's shard (template, search, search_memory, label, reg_target,
    reg_weight, template_bbox, search_bbox).  Returns the three losses averaged over ranks (0-d tensors, no autograd graph: the
    engine's fused forward with running-statistics BatchNorm; for a training step use ``train_step_sharded``)."""
    ex = ZfExchange(group)
    with torch.no_grad():
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


# ---------------------------------------------------------------------------------------------------------------------------------
# Training: gradient all-reduce (replaces nn.DataParallel's reduce_add_coalesced, scripts/train_usot.py:318)
# ---------------------------------------------------------------------------------------------------------------------------------
class GradientReducer:
    """Bucketed all-reduce of the parameter gradients, overlapped with the backward pass.

    The reference trains under ``nn.DataParallel``: every iteration the replicas' gradients are summed onto device 0 over PCIe /
    NVLink by ``reduce_add_coalesced`` AFTER backward has finished, and the parameters are re-broadcast before the next forward.
    Here every rank owns a full replica (one process per GPU) and the ~118 MB of fp32 gradients are averaged with NCCL all-reduces
    over NVLink / NVSwitch while backward is still running:

      * parameters are packed, in REVERSE registration order (roughly the order backward produces their gradients: heads first,
        stem last), into flat fp32 buckets of ``bucket_mb``; ``p.grad`` of every parameter is a VIEW into its bucket, so autograd
        accumulates straight into the communication buffer (no gather / scatter copies);
      * a post-accumulate hook per parameter counts ready gradients; when the last one of a bucket lands, ``all_reduce(AVG)`` of that
        bucket is issued asynchronously (NCCL runs it on its own stream, ordered after the kernels that produced the gradients);
      * ``finish()`` (after ``loss.backward()``) waits for the outstanding work, so the optimizer sees averaged gradients.

    Averaging (not summing) matches the reference's ``loss = torch.mean(loss)`` over the per-replica losses (scripts/train_usot.py:
    201-227).  BatchNorm statistics stay per rank, like DataParallel's per-replica statistics; ``broadcast_buffers`` copies rank 0's
    running statistics to everyone (what DataParallel keeps: the module on device 0) before a checkpoint is written."""

    def __init__(self, params, bucket_mb=25.0, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets = []       # dicts: flat, params, pending, work
        self._of = {}
        cap = int(bucket_mb * (1 << 20)) // 4
        cur, cur_n = [], 0
        for p in reversed(self.params):
            if cur and cur_n + p.numel() > cap:
                self._make_bucket(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += p.numel()
        if cur:
            self._make_bucket(cur)
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.launch_order = []   # bucket indices in the order their all-reduce was issued (diagnostics / tests)
        self.bytes_per_step = sum(b["flat"].numel() for b in self.buckets) * 4

    def _make_bucket(self, plist):
        dev, n = plist[0].device, sum(p.numel() for p in plist)
        flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in plist:
            p.grad = flat[off:off + p.numel()].view_as(p)
            self._of[p] = len(self.buckets)
            off += p.numel()
        self.buckets.append({"flat": flat, "params": plist, "pending": len(plist), "work": None})

    def zero_grad(self):
        """Zero the buckets (instead of ``optimizer.zero_grad()``, which would drop the views with set_to_none=True)."""
        for i, b in enumerate(self.buckets):
            b["flat"].zero_()
            b["pending"], b["work"] = len(b["params"]), None
            off = 0
            for p in b["params"]:   # re-attach the views if someone replaced .grad
                if p.grad is None or p.grad.data_ptr() != b["flat"].data_ptr() + off * 4:
                    p.grad = b["flat"][off:off + p.numel()].view_as(p)
                off += p.numel()
        self.launch_order = []

    def _on_grad(self, p):
        b = self.buckets[self._of[p]]
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(self._of[p])

    def _launch(self, i):
        b = self.buckets[i]
        self.launch_order.append(i)
        if self.world > 1:
            b["work"] = dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self):
        """Wait for every bucket (issuing the all-reduce of buckets whose parameters got no gradient this step), then scale to the mean."""
        for i, b in enumerate(self.buckets):
            if b["work"] is None and b["pending"] > 0 and i not in self.launch_order:
                self._launch(i)
        for b in self.buckets:
            if b["work"] is not None:
                b["work"].wait()
                b["flat"].div_(self.world)
                b["work"] = None

    def close(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


def broadcast_buffers(net, src=0, group=None):
    """Rank ``src``'s BatchNorm running statistics to every rank (DataParallel keeps device 0's)."""
    for b in net.buffers():
        dist.broadcast(b, src=src, group=group)


def train_step_sharded(net, batch, reducer, optimizer=None, loss_weights=(1.0, 1.0, 1.0), cls_ratio=0.40):
    """One data-parallel training step on this rank's shard: forward with an autograd graph (usot_b200/train.py), backward with the
    gradient all-reduce overlapped (``reducer``), optional optimizer step.  ``loss_weights`` = (lambda_1, lambda_total - lambda_1, 1.0)
    of scripts/train_usot.py:215-227.  Returns the three local losses (detached)."""
    reducer.zero_grad()
    losses = net.forward(batch["template"], batch["search"], label=batch["label"], reg_target=batch["reg_target"],
                         reg_weight=batch["reg_weight"], template_bbox=batch["template_bbox"], search_memory=batch.get("search_memory"),
                         search_bbox=batch.get("search_bbox"), cls_ratio=cls_ratio)
    loss = loss_weights[0] * losses[0] + loss_weights[2] * losses[2]
    if losses[1] is not None:
        loss = loss + loss_weights[1] * losses[1]
    loss.backward()
    reducer.finish()
    if optimizer is not None:
        optimizer.step()
    return tuple(None if v is None else v.detach() for v in losses)


class GraphedTrainStep:
    """The whole training step -- forward with the autograd graph, backward, gradient all-reduce, optimizer update -- captured ONCE in a CUDA
    graph and replayed per step.  The eager step issues ~2 200 kernel launches of this library plus ~3 000 small torch ops from Python and is
    HOST-bound (measured: 119 ms of host enqueue time per 120 ms step, profiles/r02_train_step_profile_*.json); replaying a graph removes
    the host from the critical path.  Everything the step launches is capture-safe by construction: kernels go to torch's current
    (capturing) stream, scratch memory comes from the stream-ordered allocator (cudaMallocAsync / cudaFreeAsync are captured as graph memory
    nodes), no call synchronises or reads a device value on the host, and one-time host set-up (kernel attributes, tensor-map encodes of new
    shapes) happens during the warm-up steps that precede the capture.

    Inputs are copied into static buffers before each replay; the returned losses are static tensors overwritten by the next replay."""

    def __init__(self, net, reducer, optimizer, example_batch, loss_weights=(1.0, 1.0, 1.0), cls_ratio=0.40, warmup=3):
        self.static = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in example_batch.items()}
        self._args = (net, reducer, optimizer, loss_weights, cls_ratio)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                train_step_sharded(net, self.static, reducer, optimizer, loss_weights, cls_ratio)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        from . import _lib
        before = {k: v["launches"] for k, v in _lib.profile_read().items()}
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = train_step_sharded(net, self.static, reducer, optimizer, loss_weights, cls_ratio)
        # launches of this library recorded into the graph, per kernel family: the capture itself executed nothing, every replay runs them all
        self.launches = {k: v["launches"] - before[k] for k, v in _lib.profile_read().items()}
        _lib.profile_count({k: -n for k, n in self.launches.items()})

    def close(self):
        """Release the captured graph (do this before torch.distributed.destroy_process_group when the graph holds NCCL work)."""
        if self.graph is not None:
            torch.cuda.synchronize()
            self.graph.reset()
            self.graph = None

    def __call__(self, batch=None):
        if batch is not None:
            for k, v in batch.items():
                if torch.is_tensor(v) and v.data_ptr() != self.static[k].data_ptr():
                    self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        from . import _lib
        _lib.profile_count(self.launches)
        return self.losses
