#!/usr/bin/env python
"""Benchmark of the USOT per-frame forward path on B200 (BASELINE.json metric: search-crops/s at batch 256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference|cudnn] [--config 2|3|4|5] [--batch B] [--precision fp16x3]

Default workload (BASELINE.json configs[1]): 256 synthetic 255x255x3 search crops per GPU per step
through the public call ``USOT.track(x)`` -- ResNet-50 backbone + neck, cls/reg encoders, fused depth-wise xcorr
(template batch 1 broadcast over the crops), cls/reg towers and prediction heads -- with seeded synthetic weights
(usot_b200/synth.py).  One "step" = one such call.  Multi-GPU: one process per GPU (torchrun), crops sharded across
ranks, no data-path collective (weak scaling: 256 crops per GPU).  At N > 1 the line also carries a ``collective`` record: the
ONE collective this path has -- the NCCL all-gather of the template features z_f in the cycle-memory forward (config 4) -- timed
on the same ranks (bytes, overlapped vs serialised vs no-collective step time, and the rank-0 loss compared with a single-GPU run).

Other configurations of BASELINE.json (each prints ONE JSON line of the same shape; the default line is the headline):
    --config 3   full USOT* head: track(x, memory) at batch 64 with a 7-entry memory queue per crop
    --config 4   cycle-memory training forward, 16 samples + 3 memory frames per GPU (the all-gather of z_f on the data path)
    --config 5   batch sweep B in {1, 8, 64, 512}, usot_b200 (fp16x3 / fp16) next to PyTorch + cuDNN (TF32 / fp32) on the same GPU

Output: ONE JSON line (rank 0).  ``value`` = crops/s with inputs resident in HBM; ``e2e`` = the same through the public
API with pinned HOST inputs, H2D copy and D2H of the score/box maps inside the timed region; ``roofline`` = the dense
conv kernel family (tensor bound) measured with per-launch CUDA events in a profiled pass of the same step, plus
``xcorr_roofline`` (HBM bound) for the fused GroupDW kernel; ``cpu_baseline`` = the reference's own modules (baseline/_ref,
staged unmodified by baseline/stage_reference.py; kind "reference") or, when that tree is absent, the CPU oracle port (kind
"port") on the host cores.  ``--impl reference`` times that CPU path alone on the SAME config (every step = a bounded sample of
64 crops of the batch); ``--impl cudnn`` times the same network run by PyTorch + cuDNN on the GPU.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_BACKBONE_NECK = 28.696  # per 255x255 crop, SURVEY.md §8d (exact from the conv shapes)
REF_STAGED = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "MANIFEST.json"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d["hbm_gbs"], tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), tflops_burst=d["bf16_tflops"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tflops_sustained=1400.0, tflops_burst=1590.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 7] or [r for _, r in self.rows[-3:] if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in rows]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "samples": len(rows), "reasons": reasons}

    def stop(self):
        if self.proc:
            self.proc.terminate()


# ------------------------------------------------------------------------------------------------------------------------------
# CPU legs: the reference's own modules (baseline/_ref) or the oracle port
# ------------------------------------------------------------------------------------------------------------------------------
class CpuReference:
    """USOT.track / template on the host cores.  kind "reference": the unmodified reference modules of baseline/_ref
    (lib/models/models.py:173-198) with harness-side ``.cuda()`` no-ops (baseline/run_reference_cpu.py); the template uses the
    reference's pr_pool=False centre crop because its PrRoIPool has no CPU implementation (lib/models/prroi_pool/functional.py:62-63;
    BASELINE.md §4) -- same FLOPs.  kind "port": oracle/usot_oracle.py (a torch-CPU restatement, pinned 0.0 against the reference)."""

    def __init__(self, nq=0):
        from usot_b200.synth import synthetic_inputs, synthetic_state_dict
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = synthetic_state_dict("damp025")
        self.inputs = synthetic_inputs
        self.nq = nq
        self.restore = lambda: None
        if REF_STAGED:
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            import run_reference_cpu
            self.restore = run_reference_cpu.install(with_prroi=False)
            ref_models = run_reference_cpu.import_reference("lib.models.models")
            net = ref_models.USOT()
            net.load_state_dict(self.sd, strict=True)
            self.net = net.eval()
            self.kind = "reference"
            self.what = f"unmodified reference lib.models.models.USOT (baseline/_ref), torch {torch.__version__} CPU"
        else:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import usot_oracle as O
            self.O = O
            self.kind = "port"
            self.what = f"oracle/usot_oracle.py, torch {torch.__version__} CPU"

    def prepare(self, batch, chunk=64):
        z, x, tb, _ = self.inputs(7, batch)
        self.chunks = list(x.split(chunk))
        g = torch.Generator().manual_seed(9)
        self.mem = [torch.randn(c.shape[0] * self.nq, 256, 7, 7, generator=g) for c in self.chunks] if self.nq else None
        with torch.no_grad():
            if self.kind == "reference":
                self.net.pr_pool = False
                self.net.template(z)
            else:
                self.zf = self.O.template(self.sd, z, tb)

    def step(self):
        with torch.no_grad():
            for i, c in enumerate(self.chunks):
                if self.kind == "reference":
                    if self.nq:
                        self.net.track(c, template_mem=self.mem[i], score_mem=torch.full((c.shape[0], self.nq), 0.9))
                    else:
                        self.net.track(c)
                else:
                    if self.nq:
                        self.O.track(self.sd, self.zf, c, self.mem[i], torch.full((c.shape[0], self.nq), 0.9))
                    else:
                        self.O.track(self.sd, self.zf, c)

    def time(self, reps, warmup):
        for _ in range(warmup):
            self.step()
        times = []
        for _ in range(reps):
            t = time.perf_counter()
            self.step()
            times.append(time.perf_counter() - t)
        return times


def workload_config(args):
    b = args.batch
    if args.config == 3:
        w = (f"USOT.track(x, template_mem, score_mem): batch={b} synthetic 255x255x3 search crops per GPU with a 7-entry memory queue per crop "
             "(full USOT* head: offline + memory branch, Conf_Fusion over N_q = 7) (BASELINE.json configs[2])")
    elif args.config == 4 and getattr(args, "train_step", False):
        w = (f"cycle-memory TRAINING STEP (scripts/train_usot.py:196-236{', eager' if getattr(args, 'no_graph', False) else ', one CUDA-graph replay per step' if (getattr(args, 'gpus', 1) == 1 or getattr(args, 'graph_nccl', False)) else ', eager'}): USOT.forward in train() mode + loss.backward() + gradient all-reduce + SGD, "
             f"{b} samples per GPU, 3 memory frames (= {b} templates + {4 * b} search-size crops per GPU per step) (BASELINE.json configs[3])")
    elif args.config == 4:
        w = (f"USOT.forward(...): cycle-memory training forward, {b} samples per GPU, 3 memory frames (= {b} templates + {4 * b} search-size crops per "
             "GPU per step), template features all-gathered across ranks over NCCL (BASELINE.json configs[3])")
    elif args.config == 5:
        w = "USOT.track(x) batch sweep B in {1, 8, 64, 512} per GPU, independent replicas per GPU, vs PyTorch + cuDNN on the same GPU (BASELINE.json configs[4])"
    else:
        w = (f"USOT.track(x): batch={b} synthetic 255x255x3 search crops per GPU, ResNet-50 backbone+neck -> cls/reg encoders -> "
             "fused depthwise xcorr (template batch 1) -> cls/reg towers+heads (BASELINE.json configs[1])")
    return {"workload": w, "batch_per_gpu": b, "search_size": 255, "template_size": 127, "precision": args.precision,
            **({"tunables": args.tunable} if getattr(args, "tunable", None) else {}),
            "l2": "inputs and activations of one step exceed the 126 MB L2 (200 MB / >5 GB at batch 256); no explicit flush"}


def run_reference(args, rank):
    """The reference arm: the reference's own CPU implementation of the same call on the same config, all host threads."""
    if rank != 0:
        return
    if args.config in (4, 5):
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU reference arm covers configs 2 and 3 (track); config {args.config} has no CPU arm"}), flush=True)
        return
    cpu = CpuReference(nq=7 if args.config == 3 else 0)
    sample = min(args.batch, 64)   # bounded sample per step: 64 crops of the batch (~2 s on 16 host threads), so K + W steps end within minutes
    cpu.prepare(sample)
    times = cpu.time(args.steps, args.warmup)
    sec = sum(times) / len(times)
    v = sample / sec
    line = {
        "impl": "reference", "metric": "search_crops_per_sec", "value": v, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args),
        "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cpu.cores, "kind": cpu.kind,
                         "sample": f"every step = USOT.track on a bounded sample of {sample} crops of the batch-{args.batch} workload, {args.steps} steps after {args.warmup} warm-up; {cpu.what}"},
        "e2e": {"value": v, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------------------
# PyTorch + cuDNN arm on the GPU (config 5 baseline): the same network through torch.nn.functional convs
# ------------------------------------------------------------------------------------------------------------------------------
class CudnnNet:
    """The reference modules on the GPU when baseline/_ref is staged (unmodified lib.models, cuDNN convs), else the oracle's
    functional restatement with CUDA tensors.  Offline track() only (the reference's PrRoIPool extension does not build)."""

    def __init__(self, dev):
        from usot_b200.synth import synthetic_inputs, synthetic_state_dict
        self.dev = dev
        self.sd = synthetic_state_dict("damp025")
        z, _, tb, _ = synthetic_inputs(7, 1)
        torch.backends.cudnn.benchmark = True
        if REF_STAGED:
            sys.path.insert(0, os.path.join(ROOT, "baseline"))
            import run_reference_cpu
            ref_models = run_reference_cpu.import_reference("lib.models.models")
            net = ref_models.USOT()
            net.load_state_dict(self.sd, strict=True)
            self.net = net.eval().to(dev)
            self.net.pr_pool = False
            with torch.no_grad():
                self.net.template(z.to(dev))
            self.what = "unmodified reference lib.models (baseline/_ref) on PyTorch + cuDNN"
        else:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import usot_oracle as O
            self.O = O
            self.sd_dev = {k: v.to(dev) for k, v in self.sd.items()}
            with torch.no_grad():
                self.zf = O.template(self.sd, z, tb).to(dev)
            self.net = None
            self.what = "oracle functional restatement on PyTorch + cuDNN"

    def track(self, x):
        with torch.no_grad():
            if self.net is not None:
                return self.net.track(x)
            return self.O.track(self.sd_dev, self.zf, x)


def gpu_time(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cudnn"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5], help="BASELINE.json configs[] index + 1 (default 2 = the headline)")
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--precision", default=os.environ.get("USOT_B200_PRECISION", "fp16x3"), choices=["fp32", "fp16x3", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-collective", action="store_true", help="N > 1: skip the cycle-forward all-gather record")
    ap.add_argument("--graph-nccl", action="store_true", help="--train-step at N > 1: capture the step (NCCL all-reduces included) in a CUDA graph")
    ap.add_argument("--no-graph", action="store_true", help="--train-step: run the step eagerly instead of replaying a captured CUDA graph")
    ap.add_argument("--train-step", action="store_true", help="config 4: time the whole training step (forward + backward + gradient all-reduce + SGD), train()-mode BN")
    ap.add_argument("--tunable", action="append", default=[], help="name=value performance knob (usot_set_tunable), repeatable; A/B runs only")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = {2: 256, 3: 64, 4: 16, 5: 512}[args.config]
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: usot_b200 has no CPU fallback")
    import torch.distributed as dist
    from usot_b200 import USOT, _lib, build
    from usot_b200.synth import synthetic_inputs, synthetic_state_dict
    build.build()
    for kv in args.tunable:
        name, val = kv.split("=")
        _lib.check(_lib.load().usot_set_tunable(name.encode(), int(val)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cleanup = []

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, before_stop=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        if before_stop is not None:
            before_stop()   # (e2e leg: the stop event waits for the last step's read-back, which runs on the copy stream)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, time.time()

    def finish(line):
        if rank == 0:
            print(json.dumps(line), flush=True)
        for g in cleanup:   # captured graphs (they may hold NCCL work) go before the process group
            g.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()

    sd = synthetic_state_dict("damp025")

    # ---------------------------------------------------------------- config 5: batch sweep vs cuDNN ------------------------------
    if args.config == 5:
        nets = {}
        for prec in ("fp16x3", "fp16"):
            n = USOT(precision=prec)
            n.load_state_dict(sd)
            nets[prec] = n.eval().cuda()
        cud = CudnnNet(dev)
        rows = []
        for b in (1, 8, 64, 512):
            z, x, tb, _ = synthetic_inputs(7 + rank, b)
            xc = x.to(dev)
            row = {"batch_per_gpu": b}
            for prec, net in nets.items():
                net.template(z.to(dev), tb.to(dev))
                reps = 5 if b >= 64 else 30
                for _ in range(3):
                    net.track(xc)
                ms, _, _ = timed(lambda: net.track(xc), reps)
                row[f"usot_b200_{prec}_crops_s"] = round(b * world * reps / ms * 1e3, 1)
                row[f"usot_b200_{prec}_ms"] = round(ms / reps, 3)
            for tf32 in (True, False):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                reps = 3 if b >= 64 else 15
                for _ in range(3):
                    cud.track(xc)
                ms, _, _ = timed(lambda: cud.track(xc), reps)
                key = "cudnn_tf32" if tf32 else "cudnn_fp32"
                row[f"{key}_crops_s"] = round(b * world * reps / ms * 1e3, 1)
                row[f"{key}_ms"] = round(ms / reps, 3)
            rows.append(row)
        top = rows[-1]
        finish({"metric": "search_crops_per_sec", "value": top["usot_b200_fp16x3_crops_s"], "unit": "crops/s", "n_gpus": world, "steps": 5, "warmup": 3,
                "ms_per_step": top["usot_b200_fp16x3_ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f16x3 (split-fp16 tcgen05, fp32-equivalent, f32 accumulate)", "data": "synthetic", "config": workload_config(args),
                "sweep": rows, "cudnn_arm": cud.what, "value_is": "whole-job crops/s of the batch-512 row in fp16x3 (device-resident inputs)"})
        return

    # ---------------------------------------------------------------- cuDNN arm of config 2 --------------------------------------
    if args.impl == "cudnn":
        cud = CudnnNet(dev)
        _, x, _, _ = synthetic_inputs(7 + rank, args.batch)
        xc = x.to(dev)
        out = {}
        for tf32 in (True, False):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            for _ in range(args.warmup):
                cud.track(xc)
            ms, _, _ = timed(lambda: cud.track(xc), args.steps)
            out["tf32" if tf32 else "fp32"] = (args.batch * world * args.steps / ms * 1e3, ms / args.steps)
        finish({"impl": "cudnn", "metric": "search_crops_per_sec", "value": out["fp32"][0], "unit": "crops/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": out["fp32"][1], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (cuDNN, TF32 off: the setting that meets the 1e-3 parity bar)", "data": "synthetic", "config": workload_config(args),
                "tf32": {"value": out["tf32"][0], "ms_per_step": out["tf32"][1], "note": "PyTorch's conv default (allow_tf32=True); misses the parity bar"},
                "what": cud.what, "gpu_launches": 0})
        return

    # ---------------------------------------------------------------- ours: configs 2, 3, 4 --------------------------------------
    M = 3
    net = USOT({"mem_size": M, "pr_pool": True} if args.config == 4 else None, precision=args.precision)
    net.load_state_dict(sd)
    net = net.eval().cuda()
    B = args.batch
    z, x_host, tb, sb = synthetic_inputs(7 + rank, B, n_templates=B if args.config == 4 else 1)  # every rank owns a different shard
    x_host = x_host.pin_memory()
    x_dev = x_host.to(dev)
    crops_per_step = B
    h2d_extra = 0
    mem = score = None
    train_batch = None
    graphed = None
    if args.config == 3:
        NQ = 7
        net.template(z.to(dev), tb.to(dev))
        src = synthetic_inputs(31 + rank, 16)
        feats = net.extract_memory_feature(ori_x=src[1].to(dev), search_bbox=src[3].to(dev))
        pick = torch.tensor([(b * NQ + q) * 5 % 16 for b in range(B) for q in range(NQ)], device=dev)
        mem = feats[pick].contiguous(memory_format=torch.channels_last)   # the tracker's device-resident memory queue
        score = torch.full((B, NQ), 0.9, device=dev)
    elif args.config == 4:
        from usot_b200.dist import ZfExchange
        g = torch.Generator().manual_seed(200 + rank)
        label = torch.zeros(B, 25, 25)
        label[:, 10:15, 10:15] = 1.0
        rw = torch.zeros(B, 25, 25)
        rw[:, 11:14, 11:14] = 1.0
        smem_host = (torch.rand(B, M, 3, 255, 255, generator=g) * 255.0).pin_memory()
        train_batch = dict(template=z.to(dev), search=x_dev, search_memory=smem_host.to(dev), label=label.to(dev),
                           reg_target=(torch.rand(B, 25, 25, 4, generator=g) * 40 + 5).to(dev), reg_weight=rw.to(dev), template_bbox=tb.to(dev),
                           search_bbox=sb.to(dev))
        crops_per_step = B * (1 + M)
        h2d_extra = smem_host.numel() * 4 + z.numel() * 4
        exchange = ZfExchange() if world > 1 else None
        if args.train_step:   # a REAL training step: train()-mode BatchNorm, backward, bucketed gradient all-reduce, SGD update
            from usot_b200.dist import GradientReducer
            net.train()
            reducer = GradientReducer(net.parameters(), bucket_mb=25.0)
            optimizer = torch.optim.SGD(net.parameters(), lr=1e-6, momentum=0.9, weight_decay=1e-4)
            # The eager step is host-bound (Python + ~5 000 launches): capture it once, replay per step.  With N > 1 the captured graph contains
            # the NCCL all-reduces; that works (2 GPUs: 85.8 vs 105.7 ms per step, profiles/r02_bench_2gpu_train_*.json) but the process hung in
            # destroy_process_group with the graph alive, so multi-GPU runs stay eager unless --graph-nccl is given (graph reset before teardown).
            if not args.no_graph and (world == 1 or args.graph_nccl):
                from usot_b200.dist import GraphedTrainStep
                graphed = GraphedTrainStep(net, reducer, optimizer, train_batch)
                cleanup.append(graphed)
    else:
        net.template(z.to(dev), tb.to(dev))

    def fwd(batch, ex):
        with torch.no_grad():   # forward only: the engine's fused graph (with autograd enabled forward() builds the training graph instead)
            return net.forward(batch["template"], batch["search"], label=batch["label"], reg_target=batch["reg_target"], reg_weight=batch["reg_weight"],
                               template_bbox=batch["template_bbox"], search_memory=batch["search_memory"], search_bbox=batch["search_bbox"],
                               cls_ratio=0.4, zf_exchange=ex)

    def run(x):
        if args.config == 4:
            tb_ = dict(train_batch, search=x)
            if args.train_step:
                if graphed is not None:
                    return graphed(tb_)
                from usot_b200.dist import train_step_sharded
                return train_step_sharded(net, tb_, reducer, optimizer)
            return fwd(tb_, exchange)
        return net.track(x, mem, score)

    def step_resident():
        run(x_dev)

    # End-to-end leg: every step copies ITS crops from pinned host memory and reads its score/box maps (or losses) back.  The
    # input copy of step i+1 runs on a copy stream into the other staging buffer while step i computes (what a caller of the
    # public API does to keep PCIe off the critical path); all copies are inside the timed region.
    # The caller reads step i's maps while step i+1 is already enqueued (two pinned result sets, an event per set): every step's result
    # still crosses PCIe and is waited for inside the timed region, but the host's launch work no longer sits between two steps.
    x_stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    if args.config == 4:
        out_sets = [[torch.empty(3).pin_memory()]]
    else:
        out_sets = []
        for _ in range(2):
            o = [torch.empty((B, 1, 25, 25)).pin_memory(), torch.empty((B, 4, 25, 25)).pin_memory()]
            if args.config == 3:
                o.append(torch.empty((B, 1, 25, 25)).pin_memory())
            out_sets.append(o)
    out_host = out_sets[0]
    landed = [torch.cuda.Event(), torch.cuda.Event()]
    computed = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "primed": False}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])  # the compute that last read this buffer is done
            x_stage[slot].copy_(x_host, non_blocking=True)
            if args.config == 4:  # the memory frames and templates of a training step come from the host as well
                train_batch["search_memory"].copy_(smem_host, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        cur = state["i"] & 1
        if not state["primed"]:
            consumed[0].record(); consumed[1].record()
            prefetch(cur)
            state["primed"] = True
        if args.config != 4:
            prefetch(cur ^ 1)  # next step's input, overlapped with this step's compute
        torch.cuda.current_stream().wait_event(ready[cur])
        res = run(x_stage[cur])
        consumed[cur].record()
        if args.config == 4:
            out_host[0].copy_(torch.stack([res[0], res[1], res[2]]), non_blocking=True)
            prefetch(cur ^ 1)  # (the single memory-frame buffer is free again only now)
            torch.cuda.current_stream().synchronize()  # the caller consumes the losses every step
        else:
            oh = out_sets[cur]
            computed[cur].record()
            with torch.cuda.stream(copy_stream):   # the read-back rides the copy stream: the next step's kernels do not queue behind it
                copy_stream.wait_event(computed[cur])
                for k in range(3 if args.config == 3 else 2):
                    res[k].record_stream(copy_stream)
                    oh[k].copy_(res[k], non_blocking=True)
                landed[cur].record(copy_stream)
            if state["i"] > 0:
                landed[cur ^ 1].synchronize()  # the caller consumes the PREVIOUS step's maps while this step runs (the last one at the closing sync)
        state["i"] += 1

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        step_resident()
    _lib.profile_reset(False)
    ms, t0, t1 = timed(step_resident, args.steps)
    launches = sum(f["launches"] for f in _lib.profile_read().values()) // args.steps
    clocks = sampler.window(t0, t1) if sampler else None
    for _ in range(2):
        step_e2e()
    def e2e_tail():
        if args.config != 4:
            torch.cuda.current_stream().wait_event(landed[(state["i"] - 1) & 1])

    ms_e2e, _, _ = timed(step_e2e, args.steps, e2e_tail)

    # profiled pass of the same step: per-launch CUDA events on the launching stream, per kernel family
    _lib.profile_reset(True)
    nprof = 2
    for _ in range(nprof):
        step_resident()
    prof = _lib.profile_read()
    _lib.profile_reset(False)
    if sampler:
        sampler.stop()

    # ---- N > 1: the one collective of this path (z_f all-gather of the cycle-memory forward), on the same ranks ----
    collective = collective_train = None
    if world > 1 and not args.no_collective:
        if args.config == 4 and args.train_step:
            collective = grad_allreduce_record(dist, net, reducer, dict(train_batch), timed, world)
        else:
            collective = collective_record(args, dist, dev, rank, world, sd, timed)
            if args.config == 2:
                # ... and the collective of the TRAINING step (gradient all-reduce), measured on the same ranks with a short eager run of
                # BASELINE config 4 as a training step, so that the scaling record carries both collectives of this path.  A failure here must
                # not cost the headline line: it is reported as a string instead.
                try:
                    collective_train = train_collective_record(args, dist, dev, rank, world, sd, timed)
                except Exception as e:  # noqa: BLE001
                    collective_train = {"error": f"{type(e).__name__}: {e}"[:300]}

    line = None
    if rank == 0:
        pk = peaks()
        total = crops_per_step * world
        value = total * args.steps / (ms / 1e3)
        e2e = total * args.steps / (ms_e2e / 1e3)
        conv, xc = prof["conv"], prof["groupdw_xcorr"]
        conv_tflops = conv["flops"] / (conv["ms"] / 1e3) / 1e12 if conv["ms"] > 0 else 0.0
        xc_gbs = xc["bytes"] / (xc["ms"] / 1e3) / 1e9 if xc["ms"] > 0 else 0.0
        pr = prof["pred_conv"]
        pr_gbs = pr["bytes"] / (pr["ms"] / 1e3) / 1e9 if pr["ms"] > 0 else 0.0
        step_ms_prof = sum(f["ms"] for f in prof.values()) / nprof
        traffic = {}
        for name in ("r02c_roofline_traffic.json", "r02b_roofline_traffic.json", "r02_roofline_traffic.json", "r01b_roofline_traffic.json"):
            tp = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tp):
                with open(tp) as f:
                    traffic = json.load(f)
                traffic_src = "profiles/" + name
                break
        mma_per_flop = {"fp32": 0, "fp16x3": 3, "fp16": 1}[args.precision]
        line = {
            "metric": "search_crops_per_sec", "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16x3": "f16x3 (split-fp16 tcgen05, fp32-equivalent, f32 accumulate)", "fp16": "f16 (f32 accumulate)"}[args.precision],
            "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e, "unit": "crops/s", "h2d_bytes_per_step": x_host.numel() * 4 + h2d_extra,
                    "d2h_bytes_per_step": sum(t.numel() * 4 for t in out_host), "ms_per_step": ms_e2e / args.steps,
                    "pipeline": "input of step i+1 copied while step i computes; maps of step i read while step i+1 is enqueued (double-buffered pinned inputs / results)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv), all launches of one step" if mma_per_flop else "conv_simt_kernel, all launches of one step",
                         "bound": "tensor", "achieved": conv_tflops, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                         "frac": conv_tflops / pk["tflops_sustained"],
                         "traffic": traffic.get("conv_tc_kernel", {}).get("traffic_bytes_per_launch"),
                         "traffic_note": f"dram bytes of the dominant launch (layer3.0.downsample) from {traffic_src if traffic else 'n/a'} (ncu --set full); algorithmic bytes of that launch: 1.51 GB",
                         "peak_source": pk["source"] + ", bf16 dense sustained",
                         "achieved_is": "ALGORITHMIC conv FLOPs / summed CUDA-event time of the family's launches (profiled pass of the same step)",
                         "executed_mma_tflops": conv_tflops * mma_per_flop, "executed_mma_frac": conv_tflops * mma_per_flop / pk["tflops_sustained"],
                         "mma_per_algorithmic_flop": mma_per_flop,
                         "launches_per_step": conv["launches"] // nprof, "share_of_step": conv["ms"] / nprof / step_ms_prof if step_ms_prof else None,
                         "algorithmic_gflop_per_step": conv["flops"] / nprof / 1e9},
            "xcorr_roofline": {"kernel": "groupdw_ffma2_kernel (fused 3-scale depthwise xcorr, TMA ring + packed fma.rn.f32x2)", "bound": "hbm", "achieved": xc_gbs, "peak": pk["hbm_gbs"],
                               "unit": "GB/s", "frac": xc_gbs / pk["hbm_gbs"],
                               "traffic": traffic.get("groupdw_ffma2_kernel", traffic.get("groupdw_tma_kernel", {})).get("traffic_bytes_per_launch"),
                               "launches_per_step": xc["launches"] // nprof, "algorithmic_mb_per_launch": xc["bytes"] / max(xc["launches"], 1) / 1e6,
                               "measured": "in-step (CUDA events around each launch inside the profiled pass of the same step), not an isolated launch"},
            "pred_roofline": {"kernel": "pred_gemm_kernel<4>/<1> (bbox_pred / cls_pred 3x3 heads, TMA-streamed per image)", "bound": "hbm",
                              "achieved": pr_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": pr_gbs / pk["hbm_gbs"],
                              "traffic": traffic.get("pred_gemm_kernel", {}).get("traffic_bytes_per_launch"),
                              "launches_per_step": pr["launches"] // nprof, "algorithmic_mb_per_launch": pr["bytes"] / max(pr["launches"], 1) / 1e6},
            "backbone_flop_frac": value / world * GFLOP_BACKBONE_NECK * 1e9 / (pk["tflops_sustained"] * 1e12) if args.config == 2 else None,
            "kernel_ms_per_step": {k: v["ms"] / nprof for k, v in prof.items() if v["launches"]},
        }
        if collective is not None:
            line["collective"] = collective
        if collective_train is not None:
            line["collective_train"] = collective_train
        if not args.no_cpu_baseline and args.config in (2, 3):
            sample = 64
            cpu = CpuReference(nq=7 if args.config == 3 else 0)
            cpu.prepare(sample)
            times = cpu.time(reps=2, warmup=1)
            cpu.restore()
            line["cpu_baseline"] = {"value": sample / min(times), "unit": "crops/s", "cores": cpu.cores, "kind": cpu.kind,
                                    "sample": f"USOT.track on {sample} crops of the same workload, best of 2 after 1 warm-up ({cpu.cores} threads); {cpu.what}"}
    finish(line)


def grad_allreduce_record(dist, net, reducer, batch, timed, world):
    """The collective that limits the 1 -> N TRAINING curve: the bucketed all-reduce of the fp32 gradients (replaces DataParallel's
    reduce_add_coalesced, scripts/train_usot.py:318).  Step time with the all-reduce overlapped with backward vs the same step with the
    all-reduce skipped (= its exposed time), and the all-reduce of all buckets alone (bus bandwidth, ring convention 2(N-1)/N)."""
    from usot_b200.dist import train_step_sharded
    steps = 5
    for _ in range(2):
        train_step_sharded(net, batch, reducer, None)
    t_with, _, _ = timed(lambda: train_step_sharded(net, batch, reducer, None), steps)
    reducer.world = 1                      # same step, no communication
    for _ in range(2):
        train_step_sharded(net, batch, reducer, None)
    t_without, _, _ = timed(lambda: train_step_sharded(net, batch, reducer, None), steps)
    reducer.world = world

    def allreduce_all():
        works = [dist.all_reduce(b["flat"], async_op=True) for b in reducer.buckets]
        for w in works:
            w.wait()

    for _ in range(3):
        allreduce_all()
    t_ar, _, _ = timed(allreduce_all, 10)
    t_ar /= 10
    nbytes = reducer.bytes_per_step
    return {"op": "bucketed all_reduce(mean) of the fp32 parameter gradients over NCCL, overlapped with backward (usot_b200.dist.GradientReducer)",
            "nranks": world, "bytes_per_rank_per_step": nbytes, "buckets": len(reducer.buckets),
            "ms_per_step_overlapped": t_with / steps, "ms_per_step_no_collective": t_without / steps,
            "exposed_ms": (t_with - t_without) / steps, "allreduce_alone_ms": t_ar,
            "allreduce_bus_gbs": 2.0 * (world - 1) / world * nbytes / (t_ar / 1e3) / 1e9}


def train_collective_record(args, dist, dev, rank, world, sd, timed):
    """BASELINE config 4 as a TRAINING step on these ranks (16 samples + 3 memory frames per GPU, train()-mode BatchNorm, eager): step time
    with the bucketed gradient all-reduce overlapped with backward vs without it, and the all-reduce of all buckets alone."""
    from usot_b200 import USOT
    from usot_b200.dist import GradientReducer
    from usot_b200.synth import synthetic_inputs
    B, M = 16, 3
    net = USOT({"mem_size": M, "pr_pool": True}, precision=args.precision)
    net.load_state_dict(sd)
    net = net.cuda().train()
    z, x, tb, sb = synthetic_inputs(300 + rank, B, n_templates=B)
    g = torch.Generator().manual_seed(400 + rank)
    label = torch.zeros(B, 25, 25)
    label[:, 10:15, 10:15] = 1.0
    rw = torch.zeros(B, 25, 25)
    rw[:, 11:14, 11:14] = 1.0
    batch = dict(template=z, search=x, search_memory=torch.rand(B, M, 3, 255, 255, generator=g) * 255.0, label=label,
                 reg_target=torch.rand(B, 25, 25, 4, generator=g) * 40 + 5, reg_weight=rw, template_bbox=tb, search_bbox=sb)
    batch = {k: v.to(dev) for k, v in batch.items()}
    reducer = GradientReducer(net.parameters(), bucket_mb=25.0)
    rec = grad_allreduce_record(dist, net, reducer, batch, timed, world)
    rec["workload"] = f"cycle-memory training step, {B} samples + {M} memory frames per GPU, global batch {B * world}, eager (no CUDA graph)"
    rec["samples_per_s"] = B * world / rec["ms_per_step_overlapped"] * 1e3
    reducer.close()
    return rec


def collective_record(args, dist, dev, rank, world, sd, timed):
    """BASELINE config 4 on these ranks: cycle-memory forward, 16 samples + 3 memory frames per GPU.  Times the step (a) with the NCCL
    all-gather of z_f overlapped on a side stream (the product path, usot_b200.dist.ZfExchange), (b) with the same all-gather
    serialised on the compute stream, (c) with no collective at all (every rank keeps its own z_f: the same arithmetic, since a rank
    consumes only its own rows of the gathered tensor), the all-gather alone, and checks that rank 0's losses with the collective are
    bit-identical to the no-collective (= single-GPU) run of the same shard."""
    from usot_b200 import USOT
    from usot_b200.dist import ZfExchange
    from usot_b200.synth import synthetic_inputs
    B, M = 16, 3
    net = USOT({"mem_size": M, "pr_pool": True}, precision=args.precision)
    net.load_state_dict(sd)
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

    class SerialExchange(ZfExchange):   # the same collective, but on the compute stream and waited for immediately (no overlap)
        def __call__(self, zf_local):
            n = zf_local.shape[0]
            self.gathered = torch.empty((self.world * n,) + tuple(zf_local.shape[1:]), dtype=zf_local.dtype, device=zf_local.device)
            dist.all_gather_into_tensor(self.gathered, zf_local.contiguous(), group=self.group)
            torch.cuda.current_stream().synchronize()
            return lambda: self.gathered[self.rank * n:(self.rank + 1) * n]

    def fwd(ex):
        with torch.no_grad():
            return net.forward(batch["template"], batch["search"], label=batch["label"], reg_target=batch["reg_target"], reg_weight=batch["reg_weight"],
                               template_bbox=batch["template_bbox"], search_memory=batch["search_memory"], search_bbox=batch["search_bbox"],
                               cls_ratio=0.4, zf_exchange=ex)

    ov, se = ZfExchange(), SerialExchange()
    steps = 10
    res = {}
    for name, ex in (("no_collective", None), ("overlapped", ov), ("serialised", se)):
        for _ in range(3):
            fwd(ex)
        ms, _, _ = timed(lambda: fwd(ex), steps)
        res[name] = ms / steps
    with_c = torch.stack(list(fwd(ov)))
    without = torch.stack(list(fwd(None)))
    equal = bool(torch.equal(with_c, without))
    zf = torch.randn(B, 7, 7, 256, device=dev)
    out = torch.empty(world * B, 7, 7, 256, device=dev)
    for _ in range(5):
        dist.all_gather_into_tensor(out, zf)
    ms_ag, _, _ = timed(lambda: dist.all_gather_into_tensor(out, zf), 50)
    ms_ag /= 50
    send = B * 49 * 256 * 4
    return {"op": "all_gather_into_tensor of z_f (template features) over NCCL, cycle-memory forward (BASELINE config 4)",
            "workload": f"{B} samples + {M} memory frames per GPU, global batch {B * world}", "nranks": world,
            "bytes_sent_per_rank": send, "bytes_gathered_per_rank": send * world,
            "ms_per_step_overlapped": res["overlapped"], "ms_per_step_serialised": res["serialised"], "ms_per_step_no_collective": res["no_collective"],
            "exposed_ms_overlapped": res["overlapped"] - res["no_collective"], "exposed_ms_serialised": res["serialised"] - res["no_collective"],
            "allgather_alone_ms": ms_ag, "allgather_bus_gbs": send * (world - 1) / (ms_ag / 1e3) / 1e9,
            "samples_per_s": B * world / res["overlapped"] * 1e3, "search_crops_per_s": B * world * (1 + M) / res["overlapped"] * 1e3,
            "rank0_losses_equal_single_gpu_run": equal, "rank0_losses": [float(v) for v in with_c]}


if __name__ == "__main__":
    main()
