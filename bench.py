#!/usr/bin/env python
"""Benchmark of the USOT per-frame forward path on B200 (BASELINE.json metric: search-crops/s at batch 256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 256] [--precision fp16x3]

Workload (BASELINE.json configs[1]): 256 synthetic 255x255x3 search crops per GPU per step
through the public call ``USOT.track(x)`` -- ResNet-50 backbone + neck, cls/reg encoders, fused depth-wise xcorr
(template batch 1 broadcast over the crops), cls/reg towers and prediction heads -- with seeded synthetic weights
(usot_b200/synth.py).  One "step" = one such call.  Multi-GPU: one process per GPU (torchrun), crops sharded across
ranks, no data-path collective (weak scaling: 256 crops per GPU).

Output: ONE JSON line (rank 0).  ``value`` = crops/s with inputs resident in HBM; ``e2e`` = the same through the public
API with pinned HOST inputs, H2D copy and D2H of the score/box maps inside the timed region; ``roofline`` = the dense
conv kernel family (tensor bound) measured with per-launch CUDA events in a profiled pass of the same step, plus
``xcorr_roofline`` (HBM bound) for the fused GroupDW kernel; ``cpu_baseline`` = the CPU oracle (a torch-CPU restatement
of the reference modules -- the reference is Python and /root/reference does not exist on the GPU box) on the host cores.
``--impl reference`` times that CPU path alone, with the same metric/config keys.
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


def cpu_oracle_leg(batch, reps, warmup=1):
    """The CPU restatement of the reference modules on this box's host cores (checker code, timed as the baseline)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import usot_oracle as O
    from usot_b200.synth import synthetic_inputs, synthetic_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synthetic_state_dict("damp025")
    z, x, tb, _ = synthetic_inputs(7, batch)
    with torch.no_grad():
        zf = O.template(sd, z, tb)
        for _ in range(warmup):
            O.track(sd, zf, x)
        times = []
        for _ in range(reps):
            t = time.perf_counter()
            O.track(sd, zf, x)
            times.append(time.perf_counter() - t)
    return cores, times


def run_reference(args, rank):
    if rank != 0:
        return
    sample = 8
    cores, times = cpu_oracle_leg(sample, args.steps, args.warmup)
    sec = sum(times) / len(times)
    v = sample / sec
    line = {
        "impl": "reference", "metric": "search_crops_per_sec", "value": v, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, sample_note=f"each step = USOT.track on a bounded sample of {sample} crops of the batch-{args.batch} workload"),
        "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cores, "kind": "port",
                         "sample": f"{sample} crops/step x {args.steps} steps, torch {torch.__version__} CPU, oracle/usot_oracle.py"},
        "e2e": {"value": v, "unit": "crops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, sample_note=None):
    c = {"workload": f"USOT.track(x): batch={args.batch} synthetic 255x255x3 search crops per GPU, ResNet-50 backbone+neck -> cls/reg encoders -> "
                     "fused depthwise xcorr (template batch 1) -> cls/reg towers+heads (BASELINE.json configs[1])",
         "batch_per_gpu": args.batch, "search_size": 255, "template_size": 127, "precision": args.precision,
         **({"tunables": args.tunable} if getattr(args, "tunable", None) else {}),
         "l2": "inputs (200 MB/step) and activations (>5 GB/step) exceed the 126 MB L2; no explicit flush"}
    if sample_note:
        c["sample"] = sample_note
    return c


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--precision", default=os.environ.get("USOT_B200_PRECISION", "fp16x3"), choices=["fp32", "fp16x3", "fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tunable", action="append", default=[], help="name=value performance knob (usot_set_tunable), repeatable; A/B runs only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

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

    net = USOT(precision=args.precision)
    net.load_state_dict(synthetic_state_dict("damp025"))
    net = net.eval().cuda()
    z, x_host, tb, _ = synthetic_inputs(7 + rank, args.batch)  # every rank owns a different shard of crops
    x_host = x_host.pin_memory()
    x_dev = x_host.to(dev)
    net.template(z.to(dev), tb.to(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, time.time()

    def step_resident():
        net.track(x_dev)

    # End-to-end leg: every step copies ITS 200 MB of crops from pinned host memory and reads its score/box maps back.  The
    # input copy of step i+1 runs on a copy stream into the other staging buffer while step i computes (what a caller of the
    # public API does to keep PCIe off the critical path); all copies are inside the timed region.
    x_stage = [torch.empty_like(x_dev), torch.empty_like(x_dev)]
    out_host = [torch.empty((args.batch, 1, 25, 25)).pin_memory(), torch.empty((args.batch, 4, 25, 25)).pin_memory()]
    copy_stream = torch.cuda.Stream(device=dev)
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0, "primed": False}

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])  # the compute that last read this buffer is done
            x_stage[slot].copy_(x_host, non_blocking=True)
            ready[slot].record(copy_stream)

    def step_e2e():
        cur = state["i"] & 1
        if not state["primed"]:
            consumed[0].record(); consumed[1].record()
            prefetch(cur)
            state["primed"] = True
        prefetch(cur ^ 1)  # next step's input, overlapped with this step's compute
        torch.cuda.current_stream().wait_event(ready[cur])
        cls, bbox, _, _ = net.track(x_stage[cur])
        consumed[cur].record()
        out_host[0].copy_(cls, non_blocking=True)
        out_host[1].copy_(bbox, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the caller consumes the maps every step
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
    ms_e2e, _, _ = timed(step_e2e, args.steps)

    # profiled pass of the same step: per-launch CUDA events on the launching stream, per kernel family
    _lib.profile_reset(True)
    nprof = 2
    for _ in range(nprof):
        step_resident()
    prof = _lib.profile_read()
    _lib.profile_reset(False)
    if sampler:
        sampler.stop()

    if rank == 0:
        pk = peaks()
        total = args.batch * world
        value = total * args.steps / (ms / 1e3)
        e2e = total * args.steps / (ms_e2e / 1e3)
        conv, xc = prof["conv"], prof["groupdw_xcorr"]
        conv_tflops = conv["flops"] / (conv["ms"] / 1e3) / 1e12 if conv["ms"] > 0 else 0.0
        xc_gbs = xc["bytes"] / (xc["ms"] / 1e3) / 1e9 if xc["ms"] > 0 else 0.0
        pr = prof["pred_conv"]
        pr_gbs = pr["bytes"] / (pr["ms"] / 1e3) / 1e9 if pr["ms"] > 0 else 0.0
        step_ms_prof = sum(f["ms"] for f in prof.values()) / nprof
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "r01b_roofline_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f)
        mma_per_flop = {"fp32": 0, "fp16x3": 3, "fp16": 1}[args.precision]
        line = {
            "metric": "search_crops_per_sec", "value": value, "unit": "crops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "fp16x3": "f16x3 (split-fp16 tcgen05, fp32-equivalent, f32 accumulate)", "fp16": "f16 (f32 accumulate)"}[args.precision],
            "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e, "unit": "crops/s", "h2d_bytes_per_step": x_host.numel() * 4 * 1,
                    "d2h_bytes_per_step": sum(t.numel() * 4 for t in out_host), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv), all launches of one step" if mma_per_flop else "conv_simt_kernel, all launches of one step",
                         "bound": "tensor", "achieved": conv_tflops, "peak": pk["tflops_sustained"], "unit": "TFLOP/s",
                         "frac": conv_tflops / pk["tflops_sustained"],
                         "traffic": traffic.get("conv_tc_kernel", {}).get("traffic_bytes_per_launch"),
                         "traffic_note": "dram bytes of the dominant launch (layer3.0.downsample) from profiles/r01b_roofline_traffic.json (ncu --set full); equals its algorithmic bytes",
                         "peak_source": pk["source"] + ", bf16 dense sustained",
                         "achieved_is": "ALGORITHMIC conv FLOPs / summed CUDA-event time of the family's launches (profiled pass of the same step)",
                         "executed_mma_tflops": conv_tflops * mma_per_flop, "executed_mma_frac": conv_tflops * mma_per_flop / pk["tflops_sustained"],
                         "mma_per_algorithmic_flop": mma_per_flop,
                         "launches_per_step": conv["launches"] // nprof, "share_of_step": conv["ms"] / nprof / step_ms_prof if step_ms_prof else None,
                         "algorithmic_gflop_per_step": conv["flops"] / nprof / 1e9},
            "xcorr_roofline": {"kernel": "groupdw_ffma2_kernel (fused 3-scale depthwise xcorr, TMA ring + packed fma.rn.f32x2)", "bound": "hbm", "achieved": xc_gbs, "peak": pk["hbm_gbs"],
                               "unit": "GB/s", "frac": xc_gbs / pk["hbm_gbs"],
                               "traffic": traffic.get("groupdw_ffma2_kernel", traffic.get("groupdw_tma_kernel", {})).get("traffic_bytes_per_launch"),
                               "launches_per_step": xc["launches"] // nprof, "algorithmic_mb_per_launch": xc["bytes"] / max(xc["launches"], 1) / 1e6},
            "pred_roofline": {"kernel": "pred_gemm_kernel<4>/<1> (bbox_pred / cls_pred 3x3 heads, TMA-streamed per image)", "bound": "hbm",
                              "achieved": pr_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": pr_gbs / pk["hbm_gbs"],
                              "traffic": traffic.get("pred_gemm_kernel", {}).get("traffic_bytes_per_launch"),
                              "launches_per_step": pr["launches"] // nprof, "algorithmic_mb_per_launch": pr["bytes"] / max(pr["launches"], 1) / 1e6},
            "backbone_flop_frac": value / world * GFLOP_BACKBONE_NECK * 1e9 / (pk["tflops_sustained"] * 1e12),
            "kernel_ms_per_step": {k: v["ms"] / nprof for k, v in prof.items() if v["launches"]},
        }
        if not args.no_cpu_baseline:
            sample = 8
            cores, times = cpu_oracle_leg(sample, reps=2)
            line["cpu_baseline"] = {"value": sample / min(times), "unit": "crops/s", "cores": cores, "kind": "port",
                                    "sample": f"USOT.track on {sample} crops, best of 2 after 1 warm-up, torch {torch.__version__} CPU ({cores} threads), oracle/usot_oracle.py"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
