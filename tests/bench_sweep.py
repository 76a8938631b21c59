"""BASELINE config 5 (script, not a test): throughput sweep of USOT.track() over batch sizes, usot_b200 vs the same network
run by PyTorch + cuDNN on the same GPU (the oracle's functional restatement of the reference modules with CUDA tensors,
cudnn.benchmark on, TF32 allowed = PyTorch's conv default, and TF32 off = the fp32-parity setting).

    python tests/bench_sweep.py [--batches 1,8,64,256] [--nq 0|7]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import usot_oracle as O  # noqa: E402
from usot_b200 import USOT  # noqa: E402
from usot_b200.synth import synthetic_inputs, synthetic_state_dict  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,8,64,256")
    ap.add_argument("--nq", type=int, default=0)
    ap.add_argument("--tunable", action="append", default=[], help="name=value (usot_set_tunable), repeatable; A/B runs")
    ap.add_argument("--ours-only", action="store_true", help="skip the PyTorch + cuDNN legs")
    args = ap.parse_args()
    from usot_b200 import _lib
    for kv in args.tunable:
        name, val = kv.split("=")
        _lib.check(_lib.load().usot_set_tunable(name.encode(), int(val)))
    sd = synthetic_state_dict("damp025")
    sd_cuda = {k: v.cuda() for k, v in sd.items()}
    nets = {}
    for prec in ("fp16x3", "fp16", "fp32"):
        n = USOT(precision=prec)
        n.load_state_dict(sd)
        nets[prec] = n.eval().cuda()
    torch.backends.cudnn.benchmark = True
    rows = []
    for b in [int(x) for x in args.batches.split(",")]:
        z, x, tb, sb = synthetic_inputs(7, b)
        xc, zc, tbc = x.cuda(), z.cuda(), tb.cuda()
        mem = score = None
        row = {"batch": b, "nq": args.nq}
        for prec, net in nets.items():
            net.template(zc, tbc)
            if args.nq:
                mem = net.extract_memory_feature(ori_x=xc[:1].repeat(b * args.nq, 1, 1, 1) if b * args.nq <= 64 else xc[:1].repeat(64, 1, 1, 1).repeat((b * args.nq + 63) // 64, 1, 1, 1)[:b * args.nq],
                                                 search_bbox=sb[:1].repeat(b * args.nq, 1).cuda())
                score = torch.full((b, args.nq), 0.9).cuda()
            ms = timed(lambda: net.track(xc, mem, score), 5 if b >= 64 else 20)
            row[f"ours_{prec}_ms"] = round(ms, 3)
            row[f"ours_{prec}_crops_s"] = round(b / ms * 1e3, 1)
        if args.ours_only:
            print(json.dumps(row), flush=True)
            continue
        with torch.no_grad():
            zf = O.template(sd, z, tb).cuda()
            mem_o = None if not args.nq else mem.contiguous()
            for tf32 in (True, False):
                torch.backends.cudnn.allow_tf32 = tf32
                torch.backends.cuda.matmul.allow_tf32 = tf32
                ms = timed(lambda: O.track(sd_cuda, zf, xc, mem_o, score), 3 if b >= 64 else 10)
                key = "torch_cudnn_tf32" if tf32 else "torch_cudnn_fp32"
                row[f"{key}_ms"] = round(ms, 3)
                row[f"{key}_crops_s"] = round(b / ms * 1e3, 1)
        print(json.dumps(row), flush=True)
        rows.append(row)


if __name__ == "__main__":
    main()
