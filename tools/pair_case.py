"""CTA-pair (tcgen05.mma.cta_group::2) conv launches vs the one-CTA kernel: results must be identical bit for bit.

    python tools/pair_case.py ops   fp16x3|fp16     stand-alone conv launches through the C ABI (fp32-output epilogue)
    python tools/pair_case.py model fp16x3|fp16     USOT.track() at batch 64 (+ a memory queue): split-plane epilogue, TMA residual

Used on the B200 (tools/gpurun_job_pair.sh); tests/test_gpu_tunables.py holds the same checks as pytest cases."""
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "tests"), os.path.join(HERE, "..", "oracle")]
from usot_b200 import _lib, ops  # noqa: E402


ALL = int(os.environ.get("PAIR_ALL", "7"))     # tc_cta_pair value compared with 0 for bit-identity (bit 2: every eligible layer)
SEL = int(os.environ.get("PAIR_SEL", "3"))     # tc_cta_pair value timed against 0


def set_knob(name, v):
    _lib.check(_lib.load().usot_set_tunable(name.encode(), v))


OPS = {  # name: (cin, cout, k, stride, pad, dil, h, residual, relu, n)
    "l3_down": (512, 1024, 3, 1, 1, 1, 31, False, False, 16),
    "l3_conv3_res": (256, 1024, 1, 1, 0, 1, 31, True, True, 16),
    "l3_conv2_dil2": (256, 256, 3, 1, 2, 2, 31, False, True, 24),
    "tower_odd_groups": (256, 256, 3, 1, 1, 1, 25, False, True, 25),
    "encoder": (256, 256, 3, 1, 0, 1, 31, False, True, 24),
    "l1_conv3_res": (64, 256, 1, 1, 0, 1, 63, True, True, 8),
    "l2_conv2_s2": (128, 128, 3, 2, 1, 1, 63, False, True, 32),
    "l2_down_s2": (256, 512, 3, 2, 0, 1, 63, False, False, 16),
    "l3_conv1": (1024, 256, 1, 1, 0, 1, 31, False, True, 24),
    "ragged_last_group": (256, 256, 3, 1, 1, 1, 31, False, True, 30),
}


def run_ops(prec):
    bad = 0
    for name, (cin, cout, k, s, p, d, h, res, relu, n) in OPS.items():
        g = torch.Generator(device="cuda").manual_seed(len(name) * 131 + cin)
        x = torch.randn(n, h, h, cin, device="cuda", generator=g).relu_()
        w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
        sc = torch.rand(cout, device="cuda", generator=g) + 0.5
        sh = torch.randn(cout, device="cuda", generator=g) * 0.1
        ho = (h + 2 * p - d * (k - 1) - 1) // s + 1
        r = torch.randn(n, ho, ho, cout, device="cuda", generator=g) if res else None
        set_knob("tc_cta_pair", 0)
        y0 = ops.conv2d_nhwc(x, w, sc, sh, s, p, d, r, relu, prec)
        set_knob("tc_cta_pair", ALL)
        y1 = ops.conv2d_nhwc(x, w, sc, sh, s, p, d, r, relu, prec)
        torch.cuda.synchronize()
        same = torch.equal(y0, y1)
        ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, s, p, d).permute(0, 2, 3, 1)
        ref = ref * sc.double() + sh.double()
        if res:
            ref = ref + r.double()
        if relu:
            ref = ref.relu()
        err = float((y1.double() - ref).abs().max() / ref.abs().max())
        print(f"{prec} {name}: identical={same} rel_err_vs_fp64={err:.2e} max_abs_diff={float((y0 - y1).abs().max()):.3e}", flush=True)
        bad += 0 if same else 1
    return bad


def run_model(prec):
    import usot_oracle as O
    from helpers import load_weights
    from usot_b200 import USOT
    net = USOT(precision=prec)
    net.load_state_dict(load_weights("damp025"))
    net = net.eval().cuda()
    set_knob("graph_max_batch", 0)
    bad = 0
    for batch, nq in ((64, 0), (37, 0), (64, 3)):
        z, x, tb, sb = O.synth_inputs(95, batch=2)
        xb = torch.cat([x * (1.0 + 0.01 * i) + i for i in range((batch + 1) // 2)])[:batch].cuda()
        outs = []
        for knob in (0, ALL):
            set_knob("tc_cta_pair", knob)
            net.template(z[:1].cuda(), tb[:1].cuda())
            if nq:
                mem = net.extract_memory_feature(ori_x=xb[:1].repeat(batch * nq, 1, 1, 1), search_bbox=sb[:1].repeat(batch * nq, 1).cuda())
                score = torch.full((batch, nq), 0.9).cuda()
                out = net.track(xb, mem, score)
            else:
                out = net.track(xb)
            torch.cuda.synchronize()
            outs.append([t.clone() for t in out if torch.is_tensor(t)])
        same = all(torch.equal(a, b) for a, b in zip(*outs))
        print(f"{prec} track batch={batch} nq={nq}: identical={same} tensors={len(outs[0])}", flush=True)
        bad += 0 if same else 1
    # timing of the batch-256 step, both ways (CUDA events; informal: bench.py --tunable tc_cta_pair=3 is the measurement)
    z, x, tb, sb = O.synth_inputs(96, batch=2)
    xb = x[:1].repeat(256, 1, 1, 1).cuda()
    for knob in (0, SEL, 0, SEL):
        set_knob("tc_cta_pair", knob)
        net.template(z[:1].cuda(), tb[:1].cuda())
        for _ in range(3):
            net.track(xb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            net.track(xb)
        e1.record()
        torch.cuda.synchronize()
        print(f"{prec} track batch=256 tc_cta_pair={knob}: {e0.elapsed_time(e1) / 8:.3f} ms/step", flush=True)
    return bad


if __name__ == "__main__":
    what, prec = sys.argv[1], sys.argv[2]
    t0 = time.time()
    bad = run_ops(prec) if what == "ops" else run_model(prec)
    print(f"{what} {prec}: {'OK' if bad == 0 else f'{bad} MISMATCHES'} ({time.time() - t0:.1f} s)")
    sys.exit(1 if bad else 0)
