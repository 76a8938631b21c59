"""SHA-256 of the outputs of fixed conv / track() calls: run under two builds of the library and diff the lines to prove that a kernel change
kept every result bit for bit.      python tools/output_hashes.py [path/to/other/libusot_b200.so]"""
import hashlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, ".."), os.path.join(HERE, "..", "tests"), os.path.join(HERE, "..", "oracle")]
from usot_b200 import _lib  # noqa: E402

if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.abspath(sys.argv[1])   # (before the first load())
from usot_b200 import ops  # noqa: E402


def h(t):
    return hashlib.sha256(t.detach().float().cpu().contiguous().numpy().tobytes()).hexdigest()[:16]


CASES = {  # name: (cin, cout, k, stride, pad, dil, h, residual, relu, n)
    "l3_down": (512, 1024, 3, 1, 1, 1, 31, False, False, 16),
    "l3_conv3_res": (256, 1024, 1, 1, 0, 1, 31, True, True, 16),
    "l1_conv3_res": (64, 256, 1, 1, 0, 1, 63, True, True, 8),
    "l2_conv2_s2": (128, 128, 3, 2, 1, 1, 63, False, True, 32),
    "small": (256, 256, 3, 1, 1, 1, 25, False, True, 3),
}
for prec in ("fp16x3", "fp16"):
    for split_out in ("0", "2"):
        os.environ["USOT_DEBUG_SPLIT_OUT"] = split_out
        for name, (cin, cout, k, s, p, d, hh, res, relu, n) in CASES.items():
            g = torch.Generator(device="cuda").manual_seed(len(name) * 131 + cin)
            x = torch.randn(n, hh, hh, cin, device="cuda", generator=g).relu_()
            w = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
            sc = torch.rand(cout, device="cuda", generator=g) + 0.5
            sh = torch.randn(cout, device="cuda", generator=g) * 0.1
            ho = (hh + 2 * p - d * (k - 1) - 1) // s + 1
            r = torch.randn(n, ho, ho, cout, device="cuda", generator=g) if res else None
            print(f"conv {prec} split_out={split_out} {name}: {h(ops.conv2d_nhwc(x, w, sc, sh, s, p, d, r, relu, prec))}", flush=True)
os.environ.pop("USOT_DEBUG_SPLIT_OUT", None)

import usot_oracle as O  # noqa: E402
from helpers import load_weights  # noqa: E402
from usot_b200 import USOT  # noqa: E402
for prec in ("fp16x3", "fp16"):
    net = USOT(precision=prec)
    net.load_state_dict(load_weights("damp025"))
    net = net.eval().cuda()
    _lib.check(_lib.load().usot_set_tunable(b"graph_max_batch", 0))
    for batch, nq in ((3, 0), (37, 0), (64, 3)):
        z, x, tb, sb = O.synth_inputs(95, batch=2)
        xb = torch.cat([x * (1.0 + 0.01 * i) + i for i in range((batch + 1) // 2)])[:batch].cuda()
        net.template(z[:1].cuda(), tb[:1].cuda())
        if nq:
            mem = net.extract_memory_feature(ori_x=xb[:1].repeat(batch * nq, 1, 1, 1), search_bbox=sb[:1].repeat(batch * nq, 1).cuda())
            out = net.track(xb, mem, torch.full((batch, nq), 0.9).cuda())
        else:
            out = net.track(xb)
        print(f"track {prec} batch={batch} nq={nq}: " + " ".join(h(t) for t in out if torch.is_tensor(t)), flush=True)
