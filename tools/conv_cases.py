"""Run representative conv launches through the C ABI (used under ncu; see profiles/)."""
import sys
import torch
sys.path.insert(0, ".")
from usot_b200 import ops

CASES = {  # name: (cin, cout, k, stride, pad, dil, h, residual, relu)
    "l3_conv3": (256, 1024, 1, 1, 0, 1, 31, True, True),
    "l3_down": (512, 1024, 3, 1, 1, 1, 31, False, False),
    "l1_conv3": (64, 256, 1, 1, 0, 1, 63, True, True),
    "l3_conv2": (256, 256, 3, 1, 2, 2, 31, False, True),
    "l3_conv1": (1024, 256, 1, 1, 0, 1, 31, False, True),
    "l3_conv3_nores": (256, 1024, 1, 1, 0, 1, 31, False, True),
    "l1_conv3_nores": (64, 256, 1, 1, 0, 1, 63, False, True),
    "tower": (256, 256, 3, 1, 1, 1, 25, False, True),
}
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
prec = sys.argv[3] if len(sys.argv) > 3 else "fp16x3"
for name in sys.argv[1].split(","):
    cin, cout, k, s, p, d, h, res, relu = CASES[name]
    x = torch.randn(n, h, h, cin, device="cuda")
    w = torch.randn(cout, cin, k, k, device="cuda") / (cin * k * k) ** 0.5
    sc, sh = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    ho = (h + 2 * p - d * (k - 1) - 1) // s + 1
    r = torch.randn(n, ho, ho, cout, device="cuda") if res else None
    for _ in range(2):
        y = ops.conv2d_nhwc(x, w, sc, sh, s, p, d, r, relu, prec)
    torch.cuda.synchronize()
    print(name, tuple(y.shape))
