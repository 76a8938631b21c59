"""Weight-gradient launches of the training path at the BASELINE config-4 shard (64 search-size crops) through the C ABI -- used under ncu.
    python tools/wgrad_case.py [l3_conv2|l3_conv3|l1_conv3|tower] [precision]"""
import sys
import torch
sys.path.insert(0, ".")
from usot_b200 import train

CASES = {  # name: (n, h, cin, cout, k, stride, pad, dil)
    "l3_conv2": (64, 31, 256, 256, 3, 1, 2, 2),
    "l3_conv3": (64, 31, 256, 1024, 1, 1, 0, 1),
    "l1_conv3": (64, 63, 64, 256, 1, 1, 0, 1),
    "tower": (64, 25, 256, 256, 3, 1, 1, 1),
}
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
for name in (sys.argv[1] if len(sys.argv) > 1 else "l3_conv2").split(","):
    n, h, cin, cout, k, s, p, d = CASES[name]
    ho = (h + 2 * p - d * (k - 1) - 1) // s + 1
    x = torch.randn(n, h, h, cin, device="cuda")
    gy = torch.randn(n, ho, ho, cout, device="cuda") * 1e-5
    for _ in range(2):
        gw = train.conv_wgrad(x, gy, (cout, cin, k, k), s, (p, p), (d, d), prec)
    torch.cuda.synchronize()
    print(name, tuple(gw.shape), float(gw.abs().max()))
