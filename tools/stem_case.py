"""Fused stem + max-pool (space-to-depth pre-pass + TMA-fed tcgen05 GEMM with the pooling epilogue) at the BASELINE batch through the
C ABI -- used under ncu.   python tools/stem_case.py [fp16x3|fp16] [batch]"""
import sys
import torch
sys.path.insert(0, ".")
from usot_b200 import ops
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
g = torch.Generator().manual_seed(3)
x = (torch.rand(B, 3, 255, 255, generator=g) * 255.0).cuda()
w = torch.randn(64, 3, 7, 7, generator=g) * 0.05
scale, shift = torch.rand(64, generator=g) * 0.02 + 0.005, torch.randn(64, generator=g) * 0.1
for _ in range(3):
    y = ops.stem_maxpool(x, w, scale, shift, prec)
torch.cuda.synchronize()
print(tuple(y.shape), float(y.mean()))
