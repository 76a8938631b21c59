"""Fused GroupDW at the BASELINE batch (256 crops, template batch 1) through the C ABI -- used under ncu."""
import sys
import torch
sys.path.insert(0, ".")
from usot_b200 import ops
B, C, F = 256, 256, 31
mk = lambda *s: torch.randn(*s, device="cuda")
x = [mk(B, F - 2, F - 2, C), mk(B, F - 4, F - 2, C), mk(B, F - 2, F - 4, C)]
z = [mk(1, 5, 5, C), mk(1, 3, 5, C), mk(1, 5, 3, C)]
w = torch.tensor([1.0, 0.5, 1.5], device="cuda")
for _ in range(3):
    y = ops.groupdw_xcorr(x, z, w, B)
torch.cuda.synchronize()
print(tuple(y.shape))
