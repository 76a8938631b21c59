"""Prediction convs (bbox_pred 256->4 with exp epilogue, cls_pred 256->1) at the BASELINE batch (256 maps of 25x25x256)
through the C ABI -- used under ncu."""
import sys
import torch
sys.path.insert(0, ".")
from usot_b200 import ops
B, C, R = 256, 256, 25
x = torch.randn(B, R, R, C, device="cuda")
w4, b4 = torch.randn(4, C, 3, 3, device="cuda") * 0.02, torch.randn(4, device="cuda") * 0.1
w1, b1 = torch.randn(1, C, 3, 3, device="cuda") * 0.02, torch.randn(1, device="cuda") * 0.1
adjust, bias4 = torch.tensor([0.7], device="cuda"), torch.randn(4, device="cuda") * 0.3
for _ in range(3):
    y4 = ops.pred_conv(x, w4, b4, mode=1, adjust=adjust, bias4=bias4)
    y1 = ops.pred_conv(x, w1, b1, mode=0, mul=0.1)
torch.cuda.synchronize()
print(tuple(y4.shape), tuple(y1.shape))
