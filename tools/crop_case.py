"""256 crops of 255x255 from one 480x640 uint8 frame through the C ABI (usot_crop_resize) -- used under ncu."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from usot_b200 import tracker_ops
rng = np.random.default_rng(0)
fr = tracker_ops.upload_frame(rng.integers(0, 256, (480, 640, 3), dtype=np.uint8))
rows = torch.tensor([[0, 40 + (i % 64) * 5, 30 + (i // 64) * 40, 260 + (i % 7) * 13] for i in range(256)], dtype=torch.int32).cuda()
fill = torch.full((256, 3), 117, dtype=torch.uint8).cuda()
for _ in range(3):
    out = tracker_ops.crop_resize(fr, rows, fill, 255)
torch.cuda.synchronize()
print(tuple(out.shape))
