"""Small end-to-end case for compute-sanitizer: template + track with a memory queue (batch 5, N_q = 2) in fp16x3 and fp16 --
covers the multi-image conv tiles, the fused Conf_Fusion epilogue, the fused stem + max-pool and the GroupDW / pred kernels.
`python tools/sanitize_case.py 20` runs batch 20 with tc_cta_pair = 7: every eligible conv launch as a CTA pair (cta_group::2, remote barriers,
multicast commits), an odd number of image groups included."""
import sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "oracle"); sys.path.insert(0, "tests")
import usot_oracle as O
from helpers import load_weights
from usot_b200 import USOT, _lib
_lib.check(_lib.load().usot_set_tunable(b"graph_max_batch", 0))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
if B > 5:
    _lib.check(_lib.load().usot_set_tunable(b"tc_cta_pair", 7))
for prec in ("fp16x3", "fp16"):
    n = USOT(precision=prec)
    n.load_state_dict(load_weights("damp025"))
    n = n.eval().cuda()
    z, x, tb, sb = O.synth_inputs(5, batch=B)
    n.template(z.cuda(), tb.cuda())
    feats = n.extract_memory_feature(ori_x=x[:4].cuda(), search_bbox=sb[:4].cuda())
    pick = torch.tensor([(b * 2 + q) % 4 for b in range(B) for q in range(2)]).cuda()
    mem = feats[pick].contiguous(memory_format=torch.channels_last)
    out = n.track(x.cuda(), mem, torch.full((B, 2), 0.9).cuda())
    torch.cuda.synchronize()
    print(prec, [float(t.abs().max()) for t in out[:3]])
