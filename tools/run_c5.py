"""One rank's share of BASELINE config 5 (64 videos x 512 texts, R = 288, W = 64) fwd + bwd a few times: ncu target.
usage: run_c5.py [Bv Bt R W]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import ops, synth, _lib

Bv, Bt, R, W = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (64, 512, 288, 64)
b = synth.make_batch(max(Bv, Bt), R, W, 256, seed=4).to("cuda")
im, s = b.im[:Bv].to(torch.bfloat16).contiguous(), b.s[:Bt].to(torch.bfloat16).contiguous()
imm, sm = b.im_mask[:Bv].contiguous(), b.s_mask[:Bt].contiguous()
g = torch.randn(Bv, Bt, device="cuda")
for it in range(2):
    sc, st = ops.rw_scores_fwd_raw(im, s, imm, sm, 20.0, 1, "auto")
    ops.rw_scores_bwd_raw(st, g)
torch.cuda.synchronize()
_lib.profile_enable(True)
for it in range(5):
    sc, st = ops.rw_scores_fwd_raw(im, s, imm, sm, 20.0, 1, "auto")
    ops.rw_scores_bwd_raw(st, g)
torch.cuda.synchronize()
print({k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in _lib.profile_read().items()}, "us per launch")
