"""BASELINE config 3 on one GPU (1024 videos x 1024 texts, R = 80, W = 32) fwd + bwd a few times: ncu target.
usage: run_c3.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import ops, synth, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = synth.make_batch(B, 80, 32, 256, seed=2).to("cuda")
im, s = b.im.to(torch.bfloat16), b.s.to(torch.bfloat16)
g = torch.randn(B, B, device="cuda") * 1e-3
for it in range(2):
    sc, st = ops.rw_scores_fwd_raw(im, s, b.im_mask, b.s_mask, 20.0, 1, "auto")
    ops.rw_scores_bwd_raw(st, g, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
_lib.profile_enable(True)
for it in range(3):
    sc, st = ops.rw_scores_fwd_raw(im, s, b.im_mask, b.s_mask, 20.0, 1, "auto")
    ops.rw_scores_bwd_raw(st, g, out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print({k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in _lib.profile_read().items()}, "us per launch")
