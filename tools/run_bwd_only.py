import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import ops, synth
B = 128
b = synth.make_batch(B, 80, 32, 256, seed=1).to("cuda")
im, s = b.im.to(torch.bfloat16), b.s.to(torch.bfloat16)
g = torch.randn(B, B, device="cuda")
sc, st = ops.rw_scores_fwd_raw(im, s, b.im_mask, b.s_mask, 20.0, 1, "auto")
for it in range(4):
    ops.rw_scores_bwd_raw(st, g)
torch.cuda.synchronize()
