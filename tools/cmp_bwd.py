"""Gradient error of the tcgen05 backward against the fp32 generic kernels on the same inputs (GPU only).
usage: cmp_bwd.py [B] [R] [Bt]   (DVLP_TC_BWD=1 selects the first tcgen05 backward)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
R = int(sys.argv[2]) if len(sys.argv) > 2 else 80
Bt = int(sys.argv[3]) if len(sys.argv) > 3 else B
b = synth.make_batch(max(B, Bt), R, 32, 256, seed=1).to("cuda")
im, s = b.im[:B].to(torch.bfloat16).contiguous(), b.s[:Bt].to(torch.bfloat16).contiguous()
imm, sm = b.im_mask[:B].contiguous(), b.s_mask[:Bt].contiguous()
for focal in (1, 0):
    res = {}
    for impl in ("generic", "tcgen05"):
        sc, st = ops.rw_scores_fwd_raw(im, s, imm, sm, 20.0, focal, impl)
        p = torch.softmax(sc * 20.0, dim=1)
        g = (p - torch.eye(B, Bt, device="cuda")) * (20.0 / B)      # a realistic d_scores (RWALoss-like)
        d_im, d_s = ops.rw_scores_bwd_raw(st, g if impl == "generic" else res["g"])
        if impl == "generic":
            res["g"] = g
        res[impl] = (sc, d_im, d_s)
    rel = lambda x, y: float((x - y).norm() / y.norm())
    print(f"focal={focal} B={B} Bt={Bt} R={R}: scores max abs {float((res['tcgen05'][0] - res['generic'][0]).abs().max()):.2e}  "
          f"d_im rel {rel(res['tcgen05'][1], res['generic'][1]):.3e}  d_s rel {rel(res['tcgen05'][2], res['generic'][2]):.3e}  "
          f"finite {bool(torch.isfinite(res['tcgen05'][1]).all() and torch.isfinite(res['tcgen05'][2]).all())}")
