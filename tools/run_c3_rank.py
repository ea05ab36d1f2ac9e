"""One rank's share of config 3 on a single GPU (128 videos x 1024 texts): local-path kernels only, for ncu / timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import ops, synth, _lib

Bv, Bt = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
b = synth.make_batch(Bt, 80, 32, 256, seed=2).to("cuda")
im, s = b.im[:Bv].to(torch.bfloat16).contiguous(), b.s.to(torch.bfloat16)
imm, sm = b.im_mask[:Bv].contiguous(), b.s_mask
for it in range(3):
    sc, st = ops.rw_scores_fwd_raw(im, s, imm, sm, 20.0, 1, "auto")
    part, dsc = ops.rwa_tail_raw(sc, 20.0, 0, Bt, True)
    ops.rw_scores_bwd_raw(st, dsc)
torch.cuda.synchronize()
_lib.profile_enable(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(10):
    sc, st = ops.rw_scores_fwd_raw(im, s, imm, sm, 20.0, 1, "auto")
    part, dsc = ops.rwa_tail_raw(sc, 20.0, 0, Bt, True)
    ops.rw_scores_bwd_raw(st, dsc)
e1.record()
torch.cuda.synchronize()
print({k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in _lib.profile_read().items()}, "us per launch;",
      round(e0.elapsed_time(e1) * 100, 1), "us per local step")
