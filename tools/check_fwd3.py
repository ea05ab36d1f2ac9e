"""GPU check of the large-shape forward kernel (pair_fwd3.cu) against reference fixtures and the oracle."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from demovlp_b200 import _lib, ops, synth
from oracle import rw_oracle as O
from tests.util import batch_for, load_golden

lib = _lib.load()
torch.manual_seed(0)
worst = 0.0
for name in ["loss_pt_r30", "loss_didemo_small", "loss_ft_ragged", "scores_odd_3x4_r17_w9", "scores_mc_5x5"]:
    g = load_golden(name)
    b, focal = batch_for(g)
    d = b.to("cuda")
    R, W = d.im.shape[1], d.s.shape[1]
    for dtype in (torch.bfloat16, torch.float32):
        sc, _ = ops.rw_scores_fwd_raw(d.im.to(dtype).contiguous(), d.s.to(dtype).contiguous(), d.im_mask, d.s_mask, 20.0,
                                      int(focal == "equal"), "auto", train=False)
        torch.cuda.synchronize()
        err = np.abs(sc.cpu().numpy() - g["scores"]).max()
        worst = max(worst, err)
        print(f"{name:28s} R={R:3d} W={W:3d} path={lib.dvlp_rw_path(R, W, 256, 0)} {str(dtype):15s} max|err|={err:.3e}", flush=True)
# oracle comparisons on shapes without fixtures (ragged text counts, prob focal)
for (Bv, R, W, Bt) in [(5, 288, 64, 7), (3, 240, 99, 4), (6, 30, 99, 9), (4, 100, 20, 6), (7, 48, 64, 5), (3, 130, 128, 3)]:
    b = synth.make_batch(max(Bv, Bt), R, W, 256, seed=41)
    im, s, imm, sm = b.im[:Bv], b.s[:Bt], b.im_mask[:Bv].contiguous(), b.s_mask[:Bt].contiguous()
    for focal in ("equal", "prob"):
        want = O.pair_scores(im, s, imm, sm, focal=focal)
        got, _ = ops.rw_scores_fwd_raw(im.cuda().to(torch.bfloat16).contiguous(), s.cuda().to(torch.bfloat16).contiguous(),
                                       imm.cuda(), sm.cuda(), 20.0, int(focal == "equal"), "auto", train=False)
        torch.cuda.synchronize()
        err = float((got.cpu() - want).abs().max())
        worst = max(worst, err)
        print(f"oracle Bv={Bv} Bt={Bt} R={R:3d} W={W:3d} path={lib.dvlp_rw_path(R, W, 256, 0)} focal={focal:5s} max|err|={err:.3e}", flush=True)
print("WORST", worst)
# timing
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
    return min(ts)
for (Bv, Bt, R, W) in [(64, 512, 288, 64), (1000, 1000, 240, 99), (128, 128, 30, 99), (1000, 1000, 80, 32)]:
    b = synth.make_batch(max(Bv, Bt), R, W, 256, seed=4).to("cuda")
    im, s = b.im[:Bv].to(torch.bfloat16).contiguous(), b.s[:Bt].to(torch.bfloat16).contiguous()
    imm, sm = b.im_mask[:Bv].contiguous(), b.s_mask[:Bt].contiguous()
    ms = timeit(lambda: ops.rw_scores_fwd_raw(im, s, imm, sm, 20.0, 1, "auto", train=False))
    print(f"time fwd Bv={Bv} Bt={Bt} R={R} W={W} path={lib.dvlp_rw_path(R, W, 256, 0)}: {ms:.3f} ms  "
          f"{Bv * Bt / ms / 1e3:.2f} M pairs/s  {6 * R * W * 256 * Bv * Bt / ms / 1e9:.1f} TFLOP/s (fwd algorithmic)", flush=True)
