"""GPU check of the large-shape backward (pair_bwd3.cu + grad_gemm3.cu) against the oracle's autograd."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import _lib, ops, synth
from oracle import rw_oracle as O
from tests.util import rel_err

lib = _lib.load()
shapes = [(5, 288, 64, 7), (3, 240, 99, 4), (6, 30, 99, 9), (4, 100, 20, 6), (7, 48, 64, 5), (3, 130, 128, 3), (16, 288, 64, 16)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
worst = 0.0
for (Bv, R, W, Bt) in shapes:
    b = synth.make_batch(max(Bv, Bt), R, W, 256, seed=43)
    im, s = b.im[:Bv].to(torch.bfloat16), b.s[:Bt].to(torch.bfloat16)
    imm, sm = b.im_mask[:Bv].contiguous(), b.s_mask[:Bt].contiguous()
    w = torch.randn(Bv, Bt, generator=torch.Generator().manual_seed(5))
    for focal in ("equal", "prob"):
        ic = im.float().clone().requires_grad_(True)
        sc = s.float().clone().requires_grad_(True)
        want = O.pair_scores(ic, sc, imm, sm, focal=focal)
        (want * w).sum().backward()
        x = im.cuda().requires_grad_(True)
        y = s.cuda().requires_grad_(True)
        got = ops.rw_scores(x, y, imm.cuda(), sm.cuda(), focal_equal=(focal == "equal"), impl="auto")
        (got * w.cuda()).sum().backward()
        torch.cuda.synchronize()
        es = float((got.detach().cpu() - want.detach()).abs().max())
        ei, et = rel_err(x.grad.float().cpu(), ic.grad), rel_err(y.grad.float().cpu(), sc.grad)
        fin = bool(torch.isfinite(x.grad).all() and torch.isfinite(y.grad).all())
        worst = max(worst, ei, et)
        print(f"Bv={Bv} Bt={Bt} R={R:3d} W={W:3d} path={lib.dvlp_rw_path(R, W, 256, 1)} focal={focal:5s} score err {es:.2e}  "
              f"d_im rel {ei:.3e}  d_s rel {et:.3e} finite={fin}", flush=True)
print("WORST grad rel err", worst)

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
    return min(ts)

_lib.profile_enable(True)
for (Bv, Bt, R, W) in [(64, 512, 288, 64), (32, 64, 240, 99), (128, 128, 30, 99)]:
    b = synth.make_batch(max(Bv, Bt), R, W, 256, seed=4).to("cuda")
    im, s = b.im[:Bv].to(torch.bfloat16).contiguous(), b.s[:Bt].to(torch.bfloat16).contiguous()
    imm, sm = b.im_mask[:Bv].contiguous(), b.s_mask[:Bt].contiguous()
    g = torch.randn(Bv, Bt, device="cuda")
    def fb():
        sc, st = ops.rw_scores_fwd_raw(im, s, imm, sm, 20.0, 1, "auto")
        ops.rw_scores_bwd_raw(st, g)
    ms = timeit(fb)
    _lib.profile_read()
    fb(); torch.cuda.synchronize()
    prof = {k: round(v[0], 3) for k, v in _lib.profile_read().items()}
    print(f"time fwd+bwd Bv={Bv} Bt={Bt} R={R} W={W}: {ms:.3f} ms  {Bv * Bt / ms / 1e3:.2f} M pairs/s  "
          f"{18 * R * W * 256 * Bv * Bt / ms / 1e9:.1f} TFLOP/s algorithmic   kernels(ms): {prof}", flush=True)
