"""GPU-side diagnostics for the tcgen05 backward: compares with the generic fp32 backward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import ops, synth

def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp(min=1e-30))

def main():
    shapes = [(4, 80, 32, 4), (8, 80, 32, 8), (5, 80, 32, 7), (32, 80, 32, 32), (16, 32, 32, 24), (16, 64, 32, 8), (16, 48, 32, 12), (128, 80, 32, 128)]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    for (Bv, R, W, Bt) in shapes:
        b = synth.make_batch(Bv, R, W, 256, seed=0, Bt=Bt).to("cuda")
        im, s = b.im.to(torch.bfloat16), b.s.to(torch.bfloat16)
        g = torch.randn(Bv, Bt, device="cuda", generator=torch.Generator("cuda").manual_seed(3))
        for focal in (1, 0):
            sc_r, st_r = ops.rw_scores_fwd_raw(im, s, b.im_mask, b.s_mask, 20.0, focal, "generic")
            di_r, ds_r = ops.rw_scores_bwd_raw(st_r, g)
            try:
                sc, st = ops.rw_scores_fwd_raw(im, s, b.im_mask, b.s_mask, 20.0, focal, "tcgen05")
                di, ds = ops.rw_scores_bwd_raw(st, g)
                torch.cuda.synchronize()
            except Exception as ex:
                print(f"shape {(Bv,R,W,Bt)} focal={focal}: EXCEPTION {ex}")
                return 1
            print(f"shape {(Bv,R,W,Bt)} focal={focal}: d_im rel {rel(di, di_r):.3e}  d_s rel {rel(ds, ds_r):.3e} "
                  f"nan {int(torch.isnan(di).sum()) + int(torch.isnan(ds).sum())} |d_im| {float(di_r.norm()):.3e} got {float(di.norm()):.3e} "
                  f"|d_s| {float(ds_r.norm()):.3e} got {float(ds.norm()):.3e}")
            if rel(di, di_r) > 2e-2 or rel(ds, ds_r) > 2e-2:
                pv = [(rel(di[v], di_r[v])) for v in range(min(Bv, 8))]
                pt = [(rel(ds[t], ds_r[t])) for t in range(min(Bt, 8))]
                print("   per-video rel:", [f"{x:.2e}" for x in pv])
                print("   per-text  rel:", [f"{x:.2e}" for x in pt])
                print("   d_im[0,0,:6] got", di[0, 0, :6].tolist())
                print("   d_im[0,0,:6] ref", di_r[0, 0, :6].tolist())
                print("   d_s[0,0,:6] got", ds[0, 0, :6].tolist())
                print("   d_s[0,0,:6] ref", ds_r[0, 0, :6].tolist())
                # per-region / per-word error profile of item 0
                print("   d_im[0] per-region rel:", [f"{rel(di[0, r], di_r[0, r]):.1e}" for r in range(0, R, max(1, R // 10))])
                print("   d_s[0] per-word rel:", [f"{rel(ds[0, w], ds_r[0, w]):.1e}" for w in range(0, W, 4)])
                return 2
    return 0

if __name__ == "__main__":
    sys.exit(main())
