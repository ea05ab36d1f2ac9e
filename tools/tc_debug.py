"""GPU-side diagnostics for the tcgen05 path: compares it with the generic fp32 kernels and prints
where (which videos / text slots / tiles) the two disagree.  Development aid, not a test."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import ops, synth

def main():
    shapes = [(8, 80, 32, 8), (32, 80, 32, 32), (5, 80, 32, 5), (128, 80, 32, 128), (16, 32, 32, 24), (16, 64, 32, 8), (16, 48, 32, 12)]
    if len(sys.argv) > 1:
        shapes = shapes[:int(sys.argv[1])]
    for (Bv, R, W, Bt) in shapes:
        b = synth.make_batch(Bv, R, W, 256, seed=0, Bt=Bt).to("cuda")
        im, s = b.im.to(torch.bfloat16), b.s.to(torch.bfloat16)
        for focal in (1, 0):
            ref = ops.rw_scores(im, s, b.im_mask, b.s_mask, focal_equal=focal, impl="generic")
            try:
                got = ops.rw_scores(im, s, b.im_mask, b.s_mask, focal_equal=focal, impl="tcgen05")
                torch.cuda.synchronize()
            except Exception as ex:
                print(f"shape {(Bv,R,W,Bt)} focal={focal}: EXCEPTION {ex}")
                return 1
            err = (got - ref).abs()
            print(f"shape {(Bv,R,W,Bt)} focal={focal}: max err {err.max().item():.3e} mean {err.mean().item():.3e} "
                  f"nan {int(torch.isnan(got).sum())} ref range [{ref.min().item():.3f},{ref.max().item():.3f}]")
            if not (err.max() < 2e-3):
                print(" per-video max err:", [f"{x:.1e}" for x in err.max(1).values.tolist()[:16]])
                print(" per-text  max err:", [f"{x:.1e}" for x in err.max(0).values.tolist()[:16]])
                print(" got[0,:8]", got[0, :8].tolist())
                print(" ref[0,:8]", ref[0, :8].tolist())
    return 0

if __name__ == "__main__":
    sys.exit(main())
