"""Per-phase timeline of the tcgen05 backward kernel for CTA 0 (clock64 stamps written by the TR()
macros in csrc/pair_tc_bwd.cu through the debug hook dvlp_debug_set_trace).  GPU only; debugging aid."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demovlp_b200 import _lib, ops, synth

ROLES = ["tma", "mma", "grpA(w4)", "grpB(w8)", "grpB(w12)"]
if os.environ.get("DVLP_TC_BWD", "2") == "1":
    EV_MMA = ["S issue", "S committed", "U ready", "Y committed", "UD ready", "dG committed", "P ready", "grad committed"]
    EV_GRP = ["start", "S done", "pass1 done", "norm bar", "U written", "Y done", "UD written", "t sums", "dG done",
              "P written", "grad done", "flushed"]
else:   # pair_bwd2.cu
    ROLES = ["tma", "mma", "cw(g0,q0)", "cw(last)", "-"]
    EV_MMA = ["U ready", "Y issued", "UD ready", "dG issued", "next S issued"]
    EV_GRP = ["start", "vec full", "S done", "U arrive", "Y done", "UD arrive", "tB bar", "P computed", "dG done",
              "P arrive", "flushed", "loop top"]


def main():
    global EV_MMA, EV_GRP
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    which = sys.argv[2] if len(sys.argv) > 2 else "bwd"
    if which == "fwd":   # pair_fwd2.cu built with -DFWD2_TRACE (tools/variants.sh pair_fwd2 trace "-DFWD2_TRACE")
        EV_MMA = ["S ahead issued", "Gbd full", "U ready", "Y issued"]
        EV_GRP = ["loop top", "vec full", "S loaded", "pass0 done", "deferred done", "1a done", "bar passed", "2a done",
                  "1b done", "U arrive", "num2 done"]
    lib = _lib.load()
    lib.dvlp_debug_set_trace.argtypes = [C.c_void_p]
    lib.dvlp_debug_set_trace.restype = None
    b = synth.make_batch(B, 80, 32, 256, seed=1).to("cuda")
    im, s = b.im.to(torch.bfloat16), b.s.to(torch.bfloat16)
    g = torch.randn(B, B, device="cuda") * 1e-3
    sc, st = ops.rw_scores_fwd_raw(im, s, b.im_mask, b.s_mask, 20.0, 1, "tcgen05")
    for _ in range(2):
        ops.rw_scores_bwd_raw(st, g)
    trace = torch.zeros(5 * 32 * 16, dtype=torch.int64, device="cuda")
    lib.dvlp_debug_set_trace(trace.data_ptr())
    if which == "bwd":
        ops.rw_scores_bwd_raw(st, g)
    else:
        ops.rw_scores_fwd_raw(im, s, b.im_mask, b.s_mask, 20.0, 1, "tcgen05")
    torch.cuda.synchronize()
    lib.dvlp_debug_set_trace(None)
    t = trace.cpu().view(5, 32, 16)
    t0 = int(t[t > 0].min())
    for k in (6, 7, 8):
        base = int(t[2, k, 0]) if t[2, k, 0] > 0 else t0
        print(f"--- tile {k} (cycles relative to group A start of this tile; tile period "
              f"{int(t[2, k + 1, 0] - t[2, k, 0])} cycles)")
        for role in range(1, 5):
            names = EV_MMA if role == 1 else EV_GRP
            row = []
            for e, nm in enumerate(names):
                v = int(t[role, k, e])
                if v > 0:
                    row.append(f"{nm}={v - base}")
            print(f"  {ROLES[role]:10s} " + "  ".join(row))


if __name__ == "__main__":
    main()
