"""Mint the BASELINE config-4 fixtures (1000 x 1000 eval matrix) from the UNMODIFIED reference (build container only).

    python oracle/make_golden_eval1000.py [R W]      # default 80 32; writes tests/golden/eval_1000_r{R}_w{W}.npz

TEST INFRASTRUCTURE.  The reference's own eval route: RWALoss.get_sim_by_segment (model/loss.py:73-103, 8 x 8 tiles) on
CPU tensors, plus sim_matrix(text, video) with the orientation quirk of trainer/trainer_dist.py:369-385, then the
internal rank computation of t2v_metrics (model/metric.py:22-38).  ~2.5 min (R=80, W=32) / ~11 min (R=240, W=99) on 8
cores.  Inputs are regenerated from the seed at test time (synth.make_batch(1000, R, W, 256, seed=3, kappa=0.12))."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demovlp_b200 import synth          # noqa: E402
from oracle import ref_loader           # noqa: E402


def main():
    R, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (80, 32)
    torch.set_num_threads(os.cpu_count())
    L = ref_loader.loss_module()
    M = ref_loader.metric_module()
    sim_matrix = ref_loader.sim_matrix_fn()
    b = synth.make_batch(1000, R, W, 256, seed=3, kappa=0.12)
    t0 = time.time()
    mod = L.RWALoss(focal_type="equal")
    with torch.no_grad():
        local = mod.get_sim_by_segment(b.im, b.s, b.im_mask, b.s_len, b.s_mask, device="cpu")
        gsim = sim_matrix(b.g_text, b.g_vid).numpy()
        full = gsim + local
    t2v, v2t = M.t2v_metrics(full), M.v2t_metrics(full)
    d = -full
    sd = np.sort(d, axis=1)
    rows, cols = np.where((sd - np.diag(d)[:, None]) == 0)
    _, first = np.unique(rows, return_index=True)
    keys = sorted(t2v)
    out = os.path.join(ROOT, "tests", "golden", f"eval_1000_r{R}_w{W}.npz")
    np.savez_compressed(out, meta=np.array([1000, R, W, 256, 3, 0.12]),
                        checksum=np.array([float(b.im.double().sum()), float(b.s.double().sum())]),
                        # both addends are float32-exact; the float64 matrix the trainer builds is their float64 sum
                        local=local.astype(np.float32), gsim=gsim.astype(np.float32),
                        exact=np.array([bool((local.astype(np.float32).astype(np.float64) + gsim.astype(np.float64) == full).all())]),
                        t2v_ranks=cols[first].astype(np.int32), keys=np.array(keys),
                        t2v=np.array([float(t2v[k]) for k in keys]), v2t=np.array([float(v2t[k]) for k in keys]))
    print("wrote", out, f"{time.time() - t0:.0f}s", t2v)


if __name__ == "__main__":
    main()
