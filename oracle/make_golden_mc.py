"""Mint multiple-choice fixtures from the UNMODIFIED reference (build container only).

    python oracle/make_golden_mc.py        # writes tests/golden/mc_*.npz

TEST INFRASTRUCTURE.  Replays trainer/trainer_dist.py:792-829 per question: the video is expanded n_options times,
``get_sim`` scores the 5 x 5 problem, ``sim_matrix(text, video)`` gives the global part and the prediction is
``sims_batch[0].max(0)[1]``.  Question q uses video 5q and, as its options, texts 5q .. 5q+4 of synth.make_batch(5 Q, R, W, 256, seed) in a seeded
shuffled order (stored as ``perm``)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demovlp_b200 import synth          # noqa: E402
from oracle import ref_loader           # noqa: E402

CASES = [("mc_q64_r80_w32", 64, 80, 32, 21), ("mc_q16_r240_w99", 16, 240, 99, 22), ("mc_q24_r30_w99", 24, 30, 99, 23)]
N_OPT = 5


def main():
    L = ref_loader.loss_module()
    sim_matrix = ref_loader.sim_matrix_fn()
    mod = L.RWALoss(focal_type="equal")
    for name, Q, R, W, seed in CASES:
        b = synth.make_batch(Q * N_OPT, R, W, 256, seed, 0.2)
        # shuffle the options of every question (seeded) so that the right answer is not always option 0
        gen = torch.Generator().manual_seed(seed)
        perm = torch.stack([q * N_OPT + torch.randperm(N_OPT, generator=gen) for q in range(Q)])
        local0, glob0, pred = [], [], []
        with torch.no_grad():
            for q in range(Q):
                lo = q * N_OPT
                obj = b.im[lo:lo + 1].expand(N_OPT, -1, -1)
                om = b.im_mask[lo:lo + 1].expand(N_OPT, -1)
                txt, tm, tl = b.s[perm[q]], b.s_mask[perm[q]], b.s_len[perm[q]]
                g = sim_matrix(b.g_text[perm[q]], b.g_vid[lo:lo + 1].expand(N_OPT, -1))
                loc = mod.get_sim(obj, txt, om, tl, tm)
                sims = g + loc
                local0.append(loc[0].numpy()); glob0.append(g[0].numpy()); pred.append(int(sims[0].max(0)[1]))
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                            meta=np.array([Q, R, W, 256, seed, 0.2, N_OPT], dtype=np.float64),
                            checksum=np.array([float(b.im.double().sum()), float(b.s.double().sum())]),
                            local_row0=np.stack(local0).astype(np.float32), global_row0=np.stack(glob0).astype(np.float32),
                            perm=perm.numpy(), pred=np.array(pred, dtype=np.int64))
        print(name, "pred histogram", np.bincount(pred, minlength=N_OPT))


if __name__ == "__main__":
    main()
