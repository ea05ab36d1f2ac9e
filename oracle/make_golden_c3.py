"""Mint the BASELINE config-3 fixture (global B = 1024) from the UNMODIFIED reference (build container only).

    python oracle/make_golden_c3.py [--rows-per-block 16]      # writes tests/golden/c3_global.npz

TEST INFRASTRUCTURE.  One shot through the reference needs an 86 GB (B, B, R, d) tensor, so the reference is
driven in video-row blocks exactly as SURVEY.md 8(d) c3 prescribes: RWALoss is a row-wise softmax, hence
``local = (1/B) sum_i KL_i`` splits over row blocks.  For each block the reference's own ``get_sim``
(model/loss.py:64-71) gives the score rows, model/loss.py:108-114 is applied with the matching rows of ``eye``, and
autograd accumulates ``d_s`` over the blocks -- the per-rank decomposition of the sharded loss
(trainer/trainer_dist.py:13-31 semantics: every rank holds the loss of the concatenated batch and the gradient rows
of its own shard).  NormSoftmaxLoss + sim_matrix run on the whole 1024 x 1024 CLS matrix in one piece.

The fixture keeps what a test can compare without storing 100 MB of gradients: the loss triple, a few full score /
gradient rows per rank shard (8 shards of 128), per-shard Frobenius norms and per-shard dot products with a seeded
probe tensor (an error anywhere in a shard moves its probe).
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demovlp_b200 import synth          # noqa: E402
from oracle import ref_loader           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "c3_global.npz")
SHARDS, KEEP_ROW = 8, 5                  # rows kept in full: index shard*b + KEEP_ROW of every shard


def probes(B, R, W, d):
    g = torch.Generator().manual_seed(1234)
    return (torch.randn(B, R, d, generator=g), torch.randn(B, W, d, generator=g),
            torch.randn(B, d, generator=g), torch.randn(B, d, generator=g))


def shard_stats(grad, probe, shards=SHARDS):
    b = grad.shape[0] // shards
    norms = [float(grad[r * b:(r + 1) * b].double().norm()) for r in range(shards)]
    dots = [float((grad[r * b:(r + 1) * b].double() * probe[r * b:(r + 1) * b].double()).sum()) for r in range(shards)]
    return np.array(norms), np.array(dots)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-block", type=int, default=16)
    ap.add_argument("--B", type=int, default=1024)
    ap.add_argument("--out", default=OUT)
    args = ap.parse_args()
    assert ref_loader.available(), "needs /root/reference"
    torch.set_num_threads(os.cpu_count())
    L = ref_loader.loss_module()
    sim_matrix = ref_loader.sim_matrix_fn()
    cfg = dict(synth.CONFIGS["c3"]); cfg["B"] = args.B
    B, R, W, d, seed = cfg["B"], cfg["R"], cfg["W"], cfg["d"], cfg["seed"]
    b = synth.make_batch(B, R, W, d, seed, 0.2)
    lam = 20.0
    rwa = L.RWALoss(focal_type="equal")
    im = b.im.clone().requires_grad_(True)
    s = b.s.clone().requires_grad_(True)
    scores = torch.empty(B, B)
    local = 0.0
    t0 = time.time()
    blk = args.rows_per_block
    for r0 in range(0, B, blk):
        r1 = min(B, r0 + blk)
        sc = rwa.get_sim(im[r0:r1], s, b.im_mask[r0:r1], b.s_len, b.s_mask)        # reference scoring, (blk, B)
        # model/loss.py:108-114 on these rows (eye rows r0..r1)
        eye = torch.zeros(r1 - r0, B)
        eye[torch.arange(r1 - r0), torch.arange(r0, r1)] = 1.0
        p = torch.softmax(sc * lam, dim=1)
        kl = (p * (torch.log(p) - torch.log(eye + 1e-6))).sum(1)
        part = kl.sum() / B
        part.backward()
        local += float(part)
        scores[r0:r1] = sc.detach()
        print(f"rows {r0}-{r1} local so far {local:.6f}  ({time.time() - t0:.0f}s)", flush=True)
    gt = b.g_text.clone().requires_grad_(True)
    gv = b.g_vid.clone().requires_grad_(True)
    gsim = sim_matrix(gt, gv)
    glob = L.NormSoftmaxLoss()(gsim)
    glob.backward()
    glob = float(glob)
    p_im, p_s, p_gt, p_gv = probes(B, R, W, d)
    bsh = B // SHARDS
    keep = [r * bsh + KEEP_ROW for r in range(SHARDS)]
    n_im, dot_im = shard_stats(im.grad, p_im)
    n_s, dot_s = shard_stats(s.grad, p_s)
    n_gt, dot_gt = shard_stats(gt.grad, p_gt)
    n_gv, dot_gv = shard_stats(gv.grad, p_gv)
    np.savez_compressed(
        args.out,
        meta=np.array([B, R, W, d, seed, 0.2, 1, SHARDS, KEEP_ROW, blk], dtype=np.float64),
        checksum=np.array([float(b.im.double().sum()), float(b.s.double().sum()), float(b.im_mask.double().sum()),
                           float(b.s_mask.double().sum()), float(b.g_text.double().sum()), float(b.g_vid.double().sum())]),
        losses=np.array([glob + local, glob, local], dtype=np.float64),
        keep_rows=np.array(keep), score_rows=scores[keep].numpy().astype(np.float32),
        score_row_sums=scores.double().sum(1).numpy(), score_diag=scores.diag().numpy().astype(np.float32),
        d_im_rows=im.grad[keep].numpy(), d_s_rows=s.grad[keep].numpy(),
        d_g_text_rows=gt.grad[keep].numpy(), d_g_vid_rows=gv.grad[keep].numpy(),
        d_im_norms=n_im, d_im_probe=dot_im, d_s_norms=n_s, d_s_probe=dot_s,
        d_g_text_norms=n_gt, d_g_text_probe=dot_gt, d_g_vid_norms=n_gv, d_g_vid_probe=dot_gv)
    print("wrote", args.out, "losses", glob + local, glob, local, f"{time.time() - t0:.0f}s")


if __name__ == "__main__":
    main()
