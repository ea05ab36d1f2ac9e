"""Mint golden fixtures by running the UNMODIFIED reference (build container only).

    python oracle/make_golden.py          # writes tests/golden/*.npz

The reference has no golden vectors of its own, so the
fixtures are outputs of the reference code itself -- model/loss.py
(GlobalLocalLoss fwd + autograd bwd, get_sim, get_sim_by_segment),
model/model.py:sim_matrix and model/metric.py (t2v_metrics / v2t_metrics) --
on the seeded inputs of demovlp_b200/synth.py.  Inputs are regenerated from
the seed at test time; an input checksum stored beside the outputs catches
any drift of the generator.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from demovlp_b200 import synth          # noqa: E402
from oracle import ref_loader           # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

# (name, B, R, W, Bt, seed, kappa, focal, mask64, store_grads)
LOSS_CASES = [
    ("c1_equal",      32, 80, 32, None, 0, 0.2, "equal", False, True),    # config 1, mid-training regime
    ("c1_equal_k05",  32, 80, 32, None, 0, 0.5, "equal", False, False),   # converged regime (local loss ~ 0)
    ("c1_equal_k0",   32, 80, 32, None, 0, 0.0, "equal", False, False),   # untrained regime
    ("c1_prob",       16, 80, 32, None, 0, 0.2, "prob",  False, True),
    ("c1_mask64",      8, 80, 32, None, 0, 0.2, "equal", True,  True),    # float64 video mask (MSRVTT_dataset.py:522)
    ("ft_ragged",      4, 240, 99, None, 5, 0.2, "equal", False, True),   # real fine-tune shape, R=8x30, W=99
    ("pt_r30",         8, 30, 99, None, 6, 0.2, "equal", False, True),    # real pre-train shape, R=30, W=99
    ("didemo_small",   4, 288, 64, None, 4, 0.2, "equal", False, True),   # config 5 shape, few items
]
SCORE_CASES = [
    # non-square / tiny (MC eval uses 5x5, trainer_dist.py:813-826)
    ("mc_5x5",   5, 80, 32, 5, 7, 0.2, "equal"),
    ("rect_7x12", 7, 80, 32, 12, 8, 0.2, "equal"),
    ("odd_3x4_r17_w9", 3, 17, 9, 4, 9, 0.2, "equal"),
    ("c2_corner_16x24", 16, 80, 32, 24, 1, 0.2, "equal"),
]


def checksum(b):
    return np.array([float(b.im.double().sum()), float(b.s.double().sum()),
                     float(b.im_mask.double().sum()), float(b.s_mask.double().sum()),
                     float(b.g_text.double().sum()), float(b.g_vid.double().sum())])


def main():
    assert ref_loader.available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    L = ref_loader.loss_module()
    M = ref_loader.metric_module()
    sim_matrix = ref_loader.sim_matrix_fn()
    torch.set_num_threads(os.cpu_count())

    for name, B, R, W, Bt, seed, kappa, focal, m64, keep_grads in LOSS_CASES:
        b = synth.make_batch(B, R, W, 256, seed, kappa,
                             mask_dtype=torch.float64 if m64 else torch.float32)
        im = b.im.clone().requires_grad_(True)
        s = b.s.clone().requires_grad_(True)
        gt = b.g_text.clone().requires_grad_(True)
        gv = b.g_vid.clone().requires_grad_(True)
        mod = L.GlobalLocalLoss(focal_type=focal)
        gsim = sim_matrix(gt, gv)
        gsim.retain_grad()
        loss, gl, ll = mod(gsim, im, s, b.im_mask, b.s_len, b.s_mask)
        loss.backward()
        with torch.no_grad():
            scores = mod.local_loss.get_sim(b.im, b.s, b.im_mask, b.s_len, b.s_mask)
        grads = dict(d_im=im.grad.numpy(), d_s=s.grad.numpy(),
                     d_g_text=gt.grad.numpy(), d_g_vid=gv.grad.numpy()) if keep_grads else {}
        np.savez_compressed(
            os.path.join(OUT, f"loss_{name}.npz"),
            meta=np.array([B, R, W, 256, seed, kappa, int(focal == "equal"), int(m64)], dtype=np.float64),
            checksum=checksum(b), scores=scores.numpy().astype(np.float32),
            losses=np.array([loss.item(), gl.item(), ll.item()], dtype=np.float64),
            global_sim=gsim.detach().numpy(), d_global_sim=gsim.grad.numpy(),
            grad_norms=np.array([float(im.grad.double().norm()), float(s.grad.double().norm())]),
            **grads)
        print(name, "loss", loss.item(), gl.item(), ll.item())

    for name, Bv, R, W, Bt, seed, kappa, focal in SCORE_CASES:
        b = synth.make_batch(Bv, R, W, 256, seed, kappa, Bt=Bt)
        mod = L.RWALoss(focal_type=focal)
        with torch.no_grad():
            scores = mod.get_sim(b.im, b.s, b.im_mask, b.s_len, b.s_mask)
        np.savez_compressed(os.path.join(OUT, f"scores_{name}.npz"),
                            meta=np.array([Bv, R, W, 256, seed, kappa, int(focal == "equal"), Bt], dtype=np.float64),
                            checksum=checksum(b), scores=scores.numpy().astype(np.float32))
        print(name, tuple(scores.shape))

    # eval path (config 4, cut down to 96 x 96 so the reference finishes in seconds here):
    # get_sim_by_segment -> + sim_matrix(text, video) -> t2v / v2t metrics
    b = synth.make_batch(96, 80, 32, 256, 3, 0.12)
    mod = L.RWALoss(focal_type="equal")
    with torch.no_grad():
        local = mod.get_sim_by_segment(b.im, b.s, b.im_mask, b.s_len, b.s_mask, device="cpu")
        full = sim_matrix(b.g_text, b.g_vid).numpy() + local
    t2v = M.t2v_metrics(full)
    v2t = M.v2t_metrics(full)
    keys = sorted(t2v)
    # ranks as the reference computes them internally (metric.py:25-38): first equal column
    d = -full
    sd = np.sort(d, axis=1)
    gtd = np.diag(d)[:, None]
    rows, cols = np.where((sd - gtd) == 0)
    _, first = np.unique(rows, return_index=True)
    ranks = cols[first]
    np.savez_compressed(os.path.join(OUT, "eval_96.npz"), checksum=checksum(b),
                        local=local, full=full, t2v_ranks=ranks.astype(np.int64),
                        keys=np.array(keys),
                        t2v=np.array([float(t2v[k]) for k in keys]),
                        v2t=np.array([float(v2t[k]) for k in keys]))
    print("eval_96", t2v)

    # metric-only fixtures on crafted matrices (regenerated from METRIC_MATS at test time):
    # ties, multi-caption (several captions per video, MSRVTT full-split layout)
    pack = {}
    for k, m in metric_mats().items():
        t = M.t2v_metrics(m)
        v = M.v2t_metrics(m)
        pack[k + "_sum"] = np.array([m.sum()])
        pack[k + "_t2v"] = np.array([float(t[x]) for x in keys])
        pack[k + "_v2t"] = np.array([float(v[x]) for x in keys])
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), keys=np.array(keys), **pack)
    print("metrics fixtures written")


def metric_mats():
    rng = np.random.default_rng(11)
    return {
        "rand_1000": rng.standard_normal((1000, 1000)),
        "ties_60": np.round(rng.standard_normal((60, 60)), 1),
        "multi_200x10": rng.standard_normal((200, 10)),
        "multi_ties_120x6": np.round(rng.standard_normal((120, 6)), 1),
        "const_16": np.zeros((16, 16)),
    }


if __name__ == "__main__":
    main()
