"""Shared helpers for the test-suite: golden fixtures and seeded inputs."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from demovlp_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix):
    return sorted(os.path.basename(p)[len(prefix):-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(fname):
    return dict(np.load(os.path.join(GOLDEN, fname + ".npz"), allow_pickle=False))


def checksum(b):
    return np.array([float(b.im.double().sum()), float(b.s.double().sum()),
                     float(b.im_mask.double().sum()), float(b.s_mask.double().sum()),
                     float(b.g_text.double().sum()), float(b.g_vid.double().sum())])


def batch_for(g):
    """Regenerate the seeded inputs a golden fixture was made from and verify them."""
    meta = g["meta"]
    B, R, W, d, seed = (int(x) for x in meta[:5])
    kappa = float(meta[5])
    focal = "equal" if int(meta[6]) else "prob"
    if "losses" in g:
        m64, Bt = bool(int(meta[7])), None
    else:
        m64, Bt = False, int(meta[7])
    b = synth.make_batch(B, R, W, d, seed, kappa,
                         mask_dtype=torch.float64 if m64 else torch.float32, Bt=Bt)
    np.testing.assert_allclose(checksum(b), g["checksum"], rtol=0, atol=1e-6,
                               err_msg="seeded generator drifted from the fixture's inputs")
    return b, focal


def rel_err(a, b):
    a = torch.as_tensor(a).double(); b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))
