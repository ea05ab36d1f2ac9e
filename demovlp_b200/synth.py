"""Deterministic synthetic inputs for the region-word alignment path.

Shapes and seeds follow SURVEY.md section 8(d): a shared per-item component
``kappa * z_b`` makes matched (video, text) pairs dominate like trained
embeddings; every tensor is rounded to bf16 ("bf16 in") and handed out both
as bf16 and as the exactly-equal fp32 upcast so the oracle and the CUDA path
see identical numbers.

Mask conventions mirror the reference call site (trainer/trainer_dist.py:156-159):
additive, 0 for a valid position and -100 for a padded one.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

CONFIGS = {
    # name: (B, R, W, d, seed)  -- BASELINE.json configs 1..5
    "c1": dict(B=32, R=80, W=32, d=256, seed=0),
    "c2": dict(B=128, R=80, W=32, d=256, seed=1),
    "c3": dict(B=1024, R=80, W=32, d=256, seed=2),
    "c4": dict(B=1000, R=80, W=32, d=256, seed=3),
    "c5": dict(B=512, R=288, W=64, d=256, seed=4),
}


@dataclass
class Batch:
    im: torch.Tensor        # (B, R, d) fp32, bf16-representable
    s: torch.Tensor         # (B, W, d) fp32, bf16-representable
    im_mask: torch.Tensor   # (B, R) additive {0, -100}
    s_mask: torch.Tensor    # (B, W) additive {0, -100}
    s_len: torch.Tensor     # (B,) int64 (unused by the scoring, kept for the signature)
    g_text: torch.Tensor    # (B, d) global text CLS embedding
    g_vid: torch.Tensor     # (B, d) global video CLS embedding

    def to(self, device):
        return Batch(*[getattr(self, f).to(device) for f in self.__dataclass_fields__])

    def rows(self, lo, hi):
        return Batch(*[getattr(self, f)[lo:hi] for f in self.__dataclass_fields__])


def make_batch(B, R, W, d=256, seed=0, kappa=0.2, mask_dtype=torch.float32,
               Bt=None) -> Batch:
    """Seeded batch; ``Bt`` (defaults to B) allows non-square video/text counts."""
    Bt = B if Bt is None else Bt
    g = torch.Generator().manual_seed(seed)
    n = max(B, Bt)
    z = torch.randn(n, 1, d, generator=g)
    im = torch.randn(B, R, d, generator=g) + kappa * z[:B]
    s = torch.randn(Bt, W, d, generator=g) + kappa * z[:Bt]
    g_text = torch.randn(Bt, d, generator=g) + kappa * z[:Bt, 0]
    g_vid = torch.randn(B, d, generator=g) + kappa * z[:B, 0]
    keep = torch.rand(B, R, generator=g) < 0.8
    keep[:, 0] = True
    lo = min(4, W)
    s_len = torch.randint(lo, W + 1, (Bt,), generator=g)
    valid = torch.arange(W)[None, :] < s_len[:, None]
    im_mask = ((keep.to(torch.float32) - 1.0) * 100.0).to(mask_dtype)
    s_mask = (valid.to(torch.float32) - 1.0) * 100.0
    r = lambda t: t.to(torch.bfloat16).to(torch.float32)
    return Batch(r(im), r(s), im_mask, s_mask, s_len, r(g_text), r(g_vid))


def make_config(name, kappa=0.2, **over) -> Batch:
    c = dict(CONFIGS[name]); c.update(over)
    return make_batch(c["B"], c["R"], c["W"], c["d"], c["seed"], kappa)
