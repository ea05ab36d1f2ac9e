"""Comparison of sharded config-3 results with the reference-minted fixture tests/golden/c3_global.npz
(oracle/make_golden_c3.py: the UNMODIFIED reference driven in 16-row blocks).

Shared by the 1-GPU shard-walk test (tests/test_gpu_c3.py) and by bench.py's post-run ``parity_check`` (every rank
compares the rows it owns).  Pure torch/numpy on CPU tensors: no kernels are called here.
"""
from __future__ import annotations

import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c3_global.npz")
SCORE_TOL, GRAD_TOL = 2e-3, 1e-2          # BASELINE.json north star: 2e-3 absolute, 1e-2 relative


def available():
    return os.path.isfile(GOLDEN)


def load():
    return dict(np.load(GOLDEN, allow_pickle=False))


def probes(B, R, W, d):
    """The seeded probe tensors of oracle/make_golden_c3.py (same generator, same order)."""
    g = torch.Generator().manual_seed(1234)
    return (torch.randn(B, R, d, generator=g), torch.randn(B, W, d, generator=g),
            torch.randn(B, d, generator=g), torch.randn(B, d, generator=g))


def _rel(a, b):
    a = torch.as_tensor(a).double(); b = torch.as_tensor(b).double()
    return float((a - b).norm() / b.norm().clamp(min=1e-30))


def check_rows(g, lo, hi, losses=None, scores=None, d_im=None, d_s=None, d_g_text=None, d_g_vid=None,
               grad_tol=GRAD_TOL, cls_tol=1e-3):
    """Compare what a rank that owns global rows [lo, hi) holds with the fixture.

    ``scores`` (hi-lo, B), ``d_im`` (hi-lo, R, d), ``d_s`` (hi-lo, W, d), CLS gradients (hi-lo, d): CPU tensors (any
    float dtype) or None to skip.  Returns (ok, report): per 128-row fixture shard the Frobenius norm and the
    seeded-probe dot product of each gradient, the fully stored rows, score rows / row sums / diagonal, and the loss
    triple.  An error anywhere inside a shard moves its probe: tolerance = 4 sigma of the probe of a gradient whose
    relative error is ``grad_tol``."""
    B, R, W, d = (int(x) for x in g["meta"][:4])
    shards = int(g["meta"][7])
    bsh = B // shards
    rep, ok = {}, True

    def flag(name, err, tol):
        nonlocal ok
        good = bool(err < tol)
        rep[name] = {"err": float(err), "tol": float(tol), "ok": good}
        ok = ok and good

    if losses is not None:
        want = g["losses"]
        flag("loss_triple_abs", float(np.abs(np.asarray(losses, dtype=np.float64) - want).max()), SCORE_TOL)
    keep = [int(k) for k in g["keep_rows"]]
    if scores is not None:
        sc = torch.as_tensor(scores).double()
        errs = [float((sc[k - lo] - torch.from_numpy(g["score_rows"][i]).double()).abs().max())
                for i, k in enumerate(keep) if lo <= k < hi]
        if errs:
            flag("score_rows_abs", max(errs), SCORE_TOL)
        sums = torch.from_numpy(g["score_row_sums"][lo:hi])
        flag("score_row_sums_abs_per_entry", float((sc.sum(1) - sums).abs().max()) / B, SCORE_TOL)
        diag = torch.from_numpy(g["score_diag"][lo:hi]).double()
        flag("score_diag_abs", float((sc[torch.arange(hi - lo), torch.arange(lo, hi)] - diag).abs().max()), SCORE_TOL)
    pr = probes(B, R, W, d)
    for name, grad, probe, tol in (("d_im", d_im, pr[0], grad_tol), ("d_s", d_s, pr[1], grad_tol),
                                   ("d_g_text", d_g_text, pr[2], cls_tol), ("d_g_vid", d_g_vid, pr[3], cls_tol)):
        if grad is None:
            continue
        gr = torch.as_tensor(grad).double()
        rows = [(i, k) for i, k in enumerate(keep) if lo <= k < hi]
        if rows:
            flag(name + "_rows_rel", max(_rel(gr[k - lo], g[name + "_rows"][i]) for i, k in rows), tol)
        worst_n, worst_p = 0.0, 0.0
        for sh in range(lo // bsh, (hi + bsh - 1) // bsh):
            a, b_ = max(lo, sh * bsh), min(hi, (sh + 1) * bsh)
            if (a, b_) != (sh * bsh, (sh + 1) * bsh):
                continue                                   # only whole fixture shards carry norms / probes
            blk = gr[a - lo:b_ - lo]
            wn = float(g[name + "_norms"][sh])
            worst_n = max(worst_n, abs(float(blk.norm()) - wn) / max(wn, 1e-30))
            pb = probe[a:b_].double()
            sigma = tol * wn * float(pb.norm()) / (blk.numel() ** 0.5)
            worst_p = max(worst_p, abs(float((blk * pb).sum()) - float(g[name + "_probe"][sh])) / max(4 * sigma, 1e-30))
        flag(name + "_shard_norm_rel", worst_n, tol)
        flag(name + "_shard_probe_over_4sigma", worst_p, 1.0)
    return ok, rep
