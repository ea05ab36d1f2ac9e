"""Host-side mirror of the reference's retrieval metrics (model/metric.py).

``t2v_metrics`` / ``v2t_metrics`` keep the reference signatures and return the same
dict; the rank search (model/metric.py:22-66, 142-189: a full sort per row plus an
exact-equality lookup) runs as counting kernels on the GPU -- integer/comparison work,
bit-exact with the reference by construction.  ``cols2metrics`` (:298-308) is a
handful of numpy reductions over N ranks and stays on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _upload(sims):
    if not torch.cuda.is_available():
        raise RuntimeError("demovlp_b200.metric needs a CUDA device; there is no CPU fallback")
    a = np.ascontiguousarray(np.asarray(sims, dtype=np.float64))
    assert a.ndim == 2, "expected a matrix"
    return torch.from_numpy(a).cuda()


def t2v_ranks(sims) -> np.ndarray:
    """Rank (0 = best) of the ground-truth video per text query; int64 numpy."""
    x = sims if isinstance(sims, torch.Tensor) and sims.is_cuda else _upload(sims)
    x = x.to(torch.float64).contiguous()
    nq, nv = x.shape
    out = torch.empty(nq, dtype=torch.int32, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().dvlp_ranks_t2v(_lib.ptr(x), nq, nv, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "ranks_t2v")
    return out.cpu().numpy().astype(np.int64)


def v2t_ranks(sims) -> np.ndarray:
    """Tie-averaged rank of the closest ground-truth caption per video; float64 numpy."""
    x = sims if isinstance(sims, torch.Tensor) and sims.is_cuda else _upload(sims)
    x = x.to(torch.float64).contiguous()
    nq, nv = x.shape
    out = torch.empty(nv, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().dvlp_ranks_v2t(_lib.ptr(x), nq, nv, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "ranks_v2t")
    return out.cpu().numpy()


def cols2metrics(cols, num_queries):
    """model/metric.py:298-308."""
    cols = np.asarray(cols)
    metrics = {
        "R1": 100 * float(np.sum(cols == 0)) / num_queries,
        "R5": 100 * float(np.sum(cols < 5)) / num_queries,
        "R10": 100 * float(np.sum(cols < 10)) / num_queries,
        "R50": 100 * float(np.sum(cols < 50)) / num_queries,
        "MedR": np.median(cols) + 1,
        "MeanR": np.mean(cols) + 1,
    }
    stats = np.array([metrics[x] for x in ("R1", "R5", "R10")], dtype=np.float64)
    with np.errstate(divide="ignore"):
        metrics["geometric_mean_R1-R5-R10"] = float(np.exp(np.mean(np.log(stats))))
    return metrics


def t2v_metrics(sims, query_masks=None):
    """model/metric.py:10-122.  sims: (num_queries, num_vids), x[i, j] = <text_i, video_j>."""
    sims = np.asarray(sims) if not isinstance(sims, torch.Tensor) else sims
    num_queries = sims.shape[0]
    cols = t2v_ranks(sims)
    if query_masks is not None:           # :104-112 (the reference uses the removed np.bool here)
        keep = np.asarray(query_masks).reshape(-1).astype(bool)
        assert keep.size == num_queries, "invalid query mask shape"
        cols = cols[keep]
        num_queries = int(keep.sum())
    return cols2metrics(cols, num_queries)


def v2t_metrics(sims, query_masks=None):
    """model/metric.py:125-214.  Missing captions (query_masks == 0) get similarity -1e8 (:157-162)."""
    sims = np.asarray(sims, dtype=np.float64) if not isinstance(sims, torch.Tensor) else sims
    if query_masks is not None:
        sims = np.array(sims.cpu().numpy() if isinstance(sims, torch.Tensor) else sims, dtype=np.float64)
        sims[np.logical_not(np.asarray(query_masks).reshape(-1).astype(bool)), :] = -1e8
    ranks = v2t_ranks(sims)
    return cols2metrics(ranks, sims.shape[1])
