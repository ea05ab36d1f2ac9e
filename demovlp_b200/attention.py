"""Masked attention of the ObjectTransformer's short token sequences, as a drop-in for the reference's
``attn_mask(q, k, v, mask)`` (model/object_transformer.py:93-100), which ``VarAttention.forward`` (:125-196) calls on
``(batch * heads)`` sequences: space attention (1 + R region tokens per frame), time attention (1 + F frame tokens per
region) and the CLS token's single query against all 1 + F R keys.  One fused CUDA kernel per call -- tensor cores
(mma.sync, fp32 operands split into bf16 hi + lo parts: fp32-class accuracy) for up to 64 x 64 tokens, a warp-per-key-stripe
kernel for the single-query case -- instead of two batched GEMMs, an add and a softmax through HBM; recomputing,
atomic-free backward.  fp32, head dimension 64; there is no CPU fallback.
"""
from __future__ import annotations

import torch

from . import _lib


class _AttnMask(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, mask):
        _lib.require_cuda(q, k, v, mask)
        lib = _lib.load()
        B, I, d = q.shape
        J = k.shape[1]
        if k.shape != (B, J, d) or v.shape != (B, J, d):
            raise RuntimeError(f"attn_mask: shape mismatch q {tuple(q.shape)} k {tuple(k.shape)} v {tuple(v.shape)}")
        if not lib.dvlp_attn_small_supported(I, J, d):
            raise RuntimeError("demovlp_b200.attn_mask serves head dimension 64 with up to 64 x 64 tokens, or one query "
                               f"against up to 1024 keys; got d={d}, queries={I}, keys={J}")
        qf, kf, vf = (t.detach().float().contiguous() for t in (q, k, v))
        mf = None
        if mask is not None:
            mf = mask.detach().float().expand(B, 1, J).reshape(B, J).contiguous() if mask.dim() == 3 else \
                mask.detach().float().expand(B, J).contiguous()
        out = torch.empty(B, I, d, dtype=torch.float32, device=q.device)
        lse = torch.empty(B, I, dtype=torch.float32, device=q.device)
        with torch.cuda.device(q.device):
            rc = lib.dvlp_attn_small_fwd(_lib.ptr(qf), _lib.ptr(kf), _lib.ptr(vf), _lib.ptr(mf), B, I, J, d,
                                         _lib.ptr(out), _lib.ptr(lse), _lib.stream_ptr())
        _lib.check(rc, "attn_small_fwd")
        ctx.save_for_backward(qf, kf, vf, mf if mf is not None else torch.empty(0, device=q.device), out, lse)
        ctx.has_mask = mf is not None
        ctx.dtypes = (q.dtype, k.dtype, v.dtype)
        return out.to(q.dtype)

    @staticmethod
    def backward(ctx, d_out):
        qf, kf, vf, mf, out, lse = ctx.saved_tensors
        lib = _lib.load()
        B, I, d = qf.shape
        J = kf.shape[1]
        g = d_out.float().contiguous()
        d_q, d_k, d_v = torch.empty_like(qf), torch.empty_like(kf), torch.empty_like(vf)
        with torch.cuda.device(qf.device):
            rc = lib.dvlp_attn_small_bwd(_lib.ptr(qf), _lib.ptr(kf), _lib.ptr(vf), _lib.ptr(mf if ctx.has_mask else None),
                                         _lib.ptr(out), _lib.ptr(lse), _lib.ptr(g), B, I, J, d, _lib.ptr(d_q), _lib.ptr(d_k),
                                         _lib.ptr(d_v), _lib.stream_ptr())
        _lib.check(rc, "attn_small_bwd")
        return d_q.to(ctx.dtypes[0]), d_k.to(ctx.dtypes[1]), d_v.to(ctx.dtypes[2]), None


def attn_mask(q, k, v, mask=None):
    """softmax(q k^T + mask, dim=-1) v for (b, i, d) queries and (b, j, d) keys / values; ``mask`` is the reference's
    additive (b, 1, j) tensor (or (b, j), or None).  Same call as model/object_transformer.py:93-100."""
    return _AttnMask.apply(q, k, v, mask)


def attn(q, k, v):
    """model/object_transformer.py:86-90 (no mask)."""
    return _AttnMask.apply(q, k, v, None)
