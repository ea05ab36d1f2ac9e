"""Projection heads that feed the region-word path, as drop-in modules.

The reference projects the 768-wide tower outputs to the 256-wide joint space with
``txt_proj = nn.Sequential(nn.ReLU(), nn.Linear(768, 256))`` (model/model.py:40-43, applied to every token at :88) and
``ObjectTransformer.proj = nn.Linear(768, 256, bias=False)`` (model/object_transformer.py:381, :451); the first token is
the global (CLS) embedding, the rest are the words / regions (model/model.py:88-96).  ``FusedProjection`` keeps the
``nn.Linear`` parameters (same names: ``weight``, ``bias``, so reference checkpoints load) and runs the forward as one
tcgen05 GEMM of libdemovlp_b200.so that writes the bf16 rows the scoring kernels consume; the backward uses library
GEMMs (torch.matmul).  No CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib


class _Proj(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu_in):
        _lib.require_cuda(x, weight, bias)
        lib = _lib.load()
        K = x.shape[-1]
        N = weight.shape[0]
        x2 = x.reshape(-1, K)
        if x2.dtype not in (torch.float32, torch.bfloat16):
            x2 = x2.float()
        x2 = x2.contiguous()
        M = x2.shape[0]
        wb = weight.detach().to(torch.bfloat16).contiguous()
        bf = bias.detach().float().contiguous() if bias is not None else None
        dt = _lib.dtype_code(x2)
        nbytes = lib.dvlp_proj_workspace_bytes(M, K, dt, int(relu_in))
        if nbytes == 0 or N != 256:
            raise RuntimeError(f"demovlp_b200: unsupported projection shape (tokens {M}, in {K}, out {N})")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        out = torch.empty(M, N, dtype=torch.bfloat16, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.dvlp_proj_fwd(_lib.ptr(x2), dt, M, K, N, _lib.ptr(wb), _lib.ptr(bf), int(relu_in), _lib.ptr(out),
                                   _lib.ptr(ws), nbytes, _lib.stream_ptr())
        _lib.check(rc, "proj_fwd")
        ctx.save_for_backward(x2, weight, bias if bias is not None else torch.empty(0, device=x.device))
        ctx.relu_in, ctx.has_bias, ctx.in_shape = bool(relu_in), bias is not None, x.shape
        return out.reshape(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, d_out):
        x2, weight, bias = ctx.saved_tensors
        g = d_out.reshape(-1, weight.shape[0]).float()
        a = x2.float()
        if ctx.relu_in:
            a = torch.relu(a)
        d_x = d_w = d_b = None
        if ctx.needs_input_grad[0]:
            d_x = g @ weight.float()
            if ctx.relu_in:
                d_x = d_x * (x2 > 0)
            d_x = d_x.reshape(ctx.in_shape).to(x2.dtype)
        if ctx.needs_input_grad[1]:
            d_w = (g.t() @ a).to(weight.dtype)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            d_b = g.sum(0).to(bias.dtype)
        return d_x, d_w, d_b, None


class FusedProjection(nn.Module):
    """``Linear(in_dim, 256)`` with an optional ReLU in front, returning (global, local) like model/model.py:86-96.

    ``FusedProjection(768, relu_in=True, bias=True)`` replaces ``txt_proj``; ``FusedProjection(768, relu_in=False,
    bias=False)`` replaces ``ObjectTransformer.proj``.  ``forward(tokens)`` takes (B, L + 1, in_dim) and returns
    ``(tokens[:, 0] projected, tokens[:, 1:] projected)`` in bf16 -- the dtype the scoring kernels consume."""

    def __init__(self, in_dim=768, out_dim=256, relu_in=False, bias=True):
        super().__init__()
        lin = nn.Linear(in_dim, out_dim, bias=bias)         # same initialisation and parameter names as the reference
        self.weight, self.bias = lin.weight, lin.bias
        self.relu_in = relu_in

    def project(self, tokens):
        return _Proj.apply(tokens, self.weight, self.bias, self.relu_in)

    def forward(self, tokens):
        y = self.project(tokens)
        return y[:, 0, ...], y[:, 1:, ...]
