"""torch.autograd bindings of the C-ABI kernels (device memory and streams are torch's;
the arithmetic is all in libdemovlp_b200.so)."""
from __future__ import annotations

import torch

from . import _lib


def _prep_embed(x):
    """contiguous bf16 or fp32 CUDA tensor (other float dtypes are promoted/demoted to fp32)."""
    if x.dtype not in (torch.bfloat16, torch.float32):
        x = x.float()
    return x.contiguous()


def _prep_mask(m, like_shape):
    if m is None:
        return None
    return m.detach().to(torch.float32).contiguous()


def rw_scores_fwd_raw(im, s, im_mask, s_mask, lam, focal_equal, impl, train=True):
    """Plain (non-autograd) forward: returns scores and the state the backward needs.
    ``train=False`` takes the forward-only entry point (small workspace, nothing saved; state is None)."""
    _lib.require_cuda(im, s, im_mask, s_mask)
    lib = _lib.load()
    Bv, R, d = im.shape
    Bt, W, d2 = s.shape
    if d2 != d or im_mask.shape != (Bv, R) or s_mask.shape != (Bt, W):
        raise RuntimeError(f"shape mismatch: im {tuple(im.shape)} s {tuple(s.shape)} "
                           f"im_mask {tuple(im_mask.shape)} s_mask {tuple(s_mask.shape)}")
    if im.dtype != s.dtype:
        im, s = im.float(), s.float()
    dt = _lib.dtype_code(im)
    impl = _lib.impl_code(impl)
    nbytes = (lib.dvlp_rw_workspace_bytes if train else lib.dvlp_rw_infer_workspace_bytes)(Bv, Bt, R, W, d, dt, impl)
    if nbytes == 0:
        raise RuntimeError(f"demovlp_b200: unsupported shape R={R} W={W} d={d} for impl {impl}")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=im.device)
    scores = torch.empty(Bv, Bt, dtype=torch.float32, device=im.device)
    with torch.cuda.device(im.device):
        fn = lib.dvlp_rw_scores_fwd if train else lib.dvlp_rw_scores_infer
        rc = fn(_lib.ptr(im), _lib.ptr(s), _lib.ptr(im_mask), _lib.ptr(s_mask),
                Bv, Bt, R, W, d, float(lam), int(focal_equal), dt, impl,
                _lib.ptr(scores), _lib.ptr(ws), nbytes, _lib.stream_ptr())
    _lib.check(rc, "rw_scores_fwd" if train else "rw_scores_infer")
    if not train:
        return scores, None
    return scores, (im, s, im_mask, s_mask, ws, float(lam), int(focal_equal), dt, impl)


def rw_scores_bwd_raw(state, d_scores, out_dtype=torch.float32, d_s_out=None):
    """Plain backward: (d_im, d_s) for the state returned by rw_scores_fwd_raw, fp32 unless ``out_dtype`` is
    bf16 and the tcgen05 path serves the shape (its finalisation kernel then writes bf16 directly).
    ``d_s_out``: optional preallocated fp32 (Bt, W, d) tensor to receive d_s (e.g. a peer-visible arena region)."""
    im, s, im_mask, s_mask, ws, lam, focal_equal, dt, impl = state
    lib = _lib.load()
    Bv, R, d = im.shape
    Bt, W, _ = s.shape
    d_scores = d_scores.to(torch.float32).contiguous()
    # the library decides whether its finalisation can write bf16 for this impl / shape (one source of truth)
    direct = (out_dtype == torch.bfloat16 and d_s_out is None
              and bool(lib.dvlp_rw_bwd_out_dtype_supported(impl, R, W, d, float(lam), _lib.BF16)))
    odt = torch.bfloat16 if direct else torch.float32
    d_im = torch.empty(Bv, R, d, dtype=odt, device=im.device)
    if d_s_out is not None:
        if d_s_out.dtype != torch.float32 or tuple(d_s_out.shape) != (Bt, W, d) or not d_s_out.is_contiguous():
            raise RuntimeError("d_s_out must be a contiguous fp32 (Bt, W, d) tensor")
        d_s = d_s_out
    else:
        d_s = torch.empty(Bt, W, d, dtype=odt, device=im.device)
    with torch.cuda.device(im.device):
        rc = lib.dvlp_rw_scores_bwd(_lib.ptr(im), _lib.ptr(s), _lib.ptr(im_mask), _lib.ptr(s_mask),
                                    Bv, Bt, R, W, d, lam, focal_equal, dt, impl,
                                    _lib.ptr(d_scores), _lib.ptr(d_im), _lib.ptr(d_s),
                                    _lib.BF16 if direct else _lib.F32,
                                    _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(rc, "rw_scores_bwd")
    return d_im, d_s


class _RWScores(torch.autograd.Function):
    """scores[i, j] of model/loss.py:294-330 (xattn_score_fast) with a recomputing backward."""

    @staticmethod
    def forward(ctx, im, s, im_mask, s_mask, lam, focal_equal, impl):
        train = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]      # False under no_grad (validation, retrieval)
        scores, state = rw_scores_fwd_raw(im, s, im_mask, s_mask, lam, focal_equal, impl, train=train)
        if train:
            ctx.save_for_backward(*state[:5])
            ctx.cfg = state[5:]
        return scores

    @staticmethod
    def backward(ctx, d_scores):
        state = tuple(ctx.saved_tensors) + tuple(ctx.cfg)
        d_im, d_s = rw_scores_bwd_raw(state, d_scores, out_dtype=state[0].dtype)
        return d_im.to(state[0].dtype), d_s.to(state[1].dtype), None, None, None, None, None


def rw_scores(im, s, im_mask, s_mask=None, lam=20.0, focal_equal=True, impl="auto"):
    """(n_video, n_text) region-word scores; differentiable w.r.t. im and s."""
    im, s = _prep_embed(im), _prep_embed(s)
    if s_mask is None:      # model/loss.py:307-308: a constant mask of ones
        s_mask = torch.ones(s.shape[0], s.shape[1], device=s.device)
    return _RWScores.apply(im, s, _prep_mask(im_mask, None), _prep_mask(s_mask, None), lam, focal_equal, impl)


def rw_scores_grouped(im, s, im_mask, s_mask, n_per_group, lam=20.0, focal_equal=True):
    """Forward-only scores of G groups in one launch: video g against ITS n_per_group texts (group-major ``s``).

    (G, n_per_group) fp32.  This is the multiple-choice evaluation of trainer/trainer_dist.py:792-829 batched over the
    questions (the reference expands one video n_options times and scores a 5 x 5 problem per question).  Shapes outside
    the large-shape tcgen05 kernel are served group by group through the dense forward-only entry point."""
    im, s = _prep_embed(im), _prep_embed(s)
    if s_mask is None:
        s_mask = torch.ones(s.shape[0], s.shape[1], device=s.device)
    im_mask, s_mask = _prep_mask(im_mask, None), _prep_mask(s_mask, None)
    _lib.require_cuda(im, s, im_mask, s_mask)
    lib = _lib.load()
    G, R, d = im.shape
    Bt, W, d2 = s.shape
    if d2 != d or Bt != G * n_per_group or im_mask.shape != (G, R) or s_mask.shape != (Bt, W):
        raise RuntimeError(f"shape mismatch: im {tuple(im.shape)} s {tuple(s.shape)} n_per_group {n_per_group}")
    if im.dtype != s.dtype:
        im, s = im.float(), s.float()
    dt = _lib.dtype_code(im)
    nbytes = lib.dvlp_rw_grouped_workspace_bytes(G, n_per_group, R, W, d, dt)
    if nbytes == 0 or not (0.0 < float(lam) <= 40.0):     # generic kernels, one (1 x n_per_group) problem per group
        # (shapes outside the large-shape kernel, or lambda beyond the tensor-core kernels' fixed-maximum bound)
        rows = [rw_scores_fwd_raw(im[g:g + 1], s[g * n_per_group:(g + 1) * n_per_group], im_mask[g:g + 1],
                                  s_mask[g * n_per_group:(g + 1) * n_per_group], lam, focal_equal, "auto", train=False)[0]
                for g in range(G)]
        return torch.cat(rows, 0)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=im.device)
    scores = torch.empty(G, n_per_group, dtype=torch.float32, device=im.device)
    with torch.cuda.device(im.device):
        rc = lib.dvlp_rw_scores_grouped_infer(_lib.ptr(im), _lib.ptr(s), _lib.ptr(im_mask), _lib.ptr(s_mask), G, n_per_group,
                                              R, W, d, float(lam), int(focal_equal), dt, _lib.ptr(scores), _lib.ptr(ws),
                                              nbytes, _lib.stream_ptr())
    _lib.check(rc, "rw_scores_grouped_infer")
    return scores


def rwa_tail_raw(scores, lam, row_offset, n_total, want_grad=True):
    """(sum_i KL_i / n_total, d_scores or None) for a block of score rows."""
    _lib.require_cuda(scores)
    lib = _lib.load()
    scores = scores.to(torch.float32).contiguous()
    n_rows, n_cols = scores.shape
    loss = torch.empty((), dtype=torch.float32, device=scores.device)
    d_scores = torch.empty_like(scores) if want_grad else None
    with torch.cuda.device(scores.device):
        rc = lib.dvlp_rwa_loss_fwd_bwd(_lib.ptr(scores), n_rows, n_cols, int(row_offset), int(n_total),
                                       float(lam), _lib.ptr(loss), _lib.ptr(d_scores), _lib.stream_ptr())
    _lib.check(rc, "rwa_loss_fwd_bwd")
    return loss, d_scores


class _RWATail(torch.autograd.Function):
    """sum_i KL_i / n_total of model/loss.py:108-114 for a block of rows."""

    @staticmethod
    def forward(ctx, scores, lam, row_offset, n_total):
        loss, d_scores = rwa_tail_raw(scores, lam, row_offset, n_total, ctx.needs_input_grad[0])
        if d_scores is not None:
            ctx.save_for_backward(d_scores)
        ctx.has_grad = d_scores is not None
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.has_grad:
            raise RuntimeError("rwa_loss was evaluated under no_grad")
        (d_scores,) = ctx.saved_tensors
        return d_scores * g, None, None, None


def rwa_loss(scores, lam=20.0, row_offset=0, n_total=None):
    n_total = scores.shape[1] if n_total is None else n_total
    return _RWATail.apply(scores, lam, row_offset, n_total)


class _NormSoftmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, temperature):
        _lib.require_cuda(x)
        lib = _lib.load()
        x = x.to(torch.float32).contiguous()
        N, M = x.shape
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        scratch = torch.empty(N + M, dtype=torch.float32, device=x.device)
        need = ctx.needs_input_grad[0]
        d_x = torch.empty_like(x) if need else None
        with torch.cuda.device(x.device):
            rc = lib.dvlp_normsoftmax_fwd_bwd(_lib.ptr(x), N, M, float(temperature), _lib.ptr(loss),
                                              _lib.ptr(d_x), _lib.ptr(scratch), _lib.stream_ptr())
        _lib.check(rc, "normsoftmax_fwd_bwd")
        if d_x is not None:
            ctx.save_for_backward(d_x)
        ctx.has_grad = d_x is not None
        return loss

    @staticmethod
    def backward(ctx, g):
        if not ctx.has_grad:
            raise RuntimeError("norm_softmax_loss was evaluated under no_grad")
        (d_x,) = ctx.saved_tensors
        return d_x * g, None


def norm_softmax_loss(x, temperature=0.05):
    return _NormSoftmax.apply(x, temperature)


class _SimMatrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, eps):
        _lib.require_cuda(a, b)
        lib = _lib.load()
        if a.dtype != b.dtype:
            a, b = a.float(), b.float()
        N, d = a.shape
        M, d2 = b.shape
        if d != d2:
            raise RuntimeError(f"sim_matrix: width mismatch {d} vs {d2}")
        out = torch.empty(N, M, dtype=torch.float32, device=a.device)
        scratch = torch.empty(N + M, dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            rc = lib.dvlp_sim_matrix_fwd(_lib.ptr(a), _lib.ptr(b), N, M, d, float(eps), _lib.dtype_code(a),
                                         _lib.ptr(out), _lib.ptr(scratch), _lib.stream_ptr())
        _lib.check(rc, "sim_matrix_fwd")
        ctx.save_for_backward(a, b)
        ctx.eps = float(eps)
        return out

    @staticmethod
    def backward(ctx, d_out):
        a, b = ctx.saved_tensors
        lib = _lib.load()
        N, d = a.shape
        M = b.shape[0]
        d_out = d_out.to(torch.float32).contiguous()
        # bf16 embeddings of width 256 take the mma.sync kernels, which write the gradients in bf16 directly
        odt = a.dtype if (a.dtype == torch.bfloat16 and d == 256) else torch.float32
        d_a = torch.empty(N, d, dtype=odt, device=a.device) if ctx.needs_input_grad[0] else None
        d_b = torch.empty(M, d, dtype=odt, device=a.device) if ctx.needs_input_grad[1] else None
        scratch = torch.empty(N + M, dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            rc = lib.dvlp_sim_matrix_bwd(_lib.ptr(a), _lib.ptr(b), N, M, d, ctx.eps, _lib.dtype_code(a),
                                         _lib.ptr(d_out), _lib.ptr(d_a), _lib.ptr(d_b),
                                         _lib.BF16 if odt == torch.bfloat16 else _lib.F32, _lib.ptr(scratch),
                                         _lib.stream_ptr())
        _lib.check(rc, "sim_matrix_bwd")
        return (d_a.to(a.dtype) if d_a is not None else None,
                d_b.to(b.dtype) if d_b is not None else None, None)


def sim_matrix(a, b, eps=1e-8):
    """Cosine matrix, rows = a, columns = b (model/model.py:582-590)."""
    return _SimMatrix.apply(_prep_embed(a), _prep_embed(b), eps)
