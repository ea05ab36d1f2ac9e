"""Torch (CPU) emulation of the arithmetic the CUDA kernels perform.

Not the oracle and not shipped: it restates the *kernel's* algorithm -- the
Gram-matrix form of SURVEY.md section 7.5 with optional bf16 rounding at the
points where the tcgen05 path rounds -- so that the algebra (forward identity
and the hand-derived backward of DESIGN.md) can be checked against the
oracle's autograd on CPU before any GPU time is spent.
"""
import torch

EPS = 1e-8


def _r(x, on, tag=None):
    """Round to bf16 when ``on`` is True or is a collection containing ``tag``."""
    if isinstance(on, (set, frozenset, list, tuple)):
        on = tag in on
    return x.to(torch.bfloat16).to(x.dtype) if on else x


def forward_backward(im, s, im_mask, s_mask, d_scores=None, lam=20.0, focal=True, bf16=False):
    """Returns scores (Bv,Bt) and, when d_scores is given, (d_im, d_s)."""
    Bv, R, d = im.shape
    Bt, W, _ = s.shape
    na = im.norm(dim=-1, keepdim=True); nb = s.norm(dim=-1, keepdim=True)
    alpha = 1.0 / (na + EPS); beta = 1.0 / (nb + EPS)
    Vh = _r(im * alpha, bf16, "ops"); Th = _r(s * beta, bf16, "ops")          # normalised operands (bf16 on the fast path)
    GV = _r(torch.einsum("ird,isd->irs", Vh, Vh), bf16, "gram")         # (Bv,R,R)
    GT = _r(torch.einsum("jwd,jxd->jwx", Th, Th), bf16, "gram")         # (Bt,W,W)
    S = torch.einsum("ird,jwd->ijrw", Vh, Th)                   # (Bv,Bt,R,W) fp32 accumulate
    neg = S <= 0
    l = torch.where(neg, 0.1 * S, S)
    mv = im_mask.to(S.dtype)[:, None, :, None]
    mt = s_mask.to(S.dtype)[None, :, None, :]

    def direction(ctx_dim, n_ctx, mask, G, eq):
        # ctx_dim: axis of S that is the attention context (2 = regions, 3 = words)
        q_dim = 5 - ctx_dim
        nrm = l.square().sum(q_dim, keepdim=True).sqrt() + EPS   # l2norm over the QUERY axis
        A = l / nrm
        e = torch.exp(lam * (A + mask) - lam)
        E = e.sum(ctx_dim, keepdim=True)
        H = (e * n_ctx > E) if focal else torch.ones_like(e, dtype=torch.bool)
        u = torch.where(H, e, torch.zeros_like(e))
        Q = u / u.sum(ctx_dim, keepdim=True)
        q = _r(Q, bf16, "q")
        num = (q * S).sum(ctx_dim, keepdim=True)
        Y = torch.einsum(eq, G, q)
        D = (q * Y).sum(ctx_dim, keepdim=True)
        den = D.sqrt()
        cos = torch.where(D > 1e-30, num / den, torch.zeros_like(num))
        return dict(nrm=nrm, A=A, q=q, Y=Y, D=D, den=den, cos=cos, q_dim=q_dim, ctx_dim=ctx_dim)

    dA = direction(2, R, mv, GV, "irs,ijsw->ijrw")    # i2t: context = regions, query = words
    dB = direction(3, W, mt, GT, "jwx,ijrx->ijrw")    # t2i: context = words, query = regions
    scores = dA["cos"].mean(3).squeeze(2) + dB["cos"].mean(2).squeeze(-1)
    if d_scores is None:
        return scores

    g = d_scores[:, :, None, None]

    def backward(D_, Lq):
        gam = g / Lq
        coef = gam / D_["den"]
        kap = D_["cos"] / D_["den"]
        dz = coef * D_["q"] * (S - kap * D_["Y"])
        dAtt = lam * dz
        t = (D_["A"] * dAtt).sum(D_["q_dim"], keepdim=True)
        dl = (dAtt - D_["A"] * t) / D_["nrm"]
        dS = dl * torch.where(neg, torch.full_like(S, 0.1), torch.ones_like(S)) + coef * D_["q"]
        dD = -gam * D_["cos"] / (2 * D_["D"])
        return dS, dD

    dS_A, dD_A = backward(dA, W)
    dS_B, dD_B = backward(dB, R)
    P = _r(dS_A + dS_B, bf16, "p")
    dVh = torch.einsum("ijrw,jwd->ird", P, Th)
    dTh = torch.einsum("ijrw,ird->jwd", P, Vh)
    dGV = torch.einsum("ijrw,ijsw->irs", _r(dA["q"] * dD_A, bf16, "qd"), dA["q"])
    dGT = torch.einsum("ijrw,ijrx->jwx", _r(dB["q"] * dD_B, bf16, "qd"), dB["q"])
    dVh = dVh + 2 * torch.einsum("irs,isd->ird", dGV, Vh)
    dTh = dTh + 2 * torch.einsum("jwx,jxd->jwd", dGT, Th)
    d_im = alpha * (dVh - Vh * (Vh * dVh).sum(-1, keepdim=True))
    d_s = beta * (dTh - Th * (Th * dTh).sum(-1, keepdim=True))
    return scores, d_im, d_s


def rwa_loss_fwd_bwd(scores, lam=20.0, row_offset=0, n_total=None):
    """KL rows and d_scores: loss = sum_i KL_i / n_total."""
    n_rows, n_cols = scores.shape
    n_total = n_cols if n_total is None else n_total
    logp = torch.log_softmax(lam * scores, dim=1)
    p = logp.exp()
    tgt = torch.zeros_like(scores)
    tgt[torch.arange(n_rows), torch.arange(n_rows) + row_offset] = 1.0
    gg = logp - torch.log(tgt + 1e-6)
    kl = (p * gg).sum(1, keepdim=True)
    return kl.sum() / n_total, (lam / n_total) * p * (gg - kl)
