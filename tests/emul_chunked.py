"""Torch (CPU) emulation of the LARGE-SHAPE forward planned in DESIGN.md section 8.1 (not shipped, not the oracle).

It restates, for one (video, text) pair at a time, exactly the data flow the planned tcgen05 kernel will use, so that
its algebra is pinned against the oracle on CPU before any kernel is written:

  * regions are ZERO-PADDED up to a multiple of the chunk size C and words up to the tile height Wt; the padded
    rows carry the additive mask of a padded position, the REAL counts (R, W) enter only the focal thresholds and
    the means;
  * two passes over S (resident in TMEM in the kernel): pass 1 gives the column norms nA (complete inside a chunk),
    the row sums of l^2 (-> nB) and of e1 (-> thr1) accumulated over the chunks; pass 2 recomputes e1, selects u1,
    forms e2 with nB, its per-chunk column sums (-> thr2), u2 and the numerators;
  * the Gram-form denominators are accumulated block by block:  Y1_c = sum_c' U1_c' G^V[c', c],  D1 = sum_c <u1_c, Y1_c>;
    Y2_c = G^T U2_c is chunk-local.
"""
import torch

EPS = 1e-8
PAD_MASK = -100.0        # additive mask of a padded position (the reference's own convention)


def _leaky(x):
    return torch.where(x > 0, x, 0.1 * x)


def pair_score_chunked(V, T, mv, mt, C=96, Wt=128, lam=20.0, focal=True):
    """Score of one pair.  V (R, d), T (W, d) raw embeddings, mv (R,), mt (W,) additive masks."""
    R, d = V.shape
    W = T.shape[0]
    Rp = (R + C - 1) // C * C
    Wp = (W + Wt - 1) // Wt * Wt if Wt else W
    dt = V.dtype
    Vp = torch.zeros(Rp, d, dtype=dt); Vp[:R] = V
    Tp = torch.zeros(Wp, d, dtype=dt); Tp[:W] = T
    mvp = torch.full((Rp,), PAD_MASK, dtype=dt); mvp[:R] = mv
    mtp = torch.full((Wp,), PAD_MASK, dtype=dt); mtp[:W] = mt
    # the kernels subtract the item's largest mask value (a fully masked context stays finite)
    mvp = mvp - mv.max(); mtp = mtp - mt.max()
    alpha = 1.0 / (Vp.norm(dim=1) + EPS)              # 1e8 for the zero rows: they only ever multiply exact zeros
    beta = 1.0 / (Tp.norm(dim=1) + EPS)
    Vh, Th = Vp * alpha[:, None], Tp * beta[:, None]
    GV, GT = Vh @ Vh.T, Th @ Th.T                      # per-item Gram matrices (prologue kernel)
    chunks = [slice(c, c + C) for c in range(0, Rp, C)]

    # S[w, r] in the kernel's orientation (lane = word, column = region), all chunks resident
    S = (Tp @ Vp.T) * beta[:, None] * alpha[None, :]
    L = _leaky(S)
    # ---- pass 1
    ssB = torch.zeros(Wp, dtype=dt); sumA = torch.zeros(Wp, dtype=dt)
    nA = torch.zeros(Rp, dtype=dt)
    for c in chunks:
        l = L[:, c]
        nA[c] = 1.0 / (l.square().sum(0).sqrt() + EPS)            # norm over the words: complete inside the chunk
        ssB += l.square().sum(1)
        sumA += torch.exp(lam * (l * nA[c][None, :] + mvp[c][None, :]) - lam).sum(1)
    nB = 1.0 / (ssB.sqrt() + EPS)
    thr1 = sumA / R if focal else torch.full_like(sumA, -1.0)      # REAL region count
    # ---- pass 2
    U1 = torch.zeros(Wp, Rp, dtype=dt); U2 = torch.zeros(Wp, Rp, dtype=dt)
    num1 = torch.zeros(Wp, dtype=dt); num2 = torch.zeros(Rp, dtype=dt)
    for c in chunks:
        l, s = L[:, c], S[:, c]
        e1 = torch.exp(lam * (l * nA[c][None, :] + mvp[c][None, :]) - lam)
        u1 = torch.where(e1 > thr1[:, None], e1, torch.zeros_like(e1))
        e2 = torch.exp(lam * (l * nB[:, None] + mtp[:, None]) - lam)
        thr2 = e2.sum(0) / W if focal else torch.full((C,), -1.0, dtype=dt)     # REAL word count; chunk-local
        u2 = torch.where(e2 > thr2[None, :], e2, torch.zeros_like(e2))
        U1[:, c], U2[:, c] = u1, u2
        num1 += (u1 * s).sum(1)
        num2[c] = (u2 * s).sum(0)
    # ---- Gram-form denominators, block by block
    D1 = torch.zeros(Wp, dtype=dt); D2 = torch.zeros(Rp, dtype=dt)
    for c in chunks:
        Y1 = torch.zeros(Wp, C, dtype=dt)
        for c2 in chunks:
            Y1 += U1[:, c2] @ GV[c2, c]
        D1 += (U1[:, c] * Y1).sum(1)
        Y2 = GT @ U2[:, c]
        D2[c] = (U2[:, c] * Y2).sum(0)
    cos1 = torch.where(D1 > 1e-30, num1 / D1.clamp_min(1e-30).sqrt(), torch.zeros_like(num1))
    cos2 = torch.where(D2 > 1e-30, num2 / D2.clamp_min(1e-30).sqrt(), torch.zeros_like(num2))
    # padded words / regions are queries whose cosine is exactly zero; the means run over the REAL counts
    assert float(cos1[W:].abs().sum()) == 0.0 and float(cos2[R:].abs().sum()) == 0.0
    return cos1[:W].sum() / W + cos2[:R].sum() / R


def scores_chunked(im, s, im_mask, s_mask, **kw):
    Bv, Bt = im.shape[0], s.shape[0]
    out = torch.zeros(Bv, Bt, dtype=im.dtype)
    for i in range(Bv):
        for j in range(Bt):
            out[i, j] = pair_score_chunked(im[i], s[j], im_mask[i].to(im.dtype), s_mask[j].to(im.dtype), **kw)
    return out


def forward_backward_padded(im, s, im_mask, s_mask, d_scores, C=96, Wt=128, lam=20.0, focal=True):
    """Kernel-style forward + hand-derived backward (tests/emul.py's Gram form) on ZERO-PADDED items, with what the
    planned kernel needs on top of the small-shape one:

      * real counts in the focal thresholds and in the means;
      * padded QUERIES (word rows >= W in the i2t direction, region columns >= R in the t2i direction) carry no
        gradient: their upstream factor gamma is forced to zero (their cosine is 0 / den, den > 0);
      * the inverse norm of a zero row is 1 / eps = 1e8: it only ever multiplies exact zeros on the way to REAL rows.

    Returns (scores, d_im, d_s) for the real rows."""
    Bv, R, d = im.shape
    Bt, W, _ = s.shape
    dt = im.dtype
    Rp = (R + C - 1) // C * C
    Wp = (W + Wt - 1) // Wt * Wt
    imp = torch.zeros(Bv, Rp, d, dtype=dt); imp[:, :R] = im
    sp = torch.zeros(Bt, Wp, d, dtype=dt); sp[:, :W] = s
    mvp = torch.full((Bv, Rp), PAD_MASK, dtype=dt); mvp[:, :R] = im_mask.to(dt)
    mtp = torch.full((Bt, Wp), PAD_MASK, dtype=dt); mtp[:, :W] = s_mask.to(dt)
    mvp = mvp - im_mask.to(dt).max(1, keepdim=True).values
    mtp = mtp - s_mask.to(dt).max(1, keepdim=True).values
    alpha = 1.0 / (imp.norm(dim=-1, keepdim=True) + EPS); beta = 1.0 / (sp.norm(dim=-1, keepdim=True) + EPS)
    Vh, Th = imp * alpha, sp * beta
    GV = torch.einsum("ird,isd->irs", Vh, Vh); GT = torch.einsum("jwd,jxd->jwx", Th, Th)
    S = torch.einsum("ird,jwd->ijrw", Vh, Th)
    neg = S <= 0
    l = torch.where(neg, 0.1 * S, S)
    mv = mvp[:, None, :, None]; mt = mtp[None, :, None, :]
    rvalid = (torch.arange(Rp) < R).to(dt)[None, None, :, None]       # real regions
    wvalid = (torch.arange(Wp) < W).to(dt)[None, None, None, :]       # real words

    def direction(ctx_dim, n_ctx_real, mask, G, eq):
        q_dim = 5 - ctx_dim
        nrm = l.square().sum(q_dim, keepdim=True).sqrt() + EPS
        A = l / nrm
        e = torch.exp(lam * (A + mask) - lam)
        E = e.sum(ctx_dim, keepdim=True)
        H = (e * n_ctx_real > E) if focal else torch.ones_like(e, dtype=torch.bool)
        u = torch.where(H, e, torch.zeros_like(e))                    # un-normalised, as the kernels keep it
        num = (u * S).sum(ctx_dim, keepdim=True)
        Y = torch.einsum(eq, G, u)
        D = (u * Y).sum(ctx_dim, keepdim=True)
        den = D.clamp_min(1e-30).sqrt()
        cos = torch.where(D > 1e-30, num / den, torch.zeros_like(num))
        return dict(nrm=nrm, A=A, u=u, Y=Y, D=D, den=den, cos=cos, q_dim=q_dim)

    dA = direction(2, R, mv, GV, "irs,ijsw->ijrw")     # i2t: context = regions, query = words
    dB = direction(3, W, mt, GT, "jwx,ijrx->ijrw")     # t2i: context = words, query = regions
    scores = (dA["cos"] * wvalid).sum(3).squeeze(2) / W + (dB["cos"] * rvalid).sum(2).squeeze(-1) / R
    g = d_scores[:, :, None, None]

    def backward(D_, gam):
        # gam: upstream factor per query (0 for padded queries)
        c1 = gam / D_["den"]
        kap = D_["cos"] / D_["den"]
        dAtt = lam * c1 * D_["u"] * (S - kap * D_["Y"])
        t = (D_["A"] * dAtt).sum(D_["q_dim"], keepdim=True)
        dl = (dAtt - D_["A"] * t) / D_["nrm"]
        dS = dl * torch.where(neg, torch.full_like(S, 0.1), torch.ones_like(S)) + c1 * D_["u"]
        dD = -gam * D_["cos"] / (2 * D_["D"].clamp_min(1e-30))
        return dS, dD

    dS_A, dD_A = backward(dA, g / W * wvalid)
    dS_B, dD_B = backward(dB, g / R * rvalid)
    P = dS_A + dS_B
    dVh = torch.einsum("ijrw,jwd->ird", P, Th)
    dTh = torch.einsum("ijrw,ird->jwd", P, Vh)
    dGV = torch.einsum("ijrw,ijsw->irs", dA["u"] * dD_A, dA["u"])
    dGT = torch.einsum("ijrw,ijrx->jwx", dB["u"] * dD_B, dB["u"])
    dVh = dVh + 2 * torch.einsum("irs,isd->ird", dGV, Vh)
    dTh = dTh + 2 * torch.einsum("jwx,jxd->jwd", dGT, Th)
    d_im = alpha * (dVh - Vh * (Vh * dVh).sum(-1, keepdim=True))
    d_s = beta * (dTh - Th * (Th * dTh).sum(-1, keepdim=True))
    assert torch.isfinite(d_im[:, :R]).all() and torch.isfinite(d_s[:, :W]).all()
    return scores, d_im[:, :R], d_s[:, :W]
