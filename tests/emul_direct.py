"""Torch (CPU) emulation of the algebra the LARGE-SHAPE kernels evaluate (pair_fwd3.cu / pair_bwd3.cu / grad_gemm3.cu;
DESIGN.md section 5.5) -- not shipped, not the oracle.  It differs from tests/emul_chunked.py (the round-1 plan) in the
i2t direction, which uses the DIRECT form of the cosine denominator instead of a video Gram matrix:

    C_w  = sum_r (u1 alpha)_wr V_r           (raw video rows; the weighted context of word w)
    D1_w = |C_w|^2,   Y1_wr = alpha_r <C_w, V_r>              (the "Gram" term of the backward without a Gram matrix)
    dV'  = sum P^T T + UD1^T C,   UD1 = (u1 alpha) 2 dD1       (video gradient; no dG^V, no Gram finalisation)

and it spells out the backward exactly as pair_bwd3.cu factors it (G, F without lambda; column sums sA, row sums sB;
P = a - b (nA^3 sA + nB^3 sB)).  Zero padding of regions / words as in the kernels: alpha = beta = 0 and mask term -inf on
padded positions, REAL counts in thresholds and means.
"""
import torch

EPS = 1e-8


def _leaky(x):
    return torch.where(x > 0, x, 0.1 * x)


def forward_backward_direct(im, s, im_mask, s_mask, d_scores, Rp=None, Wp=None, lam=20.0, focal=True):
    """(scores, d_im, d_s) for im (Bv, R, d), s (Bt, W, d), additive masks, upstream d_scores (Bv, Bt)."""
    Bv, R, d = im.shape
    Bt, W, _ = s.shape
    dt = im.dtype
    Rp = Rp or R
    Wp = Wp or W
    V = torch.zeros(Bv, Rp, d, dtype=dt); V[:, :R] = im
    T = torch.zeros(Bt, Wp, d, dtype=dt); T[:, :W] = s
    NEG = torch.tensor(float("-inf"), dtype=dt)
    mv = torch.full((Bv, Rp), float("-inf"), dtype=dt); mv[:, :R] = im_mask.to(dt) - im_mask.to(dt).max(1, keepdim=True).values
    mt = torch.full((Bt, Wp), float("-inf"), dtype=dt); mt[:, :W] = s_mask.to(dt) - s_mask.to(dt).max(1, keepdim=True).values
    alpha = torch.zeros(Bv, Rp, dtype=dt); alpha[:, :R] = 1.0 / (im.norm(dim=-1) + EPS)       # 0 on padded regions
    beta = torch.zeros(Bt, Wp, dtype=dt); beta[:, :W] = 1.0 / (s.norm(dim=-1) + EPS)          # 0 on padded words
    Th = T * beta[..., None]
    GT = torch.einsum("jwd,jxd->jwx", Th, Th)                                                  # text Gram (t2i keeps the Gram form)
    raw = torch.einsum("ird,jwd->ijwr", V, T)                                                  # [video, text, word, region]
    S = raw * alpha[:, None, None, :] * beta[None, :, :, None]
    l = _leaky(S)
    lg = torch.where(S > 0, torch.ones_like(S), torch.full_like(S, 0.1))
    # ---- i2t direction: context = regions (last axis), query = words
    nA = 1.0 / (l.square().sum(2, keepdim=True).sqrt() + EPS)                                  # norm over the WORDS, per (text, region)
    e1 = torch.exp(lam * (l * nA + mv[:, None, None, :]) - lam)
    thr1 = e1.sum(3, keepdim=True) / R if focal else torch.full_like(e1[..., :1], -1.0)
    u1 = torch.where(e1 > thr1, e1, torch.zeros_like(e1))
    ua = u1 * alpha[:, None, None, :]
    C = torch.einsum("ijwr,ird->ijwd", ua, V)                                                  # direct form
    D1 = C.square().sum(-1)
    num1 = (u1 * S).sum(3)
    rden1 = torch.where(D1 > 1e-30, D1.clamp_min(1e-30).rsqrt(), torch.zeros_like(D1))
    cos1 = num1 * rden1                                                                        # padded words: num1 == 0
    # ---- t2i direction: context = words, query = regions
    nB = 1.0 / (l.square().sum(3, keepdim=True).sqrt() + EPS)                                  # norm over the REGIONS, per word
    e2 = torch.exp(lam * (l * nB + mt[None, :, :, None]) - lam)
    thr2 = e2.sum(2, keepdim=True) / W if focal else torch.full_like(e2[:, :, :1], -1.0)
    u2 = torch.where(e2 > thr2, e2, torch.zeros_like(e2))
    Y2 = torch.einsum("jwx,ijxr->ijwr", GT, u2)
    D2 = (u2 * Y2).sum(2)
    num2 = (u2 * S).sum(2)
    rden2 = torch.where(D2 > 1e-30, D2.clamp_min(1e-30).rsqrt(), torch.zeros_like(D2))
    cos2 = num2 * rden2                                                                        # padded regions: num2 == 0
    scores = cos1.sum(2) / W + cos2.sum(2) / R
    # ---- backward as pair_bwd3.cu factors it
    g = d_scores.to(dt)
    c1A = (g / W)[:, :, None] * rden1                                                          # per word row
    kapA = cos1 * rden1
    c1B = (g / R)[:, :, None] * rden2                                                          # per (text, region) column
    kapB = cos2 * rden2
    Y1 = alpha[:, None, None, :] * torch.einsum("ijwd,ird->ijwr", C, V)                        # alpha_r <C_w, V_r>
    G = c1A[..., None] * u1 * (S - kapA[..., None] * Y1)
    F = c1B[:, :, None, :] * u2 * (S - kapB[:, :, None, :] * Y2)
    sA = (l * G).sum(2, keepdim=True)                                                          # over the words: per column
    sB = (l * F).sum(3, keepdim=True)                                                          # over the regions: per row
    dS = lg * lam * (G * nA + F * nB - l * (nA ** 3 * sA + nB ** 3 * sB)) + c1A[..., None] * u1 + c1B[:, :, None, :] * u2
    P = dS * alpha[:, None, None, :] * beta[None, :, :, None]
    UD1 = ua * (-(c1A * kapA))[..., None]                                                      # (u1 alpha) 2 dD1
    dVp = torch.einsum("ijwr,jwd->ird", P, T) + torch.einsum("ijwr,ijwd->ird", UD1, C)         # alpha-prescaled dV^
    dTp = torch.einsum("ijwr,ird->jwd", P, V)                                                  # beta-prescaled dT^ (S path)
    dGT = torch.einsum("ijwr,ijxr->jwx", u2 * (-0.5 * c1B * kapB)[:, :, None, :], u2)
    # finalisation (epilogue.cu): g = dX'/inv + 2 dG Xh;  d_x = inv (g - xh (xh . g))
    a_safe = torch.where(alpha > 0, alpha, torch.ones_like(alpha))
    b_safe = torch.where(beta > 0, beta, torch.ones_like(beta))
    Vh = V * alpha[..., None]
    gV = dVp / a_safe[..., None]
    gT = dTp / b_safe[..., None] + 2 * torch.einsum("jwx,jxd->jwd", dGT, Th)
    d_im = alpha[..., None] * (gV - Vh * (Vh * gV).sum(-1, keepdim=True))
    d_s = beta[..., None] * (gT - Th * (Th * gT).sum(-1, keepdim=True))
    return scores, d_im[:, :R], d_s[:, :W]
