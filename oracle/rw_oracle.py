"""CPU oracle for DemoVLP's region-word alignment path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement (torch CPU tensors, fp32 or fp64)
of the arithmetic in the reference's ``model/loss.py``, ``model/model.py``
(``sim_matrix``), ``model/metric.py`` and the eval-matrix assembly in
``trainer/trainer_dist.py``.  It exists to CHECK the CUDA path:

  * only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
    ``cpu_baseline`` / ``--impl reference`` legs may import it;
  * the shipped package ``demovlp_b200`` never imports it and has no CPU
    fallback.

Pinning.  The reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against the reference ITSELF:
``oracle/make_golden.py`` imports the unmodified reference modules from
``/root/reference`` (build container only), runs them on seeded inputs and
commits the outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks
this restatement against those fixtures everywhere, and against the live
reference whenever ``/root/reference`` is mounted.

Every function cites the reference lines it restates (paths relative to the
reference root).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

EPS = 1e-8


# --------------------------------------------------------------------------
# scoring: model/loss.py:209-338
# --------------------------------------------------------------------------
def unit_rows(x: torch.Tensor, dim: int = -1) -> torch.Tensor:
    """x / (||x||_2 + 1e-8) along ``dim``  (model/loss.py:333-338 ``l2norm``)."""
    return x / (x.square().sum(dim=dim, keepdim=True).sqrt() + EPS)


def attend(query, context, q_mask, c_mask, lam=20.0, focal="equal"):
    """One attention direction (model/loss.py:209-271 ``func_attention_fast``).

    query   (Bq, Lq, d)   context (Bc, Lc, d)
    q_mask  (Bq, Lq)      c_mask  (Bc, Lc)     additive masks
    returns the attended context (Bc, Bq, Lq, d) and the final attention
    weights (Bc, Bq, Lq, Lc) (the latter is extra, for kernel debugging).
    """
    qn = unit_rows(query)                                   # :225
    cn = unit_rows(context)                                 # :226
    sim = torch.einsum("cld,bqd->cblq", cn, qn)             # :235   (Bc,Bq,Lc,Lq)
    sim = F.leaky_relu(sim, 0.1)                            # :236
    sim = unit_rows(sim, dim=3)                             # :238   over the QUERY axis
    logit = sim.transpose(2, 3)                             # :241   (Bc,Bq,Lq,Lc)
    logit = logit + q_mask[None, :, :, None] + c_mask[:, None, None, :]   # :242
    p = torch.softmax(logit * lam, dim=-1)                  # :244-245 over the CONTEXT axis
    if focal == "equal":                                    # :251-252, :274-283
        keep = (p * p.shape[-1] - p.sum(-1, keepdim=True)) > 0
        u = torch.where(keep, p, torch.zeros_like(p))       # :257
    else:                                                   # :254
        u = p
    w = u / u.sum(-1, keepdim=True)                         # :258-259
    out = torch.einsum("cbql,cld->cbqd", w.to(cn.dtype), cn)   # :262-269
    return out, w


def cosine_rows(x1, x2):
    """sum(x1*x2) / clamp(|x1||x2|, 1e-8) over the last axis (model/loss.py:286-291).

    The reference additionally ``.squeeze()``s the result, which mis-shapes it
    when any of n_image / n_caption / L equals 1; this restatement keeps the
    full rank and is only compared with the reference on sizes > 1.
    """
    num = (x1 * x2).sum(-1)
    den = (x1.norm(2, -1) * x2.norm(2, -1)).clamp(min=EPS)
    return num / den


def pair_scores(im, s, im_mask, s_mask=None, lam=20.0, focal="equal"):
    """(n_video, n_text) region-word scores (model/loss.py:294-330 ``xattn_score_fast``).

    ``cap_lens`` of the reference is unused there (:297) and is not a parameter here.
    """
    if s_mask is None:                                      # :307-308
        s_mask = torch.ones(s.shape[0], s.shape[1]).type_as(im_mask)
    ctx_w, _ = attend(s, im, s_mask, im_mask, lam, focal)   # :311-316  (Bv,Bt,W,d)
    i2t = cosine_rows(s[None], ctx_w).mean(-1)              # :317-318  (Bv,Bt)
    ctx_r, _ = attend(im, s, im_mask, s_mask, lam, focal)   # :320-325  (Bt,Bv,R,d)
    t2i = cosine_rows(im[None], ctx_r).mean(-1)             # :326-327  (Bt,Bv)
    return t2i.T + i2t                                      # :329


# --------------------------------------------------------------------------
# loss tails: model/loss.py:105-138, :10-45 ; model/model.py:582-590
# --------------------------------------------------------------------------
def rwa_loss_from_scores(scores, lam=20.0):
    """KL(softmax(lam*scores, 1) || eye + 1e-6), row-summed, mean (model/loss.py:108-114)."""
    n = scores.shape[0]
    target = torch.eye(n).type_as(scores)
    logp = torch.log_softmax(scores * lam, dim=1)
    kl = logp.exp() * (logp - torch.log(target + 1e-6))
    return kl.sum(1).mean()


def norm_softmax_loss(x, temperature=0.05):
    """Two-way InfoNCE over a cosine matrix (model/loss.py:126-138)."""
    rows = torch.log_softmax(x / temperature, dim=1).diagonal()
    cols = torch.log_softmax(x.t() / temperature, dim=1).diagonal()
    return -rows.sum() / rows.numel() - cols.sum() / cols.numel()


def sim_matrix(a, b, eps=1e-8):
    """Cosine matrix rows=a, cols=b with max(norm, eps) (model/model.py:582-590)."""
    an = a / a.norm(dim=1, keepdim=True).clamp(min=eps)
    bn = b / b.norm(dim=1, keepdim=True).clamp(min=eps)
    return an @ bn.t()


def global_local_loss(global_sim, im, s, im_mask, s_len, s_mask, *,
                      temperature=0.05, lam=20.0, focal="equal",
                      use_local=True, use_global=True):
    """(loss, global_loss, local_loss) (model/loss.py:29-45).  ``s_len`` is unused (:297)."""
    zero = torch.tensor([0.0])
    if not use_local:
        g = norm_softmax_loss(global_sim, temperature)
        return g, g, zero
    local = rwa_loss_from_scores(pair_scores(im, s, im_mask, s_mask, lam, focal), lam)
    if not use_global:
        return local, zero, local
    g = norm_softmax_loss(global_sim, temperature)
    return g + local, g, local


# --------------------------------------------------------------------------
# eval matrix: model/loss.py:73-103, trainer/trainer_dist.py:369-385
# --------------------------------------------------------------------------
def scores_by_segment(im, s, im_mask, s_mask, lam=20.0, focal="equal", segment=8):
    """float64 numpy (n_video, n_text) built tile by tile (model/loss.py:73-103)."""
    nv, nt = im.shape[0], s.shape[0]
    out = np.zeros((nv, nt))
    for v0 in range(0, nv, segment):
        for t0 in range(0, nt, segment):
            blk = pair_scores(im[v0:v0 + segment], s[t0:t0 + segment],
                              im_mask[v0:v0 + segment], s_mask[t0:t0 + segment], lam, focal)
            out[v0:v0 + segment, t0:t0 + segment] = blk.detach().numpy()
    return out


def eval_matrix(g_text, g_vid, local_scores):
    """sim_matrix(text, video).numpy() + local  (trainer/trainer_dist.py:369-385).

    The reference adds a (text x video) matrix to a (video x text) one; this is
    only shape-valid for N x N and is kept as is.
    """
    return sim_matrix(g_text, g_vid).detach().numpy() + local_scores


# --------------------------------------------------------------------------
# retrieval metrics: model/metric.py:10-122, 125-214, 298-308
# --------------------------------------------------------------------------
def t2v_ranks(sims: np.ndarray) -> np.ndarray:
    """Rank of the ground-truth video for each text query (model/metric.py:22-66).

    The reference sorts ``-sims`` per row and takes the FIRST column equal to the
    ground-truth distance ("optimistic" tie-break, :61-66); that index is the count
    of entries strictly greater than the ground-truth similarity.
    """
    nq, nv = sims.shape
    per = nq // nv                                           # :29
    gt_col = np.arange(nq) // per
    gt = sims[np.arange(nq), gt_col]
    return (sims > gt[:, None]).sum(axis=1).astype(np.int64)


def v2t_ranks(sims: np.ndarray) -> np.ndarray:
    """Rank of the closest ground-truth caption per video (model/metric.py:142-189).

    Ties are AVERAGED (:154, :181-186): mean position of all equal entries
    = #greater + (#equal - 1) / 2; the minimum over a video's captions is kept.
    """
    st = sims.T
    nv, nc = st.shape
    per = nc // nv
    out = np.empty(nv, dtype=np.float64)
    for i in range(nv):
        row = st[i]
        best = np.inf
        for j in range(i * per, (i + 1) * per):
            greater = float((row > row[j]).sum())
            equal = float((row == row[j]).sum())
            best = min(best, greater + (equal - 1.0) / 2.0)
        out[i] = best
    return out


def cols2metrics(cols, num_queries):
    """R@K, MedR, MeanR, geometric mean (model/metric.py:298-308)."""
    cols = np.asarray(cols)
    m = {
        "R1": 100 * float(np.sum(cols == 0)) / num_queries,
        "R5": 100 * float(np.sum(cols < 5)) / num_queries,
        "R10": 100 * float(np.sum(cols < 10)) / num_queries,
        "R50": 100 * float(np.sum(cols < 50)) / num_queries,
        "MedR": np.median(cols) + 1,
        "MeanR": np.mean(cols) + 1,
    }
    with np.errstate(divide="ignore"):
        m["geometric_mean_R1-R5-R10"] = float(np.exp(np.mean(np.log(
            np.array([m["R1"], m["R5"], m["R10"]], dtype=np.float64)))))
    return m


def t2v_metrics(sims):
    return cols2metrics(t2v_ranks(np.asarray(sims)), sims.shape[0])


def v2t_metrics(sims):
    return cols2metrics(v2t_ranks(np.asarray(sims)), sims.shape[1])


# --------------------------------------------------------------------------
# multi-rank semantics: trainer/trainer_dist.py:13-31
# --------------------------------------------------------------------------
def sharded_reference(batch_fn, world, **loss_kw):
    """What P ranks running AllGather_multi + the full-batch loss would produce.

    ``batch_fn()`` returns the concatenated (global) tensors with requires_grad
    set on im / s / g_text / g_vid; the loss is evaluated once on the global
    batch (every rank computes the same scalar, trainer_dist.py:17-22) and the
    backward of the gather hands rank r rows [r*b, (r+1)*b) of the gradient
    without any reduction (:25-31).  Returns the loss triple and the list of
    per-rank gradient slices.
    """
    im, s, im_mask, s_len, s_mask, g_text, g_vid = batch_fn()
    loss, g, l = global_local_loss(sim_matrix(g_text, g_vid), im, s, im_mask,
                                   s_len, s_mask, **loss_kw)
    loss.backward()
    b = im.shape[0] // world
    grads = [dict(im=im.grad[r * b:(r + 1) * b], s=s.grad[r * b:(r + 1) * b],
                  g_text=g_text.grad[r * b:(r + 1) * b], g_vid=g_vid.grad[r * b:(r + 1) * b])
             for r in range(world)]
    return (loss.detach(), g.detach(), l.detach()), grads


def projection(tokens, weight, bias=None, relu_in=False):
    """The projection heads in front of the path (test infrastructure): ``txt_proj = Sequential(ReLU(), Linear)``
    (model/model.py:40-43, :88) with relu_in=True, ``ObjectTransformer.proj = Linear(bias=False)``
    (model/object_transformer.py:381, :451) with relu_in=False; returns (global = token 0, local = tokens 1..) as
    model/model.py:88-96 splits them."""
    import torch.nn.functional as F
    x = F.relu(tokens) if relu_in else tokens
    y = F.linear(x, weight, bias)
    return y[:, 0, ...], y[:, 1:, ...]


def attn_mask(q, k, v, mask=None):
    """Masked attention of the ObjectTransformer (test infrastructure): model/object_transformer.py:93-100 --
    sim = einsum('b i d, b j d -> b i j'); sim += mask; softmax over j; einsum with v."""
    sim = torch.einsum("bid,bjd->bij", q, k)
    if mask is not None:
        sim = sim + mask
    return torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v)
