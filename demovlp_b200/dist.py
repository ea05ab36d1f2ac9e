"""Multi-GPU (one process per GPU) version of the region-word loss.

The (B x B) pair matrix is partitioned by VIDEO ROWS: rank r keeps its b = B/P videos
and scores them against all B texts.  Text embeddings move with an all-gather, their
gradients come back with a reduce-scatter, the row-local RWALoss partial sums meet in
one scalar all-reduce.  Semantically this equals what the reference's
``AllGather_multi`` pattern (trainer/trainer_dist.py:13-31) would compute if it were
enabled in training -- every rank evaluates the loss of the CONCATENATED batch and keeps
the gradient rows of its own shard -- without any rank ever holding the
(B, B, R, W) tensor.

torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests) is the
plumbing; the arithmetic is behind the ``Backend`` hooks, which default to the CUDA
library and are replaced by an oracle-backed stub in tests/test_dist_cpu.py.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops, p2p


class CudaBackend:
    """The shipped arithmetic: C-ABI kernels via demovlp_b200.ops."""

    @staticmethod
    def scores_fwd(im, s_all, im_m, s_m_all, lam, focal_equal, impl="auto"):
        return ops.rw_scores_fwd_raw(ops._prep_embed(im), ops._prep_embed(s_all),
                                     ops._prep_mask(im_m, None), ops._prep_mask(s_m_all, None),
                                     lam, focal_equal, impl)

    @staticmethod
    def scores_bwd(state, d_scores, d_s_out=None):
        return ops.rw_scores_bwd_raw(state, d_scores, d_s_out=d_s_out)

    @staticmethod
    def rwa_tail(scores, lam, row_offset, n_total, want_grad):
        return ops.rwa_tail_raw(scores, lam, row_offset, n_total, want_grad)

    sim_matrix = staticmethod(ops.sim_matrix)
    norm_softmax = staticmethod(ops.norm_softmax_loss)


_BACKEND = CudaBackend


def set_backend(b):
    global _BACKEND
    _BACKEND = b


def gather_cat(t, group=None):
    """all_gather + cat along dim 0 in rank order (trainer_dist.py:17-22), no autograd."""
    world = dist.get_world_size(group)
    t = t.contiguous()
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out


def gather_cat_many(tensors, group=None):
    """gather_cat of several tensors as ONE collective launch where the backend can group them.

    With NCCL the all-gathers are issued inside one group (c10d's coalescing manager): one kernel launch and one
    synchronisation instead of one per tensor -- at the sizes of this path (a few MiB) every collective is
    latency-bound, so the launch count is what matters.  Other backends (gloo in the CPU tests) fall back to
    one call per tensor."""
    world = dist.get_world_size(group)
    ins = [t.contiguous() for t in tensors]
    outs = [torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in ins]
    coalesce = getattr(dist, "_coalescing_manager", None)
    if coalesce is not None and len(ins) > 1 and ins[0].is_cuda and dist.get_backend(group) == "nccl":
        with coalesce(group=group, device=ins[0].device, async_ops=False):
            for o, t in zip(outs, ins):
                dist.all_gather_into_tensor(o, t, group=group)
    else:
        for o, t in zip(outs, ins):
            dist.all_gather_into_tensor(o, t, group=group)
    return outs


class AllGather_multi(torch.autograd.Function):
    """Drop-in for trainer/trainer_dist.py:13-31: ``AllGather_multi.apply(tensor, n_gpu, args)``.

    forward = all_gather + cat in rank order; backward = this rank's slice of the incoming
    gradient, with no reduction (every rank is expected to have computed the same full loss).
    ``args`` needs ``world_size`` and ``rank`` like the reference's argparse namespace.
    """

    @staticmethod
    def forward(ctx, tensor, n_gpu, args):
        ctx.rank = args.rank
        ctx.batch_size = tensor.shape[0]
        return gather_cat(tensor)

    @staticmethod
    def backward(ctx, grad_output):
        b = ctx.batch_size
        return grad_output[b * ctx.rank:b * (ctx.rank + 1)], None, None


class _GatherRows(torch.autograd.Function):
    """Same as AllGather_multi but bound to a process group instead of an args namespace.
    ``gathered`` (optional) is the result of an all-gather that has already been done (gather_cat_many)."""

    @staticmethod
    def forward(ctx, tensor, group, gathered=None):
        ctx.rank = dist.get_rank(group)
        ctx.b = tensor.shape[0]
        return gather_cat(tensor, group) if gathered is None else gathered.view_as(gathered)

    @staticmethod
    def backward(ctx, g):
        return g[ctx.b * ctx.rank:ctx.b * (ctx.rank + 1)], None, None


class _ShardedRWA(torch.autograd.Function):
    """local_loss of the concatenated batch from row shards (model/loss.py:105-116 on P ranks)."""

    @staticmethod
    def forward(ctx, im, s, im_m, s_m, lam, focal_equal, impl, group, pre=None, xchg=None, ticket=None):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        b = s.shape[0]
        if im.shape[0] != b:
            raise RuntimeError("sharded RWALoss needs the same number of videos and texts per rank")
        if pre is None:                                     # (B, W, d): the only embedding that moves
            s_all, s_m_all = gather_cat_many([s.detach(), s_m.detach().to(torch.float32)], group)
        else:
            s_all, s_m_all = pre
        scores, state = _BACKEND.scores_fwd(im.detach(), s_all, im_m, s_m_all, lam, focal_equal, impl)
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        part, d_scores = _BACKEND.rwa_tail(scores, lam, rank * b, world * b, want)
        if xchg is not None:
            loss = xchg.sum_scalar(part)                            # peer-memory transport (p2p.py)
        else:
            loss = part.clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=group)   # rows are rank-local: one scalar
        ctx.state, ctx.d_scores, ctx.group, ctx.b = state, d_scores, group, b
        ctx.xchg, ctx.ticket = xchg, ticket
        ctx.dtypes = (im.dtype, s.dtype)
        return loss

    @staticmethod
    def backward(ctx, g):
        if ctx.d_scores is None:
            raise RuntimeError("sharded RWALoss was evaluated without grad")
        # every rank holds a partial gradient for ALL texts: sum and hand each rank its rows
        if ctx.xchg is not None:
            ctx.xchg.check(ctx.ticket)                      # the gathered texts of this forward are still in place
            d_im, _ = _BACKEND.scores_bwd(ctx.state, ctx.d_scores * g, d_s_out=ctx.xchg.reduce_buffer())
            d_s = ctx.xchg.reduce_scatter(ctx.dtypes[1] if ctx.dtypes[1] == torch.bfloat16 else torch.float32)
        else:
            d_im, d_s_all = _BACKEND.scores_bwd(ctx.state, ctx.d_scores * g)
            d_s = torch.empty((ctx.b,) + tuple(d_s_all.shape[1:]), dtype=d_s_all.dtype, device=d_s_all.device)
            dist.reduce_scatter_tensor(d_s, d_s_all.contiguous(), op=dist.ReduceOp.SUM, group=ctx.group)
        return (d_im.to(ctx.dtypes[0]), d_s.to(ctx.dtypes[1]), None, None, None, None, None, None, None, None, None)


class ShardedGlobalLocalLoss(nn.Module):
    """GlobalLocalLoss (model/loss.py:10-45) over the GLOBAL batch, one shard per rank.

    Same constructor arguments as the reference class plus ``group``.  ``forward`` keeps the
    six-argument signature; because the global cosine matrix needs the other ranks' CLS
    embeddings, the first argument is the pair ``(text_cls, video_cls)`` of THIS rank's
    (b, d) global embeddings instead of the local (b, b) ``sim_matrix`` (one-line change at
    trainer/trainer_dist.py:160).  Returns the reference's 3-tuple; all
    ranks return identical values.
    """

    def __init__(self, temperature=0.05, lambda_softmax=20, focal_type="prob", margin=0,
                 max_violation=False, use_local=True, use_global=True, coef=1000., group=None, impl="auto",
                 transport="auto"):
        super().__init__()
        self.temperature = temperature
        self.lambda_softmax = lambda_softmax
        self.focal_type = focal_type
        self.use_local = use_local
        self.use_global = use_global
        self.cof_local = coef
        self.group = group
        self.impl = impl
        if transport not in ("auto", "nccl", "p2p"):
            raise ValueError("transport must be 'auto', 'nccl' or 'p2p'")
        self.transport = transport          # "p2p": own kernels over peer memory (p2p.py); "nccl": torch.distributed
        self._xchg = None
        self._p2p_failed = False
        self._side = None                   # stream of the global (CLS) branch, see forward

    def global_part(self, text_cls, video_cls, pre=None):
        if pre is None:
            pre = gather_cat_many([text_cls.detach(), video_cls.detach()], self.group)
        t_all = _GatherRows.apply(text_cls, self.group, pre[0])
        v_all = _GatherRows.apply(video_cls, self.group, pre[1])
        return _BACKEND.norm_softmax(_BACKEND.sim_matrix(t_all, v_all), self.temperature)

    def local_part(self, im, s, im_m, s_m, pre=None, xchg=None, ticket=None):
        return _ShardedRWA.apply(im, s, im_m, s_m, self.lambda_softmax, self.focal_type == "equal",
                                 self.impl, self.group, pre, xchg, ticket)

    def _peer_exchange(self, shards, reduce_shard_shape):
        """The PeerExchange for these shard shapes, or None when the NCCL transport is to be used.  Collective:
        every rank takes the same decision (a failure anywhere sends everybody to NCCL)."""
        if self.transport == "nccl" or self._p2p_failed or dist.get_world_size(self.group) == 1:
            return None
        if not p2p.available(shards[0]):
            if self.transport == "p2p":
                raise RuntimeError("transport='p2p' needs CUDA tensors and torch symmetric memory")
            return None
        specs = [(tuple(t.shape), t.dtype) for t in shards]
        key = (tuple((tuple(sh), dt) for sh, dt in specs), tuple(reduce_shard_shape))
        if self._xchg is not None and self._xchg.key == key:
            return self._xchg
        ok, err = 1, None
        try:                         # (re)build the arena: shapes are the same on every rank, so all ranks are here
            if self._xchg is None:
                self._xchg = p2p.PeerExchange(self.group, shards[0].device)
            self._xchg.ensure(specs, reduce_shard_shape)
        except Exception as e:       # e.g. no peer access between the GPUs of this box
            ok, err = 0, e
        flag = torch.tensor([ok], device=shards[0].device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if not int(flag.item()):
            if self.transport == "p2p":
                raise RuntimeError(f"peer-memory transport unavailable: {err}")
            self._p2p_failed, self._xchg = True, None
            return None
        return self._xchg

    def forward(self, global_pair, local_im, local_s, local_im_m, local_s_l, local_s_m):
        zero = torch.tensor([0.0])
        if not self.use_local:
            g = self.global_part(*global_pair)
            return g, g, zero
        if not self.use_global:
            l = self.local_part(local_im, local_s, local_im_m, local_s_m)
            return l, zero, l
        # everything that moves in the forward (text embeddings + masks, both CLS matrices) in ONE exchange:
        # a store into the peers' arenas (own kernel, p2p.py) or one grouped NCCL all-gather
        shards = [global_pair[0].detach(), global_pair[1].detach(), local_s.detach(),
                  local_s_m.detach().to(torch.float32)]
        xchg = self._peer_exchange(shards, tuple(local_s.shape))
        ticket = None
        if xchg is not None:
            (t_all, v_all, s_all, s_m_all), ticket = xchg.allgather(shards)
        else:
            t_all, v_all, s_all, s_m_all = gather_cat_many(shards, self.group)
        # The global branch (cosine matrix of the gathered CLS embeddings + NormSoftmax, replicated on every rank) is
        # independent of the local one until the final add.  Its kernels are small and latency-bound, and so are the
        # prologue / finalisation / exchange kernels of the local branch, so it runs on a second stream: forward next
        # to the item prologue, backward (autograd replays a node on the stream of its forward) next to the gradient
        # finalisation and the reduce-scatter.  Inside a captured step the two streams become parallel graph branches.
        if t_all.is_cuda and os.environ.get("DVLP_GLOBAL_SIDE_STREAM", "1") != "0":
            cur = torch.cuda.current_stream(t_all.device)
            if self._side is None:
                self._side = torch.cuda.Stream(t_all.device)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                g = self.global_part(*global_pair, pre=(t_all, v_all))
            l = self.local_part(local_im, local_s, local_im_m, local_s_m, pre=(s_all, s_m_all), xchg=xchg, ticket=ticket)
            cur.wait_stream(self._side)
            g.record_stream(cur)
        else:
            g = self.global_part(*global_pair, pre=(t_all, v_all))
            l = self.local_part(local_im, local_s, local_im_m, local_s_m, pre=(s_all, s_m_all), xchg=xchg, ticket=ticket)
        return g + l, g, l


class ValidationFeed:
    """Device-resident replacement for the reference's validation bookkeeping (trainer/trainer_dist.py:218-391).

    The reference all-gathers seven tensors per validation batch, moves every gathered tensor to the CPU, concatenates
    there, computes ``sim_matrix`` on CPU tensors and re-uploads 8 x 8 tiles for the local scores.  Here each rank keeps
    its own batch outputs on the GPU (``add``), ``finish`` moves them ONCE (one grouped all-gather per tensor over all
    batches) and restores the reference's row order -- batch-major, rank-major inside a batch, exactly what its per-batch
    ``torch.cat(all_gather(...))`` + final ``torch.cat`` produce --, and ``similarity`` feeds the fused forward-only
    kernels and the rank kernels directly.  Single process (no process group): ``finish`` is a plain concatenation.
    """

    KEYS = ("global_text", "global_object", "local_text", "local_object", "object_mask", "text_mask", "text_length")

    def __init__(self, group=None):
        self.group = group
        self.batches = []
        self.data = None

    def add(self, global_text, global_object, local_text, local_object, object_mask, attention_mask_tail, text_length):
        """One validation batch of THIS rank (model outputs, trainer_dist.py:270-321).  ``attention_mask_tail`` is the
        tokenizer's 0/1 mask without the CLS column (``data['text']['attention_mask'][:, 1:]``, :262-267); it becomes
        the additive mask ``(m - 1) * 100`` of :323."""
        text_mask = (attention_mask_tail.to(torch.float32) - 1.0) * 100.0
        self.batches.append(tuple(t.detach() for t in (global_text, global_object, local_text, local_object,
                                                       object_mask, text_mask, text_length)))

    def finish(self):
        """Gather all batches of all ranks; returns the dict of concatenated tensors in the reference's order."""
        if not self.batches:
            raise RuntimeError("ValidationFeed.finish: no batches were added")
        mine = [torch.cat([b[k] for b in self.batches]).contiguous() for k in range(len(self.KEYS))]
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        if not multi:
            self.data = dict(zip(self.KEYS, mine))
            return self.data
        world = dist.get_world_size(self.group)
        sizes = [b[0].shape[0] for b in self.batches]             # rows per batch (the same on every rank, like the
        gathered = gather_cat_many(mine, self.group)              # reference's equal-shape all_gather requires)
        n_local = sum(sizes)
        # gathered rows are [rank][batch][row]; the reference order is [batch][rank][row]
        idx, start = [], 0
        for n in sizes:
            for r in range(world):
                idx.append(torch.arange(r * n_local + start, r * n_local + start + n))
            start += n
        perm = torch.cat(idx).to(gathered[0].device)
        self.data = {k: g.index_select(0, perm) for k, g in zip(self.KEYS, gathered)}
        return self.data

    def similarity(self, loss_module, use_local=True, rows_per_launch=256):
        """float64 numpy matrix the reference hands to its metrics (:369-385): ``sim_matrix(text, object)`` plus, when
        ``use_local``, ``local_loss.get_sim_by_segment(...)`` (video rows x text columns -- the reference adds the two
        orientations as they are, which is only meaningful for one caption per video; kept)."""
        from . import loss as L
        d = self.data if self.data is not None else self.finish()
        sims = L.sim_matrix(d["global_text"], d["global_object"]).detach().to(torch.float64)
        if use_local:
            local = loss_module.local_loss
            n_img = d["local_object"].shape[0]
            if sims.shape != (d["local_text"].shape[0], n_img) or sims.shape[0] != sims.shape[1]:
                raise RuntimeError("global (text x video) and local (video x text) matrices only add up when square")
            with torch.no_grad():
                for v0 in range(0, n_img, rows_per_launch):
                    v1 = min(v0 + rows_per_launch, n_img)
                    sims[v0:v1] += local.get_sim(d["local_object"][v0:v1], d["local_text"], d["object_mask"][v0:v1],
                                                 d["text_length"], d["text_mask"]).to(torch.float64)
        return sims.cpu().numpy()
