"""world_size-2 gloo test of the sharded loss plumbing (gather order, scalar all-reduce,
reduce-scatter of text gradients) with an oracle-backed arithmetic stub: the result must equal
the single-process oracle on the concatenated batch (trainer/trainer_dist.py:13-31 semantics)."""
import os
import socket
import sys
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    """Arithmetic stub for CPU: the oracle restatement + its autograd (tests only)."""

    @staticmethod
    def scores_fwd(im, s_all, im_m, s_m_all, lam, focal_equal, impl="auto"):
        from oracle import rw_oracle as O
        im = im.clone().requires_grad_(True)
        s_all = s_all.clone().requires_grad_(True)
        with torch.enable_grad():
            scores = O.pair_scores(im, s_all, im_m, s_m_all, lam, "equal" if focal_equal else "prob")
        return scores.detach(), (scores, im, s_all)

    @staticmethod
    def scores_bwd(state, d_scores):
        scores, im, s_all = state
        scores.backward(d_scores)
        return im.grad, s_all.grad

    @staticmethod
    def rwa_tail(scores, lam, row_offset, n_total, want_grad):
        from tests import emul
        return emul.rwa_loss_fwd_bwd(scores, lam, row_offset, n_total)

    @staticmethod
    def sim_matrix(a, b):
        from oracle import rw_oracle as O
        return O.sim_matrix(a, b)

    @staticmethod
    def norm_softmax(x, t):
        from oracle import rw_oracle as O
        return O.norm_softmax_loss(x, t)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    from demovlp_b200 import dist as D, synth
    D.set_backend(OracleBackend)
    B, b = 8, 8 // world
    full = synth.make_batch(B, 20, 8, 32, seed=13, kappa=0.2)
    sh = full.rows(rank * b, (rank + 1) * b)
    im = sh.im.clone().requires_grad_(True)
    s = sh.s.clone().requires_grad_(True)
    gt = sh.g_text.clone().requires_grad_(True)
    gv = sh.g_vid.clone().requires_grad_(True)
    mod = D.ShardedGlobalLocalLoss(focal_type="equal", coef=1.0)
    loss, g, l = mod((gt, gv), im, s, sh.im_mask, sh.s_len, sh.s_mask)
    loss.backward()

    # reference-compatible AllGather_multi: forward cat in rank order, backward = own slice
    class A: pass
    A.world_size, A.rank = world, rank
    x = (torch.arange(b * 3, dtype=torch.float32).view(b, 3) + 100 * rank).requires_grad_(True)
    y = D.AllGather_multi.apply(x, 1, A)
    (y * torch.arange(y.numel(), dtype=torch.float32).view_as(y)).sum().backward()
    # validation feed: two batches per rank; rows carry (batch, rank, row) tags
    feed = D.ValidationFeed()
    for bi, n in enumerate((3, 2)):
        tag = lambda w: (torch.arange(n, dtype=torch.float32)[:, None] + 10 * rank + 100 * bi).repeat(1, w)
        feed.add(tag(4), tag(4) + 0.5, tag(6).view(n, 3, 2), tag(8).view(n, 4, 2), tag(4), torch.ones(n, 3), tag(1)[:, 0])
    val = feed.finish()
    torch.save(dict(loss=loss.detach(), g=g.detach(), l=l.detach(), d_im=im.grad, d_s=s.grad,
                    d_gt=gt.grad, d_gv=gv.grad, gathered=y.detach(), x_grad=x.grad,
                    val_rows=val["global_text"][:, 0].clone(), val_obj=val["local_object"][:, 0, 0].clone(),
                    val_len=val["text_length"].clone(), val_mask=val["text_mask"].clone()),
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
def test_sharded_loss_equals_oracle_on_concatenated_batch():
    from demovlp_b200 import synth
    from oracle import rw_oracle as O
    world = 2
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(world, _free_port(), td), nprocs=world, join=True)
        outs = [torch.load(os.path.join(td, f"rank{r}.pt")) for r in range(world)]
    B, b = 8, 4
    full = synth.make_batch(B, 20, 8, 32, seed=13, kappa=0.2)

    def batch_fn():
        req = lambda t: t.clone().requires_grad_(True)
        return req(full.im), req(full.s), full.im_mask, full.s_len, full.s_mask, req(full.g_text), req(full.g_vid)

    (loss, g, l), grads = O.sharded_reference(batch_fn, world, focal="equal")
    for r, o in enumerate(outs):
        assert torch.allclose(o["loss"], loss, atol=1e-5), (o["loss"], loss)
        assert torch.allclose(o["g"], g, atol=1e-5)
        assert torch.allclose(o["l"], l, atol=1e-5)
        assert torch.allclose(o["d_im"], grads[r]["im"], atol=1e-6, rtol=1e-4)
        assert torch.allclose(o["d_s"], grads[r]["s"], atol=1e-6, rtol=1e-4)
        assert torch.allclose(o["d_gt"], grads[r]["g_text"], atol=1e-6, rtol=1e-4)
        assert torch.allclose(o["d_gv"], grads[r]["g_vid"], atol=1e-6, rtol=1e-4)
        # AllGather_multi semantics
        want = torch.cat([torch.arange(b * 3, dtype=torch.float32).view(b, 3) + 100 * k for k in range(world)])
        assert torch.equal(o["gathered"], want)
        w = torch.arange(want.numel(), dtype=torch.float32).view_as(want)
        assert torch.equal(o["x_grad"], w[r * b:(r + 1) * b])
        # ValidationFeed: the reference's order is batch-major, rank-major inside a batch (trainer_dist.py:252-333)
        order = torch.tensor([100 * bi + 10 * k + i for bi, n in enumerate((3, 2)) for k in range(world) for i in range(n)],
                             dtype=torch.float32)
        assert torch.equal(o["val_rows"], order) and torch.equal(o["val_obj"], order) and torch.equal(o["val_len"], order)
        assert torch.equal(o["val_mask"], torch.zeros(10, 3))          # (1 - 1) * 100
