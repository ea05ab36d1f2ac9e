"""Multi-GPU (NCCL) check of the sharded loss on the real kernels: 2 ranks, global batch 2*b, against
the single-GPU loss of the concatenated batch (the semantics of trainer/trainer_dist.py:13-31)."""
import os
import socket
import sys
import tempfile

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, b):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from demovlp_b200 import dist as D, synth
    full = synth.make_batch(world * b, 80, 32, 256, seed=2)
    sh = full.rows(rank * b, (rank + 1) * b).to(f"cuda:{rank}")
    im = sh.im.to(torch.bfloat16).requires_grad_(True)
    s = sh.s.to(torch.bfloat16).requires_grad_(True)
    gt = sh.g_text.clone().requires_grad_(True)
    gv = sh.g_vid.clone().requires_grad_(True)
    mod = D.ShardedGlobalLocalLoss(focal_type="equal", coef=1.0)
    loss, g, l = mod((gt, gv), im, s, sh.im_mask, sh.s_len, sh.s_mask)
    loss.backward()
    torch.cuda.synchronize()
    torch.save(dict(loss=loss.detach().cpu(), g=g.detach().cpu(), l=l.detach().cpu(), d_im=im.grad.float().cpu(),
                    d_s=s.grad.float().cpu(), d_gt=gt.grad.cpu(), d_gv=gv.grad.cpu()),
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_loss_two_gpus_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from demovlp_b200 import loss as L, synth
    from tests.util import rel_err
    world, b = 2, 16
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(world, port, td, b), nprocs=world, join=True)
        outs = [torch.load(os.path.join(td, f"rank{r}.pt")) for r in range(world)]
    full = synth.make_batch(world * b, 80, 32, 256, seed=2).to("cuda:0")
    im = full.im.to(torch.bfloat16).requires_grad_(True)
    s = full.s.to(torch.bfloat16).requires_grad_(True)
    gt = full.g_text.clone().requires_grad_(True)
    gv = full.g_vid.clone().requires_grad_(True)
    mod = L.GlobalLocalLoss(focal_type="equal", coef=1.0)
    loss, g, l = mod(L.sim_matrix(gt, gv), im, s, full.im_mask, full.s_len, full.s_mask)
    loss.backward()
    for r, o in enumerate(outs):
        sl = slice(r * b, (r + 1) * b)
        assert abs(o["loss"].item() - loss.item()) < 1e-4 * max(1, abs(loss.item()))
        assert abs(o["l"].item() - l.item()) < 1e-4 * max(1, abs(l.item()))
        assert rel_err(o["d_im"], im.grad.float().cpu()[sl]) < 1e-2
        assert rel_err(o["d_s"], s.grad.float().cpu()[sl]) < 1e-2
        assert rel_err(o["d_gt"], gt.grad.cpu()[sl]) < 1e-4
        assert rel_err(o["d_gv"], gv.grad.cpu()[sl]) < 1e-4


def _worker_transport(rank, world, port, out_dir, b):
    """Same sharded step through both transports: own peer-memory kernels (p2p) and NCCL."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from demovlp_b200 import dist as D, synth
    full = synth.make_batch(world * b, 80, 32, 256, seed=5)
    sh = full.rows(rank * b, (rank + 1) * b).to(f"cuda:{rank}")
    res = {}
    for transport in ("p2p", "nccl"):
        mod = D.ShardedGlobalLocalLoss(focal_type="equal", coef=1.0, transport=transport)
        for it in range(3):      # several steps: the flag epochs of the peer transport advance
            im = sh.im.to(torch.bfloat16).requires_grad_(True)
            s = sh.s.to(torch.bfloat16).requires_grad_(True)
            gt = sh.g_text.clone().requires_grad_(True)
            gv = sh.g_vid.clone().requires_grad_(True)
            loss, g, l = mod((gt, gv), im, s, sh.im_mask, sh.s_len, sh.s_mask)
            loss.backward()
        torch.cuda.synchronize()
        res[transport] = dict(loss=loss.detach().cpu(), l=l.detach().cpu(), d_im=im.grad.float().cpu(),
                              d_s=s.grad.float().cpu(), d_gt=gt.grad.cpu(), d_gv=gv.grad.cpu())
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_peer_memory_transport_matches_nccl():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from tests.util import rel_err
    world, b = 2, 16
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker_transport, args=(world, port, td, b), nprocs=world, join=True)
        outs = [torch.load(os.path.join(td, f"rank{r}.pt")) for r in range(world)]
    for o in outs:
        a, c = o["p2p"], o["nccl"]
        assert abs(a["loss"].item() - c["loss"].item()) < 1e-6 * max(1, abs(c["loss"].item()))
        # same kernels, same inputs -- up to the order of the fp32 red.global adds of the video Gram gradients when
        # the tiles of a video are walked by several CTAs (not bit-reproducible run to run)
        assert rel_err(a["d_im"], c["d_im"]) < 1e-4
        assert rel_err(a["d_s"], c["d_s"]) < 4e-3                   # one bf16 rounding, sums in a different order
        assert torch.equal(a["d_gt"], c["d_gt"]) and torch.equal(a["d_gv"], c["d_gv"])
    # every rank got the same loss bits from the rank-ordered scalar sum
    assert outs[0]["p2p"]["loss"].item() == outs[1]["p2p"]["loss"].item()


def _worker_world1(rank, world, port, out_dir, b):
    """The sharded module on a process group of ONE rank: eager step, CUDA-graph replay, and the same step with the
    global (CLS) branch kept on the caller's stream."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    from demovlp_b200 import dist as D, synth
    from demovlp_b200.graph import GraphedLossStep
    sh = synth.make_batch(b, 80, 32, 256, seed=7).to("cuda:0")
    t = dict(im=sh.im.to(torch.bfloat16), s=sh.s.to(torch.bfloat16), g_text=sh.g_text.clone(), g_vid=sh.g_vid.clone())
    keys = ("im", "s", "g_text", "g_vid")
    res = {}
    for name, side in (("side", "1"), ("inline", "0")):
        os.environ["DVLP_GLOBAL_SIDE_STREAM"] = side
        mod = D.ShardedGlobalLocalLoss(focal_type="equal", coef=1.0)
        fn = lambda x: mod((x["g_text"], x["g_vid"]), x["im"], x["s"], sh.im_mask, sh.s_len, sh.s_mask)
        x = {k: v.detach().clone().requires_grad_(True) for k, v in t.items()}
        out = fn(x)
        out[0].backward()
        torch.cuda.synchronize()
        res[name] = dict(losses=torch.stack([o.detach().float().cpu() for o in out]),
                         **{k: x[k].grad.float().cpu() for k in keys})
        if side == "1":
            gs = GraphedLossStep(fn, t, keys, warmup=2)
            for _ in range(3):
                gs.replay()
            torch.cuda.synchronize()
            res["graph"] = dict(losses=gs.losses.float().cpu(), **{k: gs.grads[k].float().cpu() for k in keys})
    torch.save(res, os.path.join(out_dir, "rank0.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sharded_module_on_one_rank_side_stream_and_graph():
    """Runs on ONE GPU (the driver's box): ShardedGlobalLocalLoss over a world of one rank must equal GlobalLocalLoss,
    with the global branch on its second stream (default), inline, and inside a captured whole-step graph, where the
    two streams become parallel branches."""
    import torch.multiprocessing as mp
    from demovlp_b200 import loss as L, synth
    from tests.util import rel_err
    b = 32
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker_world1, args=(1, port, td, b), nprocs=1, join=True)
        res = torch.load(os.path.join(td, "rank0.pt"))
    full = synth.make_batch(b, 80, 32, 256, seed=7).to("cuda:0")
    im = full.im.to(torch.bfloat16).requires_grad_(True)
    s = full.s.to(torch.bfloat16).requires_grad_(True)
    gt = full.g_text.clone().requires_grad_(True)
    gv = full.g_vid.clone().requires_grad_(True)
    out = L.GlobalLocalLoss(focal_type="equal", coef=1.0)(L.sim_matrix(gt, gv), im, s, full.im_mask, full.s_len, full.s_mask)
    out[0].backward()
    want = dict(losses=torch.stack([o.detach().float().cpu() for o in out]), im=im.grad.float().cpu(),
                s=s.grad.float().cpu(), g_text=gt.grad.float().cpu(), g_vid=gv.grad.float().cpu())
    for name in ("side", "inline", "graph"):
        got = res[name]
        assert (got["losses"] - want["losses"]).abs().max() < 1e-4 * max(1.0, float(want["losses"].abs().max())), name
        for k in ("im", "s"):
            assert rel_err(got[k], want[k]) < 1e-2, (name, k)
        for k in ("g_text", "g_vid"):
            assert rel_err(got[k], want[k]) < 1e-4, (name, k)
    # the stream placement must not change a bit of the deterministic parts
    assert torch.equal(res["side"]["losses"], res["inline"]["losses"])
    assert torch.equal(res["side"]["g_text"], res["inline"]["g_text"])
    assert torch.equal(res["graph"]["losses"], res["side"]["losses"])
