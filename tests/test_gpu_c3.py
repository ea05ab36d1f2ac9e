"""BASELINE config 3 (global B = 1024, 8x10 regions, 32 words) at FULL size on ONE GPU, against the fixture minted
from the unmodified reference in 16-row blocks (oracle/make_golden_c3.py).

The eight 128-row shards are walked sequentially through exactly the calls a rank of the sharded loss makes
(demovlp_b200/dist.py: scores forward on (b x B), RWALoss tail with its row offset, backward), the text gradients
of the shards are summed like the reduce-scatter does, and everything is compared with what the reference computes
on the concatenated batch (trainer/trainer_dist.py:13-31 semantics).  Runs on the driver's single-GPU box, so the
multi-GPU arithmetic has an oracle-backed check that is always executed."""
import numpy as np
import pytest
import torch

from demovlp_b200 import loss as L, ops, synth
from tests import c3_check
from tests.util import checksum

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shards", [8, 2])
def test_config3_shard_walk_matches_reference_fixture(shards):
    assert c3_check.available(), "tests/golden/c3_global.npz missing (python oracle/make_golden_c3.py)"
    g = c3_check.load()
    B, R, W, d, seed = (int(x) for x in g["meta"][:5])
    b = synth.make_batch(B, R, W, d, seed, float(g["meta"][5]))
    np.testing.assert_allclose(checksum(b), g["checksum"], rtol=0, atol=1e-6)
    dv = b.to("cuda")
    im, s = dv.im.to(torch.bfloat16), dv.s.to(torch.bfloat16)
    bs = B // shards
    local = 0.0
    d_s_sum = torch.zeros(B, W, d, dtype=torch.float32, device="cuda")
    d_im_all, score_rows = [], []
    for r in range(shards):
        lo, hi = r * bs, (r + 1) * bs
        scores, state = ops.rw_scores_fwd_raw(im[lo:hi].contiguous(), s, dv.im_mask[lo:hi].contiguous(), dv.s_mask,
                                              20.0, 1, "auto")
        part, d_scores = ops.rwa_tail_raw(scores, 20.0, lo, B)
        d_im, d_s = ops.rw_scores_bwd_raw(state, d_scores)           # fp32 partial gradients for ALL texts
        local += float(part)
        d_s_sum += d_s.float()
        d_im_all.append(d_im.float().cpu())
        score_rows.append(scores.cpu())
        del state
    # the CLS path is replicated on every rank: whole 1024 x 1024 cosine matrix
    gt = dv.g_text.clone().requires_grad_(True)
    gv = dv.g_vid.clone().requires_grad_(True)
    glob = L.NormSoftmaxLoss(0.05)(L.sim_matrix(gt, gv))
    glob.backward()
    losses = [float(glob) + local, float(glob), local]
    ok, rep = c3_check.check_rows(g, 0, B, losses=losses, scores=torch.cat(score_rows), d_im=torch.cat(d_im_all),
                                  d_s=d_s_sum.cpu(), d_g_text=gt.grad.cpu(), d_g_vid=gv.grad.cpu())
    print(f"[c3 shard walk x{shards}] losses {losses} want {g['losses'].tolist()}")
    for k, v in rep.items():
        print(f"   {k}: err {v['err']:.3e} tol {v['tol']:.1e} {'ok' if v['ok'] else 'FAIL'}")
    assert ok, rep
