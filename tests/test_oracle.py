"""The oracle restatement vs. the reference: committed fixtures everywhere,
the live reference modules when /root/reference is mounted (build container)."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader, rw_oracle as O
from tests.util import batch_for, golden_names, load_golden, rel_err

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))


@pytest.mark.parametrize("name", golden_names("loss_"))
def test_oracle_loss_matches_reference_fixture(name):
    g = load_golden("loss_" + name)
    b, focal = batch_for(g)
    if b.im.shape[0] * b.im.shape[1] * b.s.shape[1] > 32 * 80 * 32 and b.im.shape[0] > 8:
        pytest.skip("kept small for the CPU suite")
    im = b.im.clone().requires_grad_(True)
    s = b.s.clone().requires_grad_(True)
    gt = b.g_text.clone().requires_grad_(True)
    gv = b.g_vid.clone().requires_grad_(True)
    gsim = O.sim_matrix(gt, gv)
    gsim.retain_grad()
    loss, gl, ll = O.global_local_loss(gsim, im, s, b.im_mask, b.s_len, b.s_mask, focal=focal)
    loss.backward()
    with torch.no_grad():
        scores = O.pair_scores(b.im, b.s, b.im_mask, b.s_mask, focal=focal)
    # same ops in the same precision: agreement to reduction-order noise
    np.testing.assert_allclose(scores.numpy(), g["scores"], atol=5e-6)
    np.testing.assert_allclose([loss.item(), gl.item(), ll.item()], g["losses"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(gsim.detach().numpy(), g["global_sim"], atol=1e-6)
    assert rel_err(gsim.grad, g["d_global_sim"]) < 1e-5
    if "d_im" in g:
        assert rel_err(im.grad, g["d_im"]) < 2e-4
        assert rel_err(s.grad, g["d_s"]) < 2e-4
        assert rel_err(gt.grad, g["d_g_text"]) < 1e-5
        assert rel_err(gv.grad, g["d_g_vid"]) < 1e-5


@pytest.mark.parametrize("name", golden_names("scores_"))
def test_oracle_scores_match_reference_fixture(name):
    g = load_golden("scores_" + name)
    b, focal = batch_for(g)
    scores = O.pair_scores(b.im, b.s, b.im_mask, b.s_mask, focal=focal)
    assert scores.shape == g["scores"].shape
    np.testing.assert_allclose(scores.numpy(), g["scores"], atol=5e-6)


def test_oracle_eval_matrix_and_metrics_fixture():
    g = load_golden("eval_96")
    from demovlp_b200 import synth
    b = synth.make_batch(96, 80, 32, 256, 3, 0.12)
    local = O.scores_by_segment(b.im, b.s, b.im_mask, b.s_mask)
    assert local.dtype == np.float64 and local.shape == (96, 96)
    np.testing.assert_allclose(local, g["local"], atol=5e-6)
    # ranks / metrics are integer work: bit-exact on the reference-built matrix
    full = g["full"]
    np.testing.assert_array_equal(O.t2v_ranks(full), g["t2v_ranks"])
    keys = [str(k) for k in g["keys"]]
    t2v, v2t = O.t2v_metrics(full), O.v2t_metrics(full)
    for k, a, c in zip(keys, g["t2v"], g["v2t"]):
        assert t2v[k] == pytest.approx(a, rel=1e-12, abs=1e-12), k
        assert v2t[k] == pytest.approx(c, rel=1e-12, abs=1e-12), k


def test_oracle_metrics_on_crafted_matrices():
    from make_golden import metric_mats
    g = load_golden("metrics")
    keys = [str(k) for k in g["keys"]]
    for name, m in metric_mats().items():
        assert m.sum() == g[name + "_sum"][0]
        t2v, v2t = O.t2v_metrics(m), O.v2t_metrics(m)
        for k, a, c in zip(keys, g[name + "_t2v"], g[name + "_v2t"]):
            assert t2v[k] == pytest.approx(a, rel=1e-12, abs=1e-12), (name, k)
            assert v2t[k] == pytest.approx(c, rel=1e-12, abs=1e-12), (name, k)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_vs_live_reference_fp64():
    """fp64 on both sides: the restatement is the same function, not merely close."""
    from demovlp_b200 import synth
    L = ref_loader.loss_module()
    for focal in ("equal", "prob"):
        b = synth.make_batch(6, 24, 11, 64, 21, 0.2, Bt=6)
        im = b.im.double().requires_grad_(True); s = b.s.double().requires_grad_(True)
        im2 = b.im.double().requires_grad_(True); s2 = b.s.double().requires_grad_(True)
        gs = torch.randn(6, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(1)).clamp(-1, 1)
        g1 = gs.clone().requires_grad_(True); g2 = gs.clone().requires_grad_(True)
        ref = L.GlobalLocalLoss(focal_type=focal)(g1, im, s, b.im_mask.double(), b.s_len, b.s_mask.double())
        mine = O.global_local_loss(g2, im2, s2, b.im_mask.double(), b.s_len, b.s_mask.double(), focal=focal)
        ref[0].backward(); mine[0].backward()
        for a, c in zip(ref, mine):
            assert abs(a.item() - c.item()) < 1e-12
        assert rel_err(im2.grad, im.grad) < 1e-10
        assert rel_err(s2.grad, s.grad) < 1e-10
        assert rel_err(g2.grad, g1.grad) < 1e-12


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")
def test_oracle_sim_matrix_and_metrics_vs_live_reference():
    sm = ref_loader.sim_matrix_fn()
    M = ref_loader.metric_module()
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(13, 32, generator=g), torch.randn(9, 32, generator=g)
    a[2] = 0
    assert torch.equal(sm(a, b), O.sim_matrix(a, b))
    rng = np.random.default_rng(5)
    for m in (rng.standard_normal((40, 40)), np.round(rng.standard_normal((60, 12)), 1)):
        r, o = M.t2v_metrics(m), O.t2v_metrics(m)
        r2, o2 = M.v2t_metrics(m), O.v2t_metrics(m)
        for k in r:
            assert float(r[k]) == pytest.approx(o[k], rel=1e-12)
            assert float(r2[k]) == pytest.approx(o2[k], rel=1e-12)


def test_oracle_port_matches_config3_fixture_rows():
    """The restatement against the config-3 fixture (unmodified reference in 16-row blocks, B = 1024): the fully stored
    score rows, i.e. one video of every 128-row shard against all 1024 texts."""
    from demovlp_b200 import synth
    from tests import c3_check
    if not c3_check.available():
        pytest.skip("tests/golden/c3_global.npz missing")
    g = c3_check.load()
    B, R, W, d, seed = (int(x) for x in g["meta"][:5])
    b = synth.make_batch(B, R, W, d, seed, float(g["meta"][5]))
    keep = [int(k) for k in g["keep_rows"]]
    sc = O.pair_scores(b.im[keep], b.s, b.im_mask[keep], b.s_mask, 20.0, "equal")
    assert np.abs(sc.numpy() - g["score_rows"]).max() < 5e-6
    assert np.abs(sc.numpy()[np.arange(len(keep)), keep] - g["score_diag"][keep]).max() < 5e-6
