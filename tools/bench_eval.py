"""Timings of the BASELINE configs that are not bench.py's headline line (one JSON line each, GPU only):

  c4   MSRVTT 1k-A evaluation (trainer/trainer_dist.py:358-391): get_sim_by_segment on CPU tensors ->
       float64 (1000, 1000) matrix, + sim_matrix(text, video), t2v / v2t ranks and metrics.
  c4r  the same at the shipped configs' real shape (R = 240, W = 99: generic fp32 kernels).
  c5   DiDeMo paragraph shape (R = 288, W = 64, per-GPU B = 64): fwd + bwd of the loss on the generic fp32 kernels
       (no tcgen05 tiles for this shape yet).

    python tools/bench_eval.py [c4] [c4r] [c5]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from demovlp_b200 import _lib, loss as L, metric as M, synth


def ev(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2]


def c4(R, W, name):
    N = 1000
    b = synth.make_batch(N, R, W, 256, seed=3)
    mod = L.GlobalLocalLoss(focal_type="equal").local_loss
    im, s = b.im.to(torch.bfloat16), b.s.to(torch.bfloat16)      # CPU tensors, as the trainer holds them (:358-368)
    out = {}

    def build():
        local = mod.get_sim_by_segment(im, s, b.im_mask, b.s_len, b.s_mask, device="cuda")
        glob = L.sim_matrix(b.g_text.cuda(), b.g_vid.cuda()).double().cpu().numpy()
        out["sims"] = glob + local                                # orientation quirk of :386 kept
        return out["sims"]

    t_build = ev(build, n=3, warm=1)
    sims = out["sims"]
    t_rank = ev(lambda: (M.t2v_metrics(sims), M.v2t_metrics(sims)), n=5, warm=1)
    dev_im, dev_s, dm, ds = im.cuda(), s.cuda(), b.im_mask.cuda(), b.s_mask.cuda()
    with torch.no_grad():
        t_kernel = ev(lambda: mod.get_sim(dev_im, dev_s, dm, None, ds), n=5, warm=2)
    path = "tcgen05" if _lib.load().dvlp_rw_tcgen05_supported(R, W, 256) else "generic fp32 CUDA cores"
    print(json.dumps({"workload": name, "N": N, "regions": R, "words": W, "kernel_path": path,
                      "eval_matrix_ms_host_in_host_out": t_build * 1e3,
                      "pair_scores_per_s_end_to_end": N * N / t_build,
                      "scores_ms_device_resident": t_kernel * 1e3,
                      "pair_scores_per_s_device_resident": N * N / t_kernel,
                      "ranks_and_metrics_ms": t_rank * 1e3,
                      "t2v": M.t2v_metrics(sims), "timing": "wall clock around synchronised calls, median"}))


def c5():
    B, R, W = 64, 288, 64
    b = synth.make_batch(B, R, W, 256, seed=4).to("cuda")
    mod = L.GlobalLocalLoss(focal_type="equal")

    def step():
        im = b.im.to(torch.bfloat16).requires_grad_(True)
        s = b.s.to(torch.bfloat16).requires_grad_(True)
        gt = b.g_text.clone().requires_grad_(True)
        gv = b.g_vid.clone().requires_grad_(True)
        loss, _, _ = mod(L.sim_matrix(gt, gv), im, s, b.im_mask, b.s_len, b.s_mask)
        loss.backward()

    t = ev(step, n=5, warm=2)
    path = "tcgen05" if _lib.load().dvlp_rw_tcgen05_supported(R, W, 256) else "generic fp32 CUDA cores"
    print(json.dumps({"workload": "c5 per-GPU shard (B=64, R=288, W=64), fwd+bwd, 1 GPU", "kernel_path": path,
                      "ms_per_step": t * 1e3, "pair_scores_per_s": B * B / t,
                      "algorithmic_tflops": 18 * R * W * 256 * B * B / t / 1e12}))


if __name__ == "__main__":
    which = sys.argv[1:] or ["c4", "c4r", "c5"]
    if "c4" in which:
        c4(80, 32, "c4: 1000 x 1000 eval matrix + ranks, R=80, W=32")
    if "c4r" in which:
        c4(240, 99, "c4 real shape: 1000 x 1000 eval matrix + ranks, R=240, W=99")
    if "c5" in which:
        c5()
