"""Host-side mirror of the reference's ``model/loss.py`` for the region-word path.

Same class names, constructor parameters, ``forward`` signatures and return
conventions as the reference (model/loss.py:10-138), so the reference's
``ConfigParser.initialize('loss', module_loss)`` (parse_config_dist_multi.py:73-100)
and its trainers (trainer/trainer_dist.py:160-165, 335-340, 375-381, 820-826) work
unchanged; all arithmetic runs in the CUDA library behind include/demovlp_b200.h.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import ops


class NormSoftmaxLoss(nn.Module):
    """model/loss.py:119-138."""

    def __init__(self, temperature=0.05):
        super().__init__()
        self.temperature = temperature

    def forward(self, x):
        return ops.norm_softmax_loss(x, self.temperature)


class RWALoss(nn.Module):
    """Region-word alignment loss, model/loss.py:48-116."""

    def __init__(self, lambda_softmax=20, focal_type="prob", margin=0, max_violation=False, impl="auto"):
        super().__init__()
        self.lambda_softmax = lambda_softmax
        self.focal_type = focal_type
        self.margin = margin                 # stored and unused, as in the reference (:61-62)
        self.max_violation = max_violation
        self.impl = impl

    def get_sim(self, im, s, im_m, s_l, s_m):
        """(n_image, n_caption) scores (:64-71).  ``s_l`` is unused, as in the reference (:297)."""
        return ops.rw_scores(im, s, im_m, s_m, lam=self.lambda_softmax,
                             focal_equal=(self.focal_type == "equal"), impl=self.impl)

    def get_sim_grouped(self, im, s, im_m, s_m, n_options):
        """Batched multiple-choice scoring (trainer/trainer_dist.py:792-829): ``im`` (Q, R, d) holds the video of each of
        Q questions ONCE (the reference expands it n_options times), ``s`` (Q * n_options, W, d) the options, question-
        major.  Returns (Q, n_options) = row 0 of the reference's per-question ``local_sims_batch``; the predicted answer
        is ``(global_row + local).argmax(1)`` exactly as ``sims_batch[0].max(0)[1]`` there.  Forward only."""
        with torch.no_grad():
            return ops.rw_scores_grouped(im, s, im_m, s_m, n_options, lam=self.lambda_softmax,
                                         focal_equal=(self.focal_type == "equal"))

    def get_sim_by_segment(self, img_feats, lang_feats, img_mask, lang_length, cap_mask,
                           segment=8, device="cpu", rows_per_launch=256):
        """float64 numpy (n_image, n_caption) eval matrix (:73-103).

        The reference walks 8 x 8 tiles with a host round trip each; here the text side is
        uploaded once and the video side in blocks of ``rows_per_launch`` rows, one fused
        launch per block.  ``segment`` is accepted and ignored.  Needs a CUDA device:
        ``device="cpu"`` (the reference default) selects the current CUDA device.
        """
        if not torch.cuda.is_available():
            raise RuntimeError("demovlp_b200: get_sim_by_segment needs a CUDA device; no CPU fallback")
        dev = torch.device("cuda", torch.cuda.current_device()) if str(device) == "cpu" else torch.device(device)
        n_img, n_txt = img_feats.shape[0], lang_feats.shape[0]
        sim = np.zeros((n_img, n_txt))
        with torch.no_grad():
            la = lang_feats.to(dev)
            lam = cap_mask.to(dev)
            for v0 in range(0, n_img, rows_per_launch):
                v1 = min(v0 + rows_per_launch, n_img)
                blk = self.get_sim(img_feats[v0:v1].to(dev), la, img_mask[v0:v1].to(dev), None, lam)
                sim[v0:v1] = blk.cpu().numpy()
        return sim

    def forward(self, im, s, im_m, s_l, s_m):
        if im.shape[0] != s.shape[0]:        # the reference's eye(n_image) target needs a square matrix (:108)
            raise RuntimeError(f"RWALoss needs n_image == n_caption, got {im.shape[0]} and {s.shape[0]}")
        scores = self.get_sim(im, s, im_m, s_l, s_m)
        return ops.rwa_loss(scores, self.lambda_softmax)


class GlobalLocalLoss(nn.Module):
    """model/loss.py:10-45.  ``coef`` is stored and unused, as in the reference (:27)."""

    def __init__(self, temperature=0.05, lambda_softmax=20, focal_type="prob", margin=0,
                 max_violation=False, use_local=True, use_global=True, coef=1000.):
        super().__init__()
        self.global_loss = NormSoftmaxLoss(temperature)
        self.local_loss = RWALoss(lambda_softmax, focal_type, margin, max_violation)
        self.use_local = use_local
        self.use_global = use_global
        self.cof_local = coef

    def forward(self, global_sim, local_im, local_s, local_im_m, local_s_l, local_s_m):
        if not self.use_local:
            loss = self.global_loss(global_sim)
            return loss, loss, torch.tensor([0.0])
        if not self.use_global:
            loss = self.local_loss(local_im, local_s, local_im_m, local_s_l, local_s_m)
            return loss, torch.tensor([0.0]), loss
        global_loss = self.global_loss(global_sim)
        local_loss = self.local_loss(local_im, local_s, local_im_m, local_s_l, local_s_m)
        return global_loss + local_loss, global_loss, local_loss


def sim_matrix(a, b, eps=1e-8):
    """model/model.py:582-590 (re-exported here; also importable from demovlp_b200.model)."""
    return ops.sim_matrix(a, b, eps)
