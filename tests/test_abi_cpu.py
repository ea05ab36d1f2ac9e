"""CPU-side checks of the C-ABI boundary: the library loads, exports exactly what the
header declares, validates arguments, and refuses to compute without a GPU."""
import ctypes as C
import inspect
import os
import re

import pytest
import torch

from demovlp_b200 import _lib, loss, metric, ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "demovlp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dvlp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert names, "no symbols parsed from the header"
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/demovlp_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    assert lib.dvlp_version() >= 100


def test_error_strings_and_validation_without_gpu():
    lib = _lib.load()
    assert b"no CPU fallback" in lib.dvlp_error_string(-7)
    n = lib.dvlp_rw_workspace_bytes(128, 128, 80, 32, 256, _lib.BF16, _lib.IMPL_GENERIC)
    assert n > 0 and n % 256 == 0
    assert lib.dvlp_rw_workspace_bytes(0, 128, 80, 32, 256, _lib.BF16, 0) == 0
    z = C.c_void_p(0)
    one = C.c_void_p(256)
    # null pointers / bad shapes are rejected before any CUDA call
    assert lib.dvlp_rw_scores_fwd(z, z, z, z, 1, 1, 1, 1, 8, 20.0, 1, 0, 1, z, z, 0, z) == -3
    assert lib.dvlp_rw_scores_fwd(one, one, one, one, 0, 1, 1, 1, 8, 20.0, 1, 0, 1, one, one, 0, z) == -1
    assert lib.dvlp_rw_scores_fwd(one, one, one, one, 1, 1, 1, 1, 8, 20.0, 1, 7, 1, one, one, 0, z) == -2
    assert lib.dvlp_ranks_t2v(z, 4, 4, z, z) == -3
    assert lib.dvlp_ranks_t2v(one, 2, 4, one, z) == -1
    assert lib.dvlp_normsoftmax_fwd_bwd(one, 4, 4, 0.0, one, z, one, z) == -8
    assert lib.dvlp_rw_scores_infer(z, z, z, z, 1, 1, 1, 1, 8, 20.0, 1, 0, 1, z, z, 0, z) == -3
    small = lib.dvlp_rw_infer_workspace_bytes(128, 128, 80, 32, 256, _lib.BF16, _lib.IMPL_TCGEN05)
    assert 0 < small < lib.dvlp_rw_workspace_bytes(128, 128, 80, 32, 256, _lib.BF16, _lib.IMPL_TCGEN05)
    # peer-memory exchanges: argument checks come before any device work
    ctrl = lib.dvlp_p2p_control_bytes()
    assert ctrl >= 1024 and ctrl % 256 == 0
    src = (C.c_void_p * 1)(256)
    nb = (C.c_size_t * 1)(64)
    off = (C.c_size_t * 1)(ctrl)
    assert lib.dvlp_p2p_allgather(z, 2, 0, 1, src, nb, off, z) == -3
    assert lib.dvlp_p2p_allgather(one, 2, 2, 1, src, nb, off, z) == -8            # rank outside the world
    assert lib.dvlp_p2p_allgather(one, 2, 0, 1, src, (C.c_size_t * 1)(40), off, z) == -4   # not a multiple of 16 bytes
    assert lib.dvlp_p2p_allgather(one, 2, 0, 1, src, nb, (C.c_size_t * 1)(0), z) == -8     # inside the control block
    assert lib.dvlp_p2p_sum_scalar(one, 65, 0, one, one, z) == -8
    assert lib.dvlp_p2p_reduce_scatter(one, 2, 0, ctrl, 6, one, _lib.F32, z) == -4          # shard not a multiple of 4 floats
    assert lib.dvlp_p2p_reduce_scatter(one, 2, 0, ctrl, 8, one, 9, z) == -2
    if not torch.cuda.is_available():
        # valid arguments, no device: the library says so instead of computing anything
        assert lib.dvlp_rwa_loss_fwd_bwd(one, 4, 4, 0, 4, 20.0, one, z, z) == -7


def test_ops_refuse_cpu_tensors():
    im, s = torch.randn(2, 4, 8), torch.randn(2, 3, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.rw_scores(im, s, torch.zeros(2, 4), torch.zeros(2, 3))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sim_matrix(torch.randn(3, 8), torch.randn(3, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss.NormSoftmaxLoss()(torch.randn(3, 3))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            metric.t2v_metrics(torch.randn(4, 4).numpy())


def test_host_mirror_keeps_reference_signatures():
    """Constructor / forward parameter names are part of the boundary: the reference builds the
    loss from JSON by keyword (parse_config_dist_multi.py:73-100)."""
    want_ctor = ["temperature", "lambda_softmax", "focal_type", "margin", "max_violation",
                 "use_local", "use_global", "coef"]
    got = list(inspect.signature(loss.GlobalLocalLoss.__init__).parameters)[1:]
    assert got == want_ctor
    sig = inspect.signature(loss.GlobalLocalLoss.__init__).parameters
    assert (sig["temperature"].default, sig["lambda_softmax"].default, sig["focal_type"].default,
            sig["coef"].default) == (0.05, 20, "prob", 1000.)
    assert list(inspect.signature(loss.GlobalLocalLoss.forward).parameters)[1:] == [
        "global_sim", "local_im", "local_s", "local_im_m", "local_s_l", "local_s_m"]
    assert list(inspect.signature(loss.RWALoss.get_sim).parameters)[1:] == ["im", "s", "im_m", "s_l", "s_m"]
    seg = inspect.signature(loss.RWALoss.get_sim_by_segment).parameters
    assert list(seg)[1:8] == ["img_feats", "lang_feats", "img_mask", "lang_length", "cap_mask", "segment", "device"]
    m = loss.GlobalLocalLoss(focal_type="equal", coef=1.0)
    assert isinstance(m.local_loss, loss.RWALoss) and isinstance(m.global_loss, loss.NormSoftmaxLoss)
    assert m.to("cpu") is m          # base_trainer.py:35 calls loss.to(device)
    from oracle import ref_loader
    if ref_loader.available():
        L = ref_loader.loss_module()
        for cls in ("GlobalLocalLoss", "RWALoss", "NormSoftmaxLoss"):
            ref_params = list(inspect.signature(getattr(L, cls).__init__).parameters)
            mine = list(inspect.signature(getattr(loss, cls).__init__).parameters)
            assert mine[:len(ref_params)] == ref_params, cls
