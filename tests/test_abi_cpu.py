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


def test_kernel_path_selection_and_new_entry_points_without_gpu():
    """Which pair kernels serve a shape is host logic (no device needed): small-shape tcgen05 kernels for W = 32 /
    R in {32,48,64,80}, the large-shape streaming kernels for every other d = 256 shape they can tile (BASELINE config 5,
    the reference's shipped R = 30 / 240 x W = 99), fp32 CUDA cores otherwise; plus argument validation of the grouped
    (multiple-choice), projection and peer-check entry points."""
    lib = _lib.load()
    for R in (32, 48, 64, 80):
        assert lib.dvlp_rw_path(R, 32, 256, 1) == 2 and lib.dvlp_rw_path(R, 32, 256, 0) == 2
    for R, W in ((288, 64), (240, 99), (30, 99), (17, 9), (100, 20), (130, 128), (96, 32)):
        assert lib.dvlp_rw_path(R, W, 256, 1) == 3 and lib.dvlp_rw_path(R, W, 256, 0) == 3, (R, W)
        assert lib.dvlp_rw_tcgen05_supported(R, W, 256) == 1
    for R, W, d in ((80, 32, 128), (80, 129, 256), (320, 64, 256), (200, 32, 256)):
        assert lib.dvlp_rw_path(R, W, d, 1) == 0 and lib.dvlp_rw_tcgen05_supported(R, W, d) == 0, (R, W, d)
    # the training workspace of a large shape holds its spills; the forward-only one does not
    big = lib.dvlp_rw_workspace_bytes(64, 512, 288, 64, 256, _lib.BF16, 0)
    small = lib.dvlp_rw_infer_workspace_bytes(64, 512, 288, 64, 256, _lib.BF16, 0)
    assert 0 < small < big and big > 3 * 10 ** 9
    # bf16 gradient output exists only where the small-shape finalisation serves the shape
    assert lib.dvlp_rw_bwd_out_dtype_supported(0, 80, 32, 256, 20.0, _lib.BF16) == 1
    assert lib.dvlp_rw_bwd_out_dtype_supported(0, 288, 64, 256, 20.0, _lib.BF16) == 0
    assert lib.dvlp_rw_bwd_out_dtype_supported(0, 80, 32, 256, 50.0, _lib.BF16) == 0      # lambda beyond the tcgen05 bound
    assert lib.dvlp_rw_bwd_out_dtype_supported(1, 80, 32, 256, 20.0, _lib.F32) == 1
    z, one = C.c_void_p(0), C.c_void_p(256)
    # grouped (multiple-choice) scoring
    assert lib.dvlp_rw_grouped_workspace_bytes(64, 5, 240, 99, 256, _lib.BF16) > 0
    assert lib.dvlp_rw_grouped_workspace_bytes(64, 5, 80, 32, 128, _lib.BF16) == 0
    assert lib.dvlp_rw_scores_grouped_infer(z, z, z, z, 4, 5, 30, 99, 256, 20.0, 1, 1, z, z, 0, z) == -3
    assert lib.dvlp_rw_scores_grouped_infer(one, one, one, one, 4, 0, 30, 99, 256, 20.0, 1, 1, one, one, 0, z) == -1
    assert lib.dvlp_rw_scores_grouped_infer(one, one, one, one, 4, 5, 30, 99, 128, 20.0, 1, 1, one, one, 0, z) == -1
    # projection heads
    assert lib.dvlp_proj_workspace_bytes(100, 768, _lib.F32, 1) >= 100 * 768 * 2
    assert lib.dvlp_proj_workspace_bytes(100, 700, _lib.F32, 1) == 0                        # K not a multiple of 64
    assert lib.dvlp_proj_fwd(z, 0, 100, 768, 256, z, z, 1, z, z, 0, z) == -3
    assert lib.dvlp_proj_fwd(one, 0, 100, 768, 128, one, z, 1, one, one, 0, z) == -1        # only 256 outputs
    assert lib.dvlp_p2p_check(z, z) == -3
    assert b"timed out" in lib.dvlp_error_string(-9)


def test_projection_module_mirrors_reference_parameters():
    """FusedProjection keeps nn.Linear's parameter names / shapes (model/model.py:40-43: Sequential(ReLU, Linear(768, 256)),
    object_transformer.py:381: Linear(768, 256, bias=False)), so reference checkpoints load; CPU tensors are refused."""
    from demovlp_b200.proj import FusedProjection
    t = FusedProjection(768, 256, relu_in=True, bias=True)
    v = FusedProjection(768, 256, relu_in=False, bias=False)
    assert dict((k, tuple(p.shape)) for k, p in t.named_parameters()) == {"weight": (256, 768), "bias": (256,)}
    assert dict((k, tuple(p.shape)) for k, p in v.named_parameters()) == {"weight": (256, 768)}
    ref = torch.nn.Linear(768, 256)
    t.load_state_dict(ref.state_dict())
    with pytest.raises(RuntimeError):
        t(torch.zeros(2, 5, 768))


def test_attention_entry_points_without_gpu():
    lib = _lib.load()
    assert lib.dvlp_attn_small_supported(30, 31, 64) == 1 and lib.dvlp_attn_small_supported(8, 9, 64) == 1
    assert lib.dvlp_attn_small_supported(1, 241, 64) == 2
    assert lib.dvlp_attn_small_supported(65, 31, 64) == 0 and lib.dvlp_attn_small_supported(30, 31, 32) == 0
    z, one = C.c_void_p(0), C.c_void_p(256)
    assert lib.dvlp_attn_small_fwd(z, z, z, z, 4, 30, 31, 64, z, z, z) == -3
    assert lib.dvlp_attn_small_fwd(one, one, one, z, 4, 80, 31, 64, one, z, z) == -1
    from demovlp_b200 import attention
    with pytest.raises(RuntimeError):
        attention.attn_mask(torch.zeros(2, 3, 64), torch.zeros(2, 4, 64), torch.zeros(2, 4, 64))
