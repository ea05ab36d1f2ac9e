"""ctypes binding of the C ABI in include/demovlp_b200.h (libdemovlp_b200.so).

There is no CPU fallback: if the shared library is missing or no sm_100 GPU is
present, every op raises.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

F32, BF16 = 0, 1
IMPL_AUTO, IMPL_GENERIC, IMPL_TCGEN05 = 0, 1, 2
_IMPL_NAMES = {"auto": IMPL_AUTO, "generic": IMPL_GENERIC, "tcgen05": IMPL_TCGEN05}

_HERE = os.path.dirname(os.path.abspath(__file__))
# DVLP_LIB: a differently built copy of the same library (kernel experiments, tools/variants.sh)
LIB_PATH = os.environ.get("DVLP_LIB") or os.path.join(_HERE, "lib", "libdemovlp_b200.so")

# every symbol include/demovlp_b200.h declares: (restype, argtypes)
_p, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
SYMBOLS = {
    "dvlp_version": (_i, []),
    "dvlp_error_string": (C.c_char_p, [_i]),
    "dvlp_last_cuda_error": (_i, []),
    "dvlp_rw_tcgen05_supported": (_i, [_i, _i, _i]),
    "dvlp_rw_path": (_i, [_i, _i, _i, _i]),
    "dvlp_set_default_impl": (None, [_i]),
    "dvlp_rw_bwd_out_dtype_supported": (_i, [_i, _i, _i, _i, _f, _i]),
    "dvlp_debug_set_trace": (None, [_p]),
    "dvlp_launch_count": (C.c_uint64, []),
    "dvlp_profile_enable": (None, [_i]),
    "dvlp_profile_read": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "dvlp_rw_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "dvlp_rw_scores_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _i, _p, _p, _sz, _p]),
    "dvlp_rw_infer_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "dvlp_rw_scores_infer": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _i, _p, _p, _sz, _p]),
    "dvlp_rw_grouped_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "dvlp_rw_scores_grouped_infer": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _p, _p, _sz, _p]),
    "dvlp_rw_scores_bwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _f, _i, _i, _i, _p, _p, _p, _i, _p, _sz, _p]),
    "dvlp_rwa_loss_fwd_bwd": (_i, [_p, _i, _i, _i, _i, _f, _p, _p, _p]),
    "dvlp_sim_matrix_fwd": (_i, [_p, _p, _i, _i, _i, _f, _i, _p, _p, _p]),
    "dvlp_sim_matrix_bwd": (_i, [_p, _p, _i, _i, _i, _f, _i, _p, _p, _p, _i, _p, _p]),
    "dvlp_normsoftmax_fwd_bwd": (_i, [_p, _i, _i, _f, _p, _p, _p, _p]),
    "dvlp_proj_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "dvlp_proj_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p, _i, _p, _p, _sz, _p]),
    "dvlp_attn_small_supported": (_i, [_i, _i, _i]),
    "dvlp_attn_small_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "dvlp_attn_small_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p]),
    "dvlp_ranks_t2v": (_i, [_p, _i, _i, _p, _p]),
    "dvlp_ranks_v2t": (_i, [_p, _i, _i, _p, _p]),
    "dvlp_p2p_control_bytes": (_sz, []),
    "dvlp_p2p_check": (_i, [_p, _p]),
    "dvlp_p2p_allgather": (_i, [_p, _i, _i, _i, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), _p]),
    "dvlp_p2p_sum_scalar": (_i, [_p, _i, _i, _p, _p, _p]),
    "dvlp_p2p_reduce_scatter": (_i, [_p, _i, _i, _sz, _sz, _p, _i, _p]),
}

_lib = None


def load():
    """Load the shared library (once) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"demovlp_b200: CUDA library not built ({LIB_PATH}); run `make` or "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        lib = load()
        msg = lib.dvlp_error_string(rc).decode()
        extra = ""
        if rc == -6:
            extra = f" (cudaError {lib.dvlp_last_cuda_error()})"
        raise RuntimeError(f"demovlp_b200.{what} failed: {msg}{extra}")


def impl_code(impl) -> int:
    if isinstance(impl, int):
        return impl
    return _IMPL_NAMES[impl]


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("demovlp_b200 ops need CUDA tensors on a B200 (sm_100a); "
                               "there is no CPU fallback")


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def dtype_code(t) -> int:
    if t.dtype == torch.bfloat16:
        return BF16
    if t.dtype == torch.float32:
        return F32
    raise RuntimeError(f"unsupported dtype {t.dtype}")


def launch_count() -> int:
    return int(load().dvlp_launch_count())


def profile_enable(on: bool):
    load().dvlp_profile_enable(1 if on else 0)


def profile_read():
    """{slot: (total_ms, count)} for fwd / bwd / prep / final / grad_gemm since the last read."""
    ms = (C.c_double * 5)()
    cnt = (C.c_int * 5)()
    check(load().dvlp_profile_read(ms, cnt), "profile_read")
    names = ("fwd", "bwd", "prep", "final", "grad_gemm")
    return {n: (ms[i], cnt[i]) for i, n in enumerate(names)}
