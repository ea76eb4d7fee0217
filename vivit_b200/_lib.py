"""ctypes loader for ``libvivit_b200.so`` (the C ABI declared in ``include/vivit_b200.h``).

There is no CPU fallback: if the shared library is missing, every kernel call
raises.  Build it with ``python __graft_entry__.py build`` (or ``make -C
vivit_b200/csrc``).
"""

from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libvivit_b200.so")

P, I64, DBL, INT = c_void_p, c_int64, c_double, c_int

# name -> (restype, argtypes); mirrors include/vivit_b200.h one to one
SIGNATURES = {
    "vvt_abi_version": (INT, []),
    "vvt_last_error": (c_char_p, []),
    "vvt_launch_count": (I64, []),
    "vvt_loss_sqrt_hessian_ce": (INT, [P, P, P, I64, I64, I64, DBL, INT, P]),
    "vvt_loss_sqrt_hessian_ce_mc": (INT, [P, P, P, P, I64, I64, I64, I64, DBL, INT, P]),
    "vvt_loss_sqrt_hessian_mse": (INT, [P, I64, I64, DBL, INT, P]),
    "vvt_sqrt_backprop_linear": (INT, [P, P, P, I64, I64, I64, INT, P]),
    "vvt_conv2d_workspace_bytes": (I64, [INT] + [I64] * 8 + [INT]),
    "vvt_sqrt_backprop_conv2d": (INT, [P, P, P] + [I64] * 15 + [P, I64, INT, P]),
    "vvt_sqrt_backprop_elementwise": (INT, [P, P, P, I64, I64, INT, DBL, INT, P]),
    "vvt_maxpool2d_argmax": (INT, [P, P] + [I64] * 13 + [INT, P]),
    "vvt_sqrt_backprop_maxpool2d": (INT, [P, P, P] + [I64] * 15 + [INT, P]),
    "vvt_sqrt_backprop_avgpool2d": (INT, [P, P] + [I64] * 12 + [INT, P]),
    "vvt_v_emit_conv2d": (INT, [P, P, P] + [I64] * 16 + [P, I64, INT, P]),
    "vvt_v_emit_bias": (INT, [P, P, I64, I64, I64, INT, P]),
    "vvt_v_emit_linear": (INT, [P, P, P, I64, I64, I64, I64, INT, P]),
    "vvt_gemm": (INT, [P, P, P, I64, I64, I64, INT, INT, I64, I64, I64, DBL, DBL, I64, I64, I64, I64, P, I64, INT, P]),
    "vvt_gram_workspace_bytes": (I64, [I64, I64, I64, INT]),
    "vvt_gram_dense_accum": (INT, [P, P, I64, I64, P, I64, INT, P]),
    "vvt_gram_cross_accum": (INT, [P, P, P, I64, I64, I64, P, I64, INT, P]),
    "vvt_gram_linear_workspace_bytes": (I64, [I64, I64, I64, I64, I64, INT]),
    "vvt_gram_linear_accum": (INT, [P, P, P, I64, I64, I64, I64, INT, P, I64, INT, P]),
    "vvt_gram_cross_linear_accum": (INT, [P, P, P, P, P, I64, I64, I64, I64, I64, INT, P, I64, INT, P]),
    "vvt_scale": (INT, [P, I64, DBL, INT, P]),
    "vvt_axpy": (INT, [P, P, I64, DBL, INT, P]),
    "vvt_nccl_available": (INT, []),
    "vvt_nccl_allreduce_gram": (INT, [P, P, I64, P, I64, DBL, INT, P]),
    "vvt_center_rows": (INT, [P, P, I64, I64, INT, P]),
    "vvt_syevj_workspace_bytes": (I64, [I64, INT, INT]),
    "vvt_syevj": (INT, [P, P, P, I64, INT, P, I64, POINTER(c_int), INT, P]),
    "vvt_syevj_batched_workspace_bytes": (I64, [I64, I64, INT, INT]),
    "vvt_syevj_batched": (INT, [P, P, P, I64, I64, INT, P, I64, POINTER(c_int), INT, P]),
    "vvt_syevj_dist_workspace_bytes": (I64, [I64, INT, INT, INT]),
    "vvt_dist_arena_bytes_for": (I64, [I64]),
    "vvt_dist_arena_bytes": (I64, []),
    "vvt_dist_arena_alloc": (INT, [I64, P]),
    "vvt_dist_arena_open": (INT, [P, INT, INT]),
    "vvt_dist_arena_close_peers": (INT, []),
    "vvt_dist_arena_free": (INT, []),
    "vvt_dbg_dist_plan": (INT, [INT, INT, INT, POINTER(c_int), POINTER(c_int), INT, POINTER(c_int), POINTER(c_int)]),
    "vvt_syevj_dist": (INT, [P, P, P, P, I64, INT, P, I64, POINTER(c_int), INT, INT, P]),
    "vvt_dbg_wide_round": (INT, [P, P, P, P, I64, INT, P]),
    "vvt_filter_nonzero": (INT, [P, P, I64, DBL, DBL, POINTER(I64), INT, P]),
    "vvt_backtransform_dense": (INT, [P, P, P, P, I64, I64, I64, INT, P]),
    "vvt_backtransform_linear": (INT, [P, P, P, P, P, P, I64, I64, I64, I64, I64, P, I64, INT, P]),
    "vvt_vt_mat_prod_linear": (INT, [P, P, P, P, I64, I64, I64, I64, I64, P, I64, INT, P]),
    "vvt_scale_rows_rsqrt": (INT, [P, P, I64, I64, INT, P]),
    "vvt_dirderiv_epilogue": (INT, [P, P, P, P, P, P, I64, I64, I64, I64, I64, P, I64, INT, P]),
    "vvt_newton_coeff": (INT, [P, P, P, P, P, P, I64, I64, I64, I64, DBL, INT, P]),
    "vvt_v_apply_dense": (INT, [P, P, P, I64, I64, INT, P]),
    "vvt_v_apply_linear": (INT, [P, P, P, P, P, I64, I64, I64, I64, P, I64, INT, P]),
}

_lib = None


class KernelLibraryError(RuntimeError):
    """The CUDA kernel library is missing, stale, or a kernel call failed."""


def load() -> ctypes.CDLL:
    """Load the shared library once; raise if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KernelLibraryError(
            f"{LIB_PATH} not found. vivit_b200 has no CPU or PyTorch fallback: build the "
            "sm_100a kernels first (python __graft_entry__.py build)."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise KernelLibraryError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().vvt_last_error()
        raise KernelLibraryError(f"{what} failed with status {status}: {msg.decode() if msg else ''}")
