"""The C-ABI boundary without a GPU: ``libvivit_b200.so`` loads, exports every entry point that
``include/vivit_b200.h`` declares, the ctypes table mirrors the header one to one, and argument
validation (which runs before any CUDA call) reports errors through the status code."""

import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vivit_b200.h")


def declared_functions():
    """``{name: number of parameters}`` of every prototype in the header."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    out = {}
    for m in re.finditer(r"\b(?:int|int64_t|const char\*)\s+(vvt_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    return out


@pytest.fixture(scope="module")
def lib():
    from vivit_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def test_header_declares_the_expected_surface():
    names = declared_functions()
    assert len(names) >= 30
    for required in ("vvt_gram_linear_accum", "vvt_gram_dense_accum", "vvt_syevj", "vvt_backtransform_dense",
                     "vvt_dirderiv_epilogue", "vvt_newton_coeff", "vvt_loss_sqrt_hessian_ce"):
        assert required in names


def test_library_exports_every_declared_symbol(lib):
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/vivit_b200.h but not exported"


def test_ctypes_table_mirrors_header(lib):
    from vivit_b200 import _lib

    decl = declared_functions()
    assert set(decl) == set(_lib.SIGNATURES), set(decl) ^ set(_lib.SIGNATURES)
    for name, nargs in decl.items():
        assert len(_lib.SIGNATURES[name][1]) == nargs, name


def test_no_unexpected_exports(lib):
    from vivit_b200 import _lib

    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    vvt = {s for s in exported if s.startswith("vvt_")}
    assert vvt == set(declared_functions())


def test_argument_validation_needs_no_gpu(lib):
    assert lib.vvt_abi_version() >= 1
    # negative size -> VVT_ERR_INVALID (1), message available
    st = lib.vvt_gram_dense_accum(None, None, -1, 4, None, 0, 0, None)
    assert st == 1 and b"negative" in lib.vvt_last_error()
    # empty problems are a no-op success
    assert lib.vvt_gram_dense_accum(None, None, 0, 4, None, 0, 0, None) == 0
    assert lib.vvt_scale(None, 0, 2.0, 0, None) == 0
    # null pointers / unknown dtype are refused before any launch
    assert lib.vvt_scale(None, 8, 2.0, 0, None) == 1
    buf = ctypes.create_string_buffer(64)
    assert lib.vvt_scale(ctypes.cast(buf, ctypes.c_void_p), 8, 2.0, 7, None) == 1
    assert lib.vvt_syevj_workspace_bytes(0, 1, 0) == 0
    assert lib.vvt_syevj_workspace_bytes(1280, 1, 0) > 1280 * 1280 * 4
    assert lib.vvt_launch_count() == 0


def test_product_refuses_to_run_without_cuda():
    import torch

    from vivit_b200 import _lib, kernels

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_lib.KernelLibraryError):
        kernels.gram_dense_accum(torch.zeros(4, 4), torch.zeros(4, 8))
