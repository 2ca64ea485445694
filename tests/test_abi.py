"""CPU: the C-ABI shared library loads without a GPU, exports every symbol include/trinerflet_b200.h declares, and the
ctypes table covers exactly that set (no compute calls here)."""
import ctypes
import os
import re

from trinerflet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "trinerflet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return set(re.findall(r"\b(tnl_[a-z0-9_]+)\s*\(", src)) - {"tnl_stream_t"}


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 26
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), n
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)


def test_version_and_argument_errors_need_no_gpu():
    lib = _lib.load()
    assert lib.tnl_abi_version() == 1
    assert lib.tnl_morton3d(None, 5, None, None) == -1            # TNL_ERR_INVALID_ARGUMENT, no launch attempted
    assert b"null pointer" in lib.tnl_last_error()
    dims = _lib.MlpDims(50, 64, 64)
    assert lib.tnl_mlp_packed_bytes(ctypes.byref(dims)) == 0      # unsupported width
    dims = _lib.MlpDims(96, 64, 64)
    legacy = 2 * 2 * (96 * 64 + 64 * 16 + 32 * 64 + 64 * 64) + 2 * (8 + 16) * 64      # mma.sync fragment layout (fwd + dX)
    tc = 2 * (64 * 96 + 16 * 64 + 64 * 32 + 64 * 64 + 16 * 64)                         # tcgen05 operand tiles
    assert lib.tnl_mlp_packed_bytes(ctypes.byref(dims)) == (legacy + 255) // 256 * 256 + tc
    assert lib.tnl_march_rays_train_workspace(60000) >= 4 * 59
    assert lib.tnl_idwt_level_forward(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 12, 32, None, None) == -1


def test_no_cpu_fallback():
    import pytest
    import torch
    from trinerflet_b200 import raymarching
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(3))
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            raymarching.morton3D(torch.zeros(4, 3, dtype=torch.int32))
