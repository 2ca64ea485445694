"""TEST-ONLY front end of tests/emu/_build/libkemu.so: the product kernels compiled for the host (tests/emu/gen_kemu.py,
tests/emu/cuda_shim.h) behind the SAME C ABI as libtrinerflet_b200.so, taking numpy arrays instead of device pointers."""
import ctypes
import importlib.util
import os

import numpy as np

from trinerflet_b200 import _lib as product_lib

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


def lib():
    global _lib
    if _lib is None:
        spec = importlib.util.spec_from_file_location("gen_kemu", os.path.join(_ROOT, "tests", "emu", "gen_kemu.py"))
        gen = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(gen)
        so = ctypes.CDLL(gen.build())
        for name, (res, args) in product_lib.SIGNATURES.items():   # same table as the product binding
            fn = getattr(so, name, None)
            if fn is not None:
                fn.restype, fn.argtypes = res, args
        _lib = so
    return _lib


def p(a):
    """host pointer of a C-contiguous numpy array (None -> NULL)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "emulated kernels take contiguous arrays"
    return ctypes.c_void_p(a.ctypes.data)


def call(name, *args):
    rc = getattr(lib(), name)(*[p(a) if isinstance(a, np.ndarray) else a for a in args])
    if rc != 0:
        raise RuntimeError(f"{name} (emulated) rc={rc}: {lib().tnl_last_error().decode()}")


def f32(a):
    return np.ascontiguousarray(a, np.float32)
