"""TEST-ONLY: lets the CPU test-suite run the package's HOST code (autograd Functions, renderer, plans, feeder ...) on CPU
tensors by pointing the ctypes binding at tests/emu/_build/libkemu.so -- the product kernels compiled for the host (see
tests/kemu.py).  The product itself has no CPU path: without this monkeypatching `_lib.ptr` raises on CPU tensors and
`_lib.load()` loads the CUDA library.  Everything is undone when the pytest `monkeypatch` fixture goes out of scope.

Not emulated: the fused MLP kernels (tensor cores) -- NeRFNetwork takes its plain-torch fp32 op sequence outside autocast,
which is what these tests use -- and anything with CUDA streams / graphs (TrainStep's prefetch stream, capture())."""
import ctypes
import importlib
import pkgutil

import torch

import trinerflet_b200
from tests import kemu
from trinerflet_b200 import _lib


def _ptr(t):
    """host pointer of a CPU tensor (same contract as on the device: dense storage in the layout the ABI expects)"""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return None


def install(monkeypatch):
    so = kemu.lib()
    monkeypatch.setattr(_lib, "_lib", so)
    real_ptr, real_stream = _lib.ptr, _lib.stream
    mods = [_lib] + [importlib.import_module(m.name) for m in pkgutil.iter_modules(trinerflet_b200.__path__, "trinerflet_b200.")
                     if not m.ispkg]
    for mod in mods:
        if getattr(mod, "ptr", None) is real_ptr:
            monkeypatch.setattr(mod, "ptr", _ptr)
        if getattr(mod, "stream", None) is real_stream:
            monkeypatch.setattr(mod, "stream", _stream)
    from trinerflet_b200 import optim, raymarching, triplane_encoder
    monkeypatch.setattr(raymarching, "_cuda_f32", lambda t: t.contiguous().float())
    monkeypatch.setattr(triplane_encoder, "_require_cuda_f32", lambda t, what: None)
    monkeypatch.setattr(optim, "_require_cuda_f32", lambda p: None)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)      # wrappers that move stray CPU inputs
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: False)   # needs a driver otherwise
    return so
