"""Loads tests/emulate/_build/libtkb_emu.so -- the library's CUDA sources compiled for the CPU emulator (emu_build.py) -- with
the prototypes of the real C ABI (tinyknn_b200._lib.SIGNATURES). "Device" pointers are numpy buffers. TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import emu_build                                                   # noqa: E402

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    from tinyknn_b200._lib import SIGNATURES
    L = ctypes.CDLL(emu_build.build())
    for name, argtypes in SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = {"tkb_last_error": ctypes.c_char_p, "tkb_launch_count": ctypes.c_longlong}.get(name, ctypes.c_int)
    L.emu_last_error.restype = ctypes.c_char_p
    L.emu_stats.argtypes = [ctypes.c_void_p]
    _lib = L
    return L


def ptr(a):
    """Address of a numpy array (None -> NULL). The array must be C-contiguous and must outlive the call."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags.c_contiguous, "emulated device pointers are contiguous numpy arrays"
    return a.ctypes.data


def check(rc):
    if rc != 0:
        L = load()
        raise RuntimeError("emulated C ABI returned %d: %s | %s" % (rc, (L.tkb_last_error() or b"").decode(), (L.emu_last_error() or b"").decode()))


def stats():
    L = load()
    out = np.zeros(7, np.int64)
    L.emu_stats(out.ctypes.data)
    return dict(zip(("launches", "blocks", "fibers", "switches", "collectives", "barriers", "collectives_with_exited_lanes"), out.tolist()))


def aligned(shape, dtype, align=256):
    """numpy array whose data pointer is `align`-byte aligned (cudaMalloc gives 256)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    raw = np.zeros(n + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)
