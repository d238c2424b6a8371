"""ctypes access to oracle/_build/libdarcy_oracle.so (C restatement of the stencil path).
TEST INFRASTRUCTURE ONLY — see oracle/darcy_oracle.c."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libdarcy_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.pdes_oracle_darcy.restype = None
        _lib.pdes_oracle_sobel.restype = None
    return _lib


def sobel(img, dir, correct=True, adjoint=False):
    img = np.ascontiguousarray(img, dtype=np.float64)
    H, W = img.shape[-2:]
    n = img.size // (H * W)
    out = np.empty_like(img)
    lib().pdes_oracle_sobel(img.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p),
                            ctypes.c_long(n), H, W, int(dir), int(correct), int(adjoint))
    return out


def darcy(K, out, gw=None, use_tb=True, want_grad=True):
    """Returns (loss4 float64[4], dout float64 (B,3,H,W) or None)."""
    out = np.ascontiguousarray(out, dtype=np.float32)
    B, C, H, W = out.shape
    assert C == 3
    Kp = None
    if K is not None:
        K = np.ascontiguousarray(K, dtype=np.float32)
        Kp = K.ctypes.data_as(ctypes.c_void_p)
    gw = np.ascontiguousarray(gw if gw is not None else np.ones(4), dtype=np.float64)
    loss4 = np.zeros(4, np.float64)
    dout = np.zeros(out.shape, np.float64) if want_grad else None
    lib().pdes_oracle_darcy(Kp, out.ctypes.data_as(ctypes.c_void_p), B, H, W, int(use_tb),
                            gw.ctypes.data_as(ctypes.c_void_p), loss4.ctypes.data_as(ctypes.c_void_p),
                            dout.ctypes.data_as(ctypes.c_void_p) if want_grad else None)
    return loss4, dout
