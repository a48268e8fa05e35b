"""TEST INFRASTRUCTURE — python wrapper of the plain-C MSDeformAttn oracle (oracle/msda_ref.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpsalm_oracle.so")


def _lib():
    src = os.path.join(_HERE, "msda_ref.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return ctypes.CDLL(_SO)


def msda_ref(value, shapes, loc, aw, dtype=np.float64):
    """value [B,S,M,D], shapes [L,2] (H,W), loc [B,Lq,M,L,P,2], aw [B,Lq,M,L,P] -> [B,Lq,M*D]."""
    value = np.ascontiguousarray(value, dtype=dtype)
    loc = np.ascontiguousarray(loc, dtype=dtype)
    aw = np.ascontiguousarray(aw, dtype=dtype)
    shapes = np.ascontiguousarray(shapes, dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(shapes[:, 0] * shapes[:, 1])[:-1]]).astype(np.int64)
    B, S, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    out = np.zeros((B, Lq, M * D), dtype=dtype)
    fn = getattr(_lib(), "msda_ref_f64" if dtype == np.float64 else "msda_ref_f32")
    vp = ctypes.c_void_p
    fn.argtypes = [vp] * 6 + [ctypes.c_int] * 7
    fn.restype = None
    fn(value.ctypes.data, shapes.ctypes.data, starts.ctypes.data, loc.ctypes.data, aw.ctypes.data,
       out.ctypes.data, B, S, M, D, L, Lq, P)
    return out
