"""ctypes binding of oracle/libquanta_oracle.so (the C restatement).

TEST INFRASTRUCTURE ONLY — see oracle/quanta_oracle.c.  Used where the numpy
oracle would be too slow (full-size parity checks) and as the timed CPU
baseline (``cpu_baseline.kind == "port"``) in bench.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libquanta_oracle.so")
_lib = None

MODE_TENSOR, MODE_DIM0, MODE_BLOCK = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "quanta_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libquanta_oracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.qo_num_threads.restype = C.c_int
        _lib.qo_backend_quantize.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def num_threads():
    return int(lib().qo_num_threads())


def set_num_threads(n):
    lib().qo_set_num_threads(C.c_int(int(n)))


def quantize_affine(x, bits=8, mode=MODE_TENSOR, block=0):
    x = _f32(x)
    n = x.size
    q = np.empty(x.shape, np.uint8)
    L = lib()
    if mode == MODE_TENSOR:
        s, z = np.empty((), np.float32), np.empty((), np.float32)
        L.qo_quantize_affine_tensor(_p(x), C.c_int64(n), C.c_int(bits), _p(q), _p(s), _p(z))
    elif mode == MODE_DIM0:
        rows, cols = x.shape[0], n // x.shape[0]
        s = np.empty((1,) + x.shape[1:], np.float32)
        z = np.empty((1,) + x.shape[1:], np.float32)
        L.qo_quantize_affine_dim0(_p(x), C.c_int64(rows), C.c_int64(cols), C.c_int(bits), _p(q), _p(s), _p(z))
    else:
        assert n % block == 0
        s, z = np.empty(n // block, np.float32), np.empty(n // block, np.float32)
        L.qo_quantize_affine_block(_p(x), C.c_int64(n), C.c_int64(block), C.c_int(bits), _p(q), _p(s), _p(z))
    return q, s, z


def quantize4_block_pack(x, block=64):
    x = _f32(x)
    n = x.size
    assert n % block == 0 and block % 2 == 0
    packed = np.empty(n // 2, np.uint8)
    s, z = np.empty(n // block, np.float32), np.empty(n // block, np.float32)
    lib().qo_quantize4_block_pack(_p(x), C.c_int64(n), C.c_int64(block), _p(packed), _p(s), _p(z))
    return packed, s, z


def dequantize_affine(q, scale, zp, mode=MODE_TENSOR, p=0, packed=False, n=None):
    q = np.ascontiguousarray(q, np.uint8)
    scale, zp = _f32(scale), _f32(zp)
    if packed:
        n = q.size * 2 if n is None else n
        out = np.empty(n, np.float32)
        lib().qo_dequantize4_packed(_p(q), C.c_int64(n), C.c_int(mode), C.c_int64(p), _p(scale), _p(zp), _p(out))
        return out
    out = np.empty(q.shape, np.float32)
    lib().qo_dequantize_affine(_p(q), C.c_int64(q.size), C.c_int(mode), C.c_int64(p), _p(scale), _p(zp), _p(out))
    return out


def pack4(q):
    q = np.ascontiguousarray(q)
    if q.dtype != np.uint8:
        raise ValueError("Input tensor must be uint8")
    out = np.empty((q.size + 1) // 2, np.uint8)
    lib().qo_pack4(_p(q), C.c_int64(q.size), _p(out))
    return out


def unpack4(packed):
    p = np.ascontiguousarray(packed, np.uint8).reshape(-1)
    out = np.empty(p.size * 2, np.uint8)
    lib().qo_unpack4(_p(p), C.c_int64(p.size), _p(out))
    return out


def backend_quantize(x, bits=8, per_channel=False, symmetric=True):
    x = _f32(x)
    n = x.size
    rows = x.shape[0] if per_channel else 1
    cols = n // rows if per_channel else n
    q = np.empty(x.shape, np.uint8)
    shp = ((1,) + x.shape[1:]) if per_channel else ()
    s, z = np.empty(shp, np.float32), np.empty(shp, np.float32)
    lib().qo_backend_quantize(_p(x), C.c_int64(rows), C.c_int64(cols), C.c_int(int(per_channel)),
                              C.c_int(int(symmetric)), C.c_int(bits), _p(q), _p(s), _p(z))
    return q, s, z


def backend_dequantize(q, scale, zp, bits=8):
    q = np.ascontiguousarray(q, np.uint8)
    scale, zp = _f32(scale), _f32(zp)
    out = np.empty(q.shape, np.float32)
    lib().qo_backend_dequantize(_p(q), C.c_int64(q.size), C.c_int64(scale.size), C.c_int(bits),
                                _p(scale), _p(zp), _p(out))
    return out


def linear_dequant(x, wq, scale, zp, bias, bits, N, K, block=64):
    x = _f32(x)
    M = x.shape[0]
    wq = np.ascontiguousarray(wq, np.uint8)
    scale, zp = _f32(scale), _f32(zp)
    y = np.empty((M, N), np.float32)
    b = _f32(bias) if bias is not None else None
    lib().qo_linear_dequant(_p(x), _p(wq), _p(scale), _p(zp), _p(b) if b is not None else None, C.c_int(bits),
                            C.c_int64(M), C.c_int64(N), C.c_int64(K), C.c_int64(block), _p(y))
    return y
