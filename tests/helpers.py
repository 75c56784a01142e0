"""Comparison helpers shared by the CPU and GPU parity tests."""
import numpy as np


def bits_equal(a, b):
    """Bit-exact equality of two float32 arrays (NaNs with equal payload match)."""
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    b = np.ascontiguousarray(b, np.float32).reshape(-1)
    return a.shape == b.shape and bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))


def assert_f32_bits(a, b, what="", zero_sign_free=False):
    """float32 arrays must match bit for bit.  ``zero_sign_free`` lets -0.0 match
    +0.0 (the reference's min() returns either, see oracle/oracle_np.py)."""
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    b = np.ascontiguousarray(b, np.float32).reshape(-1)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    ua, ub = a.view(np.uint32).copy(), b.view(np.uint32).copy()
    if zero_sign_free:
        ua[ua == 0x80000000] = 0
        ub[ub == 0x80000000] = 0
    # all NaNs compare equal (payload/sign of a produced NaN is not part of the contract)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), f"{what}: NaN pattern differs"
    bad = (ua != ub) & ~na
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {a.size} float32 values differ; first at "
                           f"{int(np.argmax(bad))}: {a[np.argmax(bad)]!r} vs {b[np.argmax(bad)]!r}")


def assert_u8_equal(a, b, what=""):
    a = np.asarray(a).reshape(-1)
    b = np.asarray(b).reshape(-1)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    bad = a != b
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {a.size} codes differ; first at "
                           f"{int(np.argmax(bad))}: {a[np.argmax(bad)]} vs {b[np.argmax(bad)]}")


def ulp_diff(a, b):
    """Max distance in float32 ulps between two finite arrays."""
    a = np.ascontiguousarray(a, np.float32).reshape(-1)
    b = np.ascontiguousarray(b, np.float32).reshape(-1)

    def key(x):
        i = x.view(np.int32).astype(np.int64)
        return np.where(i < 0, -(i & 0x7FFFFFFF), i)
    return int(np.max(np.abs(key(a) - key(b)))) if a.size else 0
