"""CPU oracle (numpy) for the Quanta weight-quantization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``quanta_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs use it, and only as the checker.

This is a restatement, in plain numpy float32 arithmetic, of the algorithm the
reference implements with chains of eager torch ops.  The reference's
arithmetic lives in PyTorch ATen CPU kernels (third-party, pinned only as
``torch>=2.2.0`` in the reference's setup.py:20); the semantics restated here
were checked against torch 2.11 CPU by ``tests/golden/make_golden.py`` and are
pinned by ``tests/golden/quanta_golden.npz`` (see tests/test_oracle_golden.py):

* every intermediate is rounded to float32 (no FMA contraction),
* ``torch.round`` is round-half-to-even (``np.rint``),
* ``Tensor / int`` is a true IEEE divide by ``float32(int)``,
* ``int / Tensor`` is ``reciprocal(Tensor) * int`` (two roundings),
* float -> uint8 casts of NaN give 0 (x86 behaviour, SURVEY §8(b) "Errors").

Parity status
  rows A1-A4, B1-B2, P1-P2 : PINNED (golden vectors produced by the reference)
  rows G1-G2               : composition ``F.linear(x, dequant(q))`` — pinned
                             only through A4; the GEMM itself has no reference
  row  G3 (outlier split)  : PARITY UNPINNED — defined by this repository.

Signed zeros: the reference's ``tensor.min()`` returns ``-0.0`` or ``+0.0``
depending on ATen's vectorisation order when both are present; this oracle
orders ``-0.0 < +0.0`` (IEEE 754-2019 minimum/maximum), like the CUDA kernels.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

MODE_TENSOR = 0
MODE_DIM0 = 1
MODE_BLOCK = 2


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _key(x):
    """Order-preserving int64 key of float32 values with -0.0 < +0.0."""
    b = x.view(np.int32).astype(np.int64)
    return np.where(b < 0, -(b & 0x7FFFFFFF) - 1, b)


def _min_max(x, axis=None, keepdims=False):
    """NaN-propagating min/max with -0.0 < +0.0 (see module docstring)."""
    x = _f32(x)
    if x.size == 0:
        raise ValueError("min/max of an empty tensor")
    k = _key(x)
    imin = np.argmin(k, axis=axis, keepdims=True)
    imax = np.argmax(k, axis=axis, keepdims=True)
    if axis is None:
        mn = x.reshape(-1)[imin.reshape(-1)[0]]
        mx = x.reshape(-1)[imax.reshape(-1)[0]]
        if np.isnan(x).any():
            mn = mx = F32(np.nan)
        return F32(mn), F32(mx)
    mn = np.take_along_axis(x, imin, axis=axis)
    mx = np.take_along_axis(x, imax, axis=axis)
    nan = np.isnan(x).any(axis=axis, keepdims=True)
    mn = np.where(nan, F32(np.nan), mn).astype(np.float32)
    mx = np.where(nan, F32(np.nan), mx).astype(np.float32)
    if not keepdims:
        mn, mx = np.squeeze(mn, axis), np.squeeze(mx, axis)
    return mn, mx


def _to_u8(v):
    """float32 -> uint8 cast of already clamped values; NaN -> 0."""
    v = np.where(np.isnan(v), F32(0), v)
    return v.astype(np.int32).astype(np.uint8)


# --------------------------------------------------------------------------
# Convention A — Quanta/functional/quantization.py "linear"
# --------------------------------------------------------------------------

def quantize_affine(x, bits=8, mode=MODE_TENSOR, block=0):
    """Rows A1/A2/A3.  Follows Quanta/functional/quantization.py:185-210 (8-bit)
    and :73-99 (4-bit).

    mode TENSOR : per_channel=False branch (:197-203 / :85-91)
    mode DIM0   : per_channel=True branch  (:189-196 / :77-84), min/max over dim 0
    mode BLOCK  : the per_channel branch applied to ``x.reshape(-1, block).t()``
                   -> scale/zp of shape [N/block], codes in
                  the original (flat) order.
    Returns (q uint8 same shape as x, scale, zp); scale/zp are 0-dim for TENSOR,
    ``[1, *x.shape[1:]]`` for DIM0 and ``[N/block]`` for BLOCK.
    """
    with np.errstate(all="ignore"):
        x = _f32(x)
        L = F32(255 if bits == 8 else 15)
        if mode == MODE_TENSOR:
            mn, mx = _min_max(x)
            if mx == mn:                                 # :202-203
                mx = F32(mn + F32(1e-6))
            xv = x
        elif mode == MODE_DIM0:
            if x.ndim < 2:
                raise ValueError("per_channel needs a tensor with dim() > 1")
            mn, mx = _min_max(x, axis=0, keepdims=True)
            mx = np.where(mx == mn, (mn + F32(1e-6)).astype(np.float32), mx)   # :195-196
            xv = x
        elif mode == MODE_BLOCK:
            if block <= 0 or x.size % block:
                raise ValueError("numel must be a multiple of blocksize")
            xv = x.reshape(-1, block)
            mn, mx = _min_max(xv, axis=1, keepdims=True)
            mx = np.where(mx == mn, (mn + F32(1e-6)).astype(np.float32), mx)
        else:
            raise ValueError("bad mode")
        scale = ((mx - mn).astype(np.float32) / L).astype(np.float32)          # :205
        zp = mn                                                                # :206
        v = ((xv - mn).astype(np.float32) / scale).astype(np.float32)          # :209
        q = _to_u8(np.clip(np.rint(v), F32(0), L)).reshape(x.shape)
        if mode == MODE_BLOCK:
            scale, zp = scale.reshape(-1), np.asarray(zp).reshape(-1)
        return q, np.asarray(scale, np.float32), np.asarray(zp, np.float32)


def dequantize_affine(q, scale, zp, mode=MODE_TENSOR, block=0):
    """Row A4.  Quanta/functional/quantization.py:38 / :58:
    ``q.float() * scale + zp`` with the multiply and the add rounded separately."""
    with np.errstate(all="ignore"):
        q = np.asarray(q)
        scale, zp = _f32(scale), _f32(zp)
        if mode == MODE_BLOCK:
            qf = q.reshape(-1, block).astype(np.float32)
            out = ((qf * scale.reshape(-1, 1)).astype(np.float32) + zp.reshape(-1, 1)).astype(np.float32)
            return out.reshape(q.shape)
        qf = q.astype(np.float32)
        return ((qf * scale).astype(np.float32) + zp).astype(np.float32)


# --------------------------------------------------------------------------
# P1 / P2 — Quanta/utils/utils.py:23-48
# --------------------------------------------------------------------------

def pack4(q):
    """Row P1.  utils/utils.py:23-35: flatten, pad one zero if odd,
    ``byte[i] = q[2i] | (q[2i+1] << 4)`` in uint8 arithmetic (bits above 7 drop,
    inputs > 15 are not masked)."""
    q = np.ascontiguousarray(q)
    if q.dtype != np.uint8:
        raise ValueError("Input tensor must be uint8")
    flat = q.reshape(-1)
    if flat.size % 2:
        flat = np.concatenate([flat, np.zeros(1, np.uint8)])
    pair = flat.reshape(-1, 2)
    return (pair[:, 0] | ((pair[:, 1].astype(np.uint16) << 4) & 0xFF).astype(np.uint8)).astype(np.uint8)


def unpack4(packed):
    """Row P2.  utils/utils.py:37-48: ``out[2i] = b & 0xF ; out[2i+1] = b >> 4``;
    flat output of 2*len (keeps the pad element)."""
    p = np.ascontiguousarray(packed, dtype=np.uint8).reshape(-1)
    out = np.empty(p.size * 2, np.uint8)
    out[0::2] = p & 0x0F
    out[1::2] = (p >> 4) & 0x0F
    return out


def quantize4_block_pack(x, block=64):
    """Config 2: A3 (4-bit, blockwise) followed by P1."""
    q, s, z = quantize_affine(x, bits=4, mode=MODE_BLOCK, block=block)
    return pack4(q), s, z


# --------------------------------------------------------------------------
# Convention B — Quanta/backends/cpu/quantization.py
# --------------------------------------------------------------------------

def _isclose(a, b, rtol=1e-5, atol=1e-8):
    """torch.isclose(a, b) on float32 (ATen TensorCompare.cpp): equal, or
    finite ``|a-b|`` and ``|a-b| <= atol + |rtol*b|`` evaluated in float32."""
    a, b = _f32(a), _f32(b)
    allowed = (F32(atol) + np.abs((F32(rtol) * b).astype(np.float32))).astype(np.float32)
    actual = np.abs((a - b).astype(np.float32))
    return (a == b) | (np.isfinite(actual) & (actual <= allowed))


def backend_quantize(x, bits=8, per_channel=False, symmetric=True):
    """Row B1.  backends/cpu/quantization.py:10-59 (8-bit) and :86-135 (4-bit).
    ``scale`` is a multiplier (qmax/absmax); symmetric codes are offset by
    +128/+8; asymmetric zero-point is an integer-valued float."""
    with np.errstate(all="ignore"):
        x = _f32(x)
        Q = F32(127 if bits == 8 else 7)
        L = F32(255 if bits == 8 else 15)
        OFF = 128 if bits == 8 else 8
        if per_channel:
            if x.ndim < 2:
                raise ValueError("per_channel needs a tensor with dim() > 1")
            mn, mx = _min_max(x, axis=0, keepdims=True)
        else:
            mn, mx = _min_max(x)
            mn, mx = np.asarray(mn), np.asarray(mx)
        if bool(np.all(_isclose(mn, mx))):                                     # :38-39
            return np.zeros(x.shape, np.uint8), np.ones_like(mn), mn
        if symmetric:                                                          # :41-50
            am = np.maximum(np.abs(mn), np.abs(mx))
            am = np.where(np.isnan(mn) | np.isnan(mx), F32(np.nan), am).astype(np.float32)
            scale = ((F32(1) / am).astype(np.float32) * Q).astype(np.float32)
            zp = np.zeros_like(mn)
            v = np.clip(np.rint((x * scale).astype(np.float32)), -Q, Q)
            v = np.where(np.isnan(v), F32(0), v)
            q = (v.astype(np.int32) + OFF).astype(np.uint8)
        else:                                                                  # :51-57
            rng = (mx - mn).astype(np.float32)
            scale = ((F32(1) / rng).astype(np.float32) * L).astype(np.float32)
            zp = np.rint(((-mn) * scale).astype(np.float32)).astype(np.float32)
            v = ((x * scale).astype(np.float32) + zp).astype(np.float32)
            q = _to_u8(np.clip(np.rint(v), F32(0), L))
        return q, scale, zp


def backend_dequantize(q, scale, zp, bits=8):
    """Row B2.  backends/cpu/quantization.py:61-84 / :137-160: if every
    zero-point is close to 0 the codes are re-centred (``int8(q) - OFF``);
    then ``(q.float() - zp) / scale`` with a true divide."""
    with np.errstate(all="ignore"):
        q = np.asarray(q, np.uint8)
        scale, zp = _f32(scale), _f32(zp)
        OFF = 128 if bits == 8 else 8
        qi = q.astype(np.int32)
        if bool(np.all(_isclose(zp, np.zeros_like(zp)))):
            qi = qi.astype(np.int8).astype(np.int32) - OFF                     # int8 arithmetic, wraps
            qi = qi.astype(np.int8).astype(np.int32)
        qf = qi.astype(np.float32)
        return ((qf - zp).astype(np.float32) / scale).astype(np.float32)


# --------------------------------------------------------------------------
# Convention C — Quanta/functional/base.py (BaseQuantizer), SURVEY Appendix A.3
# --------------------------------------------------------------------------

def base_quantize(x, bits=8, per_channel=False, symmetric=True):
    """Row N4 (BaseQuantizer.quantize, base.py:11-59).  Like convention B in the symmetric case
    (``scale = max_val / abs_max`` = reciprocal * int, codes offset by 2^(bits-1)); asymmetric:
    ``scale = (2^bits - 1) / (max - min)``, ``zero_point = min`` (not rounded),
    ``q = clamp(round((x - zero_point) * scale), 0, 2^bits - 1)``.  If ``allclose(min, max)`` holds for ALL
    channels the parameters become ``(ones_like(min), min)`` and the codes are STILL computed with them
    (base.py:26-27, :43-57) — unlike convention B, which returns zero codes."""
    with np.errstate(all="ignore"):
        x = _f32(x)
        Q = F32(2 ** (bits - 1) - 1)
        L = F32(2 ** bits - 1)
        OFF = 2 ** (bits - 1)
        if per_channel:
            if x.ndim < 2:
                raise ValueError("per_channel needs a tensor with dim() > 1")   # tensor.min(dim=None, keepdim=True) raises
            mn, mx = _min_max(x, axis=0, keepdims=True)
        else:
            mn, mx = _min_max(x)
            mn, mx = np.asarray(mn), np.asarray(mx)
        if bool(np.all(_isclose(mn, mx))):                                     # :26-27
            scale, zp = np.ones_like(mn), mn
        elif symmetric:                                                        # :29-32
            am = np.maximum(np.abs(mn), np.abs(mx))
            am = np.where(np.isnan(mn) | np.isnan(mx), F32(np.nan), am).astype(np.float32)
            scale = ((F32(1) / am).astype(np.float32) * Q).astype(np.float32)
            zp = np.zeros_like(mn)
        else:                                                                  # :33-35
            rng = (mx - mn).astype(np.float32)
            scale = ((F32(1) / rng).astype(np.float32) * L).astype(np.float32)
            zp = mn
        if symmetric:                                                          # :44-49
            v = np.clip(np.rint((x * scale).astype(np.float32)), -Q, Q)
            v = np.where(np.isnan(v), F32(0), v)
            q = (v.astype(np.int32) + OFF).astype(np.uint8)
        else:                                                                  # :50-54
            v = ((x - zp).astype(np.float32) * scale).astype(np.float32)
            q = _to_u8(np.clip(np.rint(v), F32(0), L))
        return q, np.asarray(scale, np.float32), np.asarray(zp, np.float32)


def base_dequantize(q, scale, zp, bits=8, symmetric=True):
    """BaseQuantizer.dequantize (base.py:61-72): symmetric ``(int8(q) - 2^(bits-1)).float() / scale`` (int8
    arithmetic, wraps), else ``q.float() / scale + zero_point`` — true divides; which branch runs is decided by
    the quantizer's ``symmetric`` attribute, not by the zero-point values."""
    with np.errstate(all="ignore"):
        q = np.asarray(q, np.uint8)
        scale, zp = _f32(scale), _f32(zp)
        if symmetric:
            qi = (q.astype(np.int8).astype(np.int32) - 2 ** (bits - 1)).astype(np.int8).astype(np.float32)
            return (qi / scale).astype(np.float32)
        return ((q.astype(np.float32) / scale).astype(np.float32) + zp).astype(np.float32)


# --------------------------------------------------------------------------
# G1 / G2 — composition oracle for the quantized Linear layers
# --------------------------------------------------------------------------

def _round_to(a, dtype):
    """Round float32 values to bf16 / fp16 precision (round-to-nearest-even),
    returned as float32."""
    a = _f32(a)
    if dtype == "fp32":
        return a
    if dtype == "fp16":
        return a.astype(np.float16).astype(np.float32)
    if dtype == "bf16":
        b = a.view(np.uint32).astype(np.uint64)
        r = ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
        out = r.view(np.float32).copy()
        nan = np.isnan(a)
        out[nan] = np.nan
        return out
    raise ValueError(dtype)


def linear_dequant(x, q, scale, zp, bias=None, block=64, act_dtype="bf16"):
    """Rows G1/G2 (defined by composition):
    ``F.linear(x, dequantize_*bit(q, scale, zp).to(x.dtype), bias)`` with
    x[M,K] already representable in ``act_dtype``, weights dequantized by A4 in
    fp32 then rounded to ``act_dtype``, products accumulated in float64 (the
    exact value the fp32-accumulating kernels approximate)."""
    w = dequantize_affine(np.asarray(q).reshape(-1), scale, zp, MODE_BLOCK, block)
    w = _round_to(w, act_dtype).reshape(np.asarray(q).shape)
    y = x.astype(np.float64) @ w.astype(np.float64).T
    if bias is not None:
        y = y + np.asarray(bias, np.float64)
    return y


# --------------------------------------------------------------------------
# G3 — LLM.int8()-style outlier-split matmul.  PARITY UNPINNED: the reference
# only stores ``threshold`` (Quanta/nn/linear.py:20,25) and never uses it; this
# restatement is DEFINED BY THIS REPOSITORY.
# --------------------------------------------------------------------------

def rowwise_quantize_sym(w):
    """Static weight codes for G3: B1-symmetric 8-bit with one scale per OUTPUT
    row, i.e. ``quantize_8bit_cpu(w.t(), per_channel=True, symmetric=True)``
    (backends/cpu/quantization.py:29-50) transposed back and re-centred:
    returns (qw int8 [N,K] = code-128, cw float32 [N] multiplier 127/absmax)."""
    q, scale, _ = backend_quantize(_f32(w).T, bits=8, per_channel=True, symmetric=True)
    return (q.T.astype(np.int32) - 128).astype(np.int8), scale.reshape(-1)


def int8_outlier_matmul(x, qw, cw, threshold=6.0, bias=None, act_dtype="bf16"):
    """y = x @ W.T (+bias) with the LLM.int8() decomposition over static int8
    weight codes ``qw`` [N,K] and per-row multipliers ``cw`` [N]
    (see rowwise_quantize_sym):

    * outlier feature columns  J = { j : max_i |x[i,j]| > threshold }
      go through a 16-bit matmul against the dequantized weight columns
      ``round_act(qw[:,J] / cw[:,None])`` (B2 arithmetic: true divide),
    * the other columns use int8 x int8 -> int32: x is quantized per row over
      the non-outlier columns with B1-symmetric arithmetic
      (``cx[i] = rcp(absmax_i) * 127``, ``qx = clamp(rint(x*cx), -127, 127)``,
      outlier columns contribute code 0; an all-zero row gets cx = 1),
      and the int32 accumulator is rescaled as ``float(acc) / (cx[i]*cw[n])``.

    Returns (y float64 [M,N], J).  x [M,K] holds values representable in
    ``act_dtype``.
    """
    with np.errstate(all="ignore"):
        x = _f32(x)
        qw = np.asarray(qw, np.int8)
        cw = _f32(cw).reshape(-1)
        colmax = np.max(np.abs(x), axis=0)
        outl = colmax > F32(threshold)
        J = np.nonzero(outl)[0]
        xr = np.where(outl[None, :], F32(0), x)
        am = np.max(np.abs(xr), axis=1, keepdims=True).astype(np.float32)
        cx = ((F32(1) / am).astype(np.float32) * F32(127)).astype(np.float32)
        cx = np.where(am == 0, F32(1), cx).astype(np.float32)
        qx = np.clip(np.rint((xr * cx).astype(np.float32)), -127, 127).astype(np.float64)
        # integer products summed in float64: every partial sum is an integer < 2^53, hence exact
        # (|acc| <= 127 * 127 * K fits int32), and BLAS makes it fast
        acc = qx @ qw.astype(np.float64).T
        denom = (cx * cw.reshape(1, -1)).astype(np.float32)
        y = (acc.astype(np.float32) / denom).astype(np.float32).astype(np.float64)
        if J.size:
            wo = (qw[:, J].astype(np.float32) / cw.reshape(-1, 1)).astype(np.float32)
            wo = _round_to(wo, act_dtype)
            y = y + x[:, J].astype(np.float64) @ wo.astype(np.float64).T
        if bias is not None:
            y = y + np.asarray(bias, np.float64)
        return y, J


# --------------------------------------------------------------------------
# N1 — NF4 codebook quantization (Quanta/functional/quantization.py:101-118, :59-61)
# --------------------------------------------------------------------------

NF4_LEVELS = np.array([
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453,
    -0.28444138169288635, -0.18477343022823334, -0.09105003625154495, 0.0,
    0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
    0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0], dtype=np.float32)


def quantize_nf4(x, block=None):
    """``quantize_4bit_nf4`` (:101-118): abs_max = max|x| (per tensor, or per block of
    ``block`` flat elements — the reference function applied to each block), normalized =
    x / abs_max (true divide), index = argmin_l |normalized - level_l| (first index on ties; a NaN
    distance — 0/0 for an all-zero tensor — wins, i.e. index 0).  Returns (idx uint8, absmax)."""
    with np.errstate(all="ignore"):
        x = _f32(x)
        flat = x.reshape(-1) if block is None else x.reshape(-1, block)
        ax = np.abs(flat)
        am = np.max(ax, axis=-1, keepdims=True).astype(np.float32)
        am = np.where(np.any(np.isnan(ax), axis=-1, keepdims=True), F32(np.nan), am).astype(np.float32)
        normalized = (flat / am).astype(np.float32)
        dist = np.abs((normalized[..., None] - NF4_LEVELS).astype(np.float32))
        idx = np.argmin(dist, axis=-1).astype(np.uint8)
        return idx.reshape(x.shape), (am.reshape(()) if block is None else am.reshape(-1))


def dequantize_nf4(idx, absmax, block=None):
    """``dequantize_4bit(..., quant_type="nf4")`` (:59-61): levels[q] * abs_max."""
    idx = np.asarray(idx)
    lv = NF4_LEVELS[idx.astype(np.int64)]
    if block is None:
        return (lv * F32(absmax)).astype(np.float32)
    return (lv.reshape(-1, block) * _f32(absmax).reshape(-1, 1)).astype(np.float32).reshape(idx.shape)


# --------------------------------------------------------------------------
# Row N4: nf8 / fp4 / fp8 (Quanta/functional/quantization.py:120-183, :39-49, :62-69)
# --------------------------------------------------------------------------
# The decisions that depend on torch's own libm (tanh for the nf8 levels, log2 for the fp exponent
# field — neither is correctly rounded) are taken from tables derived by running the unmodified
# reference (tests/golden/make_tables_n4.py): the level values, and for each exponent field the
# smallest |x| that reaches it.  Everything else is restated arithmetic.

def _n4_tables():
    import os
    global _N4
    try:
        return _N4
    except NameError:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                            "quanta_tables_n4.npz")
        _N4 = dict(np.load(path))
        return _N4


def nf8_levels():
    return _n4_tables()["nf8_levels"].astype(np.float32)


def quantize_nf8(x, block=None):
    """``quantize_8bit_nf8`` (:170-183): as NF4 with the 256 tanh levels — the full distance matrix and
    ``argmin`` (first index on ties), NOT the thresholds the CUDA kernel uses."""
    lv = nf8_levels()
    with np.errstate(all="ignore"):
        x = _f32(x)
        flat = x.reshape(-1) if block is None else x.reshape(-1, block)
        ax = np.abs(flat)
        am = np.max(ax, axis=-1, keepdims=True).astype(np.float32)
        am = np.where(np.any(np.isnan(ax), axis=-1, keepdims=True), F32(np.nan), am).astype(np.float32)
        normalized = (flat / am).astype(np.float32).reshape(-1)
        idx = np.empty(normalized.shape, np.uint8)
        for i in range(0, normalized.size, 1 << 16):              # bounded N x 256 temporaries
            nb = normalized[i:i + (1 << 16)]
            idx[i:i + nb.size] = np.argmin(np.abs((nb[:, None] - lv).astype(np.float32)), axis=-1)
        return idx.reshape(x.shape), (am.reshape(()) if block is None else am.reshape(-1))


def dequantize_nf8(idx, absmax, block=None):
    """``dequantize_8bit(..., quant_type="nf8")`` (:39-41): levels[q] * abs_max."""
    idx = np.asarray(idx)
    lv = nf8_levels()[idx.astype(np.int64)]
    if block is None:
        return (lv * F32(absmax)).astype(np.float32)
    return (lv.reshape(-1, block) * _f32(absmax).reshape(-1, 1)).astype(np.float32).reshape(idx.shape)


def quantize_fp(x, bits):
    """``quantize_4bit_fp4`` (:120-144) / ``quantize_8bit_fp8`` (:146-168): sign | exponent field |
    mantissa.  field = clamp(round(log2(|x| + (|x| == 0)) + bias), 0, E) via the derived thresholds;
    mantissa = clamp(round(|x| / 2^(field - bias) * M - M), 0, M - 1) in float32 (M = 2 / 8: the fp4
    form ``round(v - 1)`` is the same expression with M = 1 ... clamp(0, 1))."""
    t = _n4_tables()
    bias, emax, thr = (1, 3, t["fp4_exp_thresholds"]) if bits == 4 else (7, 15, t["fp8_exp_thresholds"])
    with np.errstate(all="ignore"):
        x = _f32(x)
        a = np.abs(x)
        a0 = np.where(a == 0, F32(1.0), a).astype(np.float32)
        field = np.searchsorted(thr.astype(np.float32), a0, side="right").astype(np.int64)      # #{k : a0 >= thr[k]}
        field = np.minimum(field, emax)
        scaled = (a / np.exp2((field - bias).astype(np.float32))).astype(np.float32)
        if bits == 4:
            m = np.clip(np.rint((scaled - F32(1.0)).astype(np.float32)), 0, 1)
            code = (field << 1) | m.astype(np.int64)
            code = np.where(x < 0, code | 0x8, code)
        else:
            m = np.clip(np.rint(((scaled * F32(8.0)).astype(np.float32) - F32(8.0)).astype(np.float32)), 0, 7)
            code = (field << 3) | m.astype(np.int64)
            code = np.where(x < 0, code | 0x80, code)
        return code.astype(np.uint8)


def dequantize_fp(q, bits, bias):
    """fp4 (:62-69): (1 + m) * 2^(e - bias) * sign;  fp8 (:42-49): (1 + m / 8) * 2^(e - bias) * sign."""
    q = np.asarray(q).astype(np.int64)
    if bits == 4:
        sign, e, frac = q & 0x8, (q >> 1) & 0x3, (q & 1).astype(np.float32)
    else:
        sign, e, frac = q & 0x80, (q >> 3) & 0xF, ((q & 7).astype(np.float32) / F32(8.0)).astype(np.float32)
    val = ((F32(1.0) + frac).astype(np.float32) * np.exp2((e - int(bias)).astype(np.float32))).astype(np.float32)
    return np.where(sign != 0, -val, val).astype(np.float32)
