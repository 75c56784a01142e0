"""The CUDA backend entry the reference's dispatcher looks for.

``Quanta/backends/__init__.py:17-23`` does

    from .cuda.quantization import (quantize_8bit_cuda, dequantize_8bit_cuda,
                                    quantize_4bit_cuda, dequantize_4bit_cuda)

and forwards ``quantize_*bit(tensor, per_channel, symmetric)`` /
``dequantize_*bit(q_tensor, scale, zero_point)`` to these names for CUDA
tensors (:58-62, :80-84, :102-106, :124-128).  Semantics are those of
``Quanta/backends/cpu/quantization.py`` (convention B): ``scale`` is a
multiplier, symmetric codes are offset by +128 / +8, the asymmetric zero-point
is an integer-valued float, and the ``allclose(min, max)`` early-out and the
``allclose(zero_point, 0)`` decode switch are reproduced (on the device — no
host synchronisation).
"""
from __future__ import annotations

import torch

from ... import _host, _lib


def _quantize(tensor, per_channel, symmetric, bits):
    _host.require_cuda(tensor)
    x = tensor.detach()
    if not x.is_contiguous():                       # cpu/quantization.py:26-27
        x = x.contiguous()
    code = _host.dtype_code(x)
    if x.numel() == 0:
        raise RuntimeError("min(): cannot quantize an empty tensor")
    dev = x.device
    if per_channel:
        if x.dim() < 2:
            raise ValueError("per_channel=True needs a tensor with dim() > 1")
        rows, cols = _host.rows_cols(x)
        pshape = (1,) + tuple(x.shape[1:])
    else:
        rows, cols, pshape = 1, x.numel(), ()
    nparam = cols if per_channel else 1
    with _host.device_guard(dev):
        q = torch.empty(x.shape, dtype=torch.uint8, device=dev)
        scale = torch.empty(nparam, dtype=torch.float32, device=dev)
        zp = torch.empty(nparam, dtype=torch.float32, device=dev)
        ws = _host.quantize_workspace(dev, _host.workspace_bytes(_lib.OP_BACKEND_QUANTIZE, rows if per_channel else 1,
                                                                    cols if per_channel else 1))
        st = _lib.lib().quanta_backend_quantize(x.data_ptr(), code, rows, cols, int(bool(per_channel)),
                                                int(bool(symmetric)), bits, q.data_ptr(), scale.data_ptr(),
                                                zp.data_ptr(), ws.data_ptr(), ws.numel(), _host.stream_ptr(dev))
    _lib.check(st, "quanta_backend_quantize")
    return q, scale.reshape(pshape), zp.reshape(pshape)


def _dequantize(q_tensor, scale, zero_point, bits):
    _host.require_cuda(q_tensor, "q_tensor")
    dev = q_tensor.device
    q = q_tensor.detach()
    if q.dtype != torch.uint8:
        q = q.to(torch.uint8)
    if not q.is_contiguous():                       # cpu/quantization.py:77-78
        q = q.contiguous()
    scale = torch.as_tensor(scale, dtype=torch.float32, device=dev).contiguous()
    zp = torch.as_tensor(zero_point, dtype=torch.float32, device=dev).contiguous()
    n = q.numel()
    out = torch.empty(q.shape, dtype=torch.float32, device=dev)
    if n == 0:
        return out
    nchan = max(scale.numel(), zp.numel())
    if scale.numel() != nchan:
        scale = scale.expand(nchan).contiguous()
    if zp.numel() != nchan:
        zp = zp.expand(nchan).contiguous()
    rows, cols = _host.rows_cols(q) if nchan > 1 else (1, n)
    if nchan not in (1, cols):
        raise ValueError(f"scale of {nchan} elements does not broadcast over codes of shape {tuple(q.shape)}")
    with _host.device_guard(dev):
        ws = _host.workspace(dev, 256)
        st = _lib.lib().quanta_backend_dequantize(q.data_ptr(), rows, cols, nchan, bits, scale.data_ptr(),
                                                  zp.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                  _host.stream_ptr(dev))
    _lib.check(st, "quanta_backend_dequantize")
    return out


def quantize_8bit_cuda(tensor, per_channel=False, symmetric=True):
    """CUDA twin of quantize_8bit_cpu (backends/cpu/quantization.py:10-59)."""
    return _quantize(tensor, per_channel, symmetric, 8)


def dequantize_8bit_cuda(q_tensor, scale, zero_point):
    """CUDA twin of dequantize_8bit_cpu (backends/cpu/quantization.py:61-84)."""
    return _dequantize(q_tensor, scale, zero_point, 8)


def quantize_4bit_cuda(tensor, per_channel=False, symmetric=True):
    """CUDA twin of quantize_4bit_cpu (backends/cpu/quantization.py:86-135)."""
    return _quantize(tensor, per_channel, symmetric, 4)


def dequantize_4bit_cuda(q_tensor, scale, zero_point):
    """CUDA twin of dequantize_4bit_cpu (backends/cpu/quantization.py:137-160)."""
    return _dequantize(q_tensor, scale, zero_point, 4)
