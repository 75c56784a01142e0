"""Drop-in CUDA version of ``Quanta.functional.base.BaseQuantizer`` (Quanta/functional/base.py:5-72) — the third
``(q, scale, zero_point)`` convention of the reference:

    bq = BaseQuantizer(num_bits=8, symmetric=True)
    q, scale, zero_point = bq.quantize(tensor, per_channel=False)
    x = bq.dequantize(q, scale, zero_point)

Same constructor, method names, return layout (uint8 codes of the input's shape; 0-dim parameters per tensor,
``[1, *shape[1:]]`` per channel) and arithmetic as the reference, computed by ``quanta_base_quantize`` /
``quanta_base_dequantize`` (include/quanta_b200.h).  CUDA tensors only — there is no CPU fallback.
"""
from __future__ import annotations

import torch

from .. import _host, _lib


class BaseQuantizer:
    def __init__(self, num_bits: int = 8, symmetric: bool = True):
        if num_bits not in (4, 8):
            raise NotImplementedError("the CUDA BaseQuantizer implements num_bits = 8 and 4 (what the reference's "
                                      "own callers use, Quanta/functional/tensor_ops.py:19-29)")
        self.num_bits = num_bits
        self.symmetric = symmetric
        self.max_val = 2 ** (num_bits - 1) - 1 if symmetric else 2 ** num_bits - 1      # base.py:9

    def quantize(self, tensor, per_channel: bool = False):
        _host.require_cuda(tensor)
        x = tensor.detach()
        if not x.is_contiguous():                                  # base.py:41-42
            x = x.contiguous()
        code = _host.dtype_code(x)
        if x.numel() == 0:
            raise RuntimeError("min(): cannot quantize an empty tensor")
        if per_channel and x.dim() < 2:
            # the reference calls tensor.min(dim=None, keepdim=True) here, which raises (base.py:18-19)
            raise ValueError("per_channel=True needs a tensor with dim() > 1")
        dev = x.device
        if per_channel:
            rows, cols = _host.rows_cols(x)
            pshape = (1,) + tuple(x.shape[1:])
        else:
            rows, cols, pshape = 1, x.numel(), ()
        nparam = cols if per_channel else 1
        L = _lib.lib()
        with _host.device_guard(dev):
            q = torch.empty(x.shape, dtype=torch.uint8, device=dev)
            scale = torch.empty(nparam, dtype=torch.float32, device=dev)
            zp = torch.empty(nparam, dtype=torch.float32, device=dev)
            ws = _host.workspace(dev, _host.workspace_bytes(_lib.OP_BASE_QUANTIZE, rows, cols))
            st = L.quanta_base_quantize(x.data_ptr(), code, rows, cols, int(bool(per_channel)), int(bool(self.symmetric)),
                                        self.num_bits, q.data_ptr(), scale.data_ptr(), zp.data_ptr(), ws.data_ptr(),
                                        ws.numel(), _host.stream_ptr(dev))
        _lib.check(st, "quanta_base_quantize")
        return q, scale.reshape(pshape), zp.reshape(pshape)

    def dequantize(self, q_tensor, scale, zero_point):
        _host.require_cuda(q_tensor, "q_tensor")
        q = q_tensor.detach()
        if q.dtype != torch.uint8:
            q = q.to(torch.uint8)
        if not q.is_contiguous():                                  # base.py:65-66
            q = q.contiguous()
        dev = q.device
        out = torch.empty(q.shape, dtype=torch.float32, device=dev)
        if q.numel() == 0:
            return out
        scale = torch.as_tensor(scale, dtype=torch.float32, device=dev).reshape(-1).contiguous()
        zp = torch.as_tensor(zero_point, dtype=torch.float32, device=dev).reshape(-1).contiguous()
        nchan = max(scale.numel(), zp.numel())
        if scale.numel() != nchan:
            scale = scale.expand(nchan).contiguous()
        if zp.numel() != nchan:
            zp = zp.expand(nchan).contiguous()
        if nchan > 1:
            rows, cols = _host.rows_cols(q)
            if cols != nchan:
                raise ValueError(f"scale of {nchan} values does not broadcast over codes of shape {tuple(q.shape)} like the "
                                 "reference's per_channel results ([1, *shape[1:]])")
        else:
            rows, cols = 1, q.numel()
        with _host.device_guard(dev):
            st = _lib.lib().quanta_base_dequantize(q.data_ptr(), rows, cols, nchan, self.num_bits, int(bool(self.symmetric)),
                                                   scale.data_ptr(), zp.data_ptr(), out.data_ptr(), _host.stream_ptr(dev))
        _lib.check(st, "quanta_base_dequantize")
        return out
